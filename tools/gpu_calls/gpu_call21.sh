#!/bin/bash
O=gpurun_out/c21; mkdir -p $O
timeout 900 python -m pytest tests/test_resize_gpu.py -m gpu -q --timeout 800 > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 600 python tools/bench_resize.py --sweep > $O/bench_resize.txt 2>&1; echo "bench rc=$?" >> $O/rc.txt
CUBLAS=1 timeout 300 python tools/bench_gemm_shapes.py > $O/gemm_vs_cublas.txt 2>&1
cat $O/rc.txt; tail -n 5 $O/tests.log | cut -c1-300; cat $O/bench_resize.txt | cut -c1-250; cat $O/gemm_vs_cublas.txt | cut -c1-250
