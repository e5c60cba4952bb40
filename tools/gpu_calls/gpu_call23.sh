#!/bin/bash
O=gpurun_out/c23; mkdir -p $O
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 120 > $O/tests_gemm.log 2>&1; echo "tests_gemm rc=$?" >> $O/rc.txt
SWEEP_BN=1 timeout 500 python tools/bench_gemm_shapes.py > $O/gemm_pair_sweep.txt 2>&1; echo "sweep rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 30 $O/tests_gemm.log | cut -c1-300; grep -v "^$" $O/gemm_pair_sweep.txt | cut -c1-150
