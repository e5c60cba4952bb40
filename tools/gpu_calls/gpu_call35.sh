#!/bin/bash
O=gpurun_out/c35; mkdir -p $O
timeout 900 python -m pytest tests/test_resize_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout 800 > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 3 $O/tests.log | cut -c1-300
