#!/bin/bash
O=gpurun_out/c20; mkdir -p $O
timeout 900 python -m pytest tests/test_resize_gpu.py -m gpu -q --timeout 800 > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_resize_gpu.py -m gpu -q -x --timeout 500 -k "odd_sizes or tiny_src or one_px or odd_up or alignment" > $O/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> $O/rc.txt
timeout 600 python tools/bench_resize.py --sweep > $O/bench_resize.txt 2>&1; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 8 $O/sanitizer.log | cut -c1-300; tail -n 25 $O/tests.log | cut -c1-300; cat $O/bench_resize.txt | cut -c1-250
