#!/bin/bash
# bring-up of the persistent decode kernel: phase-by-phase report first (never asserts), then the asserting tests
set -u
mkdir -p gpurun_out/c2
O=gpurun_out/c2
MEGA_DEBUG_REPORT=1 timeout 600 python -m pytest tests/test_mega_gpu.py -m gpu -q -s -k "phases" --timeout 500 > $O/mega_report.log 2>&1; echo "report rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 500 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_ops_gpu.py -m gpu -q --timeout 800 > $O/tests_small.log 2>&1; echo "tests_small rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
grep "\[mega\]" $O/mega_report.log | head -80
tail -n 5 $O/mega_tests.log
tail -n 3 $O/tests_small.log
