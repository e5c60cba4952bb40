#!/bin/bash
# persistent decode kernel: phase report + asserting tests (small + 7B widths), whole GPU suite, bench, ncu of the kernel
set -u
mkdir -p gpurun_out/c3
O=gpurun_out/c3
MEGA_DEBUG_REPORT=1 timeout 600 python -m pytest tests/test_mega_gpu.py -m gpu -q -s -k "phases" --timeout 500 > $O/mega_report.log 2>&1; echo "report rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 500 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_ops_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout 800 > $O/tests_small.log 2>&1; echo "tests_small rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests/test_parity_7b_gpu.py -m gpu -q -s --timeout 1400 > $O/tests_7b.log 2>&1; echo "tests_7b rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 40 -c 1 -o $O/ncu_mega -f python bench.py --seconds 20 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-batch > $O/ncu_mega.log 2>&1; echo "ncu_mega rc=$?" >> $O/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --seconds 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-batch > $O/launches.log 2>&1; echo "launches rc=$?" >> $O/rc.txt
cat $O/rc.txt
grep "\[mega\]" $O/mega_report.log | awk '{ if ($0 ~ /bad [1-9]/ || $0 ~ /nan [1-9]/) print }' | head -40
tail -n 5 $O/mega_tests.log
tail -n 3 $O/tests_small.log
tail -n 3 $O/tests_7b.log
tail -c 1500 $O/bench_native.json
