#!/bin/bash
set -u
mkdir -p gpurun_out/c4
O=gpurun_out/c4
timeout 900 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 800 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 500 -k "mcq or pretrained or sharp" > $O/tests_new.log 2>&1; echo "tests_new rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
timeout 900 python tools/bench_mcq_prefill.py --batches 1,2,4,8,16 > $O/mcq_sweep.md 2> $O/mcq_sweep.err; echo "mcq rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -n 12 $O/mega_tests.log | cut -c1-300
tail -n 5 $O/tests_new.log
grep -i "livecc_b200\]" $O/*.log $O/*.err | head
tail -c 2500 $O/bench_native.json
cat $O/mcq_sweep.md
