#!/bin/bash
set -u
mkdir -p gpurun_out/c8
O=gpurun_out/c8
timeout 900 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 800 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
timeout 600 python tools/bench_mega_phases.py > $O/mega_phases.txt 2>&1; echo "phases rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -n 6 $O/mega_tests.log | cut -c1-300
cat $O/mega_phases.txt
tail -c 3000 $O/bench_native.json
grep -v CUDAEvent $O/bench_native.err | tail -5
