#!/bin/bash
set -u
mkdir -p gpurun_out/c6
O=gpurun_out/c6
timeout 900 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 800 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 60 -c 1 -o $O/ncu_mega -f python bench.py --seconds 20 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-batch > $O/ncu_mega.log 2>&1; echo "ncu_mega rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -n 8 $O/mega_tests.log | cut -c1-300
grep -i "livecc_b200\]" $O/*.log $O/*.err | head
tail -c 3000 $O/bench_native.json
