#!/bin/bash
O=gpurun_out/c13; mkdir -p $O
timeout 900 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 800 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
KV=8000 timeout 400 python tools/trace_mega.py > $O/trace_kv8000.txt 2>&1
PHASES=full CASES=1:1000,1:8000,1:17000,4:8000,8:8000 timeout 500 python tools/bench_mega_phases.py > $O/phases.txt 2>&1
cat $O/rc.txt; tail -n 3 $O/mega_tests.log | cut -c1-200
tail -11 $O/trace_kv8000.txt; grep "B=\|full" $O/phases.txt
