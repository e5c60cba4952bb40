#!/bin/bash
O=gpurun_out/c15; mkdir -p $O
timeout 900 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 800 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
KV=8000 timeout 400 python tools/trace_attn.py > $O/trace_attn.txt 2>&1
KV=8000 timeout 400 python tools/trace_mega.py > $O/trace_kv8000.txt 2>&1
PHASES=full CASES=1:1000,1:8000,1:17000,4:8000 timeout 500 python tools/bench_mega_phases.py > $O/phases.txt 2>&1
cat $O/rc.txt; tail -n 3 $O/mega_tests.log | cut -c1-200
tail -12 $O/trace_attn.txt | head -4; tail -3 $O/trace_attn.txt
tail -6 $O/trace_kv8000.txt; grep "B=\|full" $O/phases.txt
