#!/bin/bash
set -u
mkdir -p gpurun_out/c7
O=gpurun_out/c7
MEGA_DEBUG_REPORT=1 timeout 600 python -m pytest tests/test_mega_gpu.py -m gpu -q -s -k "phases and wide" --timeout 500 > $O/mega_report.log 2>&1; echo "report rc=$?" >> $O/rc.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_mega_gpu.py -m gpu -q -s -k "batched and wide" --timeout 800 > $O/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> $O/rc.txt
cat $O/rc.txt
grep "\[mega\]" $O/mega_report.log | grep -B1 "wrong\|bad [1-9]" | head -60
grep -i "error\|invalid\|out of bounds\|=========" $O/sanitizer.log | head -40
tail -n 5 $O/sanitizer.log | cut -c1-300
