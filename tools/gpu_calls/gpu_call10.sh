#!/bin/bash
set -u
O=gpurun_out/${OUT:-c10}
mkdir -p $O
timeout 900 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 800 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
CASES=1:1000,1:17000,4:8000,8:8000 timeout 500 python tools/bench_mega_phases.py > $O/phases.txt 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -n 6 $O/mega_tests.log | cut -c1-300
grep -i "livecc_b200\]\|not supported" $O/*.log $O/*.txt $O/*.err | head -5
cat $O/phases.txt
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_native.json').read().strip().splitlines()[-1])
    print(d['value'], d['phases_ms_per_chunk'], 'kernel frac', d['roofline']['frac'], d['roofline']['us_per_launch'], 'in situ', d['roofline_step']['ms_per_step'], d['roofline_step']['frac'])
    print(d.get('multi_stream'))
except Exception as e:
    print('bench parse failed', e)
PY
grep -v CUDAEvent $O/bench_native.err | tail -4
