#!/bin/bash
O=gpurun_out/c16; mkdir -p $O
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_mega.json 2> $O/bench_mega.err; echo "bench mega rc=$?" >> $O/rc.txt
LIVECC_B200_MEGA=0 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batch --no-e2e > $O/bench_perop.json 2> $O/bench_perop.err; echo "bench perop rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<PY
import json
for f in ('bench_mega','bench_perop'):
    try:
        d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['phases_ms_per_chunk'], 'kernel', d['roofline'].get('frac'), d['roofline'].get('us_per_launch'), 'in situ', d['roofline_step']['ms_per_step'], d['roofline_step']['frac'])
        print(d.get('multi_stream'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
grep -v CUDAEvent $O/bench_perop.err | tail -4
