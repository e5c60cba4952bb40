#!/bin/bash
set -u
mkdir -p gpurun_out/c9
O=gpurun_out/c9
timeout 900 python -m pytest tests/test_mega_gpu.py -m gpu -q -s --timeout 800 > $O/mega_tests.log 2>&1; echo "mega_tests rc=$?" >> $O/rc.txt
for la in 0 32 64 160; do
  LIVECC_B200_MEGA_LOOKAHEAD=$la CASES=1:1000,1:17000,4:8000 timeout 400 python tools/bench_mega_phases.py > $O/phases_la$la.txt 2>&1
done
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -n 6 $O/mega_tests.log | cut -c1-300
for la in 0 32 64 160; do grep -v "err=0" $O/phases_la$la.txt | head -3; grep "look-ahead\|B=\|qkv\|o_proj\|down\|layers\|full" $O/phases_la$la.txt; done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c9/bench_native.json').read().strip().splitlines()[-1])
print(d['value'], d['phases_ms_per_chunk'], d['roofline']['frac'], d['roofline_step']['ms_per_step'], d['roofline_step']['frac'], d.get('multi_stream'))
PY
