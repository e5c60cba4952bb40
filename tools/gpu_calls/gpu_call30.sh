#!/bin/bash
O=gpurun_out/c30; mkdir -p $O
timeout 1200 python bench.py --seconds 240 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-batch > $O/bench_240s.json 2> $O/bench_240s.err; echo "bench240 rc=$?" >> $O/rc.txt
cat $O/rc.txt
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c30/bench_240s.json").read().strip().splitlines()[-1])
print("value", d["value"], d["frames_per_s"], d["kv_len_end"], d["phases_ms_per_chunk"], "step", d["roofline_step"]["ms_per_step"], d["roofline_step"]["frac"])
PY
