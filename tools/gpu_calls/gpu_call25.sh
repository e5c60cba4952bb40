#!/bin/bash
O=gpurun_out/c25; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 6 $O/tests_all.log | cut -c1-300; tail -2 $O/smoke.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c25/bench_native.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["phases_ms_per_chunk"], "step", d["roofline_step"]["ms_per_step"], d["roofline_step"]["frac"], "gateup", d["roofline"]["us_per_launch"], d["roofline"]["frac"])
print(d.get("ingest")); print(d.get("multi_stream")); print(d["gpu_launches"], d["clocks"])
PY
