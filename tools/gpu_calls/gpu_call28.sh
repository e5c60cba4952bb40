#!/bin/bash
O=gpurun_out/c28; mkdir -p $O
timeout 900 python -m pytest tests/test_resize_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 800 > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 4 $O/tests.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c28/bench_native.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["p50_frame_latency_ms"], d["phases_ms_per_chunk"], "step", d["roofline_step"]["ms_per_step"], d["roofline_step"]["frac"], "gateup", d["roofline"]["us_per_launch"], d["roofline"]["frac"])
print(d["gpu_launches"], d["clocks"])
PY
