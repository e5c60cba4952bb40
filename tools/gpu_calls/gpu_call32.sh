#!/bin/bash
O=gpurun_out/c32; mkdir -p $O
timeout 900 python -m pytest tests/test_resize_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 800 > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-batch > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 3 $O/tests.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c32/bench_native.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["p50_frame_latency_ms"], d["p50_frame_latency_ms"], d["phases_ms_per_chunk"], "step", d["roofline_step"]["ms_per_step"], d["roofline_step"]["frac"], "cpu", d["cpu_baseline"]["value"], d["ingest"]["host_torchvision_ms"])
PY
