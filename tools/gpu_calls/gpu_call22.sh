#!/bin/bash
O=gpurun_out/c22; mkdir -p $O
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 800 > $O/tests_gemm.log 2>&1; echo "tests_gemm rc=$?" >> $O/rc.txt
CUBLAS=1 timeout 300 python tools/bench_gemm_shapes.py > $O/gemm_vs_cublas.txt 2>&1
SWEEP_BN=1 timeout 400 python tools/bench_gemm_shapes.py > $O/gemm_bn_sweep.txt 2>&1
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 800 > $O/tests_engine.log 2>&1; echo "tests_engine rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-batch > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 4 $O/tests_gemm.log | cut -c1-300; tail -n 4 $O/tests_engine.log | cut -c1-300; cat $O/gemm_vs_cublas.txt | cut -c1-250
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c22/bench_native.json").read().strip().splitlines()[-1])
print(d["value"], d["phases_ms_per_chunk"], d["roofline_step"]["ms_per_step"], d.get("ingest"))
PY
grep -v "^$" $O/gemm_bn_sweep.txt | cut -c1-120
