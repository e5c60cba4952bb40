#!/bin/bash
O=gpurun_out/c18; mkdir -p $O
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_mega_gpu.py -m gpu -q --timeout 800 > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 300 python tools/bench_decode.py --kv 8192 --steps 64 > $O/bench_decode.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
timeout 300 python tools/bench_gemm_shapes.py > $O/gemm_shapes.txt 2>&1
cat $O/rc.txt; tail -n 4 $O/tests.log | cut -c1-300
tail -6 $O/bench_decode.txt | cut -c1-250
tail -c 2500 $O/bench_native.json
grep -v "^$" $O/gemm_shapes.txt | head -30
