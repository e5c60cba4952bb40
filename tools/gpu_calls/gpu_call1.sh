#!/bin/bash
# Round-2 GPU call #1: full GPU test suite (new asserting parity tests), validation of the two opt-in paths written
# without GPU time in round 1, FA2 head-to-head, GEMM shape timings + ncu, a short native bench and the hf_gpu arm.
set -u
mkdir -p gpurun_out/c1
O=gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py tests/test_engine_gpu.py -m gpu -q -x --timeout 900 > $O/tests_small.log 2>&1; echo "tests_small rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests/test_parity_7b_gpu.py -m gpu -q -s --timeout 1400 > $O/tests_7b.log 2>&1; echo "tests_7b rc=$?" >> $O/rc.txt
LIVECC_B200_GEMM_SPLITK=1 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py -m gpu -q --timeout 500 > $O/tests_splitk.log 2>&1; echo "tests_splitk rc=$?" >> $O/rc.txt
LIVECC_B200_ATTN_PTMEM=1 timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "vit_attention or prefill_attention or attn" --timeout 500 > $O/tests_ptmem.log 2>&1; echo "tests_ptmem rc=$?" >> $O/rc.txt
timeout 300 python tools/bench_vs_fa2.py > $O/vs_fa2.md 2>&1; echo "vs_fa2 rc=$?" >> $O/rc.txt
timeout 300 python tools/bench_gemm_shapes.py > $O/gemm_shapes.txt 2>&1
LIVECC_B200_GEMM_SPLITK=1 timeout 300 python tools/bench_gemm_shapes.py > $O/gemm_shapes_splitk.txt 2>&1
for c in prefill_gateup vit_fc1 prefill_down; do
  CASE=$c timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 2 -o $O/ncu_gemm_$c -f python tools/bench_gemm_shapes.py > $O/ncu_gemm_$c.log 2>&1
done
timeout 900 python bench.py --steps 2 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
LIVECC_B200_GEMM_SPLITK=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_native_splitk.json 2> $O/bench_native_splitk.err
timeout 900 python bench.py --impl hf_gpu --steps 1 --warmup 1 > $O/bench_hf_gpu.json 2> $O/bench_hf_gpu.err; echo "hf_gpu rc=$?" >> $O/rc.txt
timeout 900 python bench.py --impl hf_gpu --liger --steps 1 --warmup 1 > $O/bench_hf_gpu_liger.json 2> $O/bench_hf_gpu_liger.err; echo "hf_gpu_liger rc=$?" >> $O/rc.txt
cat $O/rc.txt
tail -5 $O/tests_small.log $O/tests_7b.log
