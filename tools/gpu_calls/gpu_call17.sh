#!/bin/bash
O=gpurun_out/c17; mkdir -p $O
SWEEP_BN=1 timeout 500 python tools/bench_gemm_shapes.py > $O/gemm_bn_sweep.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench_native.json 2> $O/bench_native.err; echo "bench rc=$?" >> $O/rc.txt
LIVECC_B200_MEGA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 60 -c 1 -o $O/ncu_mega -f python bench.py --seconds 20 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-batch > $O/ncu_mega.log 2>&1; echo "ncu_mega rc=$?" >> $O/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm_bf16|decode_mega|attn_|vit_|flash_|sample_greedy|embed_|video_rank|mrope|layernorm|rmsnorm|cast_f32|patchify|fill_cu|splitk|gemv_" -c 6000 --csv --log-file $O/launches.csv python bench.py --seconds 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-batch > $O/launches.log 2>&1; echo "launches rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 4 $O/tests_all.log | cut -c1-300; tail -2 $O/smoke.log | cut -c1-300
cat $O/gemm_bn_sweep.txt | grep -v "^$" | head -40
tail -c 2500 $O/bench_native.json
MASK=$((1|64)) KV=8000 timeout 300 python tools/trace_attn.py > $O/trace_stage.txt 2>&1; tail -8 $O/trace_stage.txt | cut -c1-200
