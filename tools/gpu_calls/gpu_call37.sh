#!/bin/bash
# BASELINE config #3, second variant: 240 s clip at 280x280 (~100 vision tokens per frame pair = the
# reference's training budget, "24k vision tokens"), then the P-through-TMEM attention variant timed against the default.
O=gpurun_out/c37; mkdir -p $O
timeout 230 python bench.py --seconds 240 --size 280 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-batch \
    > $O/bench_240s_280.json 2> $O/bench_240s_280.err; echo "bench280 rc=$?" >> $O/rc.txt
IMPLS=2 BNS=64,128 timeout 70 python tools/bench_vit_attn.py > $O/vit_attn_default.txt 2>&1; echo "vit_attn rc=$?" >> $O/rc.txt
LIVECC_B200_ATTN_PTMEM=1 IMPLS=2 BNS=64,128 timeout 70 python tools/bench_vit_attn.py > $O/vit_attn_ptmem.txt 2>&1; echo "vit_attn_ptmem rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -c 600 $O/bench_240s_280.json; tail -3 $O/vit_attn_default.txt; tail -3 $O/vit_attn_ptmem.txt
