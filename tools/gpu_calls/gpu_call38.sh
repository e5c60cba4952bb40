#!/bin/bash
# Whole GPU suite with the P-through-TMEM attention variants forced on (ViT + prefill), then the prefill kernel timed with and
# without the variant at the streaming shape.
O=gpurun_out/c38; mkdir -p $O
LIVECC_B200_ATTN_PTMEM=1 timeout 150 python -m pytest tests -m gpu -q -x --timeout 140 -p no:cacheprovider > $O/tests_ptmem.log 2>&1; echo "tests_ptmem rc=$?" >> $O/rc.txt
tail -n 3 $O/tests_ptmem.log | cut -c1-300
LIVECC_B200_ATTN_PTMEM=0 PASTS=9000,17000 timeout 40 python tools/bench_prefill_attn.py > $O/prefill_attn_default.txt 2>&1; echo "prefill_default rc=$?" >> $O/rc.txt
LIVECC_B200_ATTN_PTMEM=1 PASTS=9000,17000 timeout 40 python tools/bench_prefill_attn.py > $O/prefill_attn_ptmem.txt 2>&1; echo "prefill_ptmem rc=$?" >> $O/rc.txt
cat $O/rc.txt; grep "impl 2" $O/prefill_attn_default.txt $O/prefill_attn_ptmem.txt
