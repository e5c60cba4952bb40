#!/bin/bash
O=gpurun_out/c14; mkdir -p $O
KV=8000 timeout 400 python tools/trace_attn.py > $O/trace_attn.txt 2>&1
tail -14 $O/trace_attn.txt
