#!/bin/bash
O=gpurun_out/c12; mkdir -p $O
KV=1000 timeout 400 python tools/trace_mega.py > $O/trace_kv1000.txt 2>&1
KV=8000 timeout 400 python tools/trace_mega.py > $O/trace_kv8000.txt 2>&1
tail -22 $O/trace_kv1000.txt; tail -22 $O/trace_kv8000.txt
