#!/bin/bash
O=gpurun_out/c34; mkdir -p $O
SWEEP_BN=1 BNS=0,64,80,96,112,128,144,160,176,192,208,224,240,256 timeout 600 python tools/bench_gemm_shapes.py > $O/gemm_width_sweep.txt 2>&1; echo "sweep rc=$?" >> $O/rc.txt
cat $O/rc.txt; grep -v "^$" $O/gemm_width_sweep.txt | grep -v "mcq_\|8192" | cut -c1-110
