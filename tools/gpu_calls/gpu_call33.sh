#!/bin/bash
O=gpurun_out/c33; mkdir -p $O
for bn in 256 -256; do
  CASE=mcq_gateup_2090 BN=$bn timeout 300 ncu --set full --clock-control none -k regex:gemm -s 2 -c 1 -o $O/ncu_gemm_mcq_bn$bn -f python tools/bench_gemm_shapes.py > $O/ncu_bn$bn.log 2>&1; echo "ncu bn=$bn rc=$?" >> $O/rc.txt
  CASE=vit_fc1 BN=$bn timeout 300 ncu --set full --clock-control none -k regex:gemm -s 2 -c 1 -o $O/ncu_gemm_fc1_bn$bn -f python tools/bench_gemm_shapes.py > $O/ncu_fc1_bn$bn.log 2>&1; echo "ncu fc1 bn=$bn rc=$?" >> $O/rc.txt
done
cat $O/rc.txt; ls -la $O | head -20
