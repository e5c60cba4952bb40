#!/bin/bash
O=gpurun_out/c24; mkdir -p $O
timeout 600 python tools/bench_vit_graph.py > $O/vit_graph.txt 2>&1; echo "vit_graph rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 6 $O/vit_graph.txt | cut -c1-300
