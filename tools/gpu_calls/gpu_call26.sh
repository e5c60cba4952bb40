#!/bin/bash
O=gpurun_out/c26; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench_n2 rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -c 1500 $O/bench_n2.json; tail -n 5 $O/bench_n2.err | cut -c1-300
