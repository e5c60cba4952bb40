#!/bin/bash
O=gpurun_out/c27; mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 800 -k "ingest or vit_graph" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm|decode_mega|attn_|vit_|flash_|sample_greedy|embed_|video_rank|mrope|layernorm|rmsnorm|cast_f32|patchify|fill_cu|splitk|gemv_|resize_" -c 8000 --csv --log-file $O/launches.csv python bench.py --seconds 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-batch > $O/launches.log 2>&1; echo "launches rc=$?" >> $O/rc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:resize_bicubic -c 1 -o $O/ncu_resize -f python tools/bench_resize.py --iters 2 > $O/ncu_resize.log 2>&1; echo "ncu_resize rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 5 $O/tests.log | cut -c1-300; tail -3 $O/launches.log | cut -c1-200; tail -3 $O/ncu_resize.log | cut -c1-200
