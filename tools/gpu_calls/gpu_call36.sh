#!/bin/bash
O=gpurun_out/c36; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 1500 > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/rc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 400 python tools/bench_mcq_prefill.py --batches 1,4,16 > $O/mcq.txt 2>&1; echo "mcq rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -n 3 $O/tests_all.log | cut -c1-300; tail -1 $O/smoke.log | cut -c1-300; tail -5 $O/mcq.txt
