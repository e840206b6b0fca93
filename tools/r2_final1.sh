#!/bin/bash
# final check on 1 GPU of exactly what the driver runs: smoke(), pytest -m gpu, bench.py (defaults) and the reference arm
O=gpurun_out/final1; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests -x -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench_ref rc=$?" >> $O/rc.txt
timeout 300 python bench.py --config d7 --steps 5 --warmup 3 --no-cpu > $O/bench_d7.json 2> $O/bench_d7.err; echo "bench_d7 rc=$?" >> $O/rc.txt
cat $O/rc.txt; tail -3 $O/smoke.log; tail -2 $O/all_gpu.log; cut -c1-400 $O/bench.json; cut -c1-400 $O/bench_ref.json
