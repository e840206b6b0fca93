#!/bin/bash
# round-2 GPU call 6d: shared-space pointers in all tcgen05 kernels (LDS/STS instead of generic LD/ST), pw_wgrad 7x7
O=gpurun_out/call6d; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests -q -m gpu -x > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 600 python bench.py --config d4 --steps 5 --warmup 2 --no-cpu > $O/bench_d4.json 2> $O/bench_d4.err; echo "bench_d4 rc=$?" >> $O/rc.txt
cat $O/rc.txt
