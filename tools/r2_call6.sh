#!/bin/bash
# round-2 GPU call 6: parallel SE backward, dw kernels (k5: 8 warps, small-map tiles), stem wgrad v2
O=gpurun_out/call6; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "dwconv_fused or backbone or checkpoint or invalidate" > $O/unit.log 2>&1; echo "unit rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 900 python bench.py --config d4 --steps 5 --warmup 2 --no-cpu --full-breakdown > $O/bench_d4.json 2> $O/bench_d4.err; echo "bench_d4 rc=$?" >> $O/rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file $O/launches.csv python tools/one_step.py 2 > $O/ncu_list.log 2>&1; echo "ncu_list rc=$?" >> $O/rc.txt
cat $O/rc.txt
