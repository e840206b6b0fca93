#!/bin/bash
# round-2 GPU call 7: final evidence -- ncu --set full captures of the kernels the report quotes, launch list, bench
O=gpurun_out/call7; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
cap() {   # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o $O/prof_$1 python tools/one_step.py 1 > $O/ncu_$1.log 2>&1
  echo "ncu_$1 rc=$?" >> $O/rc.txt
}
cap planes 'conv_planes_kernel.*256' 1 2
cap wgrad_multi wgrad_tc2_multi_kernel 0 3
cap dw_fwd dw_fwd_fused_kernel 0 5
cap dw_bwd dw_bwd_fused_kernel 10 6
cap bifpn "fuse_(fwd|bwd)" 0 4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file $O/launches.csv python tools/one_step.py 2 > $O/ncu_list.log 2>&1; echo "ncu_list rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown --no-graph > $O/bench_eager.json 2> $O/bench_eager.err; echo "bench_eager rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
