#!/bin/bash
# round-2 GPU call 7: final evidence -- ncu --set full captures of the kernels the report quotes (condensed to text ON THE
# BOX: the reports themselves exceed the 64 MiB that travel back), launch list, bench
O=gpurun_out/call7; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
cap() {   # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o /tmp/prof_$1 python tools/one_step.py 1 > $O/ncu_$1.log 2>&1
  echo "ncu_$1 rc=$?" >> $O/rc.txt
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep > $O/r02_ncu_$1.txt 2>> $O/ncu_$1.log
  python tools/ncu_stalls.py /tmp/prof_$1.ncu-rep 16 > $O/r02_ncu_$1_stalls.txt 2>> $O/ncu_$1.log
}
cap planes 'conv_planes_kernel.*256' 1 2
python tools/ncu_traffic.py "conv_planes_kernel<256> 256->256=/tmp/prof_planes.ncu-rep:0" > $O/r02_ncu_traffic.json 2>> $O/ncu_planes.log
cap wgrad_multi wgrad_tc2_multi_kernel 0 3
cap dw_fwd dw_fwd_fused_kernel 0 5
cap dw_bwd dw_bwd_fused_kernel 10 6
cap bifpn "fuse_(fwd|bwd)" 0 4
cap pw_gemm pw_gemm_kernel 0 6
cap pw_wgrad pw_wgrad_kernel 24 8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file $O/launches.csv python tools/one_step.py 2 > $O/ncu_list.log 2>&1; echo "ncu_list rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown --no-graph > $O/bench_eager.json 2> $O/bench_eager.err; echo "bench_eager rc=$?" >> $O/rc.txt
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
du -sh $O >> $O/rc.txt
cat $O/rc.txt
