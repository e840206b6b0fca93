#!/bin/bash
# round-2 GPU call 8: depthwise backward with the weight gradient as a second loop (one tap row per thread), 2 CTAs x 8 warps
O=gpurun_out/call8; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "dwconv_fused or backbone or train_mode" > $O/unit.log 2>&1; echo "unit rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none -k "regex:conv_planes_kernel" -s 17 -c 2 -o /tmp/prof_planes python tools/one_step.py 1 > $O/ncu_planes.log 2>&1; echo "ncu_planes rc=$?" >> $O/rc.txt
python tools/ncu_summary.py /tmp/prof_planes.ncu-rep > $O/r02_ncu_planes.txt 2>> $O/ncu_planes.log
python tools/ncu_traffic.py "conv_planes_kernel<256> 256->256=/tmp/prof_planes.ncu-rep:0" > $O/r02_ncu_traffic.json 2>> $O/ncu_planes.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:dw_bwd_fused_kernel" -s 10 -c 6 -o /tmp/prof_dw_bwd python tools/one_step.py 1 > $O/ncu_dw_bwd.log 2>&1; echo "ncu_dw_bwd rc=$?" >> $O/rc.txt
python tools/ncu_summary.py /tmp/prof_dw_bwd.ncu-rep > $O/r02_ncu_dw_bwd.txt 2>> $O/ncu_dw_bwd.log
python tools/ncu_stalls.py /tmp/prof_dw_bwd.ncu-rep 16 > $O/r02_ncu_dw_bwd_stalls.txt 2>> $O/ncu_dw_bwd.log
timeout 900 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
cat $O/rc.txt
