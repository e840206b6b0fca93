#!/bin/bash
# round-2 GPU call 4: resident-weight pw_gemm, block-wise NMS scan, pipeline kernels, checkpoint round trip; benches + ncu
O=gpurun_out/call4; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_pipeline.py -q -x -s -k "pointwise_gemm or planes or nms or detect or pipeline or device_ or checkpoint or invalidate or d7" > $O/unit.log 2>&1; echo "unit rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
EFFDET_B200_PERSIST=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench_nopersist.json 2> $O/bench_nopersist.err; echo "bench_nopersist rc=$?" >> $O/rc.txt
timeout 900 python bench.py --config d7 --steps 5 --warmup 2 --no-cpu --full-breakdown > $O/bench_d7.json 2> $O/bench_d7.err; echo "bench_d7 rc=$?" >> $O/rc.txt
timeout 900 python bench.py --config d7 --threshold 0.01 --steps 3 --warmup 1 --no-cpu --full-breakdown > $O/bench_d7_t001.json 2> $O/bench_d7_t001.err; echo "bench_d7_t001 rc=$?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pw_gemm -c 6 -o $O/prof_pw python tools/one_step.py 1 > $O/ncu_pw.log 2>&1; echo "ncu_pw rc=$?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_persist_kernel -s 4 -c 2 -o $O/prof_head python tools/one_step.py 1 > $O/ncu_head.log 2>&1; echo "ncu_head rc=$?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc2_multi_kernel -s 2 -c 1 -o $O/prof_wgrad python tools/one_step.py 1 > $O/ncu_wgrad.log 2>&1; echo "ncu_wgrad rc=$?" >> $O/rc.txt
cat $O/rc.txt
