#!/bin/bash
# round-2 GPU call 3: pw_gemm v2 (8/16 converter warps, deeper store staging, planes mode), planes path, faster sigmoid
O=gpurun_out/call3; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -s -k "pointwise_gemm or planes" > $O/unit_pw.log 2>&1; echo "unit_pw rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "dwconv_fused or conv1x1_input_prologue or conv2d_tensor_core or epilogue" > $O/unit_new.log 2>&1; echo "unit_new rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "backbone or drop_connect or train_mode" > $O/backbone.log 2>&1; echo "backbone rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 900 python bench.py --config d4 --steps 5 --warmup 2 --no-cpu --full-breakdown > $O/bench_d4.json 2> $O/bench_d4.err; echo "bench_d4 rc=$?" >> $O/rc.txt
timeout 900 python bench.py --config d7 --steps 5 --warmup 2 --no-cpu --full-breakdown > $O/bench_d7.json 2> $O/bench_d7.err; echo "bench_d7 rc=$?" >> $O/rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file $O/launches.csv python tools/one_step.py 2 > $O/ncu_list.log 2>&1; echo "ncu_list rc=$?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pw_gemm -c 6 -o $O/prof_pw python tools/one_step.py 1 > $O/ncu_pw.log 2>&1; echo "ncu_pw rc=$?" >> $O/rc.txt
cat $O/rc.txt
