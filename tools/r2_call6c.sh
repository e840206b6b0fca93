#!/bin/bash
# round-2 GPU call 6c: pw_wgrad v3 (lean converters, wide tiles, adaptive K), stem wgrad v3, dw constants in smem
O=gpurun_out/call6c; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "pointwise_wgrad or conv1x1_input or expand_gradients or graphed or stem or dwconv_fused" > $O/unit.log 2>&1; echo "unit rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_wgrad_kernel -s 22 -c 9 -o $O/prof_pw_wgrad python tools/one_step.py 1 > $O/ncu_pw_wgrad.log 2>&1; echo "ncu_pw_wgrad rc=$?" >> $O/rc.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pw_gemm_kernel -s 0 -c 6 -o $O/prof_pw_gemm python tools/one_step.py 1 > $O/ncu_pw_gemm.log 2>&1; echo "ncu_pw_gemm rc=$?" >> $O/rc.txt
cat $O/rc.txt
