#!/bin/bash
# round-2 GPU call: new fused depthwise kernels + train-mode parity + staged-kernel validation (COAL / PAIR) + bench
O=gpurun_out/call1; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "dwconv_fused or conv1x1_input_prologue" > $O/unit_new.log 2>&1; echo "unit_new rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "backbone or drop_connect" > $O/backbone.log 2>&1; echo "backbone rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "train_mode" > $O/train_mode.log 2>&1; echo "train_mode rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench_base.json 2> $O/bench_base.err; echo "bench rc=$?" >> $O/rc.txt
EFFDET_B200_COAL=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "conv2d or head or bifpn or backbone or model_train" > $O/coal_tests.log 2>&1; echo "coal rc=$?" >> $O/rc.txt
timeout 300 python tools/bench_head.py > $O/head_default.log 2>&1; echo "head_default rc=$?" >> $O/rc.txt
EFFDET_B200_COAL=1 timeout 300 python tools/bench_head.py > $O/head_coal.log 2>&1; echo "head_coal rc=$?" >> $O/rc.txt
EFFDET_B200_COAL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench_coal.json 2> $O/bench_coal.err; echo "bench_coal rc=$?" >> $O/rc.txt
EFFDET_B200_PAIR=1 timeout -s KILL 240 python tools/bench_head.py > $O/head_pair.log 2>&1; echo "head_pair rc=$?" >> $O/rc.txt
nvidia-smi > $O/after_pair.txt 2>&1
EFFDET_B200_PAIR=1 timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -k "head or model_train" > $O/pair_tests.log 2>&1; echo "pair rc=$?" >> $O/rc.txt
cat $O/rc.txt
