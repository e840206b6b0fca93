#!/bin/bash
# round-2 GPU call 6b: pw_wgrad v2 (register double-buffering, balanced tiles), graph step via autograd.grad
O=gpurun_out/call6b; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "pointwise_wgrad or conv1x1_input or expand_gradients or graphed" > $O/unit.log 2>&1; echo "unit rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
EFFDET_B200_PWWG=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --full-breakdown > $O/bench_nopwwg.json 2> $O/bench_nopwwg.err; echo "bench_nopwwg rc=$?" >> $O/rc.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/all_gpu.log 2>&1; echo "all_gpu rc=$?" >> $O/rc.txt
timeout 300 python tools/bw_probe.py > $O/bw_probe.txt 2>&1; echo "bw_probe rc=$?" >> $O/rc.txt
cat $O/rc.txt
