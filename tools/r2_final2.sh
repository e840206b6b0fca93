#!/bin/bash
# final check of the DEFAULT paths on 2 GPUs: bench at N=2 exactly as the driver launches it, then the DDP tests
O=gpurun_out/final2; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2.json 2> $O/bench2.err; echo "bench2 rc=$?" >> $O/rc.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/ref2.json 2> $O/ref2.err; echo "ref2 rc=$?" >> $O/rc.txt
timeout 300 python -m pytest tests/test_ddp_nccl.py -q -s > $O/ddp_test.log 2>&1; echo "ddp_test rc=$?" >> $O/rc.txt
cat $O/rc.txt; cut -c1-300 $O/bench2.json; cut -c1-300 $O/ref2.json
