#!/bin/bash
# 2-GPU call: DDP step as one CUDA graph, teardown fixed -- every step under a SHORT timeout
O=gpurun_out/ddp3; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
EFFDET_DDP_GRAPH=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/bench2_graph.json 2> $O/bench2_graph.err; echo "bench2_graph rc=$?" >> $O/rc.txt
timeout 300 python -m pytest tests/test_ddp_nccl.py -q -s > $O/ddp_test.log 2>&1; echo "ddp_test rc=$?" >> $O/rc.txt
cat $O/rc.txt; cat $O/bench2_graph.json | cut -c1-400
