#!/bin/bash
# 8-GPU check of the graphed DDP step (short timeouts: a hang must not eat the budget)
O=gpurun_out/n8; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
EFFDET_DDP_GRAPH=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > $O/bench8_graph.json 2> $O/bench8_graph.err; echo "bench8_graph rc=$?" >> $O/rc.txt
EFFDET_DDP_GRAPH=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > $O/bench8_eager.json 2> $O/bench8_eager.err; echo "bench8_eager rc=$?" >> $O/rc.txt
cat $O/rc.txt; cut -c1-300 $O/bench8_graph.json; cut -c1-300 $O/bench8_eager.json
