#!/bin/bash
# 2-GPU call: DDP step as one CUDA graph (NCCL all-reduces captured) -- test + bench (graph vs eager)
O=gpurun_out/ddp2b; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_ddp_nccl.py -q -s > $O/ddp_test.log 2>&1; echo "ddp_test rc=$?" >> $O/rc.txt
run2() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/bench2_$name.json 2> $O/bench2_$name.err; echo "bench2_$name rc=$?" >> $O/rc.txt
}
run2 graph FOO=1
run2 eager EFFDET_DDP_GRAPH=0
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench1.json 2> $O/bench1.err; echo "bench1 rc=$?" >> $O/rc.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-graph > $O/bench1_eager.json 2> $O/bench1_eager.err; echo "bench1_eager rc=$?" >> $O/rc.txt
for f in $O/bench*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config'].get('execution'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done > $O/summary.txt
cat $O/rc.txt $O/summary.txt
