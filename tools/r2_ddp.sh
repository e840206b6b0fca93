#!/bin/bash
# 2-GPU call: product DDP/NCCL gradient test + where does the fixed ~1.5 ms/step of DDP go (bucket size, NCCL CTAs)
O=gpurun_out/ddp2; mkdir -p $O; rm -f $O/rc.txt
python __graft_entry__.py > $O/build.log 2>&1; echo "build rc=$?" >> $O/rc.txt
timeout 900 python -m pytest tests/test_ddp_nccl.py -q -s > $O/ddp_test.log 2>&1; echo "ddp_test rc=$?" >> $O/rc.txt
run2() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/bench2_$name.json 2> $O/bench2_$name.err; echo "bench2_$name rc=$?" >> $O/rc.txt
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench1.json 2> $O/bench1.err; echo "bench1 rc=$?" >> $O/rc.txt
run2 default FOO=1
run2 bucket100 EFFDET_DDP_BUCKET_MB=100
run2 bucket8 EFFDET_DDP_BUCKET_MB=8
run2 ctas4 NCCL_MAX_CTAS=4
run2 ctas4_b100 NCCL_MAX_CTAS=4 EFFDET_DDP_BUCKET_MB=100
run2 nostatic EFFDET_DDP_STATIC=0
for f in $O/bench*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done > $O/summary.txt
cat $O/rc.txt $O/summary.txt
