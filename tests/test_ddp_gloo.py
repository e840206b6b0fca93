"""N>1 host logic on CPU: 2 gloo ranks, DistributedDataParallel(find_unused_parameters=True) exactly as
reference train.py:250, each rank on its own shard of the global batch.  The product has no CPU path, so the
module wrapped here is the oracle (same parameter names, same 5 dead backbone parameters); what is under test
is the data-parallel contract bench.py relies on: rank-local mean losses + DDP averaging == full-batch
gradients, identical on every rank, and unused parameters do not hang the reducer."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import effdet_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


class OracleDet(torch.nn.Module):
    def __init__(self, cfg, sd):
        super().__init__()
        self.cfg = cfg
        self.names = [k for k, v in sd.items() if v.is_floating_point() and 'running' not in k]
        self.params = torch.nn.ParameterList([torch.nn.Parameter(sd[k].clone()) for k in self.names])
        self.fixed = {k: v for k, v in sd.items() if k not in set(self.names)}

    def forward(self, images, ann):
        sd = dict(self.fixed)
        sd.update({k: p for k, p in zip(self.names, self.params)})
        return O.train_forward(sd, images, ann, self.cfg)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cfg = O.make_config('efficientdet-d0', num_classes=20, W_bifpn=64, D_bifpn=2)
    sd = O.init_state_dict(cfg, seed=0)
    net = torch.nn.parallel.DistributedDataParallel(OracleDet(cfg, sd), find_unused_parameters=True)
    images, ann = O.synthetic_batch(2 * world, size=128, num_classes=20, seed=77)
    lo, hi = rank * 2, rank * 2 + 2                     # weak scaling: a fixed shard per rank
    cl, rl = net(images[lo:hi], ann[lo:hi])
    (cl.mean() + rl.mean()).backward()
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in zip(net.module.names, net.module.params)}
    torch.save(grads, os.path.join(out, 'g%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_ddp_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(os.path.join(tmp_path, 'g0.pt'))
    g1 = torch.load(os.path.join(tmp_path, 'g1.pt'))
    # single-process full batch: the per-rank losses are means over the shard -> average over ranks
    cfg = O.make_config('efficientdet-d0', num_classes=20, W_bifpn=64, D_bifpn=2)
    sd = O.init_state_dict(cfg, seed=0)
    m = OracleDet(cfg, sd)
    images, ann = O.synthetic_batch(2 * world, size=128, num_classes=20, seed=77)
    cl, rl = m(images, ann)
    (cl.mean() + rl.mean()).backward()
    dead = 0
    for k, p in zip(m.names, m.params):
        a, b = g0[k], g1[k]
        if p.grad is None:
            dead += 1
            assert a is None or float(a.abs().max()) == 0.0
            continue
        assert torch.equal(a, b), 'ranks disagree on ' + k
        assert O.rel_err(a, p.grad) < 1e-3, (k, O.rel_err(a, p.grad))   # fp32 summation order (2 shards vs 1 batch)
    assert dead == 5            # _conv_head, _bn1.{weight,bias}, _fc.{weight,bias}  
