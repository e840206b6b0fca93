"""Fused ``clip_grad_norm_`` + ``AdamW`` step: what the reference's loop does right after
``loss.backward()`` (``train.py:115-118`` with ``optim.AdamW(model.parameters(), lr)``, ``train.py:268``) as two
multi-tensor kernel launches and no host synchronisation.

    opt = FusedClipAdamW(model.parameters(), lr=1e-4, max_norm=0.1)
    loss.backward(); opt.step(); opt.zero_grad()

Semantics follow ``torch.nn.utils.clip_grad_norm_`` (coefficient max_norm / (total_norm + 1e-6), clamped to 1, gradients
rescaled in place) and ``torch.optim.AdamW`` (decoupled weight decay, bias-corrected moments, no amsgrad).
"""
import numpy as np
import torch

from . import _native as N

_CHUNK = 65536


def _check_param(p):
    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
        raise N.EffdetNativeError('FusedClipAdamW needs contiguous CUDA float32 parameters and gradients')


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=0.1):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm)
        super().__init__(params, defaults)
        self._tables = {}
        self.last_norm_sq = None          # device scalar of the most recent step (read it only if you need it)

    def _table(self, gi, plist):
        ms, vs = [], []
        for p in plist:
            st = self.state[p]
            if 'exp_avg' not in st:
                st['step'] = 0
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            ms.append(st['exp_avg'])
            vs.append(st['exp_avg_sq'])
        # gradients are re-allocated by zero_grad(set_to_none=True) and the moments by load_state_dict():
        # the device tables are rebuilt whenever any of the four pointers of any tensor moved
        key = tuple((p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr()) for p, m, v in zip(plist, ms, vs))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        dev = plist[0].device
        numels = np.array([p.numel() for p in plist], dtype=np.int64)
        ct, co = [], []
        for t, n in enumerate(numels):
            for off in range(0, int(n), _CHUNK):
                ct.append(t)
                co.append(off)

        def up(a):
            return torch.from_numpy(a).to(dev)

        tab = dict(p=up(np.array([p.data_ptr() for p in plist], dtype=np.uint64).view(np.int64)),
                   g=up(np.array([p.grad.data_ptr() for p in plist], dtype=np.uint64).view(np.int64)),
                   m=up(np.array([t.data_ptr() for t in ms], dtype=np.uint64).view(np.int64)),
                   v=up(np.array([t.data_ptr() for t in vs], dtype=np.uint64).view(np.int64)),
                   numels=up(numels), ct=up(np.array(ct, dtype=np.int32)), co=up(np.array(co, dtype=np.int64)),
                   nchunks=len(ct))
        self._tables[gi] = (key, tab)
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        # one global norm over every parameter group, like clip_grad_norm_(model.parameters(), max_norm)
        live = []
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group['params'] if p.grad is not None]
            for p in plist:
                _check_param(p)
            if plist:
                live.append((gi, group, plist))
        if not live:
            return loss
        dev_t = live[0][2][0]
        norm_sq = torch.zeros((1,), device=dev_t.device, dtype=torch.float32)
        tabs = []
        for gi, group, plist in live:
            tab = self._table(gi, plist)
            tabs.append(tab)
            N.call('effdet_multi_sumsq', dev_t, tab['g'].data_ptr(), tab['numels'].data_ptr(), tab['ct'].data_ptr(),
                   tab['co'].data_ptr(), tab['nchunks'], _CHUNK, norm_sq.data_ptr())
        for (gi, group, plist), tab in zip(live, tabs):
            st0 = self.state[plist[0]]
            # one bias correction per launch: every live parameter of the group must be at the same step (parameters
            # that start receiving gradients later, or checkpoints with per-parameter counts, would silently get the
            # wrong correction -- torch.optim.AdamW tracks the count per parameter)
            steps = {int(self.state[p]['step']) for p in plist}
            if len(steps) > 1:
                raise N.EffdetNativeError('FusedClipAdamW: parameters of one group are at different steps %s; put late '
                                          'starters into their own param group' % sorted(steps))
            step = int(st0['step']) + 1          # int() also accepts the tensor step of a torch.optim.AdamW checkpoint
            for p in plist:
                self.state[p]['step'] = step
            b1, b2 = group['betas']
            N.call('effdet_multi_clip_adamw', dev_t, tab['p'].data_ptr(), tab['g'].data_ptr(), tab['m'].data_ptr(),
                   tab['v'].data_ptr(), tab['numels'].data_ptr(), tab['ct'].data_ptr(), tab['co'].data_ptr(),
                   tab['nchunks'], _CHUNK, norm_sq.data_ptr(), float(group['max_norm'] or 0.0), float(group['lr']),
                   float(b1), float(b2), float(group['eps']), float(group['weight_decay']), 1.0 - b1 ** step,
                   1.0 - b2 ** step, 1)
            # the kernel wrote through raw pointers: tell autograd (and the packed-weight / folded-BN caches of
            # models/_ops.py, which key on Tensor._version) that parameters and gradients changed in place
            torch.autograd.graph.increment_version(plist)
            torch.autograd.graph.increment_version([p.grad for p in plist])
        self.last_norm_sq = norm_sq
        return loss
