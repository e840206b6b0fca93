"""CUDA-graph execution of the training step (forward + FocalLoss + backward, optionally the fused optimizer).

Every kernel of the hot path is launched on the caller's stream with device-resident arguments and no host
synchronisation (the C ABI never syncs, FocalLoss reads its upstream gradients from device memory), so the ~480
launches of a step can be captured once and replayed: the host cost of a step drops from ~12-16 ms of ctypes calls to
one cudaGraphLaunch, which matters as soon as the GPU finishes a step faster than Python can issue it.

    step = GraphedTrainStep(model, images_example, annotations_example)          # model.train(); freeze_bn() first
    loss = step(images, annotations)      # copies into the static inputs, replays, returns the static loss tensor
    ... model parameters' .grad now hold this step's gradients (same tensors every step)

Build it BEFORE any eager step, or after every reference to earlier eager losses / outputs has been dropped: a live autograd
graph keeps the parameters' gradient accumulators bound to the stream it ran on (the legacy stream), and the engine
may not touch that stream while another one is capturing.

Shapes are static (the reference pads every batch to the common size and a fixed annotation count per batch can be
obtained by padding with -1 rows, which FocalLoss ignores).  Drop-connect keeps working: torch.rand inside a captured
region advances the Philox offset on every replay.  With `optimizer=FusedClipAdamW(...)` the clip + AdamW launches are
captured too; the packed-weight / folded-BN derivations are then captured as well (they must re-run every step because
the weights change without Python seeing it).
"""
import torch

from . import _ops


class GraphedTrainStep:
    """model: the EfficientDet module, or its DistributedDataParallel wrapper.  For DDP the NCCL gradient all-reduces are
    captured with the step (torch's documented recipe: NCCL >= 2.9.6, TORCH_NCCL_ASYNC_ERROR_HANDLING=0 before
    init_process_group, the DDP wrapper CONSTRUCTED inside a side-stream context, >= 11 eager warm-up iterations);
    gradients then have to flow through DDP's AccumulateGrad hooks, so the captured step calls loss.backward()."""

    def __init__(self, model, images, annotations, optimizer=None, warmup=3):
        if not images.is_cuda:
            raise _ops.N.EffdetNativeError('GraphedTrainStep needs CUDA example inputs')
        self.ddp = isinstance(model, torch.nn.parallel.DistributedDataParallel)
        if self.ddp:
            warmup = max(warmup, 11)
        self.model = model
        self.optimizer = optimizer
        self.static_images = images.clone()
        self.static_annots = annotations.clone()
        params = [p for p in model.parameters() if p.requires_grad]
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream(images.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                                   # fills allocator pools, caches, cuFuncSetAttribute
                self._eager_step(params)
        torch.cuda.current_stream(images.device).wait_stream(side)
        torch.cuda.synchronize(images.device)
        if optimizer is not None:
            _ops.invalidate_caches()                                  # capture the weight re-packing with the step
        self.graph = torch.cuda.CUDAGraph()
        for p in params:
            p.grad = None
        n0 = _ops.N.launch_count()
        try:
            with torch.cuda.graph(self.graph):
                self.static_loss = self._eager_step(params, zero=False)
        except RuntimeError as e:
            if 'legacy stream' in str(e) or 'previous error during capture' in str(e):
                raise _ops.N.EffdetNativeError(
                    'GraphedTrainStep: the autograd engine tried to synchronise with the legacy default stream during '
                    'capture.  The gradient accumulators of the parameters remember the stream of an earlier eager step '
                    'for as long as that step\'s graph is alive: drop every reference to earlier losses / outputs '
                    '(`del loss`) before building the graphed step, or run eager steps under a side stream.') from e
            raise
        self.library_launches = _ops.N.launch_count() - n0            # kernels of this library recorded into the graph
        self.params = params

    def _eager_step(self, params, zero=True):
        if zero:
            for p in params:
                p.grad = None
        cl, rl = self.model([self.static_images, self.static_annots])
        loss = cl.mean() + rl.mean()
        # torch.autograd.grad, not .backward(): AccumulateGrad nodes remember the stream of an earlier eager step (the
        # legacy stream if the caller still holds that step's loss) and would make it wait on the capturing stream --
        # cudaErrorStreamCaptureImplicit.  The returned tensors live in the graph's pool: every replay rewrites them.
        if self.ddp:
            loss.backward()
        else:
            grads = torch.autograd.grad(loss, params, allow_unused=True)
            for p, g in zip(params, grads):
                p.grad = g
        if self.optimizer is not None:
            self.optimizer.step()
        return loss.detach()

    def __call__(self, images, annotations):
        self.static_images.copy_(images, non_blocking=True)
        self.static_annots.copy_(annotations, non_blocking=True)
        self.graph.replay()
        return self.static_loss
