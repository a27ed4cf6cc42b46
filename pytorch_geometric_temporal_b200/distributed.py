"""Data-parallel plumbing: one process per GPU, replicas hold the full graph + model, windows are
sharded by rank (signal.shard_indices), ONE flat-bucket gradient all-reduce per optimizer step.

Replaces the reference's DDP-over-gloo launched through Dask (examples/indexBatching/DCRNN/pems_ddp.py:
83-85,198-207): payloads are 0.6 KB - 2.8 MB, i.e. latency-bound, so all gradients
live in one contiguous fp32 buffer (the `.grad` of every parameter is a view into it) and a step costs
exactly one NCCL all-reduce over NVLink -- no bucketing, no per-parameter launches, no copies.
torch.distributed is used for the plumbing (backend nccl on GPUs, gloo on CPU for tests)."""
import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def init_process_group(backend: Optional[str] = None):
    """Initialise from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).  Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if cuda else "gloo")
        if backend == "nccl":
            dist.init_process_group(backend, device_id=device)
        else:
            dist.init_process_group(backend)
    return rank, world, device


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


class FlatGradSync(object):
    """Owns one flat fp32 buffer holding every parameter's gradient.

    usage:  sync = FlatGradSync(model.parameters()); ...; loss.backward(); sync.all_reduce(); opt.step(); sync.zero()
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], average: bool = True):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=dev, dtype=dt)
        self.average = average
        self._offsets = []
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != dt:
                raise ValueError("all parameters must share device and dtype")
            p.grad = self.flat[off:off + p.numel()].view_as(p)  # autograd accumulates in place into the view
            self._offsets.append(off)
            off += p.numel()

    def _check_aliasing(self):
        """`optimizer.zero_grad()` / `module.zero_grad()` default to set_to_none=True, which drops the views into the flat
        buffer: autograd would then allocate fresh gradients and the collective would reduce a stale buffer.  Gradients
        found outside the buffer are copied in and re-attached (use `sync.zero()` instead of zero_grad())."""
        base, es = self.flat.data_ptr(), self.flat.element_size()
        for p, off in zip(self.params, self._offsets):
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != base + off * es:
                view.copy_(p.grad)
                p.grad = view

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def all_reduce(self, async_op: bool = False):
        """ONE collective for the whole model; grads become the mean over ranks (DDP semantics)."""
        self._check_aliasing()
        w = world_size()
        if w == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return work
        if self.average:
            self.flat.mul_(1.0 / w)
        return None

    def finish(self, work):
        work.wait()
        if self.average:
            self.flat.mul_(1.0 / world_size())

    def zero(self):
        self.flat.zero_()


class FlatAdam(object):
    """torch.optim.Adam (the optimizer of examples/indexBatching/DCRNN/pems_ddp.py:90) over the flat buffers of a FlatGradSync.

    The parameters are moved into ONE flat fp32 buffer (each `p.data` becomes a view of it, like the gradients), the moments are
    two more, and a step is ONE launch (`stmp_adam_flat`) instead of the ~35 of the capturable foreach implementation; the step
    counter lives on the device, so the launch can be captured in a CUDA graph.  `step()` also clears the gradient buffer
    (zero_grad=False keeps it) -- `sync.all_reduce(); opt.step()` is the whole tail of a data-parallel training step.
    CUDA only: there is no CPU fallback.
    """

    def __init__(self, sync: FlatGradSync, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if not sync.flat.is_cuda or sync.flat.dtype != torch.float32:
            raise RuntimeError("FlatAdam needs CUDA fp32 parameters (the update is a CUDA kernel; there is no CPU fallback)")
        self.sync = sync
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.flat = torch.empty_like(sync.flat)
        with torch.no_grad():
            for p, off in zip(sync.params, sync._offsets):
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, device=self.flat.device, dtype=torch.float32)
        self._ticket = torch.zeros(1, device=self.flat.device, dtype=torch.int32)

    def step(self, zero_grad: bool = True, grad_scale: float = 1.0):
        from . import ops
        self.sync._check_aliasing()
        ops.adam_flat(self.flat, self.sync.flat, self.exp_avg, self.exp_avg_sq, self.step_count, self._ticket, self.lr, self.betas[0],
                      self.betas[1], self.eps, self.weight_decay, grad_scale, zero_grad)
        # the kernel wrote through the flat buffer: bump the version counters the packed-weight caches are keyed on
        torch._C._increment_version(self.sync.params)

    def zero_grad(self):
        self.sync.zero()


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Make every replica start from rank `src`'s parameters/buffers with one flat broadcast."""
    if world_size() == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.broadcast(flat, src)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t).to(t.dtype))
        off += t.numel()


def reduce_scalar(value: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """Per-epoch validation-loss reduction (pems_ddp.py:160-161)."""
    if world_size() > 1:
        dist.reduce(value, dst=dst, op=dist.ReduceOp.SUM)
    return value


def masked_mae_loss(y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
    """examples/indexBatching/DCRNN/utils.py:10-18: mean(nan_to_zero(|p - y| * mask / mean(mask))), mask = (y != 0).
    On CUDA fp32 tensors this is the fused kernel pair `stmp_masked_mae_fwd/bwd` (3 launches instead of ~25)."""
    if y_pred.is_cuda and y_pred.dtype == torch.float32 and y_true.dtype == torch.float32 and y_pred.shape == y_true.shape:
        from . import ops
        return ops.masked_mae(y_pred, y_true)
    return masked_mae_loss_reference(y_pred, y_true)


def masked_mae_loss_reference(y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
    """The op-for-op form (any device): used for non-CUDA tensors and as the checker of the fused kernels."""
    mask = (y_true != 0).float()
    mask = mask / mask.mean()
    loss = torch.abs(y_pred - y_true) * mask
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return loss.mean()
