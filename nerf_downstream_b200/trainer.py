"""Data-parallel training step for the sparse networks.

One process per GPU.  Parameters and gradients live in two flat fp32 arenas:
  * autograd accumulates every parameter gradient straight into its arena slice;
  * at N > 1 the gradient arena is all-reduced bucket by bucket over NCCL (NVLink 5 / NVSwitch) from
    post-accumulate hooks, so communication overlaps the remaining backward pass — the reference gets
    the same exchange from Lightning's DDPPlugin (co3d_3d/train.py:174-186);
  * one fused SGD kernel (momentum, weight decay, 1/world scaling) updates the whole parameter arena
    (co3d_3d/configs/co3d_cls.gin:31-39, src/modules/optim.py:60-69).
Samples never cross ranks, so there is no other collective on the data path.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops


class FlatArena:
    """Re-homes a module's parameters into one contiguous fp32 buffer (+ a gradient twin)."""

    def __init__(self, module: torch.nn.Module):
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev = self.params[0].device
        # reverse order ~ the order gradients become ready during backward
        self.order = list(reversed(self.params))
        self.offsets, off = [], 0
        for p in self.order:
            if p.dtype != torch.float32:
                raise TypeError("FlatArena expects fp32 parameters")
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4  # keep every slice 16-byte aligned
        self.numel = off
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.order, self.offsets):
                view = self.data[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view_as(p)

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.order, self.offsets):  # re-attach in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view_as(p)


OPTIMIZERS = ("SGD", "ASGD", "Adam", "AdamW", "Adagrad", "Adadelta", "Adamax", "RMSprop", "Rprop")   # optim.py:57


class DataParallelTrainer:
    """`optimizer_name` "SGD" (what every reference config selects) runs the fused `spc_sgd_step` on the flat arena;
    the other names of the reference's `get_optimizer` (optim.py:57-69) run the torch optimizer of that name on the
    arena-backed parameters (`optimizer_kwargs` = its gin bindings), with the same bucketed all-reduce in front."""

    def __init__(self, model: torch.nn.Module, lr: float = 0.1, momentum: float = 0.9, weight_decay: float = 1e-4,
                 bucket_mb: float = 25.0, process_group=None, optimizer_name: str = "SGD",
                 optimizer_kwargs: Optional[dict] = None):
        if optimizer_name not in OPTIMIZERS:
            raise ValueError(f"optimizer {optimizer_name} not recognized in {list(OPTIMIZERS)}.")
        self.model = model
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.arena = FlatArena(model)
        self.momentum_buf = torch.zeros_like(self.arena.data)
        self.optimizer_name = optimizer_name
        self.torch_optimizer = None
        if optimizer_name != "SGD":
            kw = dict(optimizer_kwargs or {})
            if optimizer_name != "Rprop":
                kw.setdefault("weight_decay", weight_decay)
            self.torch_optimizer = getattr(torch.optim, optimizer_name)(self.arena.params, lr=lr, **kw)
        self.steps = 0
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.group = process_group
        self._handles = []
        self._buckets = []      # (start, end) element ranges in arena order
        self._bucket_of = {}
        self._pending = []
        self._launched: List[int] = []      # bucket ids in the order their all-reduce was launched this step
        self._last_order: List[int] = []    # ... of the last complete step
        # exposed communication: device time between the end of backward and the last bucket's completion
        # (CUDA events on the compute stream), filled when `time_exposed` is set (scripts/bench_resnet14.py)
        self.time_exposed = False
        self.exposed_ms: List[float] = []
        self._exposed_events = []
        # The backward kernels add parameter gradients straight into the arena slices (ops.register_grad_sink):
        # no temporary + one tiny accumulation launch per parameter.  The callback stands in for autograd's
        # post-accumulate hook (bucketed all-reduce at world > 1).
        self._hooks = {}
        for i, p in enumerate(self.arena.order):
            hook = self._make_hook(i) if self.world > 1 else None
            self._hooks[id(p)] = hook
            ops.register_grad_sink(p, hook if hook is not None else (lambda param: None))
        if self.world > 1:
            # replicas must START equal: DistributedDataParallel (what the reference gets from Lightning's
            # DDPPlugin, train.py:184) broadcasts rank 0's parameters and buffers at construction
            dist.broadcast(self.arena.data, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                           group=process_group)
            for buf in model.buffers():
                dist.broadcast(buf.data, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                               group=process_group)
            self._make_buckets(int(bucket_mb * 1024 * 1024 / 4))
            for i, p in enumerate(self.arena.order):
                p.register_post_accumulate_grad_hook(self._hooks[id(p)])

    # ---- bucketing ---------------------------------------------------------------
    def _make_buckets(self, bucket_elems: int):
        a = self.arena
        start, count = 0, 0
        members = []
        for i, (p, o) in enumerate(zip(a.order, a.offsets)):
            members.append(i)
            end = o + (p.numel() + 3) // 4 * 4
            if end - start >= bucket_elems or i == len(a.order) - 1:
                b = len(self._buckets)
                self._buckets.append((start, end, len(members)))
                for m in members:
                    self._bucket_of[m] = b
                members = []
                start = end
        self._pending = [n for (_, _, n) in self._buckets]

    def _make_hook(self, index: int):
        def hook(param):
            b = self._bucket_of[index]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                s, e, _ = self._buckets[b]
                self._launched.append(b)
                self._handles.append(dist.all_reduce(self.arena.grad[s:e], op=dist.ReduceOp.SUM, group=self.group,
                                                     async_op=True))
        return hook

    # ---- one step ------------------------------------------------------------------
    def backward_and_step(self, loss: torch.Tensor):
        self.arena.zero_grad()
        if self.world > 1:
            self._pending = [n for (_, _, n) in self._buckets]
            self._handles = []
            self._launched = []
        loss.backward()
        if self.world > 1:
            for b, left in enumerate(self._pending):  # parameters that received no gradient this step
                if left > 0:
                    s, e, _ = self._buckets[b]
                    self._launched.append(b)
                    self._handles.append(dist.all_reduce(self.arena.grad[s:e], op=dist.ReduceOp.SUM,
                                                         group=self.group, async_op=True))
            if self.time_exposed and self.arena.data.is_cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for h in self._handles:
                    h.wait()
                e1.record()
                self._exposed_events.append((e0, e1))
            else:
                for h in self._handles:
                    h.wait()
        if self.torch_optimizer is None:
            ops.sgd_step(self.arena.data, self.arena.grad, self.momentum_buf, self.lr, self.momentum,
                         self.weight_decay, 1.0 / self.world, self.steps == 0)
        else:
            if self.world > 1:
                self.arena.grad.mul_(1.0 / self.world)
            self.torch_optimizer.step()
        self.steps += 1
        if self.world > 1:
            self._last_order = list(self._launched)

    def abort_step(self):
        """A step failed on THIS rank (exception-safe training, segmentation_training.py:276-283): drop its
        gradients.  At world > 1 the other ranks still all-reduce every bucket, so this rank waits for the buckets it
        has already launched and contributes zeros for the rest — in the launch order of the last complete step, which
        is the order the peers use — instead of leaving them blocked in `h.wait()`; the zeroed gradients then take part
        in the peers' averaged update of this step (this rank applies the same update, replicas stay equal)."""
        if self.world > 1:
            for h in self._handles:
                h.wait()
            launched = set(self._launched)
            self.arena.grad.zero_()
            rest = [b for b in (self._last_order or range(len(self._buckets))) if b not in launched]
            rest += [b for b in range(len(self._buckets)) if b not in launched and b not in rest]
            hs = []
            for b in rest:
                s, e, _ = self._buckets[b]
                hs.append(dist.all_reduce(self.arena.grad[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for h in hs:
                h.wait()
            self._handles, self._launched = [], []
            # the peers step with (their gradients + our zeros) / world: apply the identical update here
            if self.torch_optimizer is None:
                ops.sgd_step(self.arena.data, self.arena.grad, self.momentum_buf, self.lr, self.momentum,
                             self.weight_decay, 1.0 / self.world, self.steps == 0)
            else:
                self.arena.grad.mul_(1.0 / self.world)
                self.torch_optimizer.step()
            self.steps += 1
        self.arena.zero_grad()

    def exposed_allreduce_ms(self) -> List[float]:
        """Per-step device time the compute stream spent waiting for gradient all-reduces after backward had
        finished (what the overlap did NOT hide).  Synchronises."""
        if self._exposed_events:
            torch.cuda.synchronize()
            self.exposed_ms += [a.elapsed_time(b) for a, b in self._exposed_events]
            self._exposed_events = []
        return self.exposed_ms

    def set_lr(self, lr: float):
        self.lr = lr
        if self.torch_optimizer is not None:
            for group in self.torch_optimizer.param_groups:
                group["lr"] = lr


def cosine_lr(base_lr: float, step: int, max_steps: int, eta_min: float = 0.0) -> float:
    """CosineAnnealingLR (co3d_3d/src/modules/optim.py:103-120); see `schedules.cosine`."""
    from . import schedules
    return schedules.cosine(base_lr, max_steps, eta_min).lr(step)


def poly_lr(base_lr: float, step: int, max_steps: int, power: float = 0.9) -> float:
    """PolyLR = base_lr * (1 - step / (max_steps + 1)) ** power (optim.py:181-204); see `schedules.poly`."""
    from . import schedules
    return schedules.poly(base_lr, max_steps, power).lr(step)
