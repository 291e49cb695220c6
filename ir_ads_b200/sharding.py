"""Multi-GPU host logic of the MSDeformAttn path: batch sharding and the one collective around it.

The op shards by image with NO collective inside (every output row depends only on its own
image's value); the reference gets this from DDP (detectron2/detectron2/engine/defaults.py:60-79,
total batch split by build_detection_train_loader(total_batch_size=...),
projects/vCLR_deformable_mask/configs/.../deformable_train_voc_eval_nonvoc.py:286).  The only
exchange step next to the op is the gradient all-reduce of the module's projection weights
(230 272 parameters per module), done here with one flat NCCL all-reduce per step.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_batch(total_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(first image, image count) of `rank`.  Like detectron2's loader, the total batch must divide
    evenly (build_detection_train_loader asserts total_batch_size % world_size == 0)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    if total_batch % world_size != 0:
        raise ValueError(f"total batch {total_batch} is not divisible by world size {world_size}")
    per = total_batch // world_size
    return rank * per, per


def projection_parameters(modules: Iterable[torch.nn.Module]) -> List[torch.nn.Parameter]:
    """The parameters whose gradients are exchanged: sampling_offsets / attention_weights /
    value_proj / output_proj of every MultiScaleDeformableAttention (multi_scale_deform_attn.py:193-196)."""
    out = []
    for m in modules:
        for name in ("sampling_offsets", "attention_weights", "value_proj", "output_proj"):
            lin = getattr(m, name)
            out.extend([lin.weight, lin.bias])
    return out


class GradBucket:
    """One flat buffer holding all projection-weight gradients: every ``p.grad`` is a VIEW into it
    (as DDP's gradient_as_bucket_view does), so the exchange is exactly one in-place all-reduce and
    one scale per step -- no gather / scatter copies.  The payload is ~0.9 MB per module, so the
    collective is latency- not bandwidth-bound.  Use ``zero_()`` instead of
    ``optimizer.zero_grad(set_to_none=True)`` to keep the views attached."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = list(params)
        n = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(n, dtype=first.dtype, device=first.device)
        o = 0
        for p in self.params:
            view = self.flat[o:o + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            o += p.numel()

    def zero_(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self, group=None) -> None:
        """grad <- mean over ranks (DDP semantics).  No-op without an initialised process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


class OverlappedGradSync:
    """DDP-style bucketed gradient exchange for the MSDeformAttn modules of a layer stack: one bucket per module
    (8 projection tensors, 230 272 parameters = 0.92 MB), whose mean all-reduce is LAUNCHED FROM A GRADIENT HOOK the
    moment the module's last gradient has been accumulated, so it runs (on NCCL's own stream) while autograd is still
    working on the earlier modules -- what DistributedDataParallel's reducer does for the reference
    (detectron2/detectron2/engine/defaults.py:60-79).  ``finish()`` waits for the outstanding collectives and scales
    by 1 / world_size.  Gradients are views into one flat buffer (see GradBucket), so nothing is copied.

    Without an initialised process group (or world_size 1) the hooks only count; ``finish()`` is then a no-op."""

    def __init__(self, module_params: Sequence[Sequence[torch.nn.Parameter]], group=None):
        self.group = group
        self._pending: List[int] = []
        self._works: list = []
        self._handles = []
        all_params = [p for ps in module_params for p in ps]
        self.all = GradBucket(all_params)              # one flat buffer; per-module buckets are slices of it
        o = 0
        self.slices = []
        for k, ps in enumerate(module_params):
            n = sum(p.numel() for p in ps)
            self.slices.append(self.all.flat[o:o + n])
            o += n
            self._pending.append(len(ps))
            for p in ps:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(k)))
        self._sizes = [len(ps) for ps in module_params]

    def _active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _make_hook(self, k: int):
        def hook(_param):
            self._pending[k] -= 1
            if self._pending[k] == 0 and self._active():
                self._works.append(dist.all_reduce(self.slices[k], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        return hook

    def zero_(self) -> None:
        """Start of a step: clear the gradients (the views stay attached) and re-arm the hooks."""
        self.all.zero_()
        self._pending = list(self._sizes)
        self._works = []

    def finish(self) -> None:
        """End of backward: wait for the collectives launched by the hooks, then grad <- mean over ranks."""
        for w in self._works:
            w.wait()
        if self._active():
            launched = len(self._works)
            if launched != len(self.slices):             # a module whose parameters got no gradient this step
                raise RuntimeError(f"{len(self.slices) - launched} bucket(s) never became ready: every projection "
                                   "weight must receive a gradient (find_unused_parameters=False semantics)")
            self.all.flat.div_(dist.get_world_size(self.group))
        self._works = []

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def max_over_ranks(value: float, device) -> float:
    """Step time of the job = the slowest rank's (timing rule of bench.py)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
