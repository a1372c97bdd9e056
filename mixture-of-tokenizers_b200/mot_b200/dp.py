"""Data-parallel plumbing of the path (the reference's only parallelism): tables are replicated, the
batch is sharded by rank, and the dense gradients of the three tensors the path owns are averaged across ranks.

Reference: `dist.broadcast(param, 0)` for every parameter at start (spt/train_gpt.py:1127-1128; runs/7:571-572) and one
`dist.all_reduce(param.grad, op=AVG)` per parameter per step (spt/train_gpt.py:1320-1321; runs/7:697-700).  Here the
gradients of the path live in ONE flat bucket `[embed_tokens.grad | embed_bytes.grad | mixin.grad ...]` that the
backward kernels write into directly, so the exchange is a single NCCL all-reduce (NVLS on NVSwitch) with no packing
copy.  Pure torch.distributed host logic: works with any backend (the CPU tests run it over gloo)."""
from __future__ import annotations

import os
import warnings
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_for_rank(tokens: torch.Tensor, pos: int, local: int, rank: int) -> torch.Tensor:
    """The reference's batch sharding: rank r takes tokens[pos + r*local : pos + (r+1)*local] (runs/7:468-474)."""
    return tokens[pos + rank * local: pos + (rank + 1) * local]


def broadcast_params(params: Iterable[torch.Tensor], src: int = 0, group=None) -> None:
    """Replicate the tables from rank `src` once (spt/train_gpt.py:1127-1128)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for p in params:
        dist.broadcast(p.detach(), src, group=group)


def own_allreduce_pays(group=None) -> bool:
    """Whether the library's NVLS all-reduce beats NCCL for the gradient bucket of this path (77 MB bf16 at the
    124M shape).  Measured on 8xB200 / NVSwitch (profiles/r1_experiments.md): 8 ranks 215 us vs NCCL 270 us;
    4 ranks 235 vs 214 us; 2 ranks 226 vs 174 us.  Through the switch every rank moves S*(1 + 1/n) bytes per
    direction (its own slice also travels to the switch and back), NCCL's ring 2*S*(n-1)/n: the in-switch reduction
    wins from n = 8 on.  MOT_DP_OWN=1 / MOT_DP_NCCL=1 force either."""
    if os.environ.get("MOT_DP_NCCL"):
        return False
    if os.environ.get("MOT_DP_OWN"):
        return True
    if not (dist.is_available() and dist.is_initialized()):
        return False
    return dist.get_world_size(group) >= 8


class GradBucket:
    """One flat buffer holding the gradients of `params` back to back (each slice 16-byte aligned), with a view per
    parameter.  `views()` are handed to the backward kernels as their dense-gradient outputs (mot_embed_bwd
    overwrites every row, so no zeroing is needed), `attach()` points `param.grad` at them, `all_reduce_avg()`
    averages the whole bucket across ranks in one collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype: Optional[torch.dtype] = None, symmetric=False,
                 group=None):
        """symmetric=True (CUDA, initialised NCCL group): allocate the bucket in symmetric memory with a multicast
        mapping so that all_reduce_avg() runs the library's own NVLS kernel (mot_dp_allreduce_avg) instead of NCCL.
        symmetric="auto": do that where the kernel was measured faster than NCCL (own_allreduce_pays)."""
        if symmetric == "auto":
            symmetric = own_allreduce_pays(group)
        self.params: List[torch.nn.Parameter] = list(params)
        if not self.params:
            raise ValueError("GradBucket needs at least one parameter")
        self.dtype = dtype or self.params[0].dtype
        dev = self.params[0].device
        esz = torch.empty(0, dtype=self.dtype).element_size()
        align = max(1, 16 // esz)
        self.offsets, n = [], 0
        for p in self.params:
            if p.device != dev:
                raise ValueError("GradBucket: parameters on different devices")
            self.offsets.append(n)
            n += (p.numel() + align - 1) // align * align
        self._symm = None
        self._epoch = 0
        n = (n + align - 1) // align * align
        self.flat = None
        if symmetric and dev.type == "cuda" and dist.is_available() and dist.is_initialized() \
                and not os.environ.get("MOT_DP_NCCL"):
            # Symmetric memory + multicast need NVSwitch and a driver with fabric support; where either is missing
            # the bucket is ordinary device memory and the exchange is NCCL's all-reduce (a library collective on
            # the same data, not a different code path for the kernels).
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = group if group is not None else dist.group.WORLD
                flat = symm_mem.empty(n, dtype=self.dtype, device=dev)
                flat.zero_()
                hdl = symm_mem.rendezvous(flat, grp)
                self.flat = flat
                if hdl.multicast_ptr != 0:       # NVSwitch multicast (NVLS) available
                    self._symm = hdl
            except Exception as e:  # noqa: BLE001 - any allocator / rendezvous failure means "no NVLS here"
                warnings.warn(f"GradBucket: symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL")
                self.flat = None
                self._symm = None
        if self.flat is None:
            self.flat = torch.zeros(n, dtype=self.dtype, device=dev)
        self._views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.params)]

    def views(self) -> List[torch.Tensor]:
        return self._views

    def view_of(self, param: torch.nn.Parameter) -> torch.Tensor:
        for p, v in zip(self.params, self._views):
            if p is param:
                return v
        raise KeyError("parameter is not in this bucket")

    def attach(self) -> None:
        """param.grad = its slice of the bucket (grads of a different dtype are copied in, e.g. an fp32 master-weight
        gradient into a bf16 bucket is NOT done silently: dtypes must match)."""
        for p, v in zip(self.params, self._views):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                if p.grad.dtype != self.dtype:
                    raise TypeError("GradBucket: gradient dtype differs from the bucket dtype")
                v.copy_(p.grad)
            p.grad = v

    def all_reduce_avg(self, group=None, async_op: bool = False):
        """One collective for the whole bucket.  AVG = SUM / world_size (gloo has no AVG reduce op)."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(group)
        if self._symm is not None and group is None and self.dtype in (torch.bfloat16, torch.float32):
            # one kernel per rank over the multicast mapping: reduce in the switch, average, multicast back
            from . import _lib as L
            h = self._symm
            self._epoch += 1
            dev = self.flat.device
            rc = L.lib().mot_dp_allreduce_avg(h.multicast_ptr, h.signal_pad_ptrs_dev, h.rank, h.world_size,
                                              self.flat.numel() * self.flat.element_size(),
                                              L.BF16 if self.dtype == torch.bfloat16 else L.F32, self._epoch,
                                              torch.cuda.current_stream(dev).cuda_stream)
            L.check(rc, "mot_dp_allreduce_avg")
            return None
        if dist.get_backend(group) == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=False)
        self.flat.div_(world)
        return work
