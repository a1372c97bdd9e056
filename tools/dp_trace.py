"""Timeline of one data-parallel step on every rank (library built with -DMOT_TRACE):
    MOT_LIB_SUFFIX=_trace torchrun --nproc-per-node N tools/dp_trace.py [n_slabs]
Rank 0 prints, relative to its first forward warp: end of forward / backward / finalize and, per exchange launch, when
its blocks entered, passed griddepcontrol.wait, passed the entry barrier, finished their data and passed the exit barrier."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import numpy as np
import torch
import torch.distributed as dist
import mot_b200
from mot_b200 import ops, _lib, dp

rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
d = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=d)
world = dist.get_world_size()
n_slabs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N, Dt, bd, bpt, V = 49152, 768, 48, 16, 50257
g = torch.Generator(device=d).manual_seed(1)
E_tok = torch.randn(V, Dt, generator=g, device=d).bfloat16()
E_byte = torch.randn(458, bd, generator=g, device=d).bfloat16()
g = torch.Generator(device=d).manual_seed(100 + rank)
tok = torch.randint(0, V - 1, (N,), generator=g, device=d, dtype=torch.int32)
ids = torch.randint(0, 458, (bpt, N), generator=g, device=d, dtype=torch.int32)
gout = torch.randn(N, Dt, generator=g, device=d).bfloat16()
out = torch.empty(N, Dt, dtype=torch.bfloat16, device=d)
rstd = torch.empty(N, dtype=torch.float32, device=d)
bucket = dp.GradBucket([torch.nn.Parameter(E_tok, requires_grad=False), torch.nn.Parameter(E_byte, requires_grad=False)],
                       symmetric=True, n_slabs=max(n_slabs, 1))
gt, gb = bucket.views()
spec = mot_b200.MixSpec(combine="add", slot_major=True)
desc = ops.make_desc(spec, N, E_tok, E_byte, bpt, ids=ids, ttb=None, has_lam=False, seq_len=N, dp_slabs=n_slabs if n_slabs > 1 else 0)
ws = ops.acquire_workspace(desc, d)
trace = torch.zeros((3 * 4096 + 16 * 64) * 64, dtype=torch.int64, device=d)
lib = _lib.lib()


def step():
    ops.embed_plan_async(desc, tok, ws, d)
    ops.embed_forward_out(desc, tok, ids, None, E_tok, E_byte, None, out, rstd=rstd)
    ops.embed_plan_join(ws, d)
    if n_slabs > 1:
        for k in range(n_slabs):
            ops.embed_backward_slab_out(desc, tok, ids, None, E_tok, E_byte, None, gout, out, rstd, gt, gb, None, ws.buf, k, n_slabs,
                                        reserve_sms=bucket.reserve_sms if k > 0 else 0, plan_joined=True)
            lo, hi = ops.slab_rows(V, k, n_slabs)
            bucket.exchange_async(lo * Dt, hi * Dt if k < n_slabs - 1 else bucket.flat.numel(), last=(k == n_slabs - 1))
        bucket.wait()
    else:
        ops.embed_backward_out(desc, tok, ids, None, E_tok, E_byte, None, gout, gt, gb, None, ws.buf, plan_ready=True,
                               ws_clean=True, out_saved=out, rstd=rstd, plan_joined=True)
        bucket.all_reduce_avg()
    ws.clean = True


for _ in range(20):
    step()
dist.barrier(); torch.cuda.synchronize()
for _ in range(5):      # a run of steps, the last one traced (steady state: queues are full)
    step()
lib.mot_profile_trace(trace.data_ptr())
step()
lib.mot_profile_trace(None)
step()
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(-1, 64)
f, b, z, x = t[:4096], t[4096:8192], t[8192:12288], t[12288:].reshape(16, 64, 64)
T0 = f[f[:, 0] > 0][:, 0].min()
us = lambda a: (a - T0) / 1e3   # noqa: E731
if rank == 0:
    bl = b[b[:, 0] > 0]
    zl = z[z[:, 0] > 0]
    print(f"world {world}, {n_slabs} slab(s), exchange {bucket.algo}; times in us after the first forward warp entered")
    print(f"  forward ends {us(f[f[:, 0] > 0][:, 62].max()):.1f}; last backward launch: entry {us(bl[:, 0].min()):.1f} end {us(bl[:, 62].max()):.1f}; "
          f"last finalize: entry {us(zl[:, 0].min()):.1f} end {us(zl[:, 62].max()):.1f}")
    for i in range(16):
        xi = x[i][x[i][:, 0] > 0]
        if len(xi) == 0:
            continue
        print(f"  exchange launch {i}: {len(xi)} blocks | entry {us(xi[:, 0].min()):.1f} | after pdl_wait {us(xi[:, 1].min()):.1f}..{us(xi[:, 1].max()):.1f} "
              f"| after entry barrier {us(xi[:, 2].min()):.1f}..{us(xi[:, 2].max()):.1f} | data done {us(xi[:, 3].min()):.1f}..{us(xi[:, 3].max()):.1f} "
              f"| after exit barrier {us(xi[:, 4].max()):.1f}")
dist.destroy_process_group()
