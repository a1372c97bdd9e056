"""Per-warp timeline of the main forward / backward kernel (library built with -DMOT_TRACE):
    MOT_LIB_SUFFIX=_trace python tools/trace_timeline.py [N] [Dt] [bd]
Prints, relative to the earliest stamp of the kernel: when warps enter, pass griddepcontrol.wait, have their first
batch / first data, per-occurrence intervals, and when they finish (percentiles over all warps)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import numpy as np
import torch
import mot_b200
from mot_b200 import ops, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
Dt = int(sys.argv[2]) if len(sys.argv) > 2 else 768
bd = int(sys.argv[3]) if len(sys.argv) > 3 else 48
bpt, V = 16, 50257
d = torch.device("cuda:0")
g = torch.Generator(device=d).manual_seed(1)
E_tok = torch.randn(V, Dt, generator=g, device=d).bfloat16()
E_byte = torch.randn(458, bd, generator=g, device=d).bfloat16()
tok = torch.randint(0, V - 1, (N,), generator=g, device=d, dtype=torch.int32)
ids = torch.randint(0, 458, (bpt, N), generator=g, device=d, dtype=torch.int32)
gout = torch.randn(N, Dt, generator=g, device=d).bfloat16()
out = torch.empty(N, Dt, dtype=torch.bfloat16, device=d)
rstd = torch.empty(N, dtype=torch.float32, device=d)
gt, gb = torch.empty_like(E_tok), torch.empty_like(E_byte)
spec = mot_b200.MixSpec(combine="add", slot_major=True)
desc = ops.make_desc(spec, N, E_tok, E_byte, bpt, ids=ids, ttb=None, has_lam=False, seq_len=N)
ws = ops.acquire_workspace(desc, d)
trace = torch.zeros(3 * 4096 * 64, dtype=torch.int64, device=d)
trace2 = torch.zeros(3 * 4096 * 64, dtype=torch.int64, device=d)
lib = _lib.lib()


def step():
    ops.embed_plan_async(desc, tok, ws, d)
    ops.embed_forward_out(desc, tok, ids, None, E_tok, E_byte, None, out, rstd=rstd)
    ops.embed_plan_join(ws, d)
    ops.embed_backward_out(desc, tok, ids, None, E_tok, E_byte, None, gout, gt, gb, None, ws.buf, plan_ready=True,
                           ws_clean=True, out_saved=out, rstd=rstd, plan_joined=True)
    ws.clean = True


for _ in range(20):
    step()
torch.cuda.synchronize()
lib.mot_profile_trace(trace.data_ptr())
step()
lib.mot_profile_trace(trace2.data_ptr())   # the next step, back to back: where does the time between two steps go?
step()
step()
torch.cuda.synchronize()
lib.mot_profile_trace(None)
t = trace.cpu().numpy().reshape(3, 4096, 64)
t2 = trace2.cpu().numpy().reshape(3, 4096, 64)
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{N}_{Dt}.npy"), t)




def pct(x):
    return "p0 %.2f p10 %.2f p50 %.2f p90 %.2f p100 %.2f" % tuple(np.percentile(x, [0, 10, 50, 90, 100]))


f0 = t[0][t[0][:, 0] > 0]
b0 = t[1][t[1][:, 0] > 0]
T0 = f0[:, 0].min()
print("absolute (us after the first forward warp entered): fwd last end %.2f | bwd first entry %.2f, pdl_wait returns %.2f..%.2f, last end %.2f"
      % ((f0[:, 62].max() - T0) / 1e3, (b0[:, 0].min() - T0) / 1e3, (b0[:, 1].min() - T0) / 1e3, (b0[:, 1].max() - T0) / 1e3,
         (b0[:, 62].max() - T0) / 1e3))
z0 = t[2][t[2][:, 0] > 0]
n0 = t2[0][t2[0][:, 0] > 0]
print("  finalize: first entry %.2f, pdl_wait returns %.2f..%.2f, last end %.2f | next step: fwd first entry %.2f, first pdl_wait return %.2f"
      % ((z0[:, 0].min() - T0) / 1e3, (z0[:, 1].min() - T0) / 1e3, (z0[:, 1].max() - T0) / 1e3, (z0[:, 62].max() - T0) / 1e3,
         (n0[:, 0].min() - T0) / 1e3, (n0[:, 1].min() - T0) / 1e3))
bw = t[1]
sm_end = {}
for gw_ in range(1776):
    if bw[gw_, 0] > 0:
        sm_end.setdefault(gw_ % 148, []).append((bw[gw_, 61] - T0) / 1e3)
sm_mean = np.array([np.mean(v) for v in sm_end.values()])
sm_spread = np.array([np.max(v) - np.min(v) for v in sm_end.values()])
print("  backward stream end per SM (mean over its 12 warps):", pct(sm_mean), "| spread inside an SM:", pct(sm_spread))
for name, a in (("forward", t[0]), ("backward", t[1])):
    live = a[:, 0] > 0
    a = a[live]
    t0 = a[:, 0].min()
    us = lambda col: (a[:, col] - t0) / 1e3   # noqa: E731
    print(f"== {name}: {live.sum()} warps, occurrences per warp {pct(a[:, 63])}")
    print("  kernel entry      ", pct(us(0)))
    print("  after pdl_wait    ", pct(us(1)))
    print("  slot 2            ", pct(us(2)), "(fwd: ring prologue issued; bwd: first batches loaded)")
    print("  slot 3            ", pct(us(3)), "(fwd: byte table staged; bwd: first copies issued)")
    if name == "backward":
        print("  after zero fill   ", pct(us(4)))
    n_occ = a[:, 63].astype(int)
    first = us(5)
    print("  first data        ", pct(first))
    last = np.array([(a[i, 5 + min(n_occ[i], 57) - 1] - t0) / 1e3 for i in range(len(a))])
    print("  last data         ", pct(last))
    print("  end               ", pct(us(62)))
    per = np.array([(a[i, 5 + min(n_occ[i], 57) - 1] - a[i, 5]) / 1e3 / max(min(n_occ[i], 57) - 1, 1) for i in range(len(a))])
    print("  us per occurrence ", pct(per))
    # per-occurrence interval profile over the stream (median over warps of the k-th interval)
    K = int(np.median(n_occ))
    iv = [np.median((a[:, 5 + k + 1] - a[:, 5 + k]) / 1e3) for k in range(min(K, 56) - 1)]
    print("  median interval by occurrence index:", " ".join(f"{x:.2f}" for x in iv))
