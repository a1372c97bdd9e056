"""GPU timeline of one bench step (kernel durations and the gaps between them) via torch.profiler / CUPTI.
    python tools/timeline.py [workload]   -> prints the kernels of the last profiled step in launch order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch
import bench, mot_b200
from mot_b200 import ops
from torch.profiler import profile, ProfilerActivity

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
w = dict(bench.WORKLOADS[name]); w["name"] = name
d = torch.device("cuda:0")
N, Dt, bd, bpt = w["N"], w["Dt"], w["bd"], w["bpt"]
dt = torch.bfloat16
_, _, Do = bench.algorithmic_bytes(w)
spec_kw = dict(mot_b200.RUN_VARIANTS[w["variant"]]); slot_major = spec_kw.get("slot_major", False)
spec = mot_b200.MixSpec(**spec_kw)
g = torch.Generator(device=d).manual_seed(1)
E_tok = torch.randn(bench.V_TOK, Dt, generator=g, device=d).to(dt)
E_byte = torch.randn(bench.V_BYTE, bd, generator=g, device=d).to(dt)
tok = bench.make_tokens(N, "uniform", 5).to(d)
ids = torch.randint(0, bench.V_BYTE, (bpt, N) if slot_major else (1, N * bpt), generator=g, device=d, dtype=torch.int32)
gout = torch.randn(N, Do, generator=g, device=d).to(dt)
lam = torch.tensor([0.5, 0.5], device=d) if w["variant"] in ("V3c", "V3d") else None
out = torch.empty(N, Do, dtype=dt, device=d)
gE_tok, gE_byte = torch.empty_like(E_tok), torch.empty_like(E_byte)
g_lam = torch.empty(2, device=d) if lam is not None else None
desc = ops.make_desc(spec, N, E_tok, E_byte, bpt, ids=ids, ttb=None, has_lam=lam is not None, seq_len=N)
ws = ops.acquire_workspace(desc, d)
main_stream = torch.cuda.current_stream(d)
def step():
    ops.embed_plan_async(desc, tok, ws, d)
    ops.embed_forward_out(desc, tok, ids, None, E_tok, E_byte, lam, out)
    ops.embed_plan_join(ws, d)
    ops.embed_backward_out(desc, tok, ids, None, E_tok, E_byte, lam, gout, gE_tok, gE_byte, g_lam, ws.buf, plan_ready=True, ws_clean=True)
    ws.clean = True
for _ in range(10):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
per_step = len(evs) // 10
last = evs[-per_step:]
t0 = last[0].time_range.start
prev_end = None
print(f"{'start_us':>9} {'dur_us':>8} {'gap_us':>7}  kernel")
for e in last:
    gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f} {gap:7.1f}  {e.name[:70]}")
    prev_end = e.time_range.end
print(f"step span {last[-1].time_range.end - t0:.1f} us; steady-state period {(evs[-1].time_range.end - evs[-per_step*5].time_range.start)/5:.1f} us")
