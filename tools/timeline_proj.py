"""GPU timeline of one module-level step of the concat+projection variant (torch.profiler / CUPTI)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch, mot_b200
from torch.profiler import profile, ProfilerActivity
d = torch.device("cuda:0")
Dt, bd, Do = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1024, 64, 1024)))
N, bpt = 65536, 16
m = mot_b200.MoTProjEmbedding(50257, 458, Dt, bd, Do, bpt, variant="V1").to(d).bfloat16()
tok = torch.randint(0, 50256, (N,), device=d, dtype=torch.int32)
ids = torch.randint(0, 458, (1, N * bpt), device=d, dtype=torch.int32)
gout = torch.randn(1, N, Do, device=d).bfloat16()
def step():
    for p in m.parameters(): p.grad = None
    m(tok, ids).backward(gout)
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
last = evs[-(len(evs) // 3):]
t0 = last[0].time_range.start
for e in last:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:9.1f}  {e.name[:90]}")
