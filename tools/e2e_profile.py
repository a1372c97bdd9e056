"""cProfile of the host side of one module-level step (what bench.py's e2e leg times)."""
import cProfile, pstats, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch, mot_b200
d = torch.device("cuda:0")
N, Dt, bd, bpt = 49152, 768, 48, 16
mod = mot_b200.MoTEmbedding(50257, 458, Dt, bd, bpt, variant="V3").to(d).bfloat16()
bucket = mod.attach_grad_bucket()
tok_host = torch.randint(0, 50256, (N,), dtype=torch.int32).pin_memory()
ttb_tab = torch.randint(0, 458, (50257, bpt), dtype=torch.int32).to(torch.int16).to(d)
gout = torch.randn(1, N, Dt, device=d).bfloat16()
res_host = torch.empty(458, bd, dtype=torch.bfloat16).pin_memory()
def step():
    for p_ in mod.parameters():
        p_.grad = None
    t_in = tok_host.to(d, non_blocking=True)
    b_in = mot_b200.ttb_expand(t_in, ttb_tab, out_dtype=torch.int32).view(bpt, -1)
    x = mod(t_in, b_in)
    x.backward(gout)
    res_host.copy_(mod.embed_bytes.weight.grad, non_blocking=True)
    torch.cuda.current_stream().synchronize()
for _ in range(20): step()
import time
t0 = time.perf_counter()
for _ in range(200): step()
print("us/step", (time.perf_counter() - t0) / 200 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
