"""BASELINE.json configs[4]: throughput sweep 1K..1M tokens per GPU of the ttb expansion, the pull, and the fused
byte-mix embedding fwd+bwd (V3 sum 768/48 and 1024/64; V1 concat+projection 1024/64/1024), uniform and Zipf tokens.
Prints a markdown table (committed as profiles/r1_sweep.md).  One GPU; the same calls bench.py times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import numpy as np, torch
import bench, mot_b200
from mot_b200 import ops

d = torch.device("cuda:0")
V, Vb, bpt = 50257, 458, 16
tab8 = np.load(os.path.join(ROOT, "tests", "golden", "ttb_8_left_pad.npz"))["table"]
tab = np.full((V, bpt), 456, dtype=np.int16); tab[:50256, 8:] = tab8; tab[50256] = 457   # left-padded to 16
ttb = torch.from_numpy(tab).to(d)

def timeit(fn, reps):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us

def embed_step(N, Dt, bd, dist):
    dt = torch.bfloat16
    g = torch.Generator(device=d).manual_seed(1)
    E_tok = torch.randn(V, Dt, generator=g, device=d).to(dt); E_byte = torch.randn(Vb, bd, generator=g, device=d).to(dt)
    tok = bench.make_tokens(N, dist, 7).to(d)
    ids = torch.randint(0, Vb, (bpt, N), generator=g, device=d, dtype=torch.int32)
    gout = torch.randn(N, Dt, generator=g, device=d).to(dt); out = torch.empty(N, Dt, dtype=dt, device=d)
    gE_tok, gE_byte = torch.empty_like(E_tok), torch.empty_like(E_byte)
    spec = mot_b200.MixSpec(combine="add", slot_major=True)
    desc = ops.make_desc(spec, N, E_tok, E_byte, bpt, ids=ids, ttb=None, has_lam=False, seq_len=N)
    ws = ops.acquire_workspace(desc, d)
    # keep out + rstd where the library's saved-output backward would use them (up to 4 positions per vocabulary row)
    rstd = torch.empty(N, dtype=torch.float32, device=d) if ops.embed_bwd_uses_saved(desc) else None
    def step():
        ops.embed_plan_async(desc, tok, ws, d)
        ops.embed_forward_out(desc, tok, ids, None, E_tok, E_byte, None, out, rstd=rstd)
        ops.embed_plan_join(ws, d)
        ops.embed_backward_out(desc, tok, ids, None, E_tok, E_byte, None, gout, gE_tok, gE_byte, None, ws.buf, plan_ready=True,
                               ws_clean=True, out_saved=out if rstd is not None else None, rstd=rstd)
        ws.clean = True
    return step

def proj_step(N, Dt, bd, Do):
    m = mot_b200.MoTProjEmbedding(V, Vb, Dt, bd, Do, bpt, variant="V1").to(d).bfloat16()
    tok = bench.make_tokens(N, "uniform", 7).to(d)
    ids = torch.randint(0, Vb, (1, N * bpt), device=d, dtype=torch.int32)
    gout = torch.randn(1, N, Do, device=d).bfloat16()
    def step():
        for p in m.parameters(): p.grad = None
        m(tok, ids).backward(gout)
    return step

print("| tokens/GPU | ttb_expand Mtok/s | pull_from_left Mtok/s | V3 768/48 uniform Mtok/s (us) | V3 768/48 zipf | V3 1024/64 uniform | V1 proj 1024/64/1024 Mtok/s (ms) |")
print("|---|---|---|---|---|---|---|")
for N in [1024, 4096, 16384, 65536, 262144, 1048576]:
    reps = 200 if N <= 65536 else 20
    tok = bench.make_tokens(N, "zipf", 3).to(d)
    t_exp = timeit(lambda: mot_b200.ttb_expand(tok, ttb, out_dtype=torch.int32), reps)
    by = mot_b200.ttb_expand(tok, ttb, out_dtype=torch.int32)
    t_pull = timeit(lambda: mot_b200.pull_from_left(by, bpt), reps)
    t_a = timeit(embed_step(N, 768, 48, "uniform"), reps)
    t_z = timeit(embed_step(N, 768, 48, "zipf"), reps)
    t_b = timeit(embed_step(N, 1024, 64, "uniform"), reps)
    t_p = timeit(proj_step(N, 1024, 64, 1024), max(5, reps // 10))
    f = lambda t: f"{N / t:.1f} ({t:.1f})"
    print(f"| {N} | {N / t_exp:.0f} | {N / t_pull:.0f} | {f(t_a)} | {f(t_z)} | {f(t_b)} | {N / t_p:.1f} ({t_p / 1e3:.3f}) |", flush=True)
