"""Micro-benchmarks that bracket the forward kernel: plain copy, torch gather, our gather-only, our V3."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch, torch.nn.functional as F
import mot_b200
from mot_b200 import ops
d = torch.device("cuda:0")
N, V, Dt, bd, bpt = 49152, 50257, 768, 48, 16
g = torch.Generator(device=d).manual_seed(0)
E = torch.randn(V, Dt, generator=g, device=d).bfloat16()
Eb = torch.randn(458, bd, generator=g, device=d).bfloat16()
tok = torch.randint(0, V, (N,), generator=g, device=d, dtype=torch.int32)
tok_sorted = tok.sort().values.int()
ids = torch.randint(0, 458, (bpt, N), generator=g, device=d, dtype=torch.int32)
src = torch.randn(N, Dt, generator=g, device=d).bfloat16()
dst = torch.empty_like(src)
out = torch.empty(N, Dt, dtype=torch.bfloat16, device=d)
def timeit(fn, n=50):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
big_a = torch.empty(256 << 20, dtype=torch.uint8, device=d); big_b = torch.empty_like(big_a)
print("copy 256MB->256MB      us", timeit(lambda: big_b.copy_(big_a)), "(512 MB traffic)")
print("copy 75MB (N x Dt)     us", timeit(lambda: dst.copy_(src)))
print("torch F.embedding      us", timeit(lambda: F.embedding(tok.long(), E)))
tl = tok.long()
print("torch index_select     us", timeit(lambda: torch.index_select(E, 0, tl, out=out)))
tsl = tok_sorted.long()
print("torch index_select srt us", timeit(lambda: torch.index_select(E, 0, tsl, out=out)))
for name, spec, use_ids in [("ours tok_only no norm", mot_b200.MixSpec(combine="tok_only", out_norm=False), False),
                            ("ours tok_only + norm ", mot_b200.MixSpec(combine="tok_only", out_norm=True), False),
                            ("ours V3 (sum + norm) ", mot_b200.MixSpec(combine="add", slot_major=True), True)]:
    desc = ops.make_desc(spec, N, E, Eb if use_ids else None, bpt, ids=ids if use_ids else None, ttb=None, has_lam=False)
    print(name, " us", timeit(lambda: ops.embed_forward_out(desc, tok, ids if use_ids else None, None, E, Eb if use_ids else None, None, out)))
    print(name, "sorted tok us", timeit(lambda: ops.embed_forward_out(desc, tok_sorted, ids if use_ids else None, None, E, Eb if use_ids else None, None, out)))
