"""Forward only / forward+backward of one parity case, synchronising after each kernel (debug helper)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mot_b200
import test_gpu_parity as t
name, dt, N, what = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[dt]
variant, kw, (V, Vb, bpt, Dt, bd), slot_major, use_lam = t.CASES[name]
g = torch.Generator().manual_seed(0)
d = torch.device("cuda:0")
toks = torch.randint(0, V, (N,), generator=g, dtype=torch.int32).to(d)
ids = torch.randint(0, Vb, (N, bpt), generator=g).int()
ids = (ids.t().contiguous() if slot_major else ids).to(d)
E_tok = torch.randn(V, max(Dt, 8), generator=g).to(dtype).to(d).requires_grad_(True) if Dt else None
E_byte = torch.randn(Vb, bd, generator=g).to(dtype).to(d).requires_grad_(True) if kw["combine"] != "tok_only" else None
lam = torch.tensor([0.7, 0.4], device=d, requires_grad=True) if use_lam else None
out = mot_b200.mot_embed(toks if E_tok is not None else None, ids if E_byte is not None else None, E_tok, E_byte,
                         mot_b200.MixSpec(**kw), bpt=bpt, lam=lam)
torch.cuda.synchronize()
print("fwd ok", float(out.float().abs().mean()))
if what == "bwd":
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
    print("bwd ok")
