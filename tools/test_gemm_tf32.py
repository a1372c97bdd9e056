import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch
from mot_b200 import ops
d = torch.device("cuda:0")
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for (n, K, Do) in [(33, 64, 32), (128, 32, 256), (256, 256, 256), (1000, 1280, 256)]:
    g = torch.Generator(device=d).manual_seed(n)
    x = torch.randn(n, K, generator=g, device=d); w = torch.randn(Do, K, generator=g, device=d) / K ** 0.5
    dy = torch.randn(n, Do, generator=g, device=d)
    y = torch.full((n, Do), 7.0, device=d); dx = torch.full((n, K), 7.0, device=d); dw = torch.full((Do, K), 7.0, device=d)
    ops.linear_forward_out(x, w, y); ops.linear_bwd_input_out(dy, w, dx); ops.linear_bwd_weight_out(dy, x, dw)
    torch.cuda.synchronize()
    rdx = dy.double() @ w.double(); rdw = dy.double().t() @ x.double()
    print(n, K, Do, "fwd", rel(y, x.double() @ w.double().t()), "dx", rel(dx, rdx), "dw", rel(dw, rdw))
    print("   dx[0,:6]", dx[0, :6].tolist(), "ref", rdx[0, :6].tolist())
    print("   dw[0,:6]", dw[0, :6].tolist(), "ref", rdw[0, :6].tolist())
