"""Quick GPU check of the tcgen05 projection kernels against torch.matmul (all three operand-order flavours)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
import torch
from mot_b200 import _lib as L
lib = L.lib()
d = torch.device("cuda:0")
st = lambda: torch.cuda.current_stream().cuda_stream
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for (n, K, Do) in [(128, 64, 256), (256, 128, 256), (1000, 1024, 1024), (4096, 2048, 1024), (777, 1920, 1024), (65536, 2048, 1024)]:
    g = torch.Generator(device=d).manual_seed(n)
    x = torch.randn(n, K, generator=g, device=d).bfloat16()
    w = (torch.randn(Do, K, generator=g, device=d) / K ** 0.5).bfloat16()
    dy = torch.randn(n, Do, generator=g, device=d).bfloat16()
    y = torch.empty(n, Do, dtype=torch.bfloat16, device=d)
    rc = lib.mot_linear_fwd(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), n, K, Do, 0, 0, st()); torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    print(f"n={n} K={K} Do={Do} fwd rc={rc} err={rel(y, ref):.2e}", end=" | ")
    dx = torch.empty(n, K, dtype=torch.bfloat16, device=d)
    rc = lib.mot_linear_bwd_input(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), n, K, Do, 0, None, 0, st()); torch.cuda.synchronize()
    print(f"dX rc={rc} err={rel(dx, dy.float() @ w.float()):.2e}", end=" | ")
    dw = torch.empty(Do, K, dtype=torch.float32, device=d); dwb = torch.empty(Do, K, dtype=torch.bfloat16, device=d)
    rc = lib.mot_linear_bwd_weight(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), dwb.data_ptr(), n, K, Do, 0, None, 0, st()); torch.cuda.synchronize()
    refw = dy.float().t() @ x.float()
    print(f"dW rc={rc} err={rel(dw, refw):.2e} bf16 {rel(dwb, refw):.2e}")
# timing at the runs/7 shape
n, K, Do = 65536, 2048, 1024
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
fl = 2.0 * n * K * Do
print("fwd  %.3f ms %.0f TFLOP/s | torch %.3f ms" % ((tf := t(lambda: lib.mot_linear_fwd(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), n, K, Do, 0, 0, st()))), fl / tf / 1e9, t(lambda: torch.matmul(x, w.t()))))
print("dX   %.3f ms %.0f TFLOP/s | torch %.3f ms" % ((tf := t(lambda: lib.mot_linear_bwd_input(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), n, K, Do, 0, None, 0, st()))), fl / tf / 1e9, t(lambda: torch.matmul(dy, w))))
print("dW   %.3f ms %.0f TFLOP/s | torch %.3f ms" % ((tf := t(lambda: lib.mot_linear_bwd_weight(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), dwb.data_ptr(), n, K, Do, 0, None, 0, st()))), fl / tf / 1e9, t(lambda: torch.matmul(dy.t(), x))))
