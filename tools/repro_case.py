"""Run one parity case of tests/test_gpu_parity.py by name (debug helper; e.g. under compute-sanitizer)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as t
name = sys.argv[1]
dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[sys.argv[2]]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 37
t.run_case(name, dtype, N=N)
torch.cuda.synchronize()
print("ok", name, dtype, N)
