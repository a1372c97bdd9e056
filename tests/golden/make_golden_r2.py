"""Round-2 additions to the golden fixtures, generated FROM THE REFERENCE ITSELF like make_golden.py (same method,
separate files so that the round-1 fixtures stay byte-identical):

  * runs_float_r2.npz : V3e, runs/71042_mot-in_toks-valemb.py (mixin_bytes :225-228 + the forward lines :311-314,
                        lambdas divided by their sum), fp32 and bf16
  * mixout.npz        : the output-side expands of scaled-pre-train, ByteMixoutCopy / ByteMixoutSplit
                        (train_gpt.py:483-518) with n_layer_out = 0 (no attention layers: forward = the expand alone)

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_r2.py
"""
from __future__ import annotations

import dataclasses
import os
import sys
import zlib
from typing import Literal

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

from make_golden import HERE, REF, extract, np_


def golden_v3e():
    out = {}
    T, bpt, bd, V, Vb = 24, 16, 8, 80, 458
    tag = "V3e_run71042"
    ns = dict(torch=torch, nn=nn, F=F, Tensor=Tensor)
    extract(f"{REF}/modded-nanogpt/runs/71042_mot-in_toks-valemb.py", ["norm", "mixin_bytes"], ns)
    norm, mixin_bytes = ns["norm"], ns["mixin_bytes"]
    for dt_tag, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        torch.manual_seed(zlib.crc32(tag.encode()) % 997 + (0 if dt_tag == "f32" else 1))
        Dt = bpt * bd
        embed_tokens = nn.Embedding(V, Dt).to(dtype)
        embed_bytes = nn.Embedding(Vb, bd).to(dtype)
        scalars = nn.Parameter(torch.tensor([0.4, 1.1]))  # [-2] = byte lambda, [-1] = token lambda; sum != 1
        token_inputs = torch.randint(0, V, (T,), dtype=torch.int32)
        flat = torch.randint(0, Vb, (1, T * bpt), dtype=torch.int32)
        byte_inputs = flat.view(bpt, -1).contiguous()          # runs/71042:478 `.view(16,-1)`
        # runs/71042:311-314
        norm_scalrs_sum = scalars[-1] + scalars[-2]
        x_toks = norm(embed_tokens(token_inputs)[None]) * scalars[-1] / norm_scalrs_sum
        x_bytes = norm(embed_bytes(byte_inputs).squeeze()) * scalars[-2] / norm_scalrs_sum
        x = mixin_bytes(x_toks, x_bytes)
        gout = torch.randn(x.shape).to(x.dtype)
        x.backward(gout)
        k = f"{tag}_{dt_tag}"
        out[f"{k}_tokens"] = np_(token_inputs)
        out[f"{k}_byte_inputs"] = np_(byte_inputs)
        out[f"{k}_E_tok"] = np_(embed_tokens.weight)
        out[f"{k}_E_byte"] = np_(embed_bytes.weight)
        out[f"{k}_gE_tok"] = np_(embed_tokens.weight.grad)
        out[f"{k}_gE_byte"] = np_(embed_bytes.weight.grad)
        out[f"{k}_scalars"] = np_(scalars)
        out[f"{k}_gscalars"] = np_(scalars.grad)
        out[f"{k}_out"] = np_(x)
        out[f"{k}_gout"] = np_(gout)
    np.savez_compressed(f"{HERE}/runs_float_r2.npz", **out)
    print("runs float r2:", len(out), "arrays")


def golden_mixout():
    import einops
    ns = dict(torch=torch, nn=nn, F=F, einops=einops, Tensor=Tensor, Literal=Literal, dataclass=dataclasses.dataclass)
    extract(f"{REF}/scaled-pre-train/train_gpt.py",
            ["ByteHyperparameters", "ModelDims", "norm", "ByteMixoutCopy", "ByteMixoutSplit"], ns)
    ns["ByteSelfAttn"] = None      # never constructed with n_layer_out = 0
    out = {}
    for tag, (B, S, D, bpt, dtype) in {
        "copy_f32": (2, 5, 32, 4, torch.float32),
        "copy_bf16": (1, 7, 64, 16, torch.bfloat16),
        "split_f32": (2, 5, 32, 4, torch.float32),
        "split_bf16": (1, 7, 64, 16, torch.bfloat16),
    }.items():
        torch.manual_seed(zlib.crc32(tag.encode()) % 1000 + 3)
        bp = ns["ByteHyperparameters"](bytes_per_token=bpt, n_layer_out=0)
        dims = ns["ModelDims"](model_dim=D, byte_dim=D // bpt, token_dim=D)
        mod = ns["ByteMixoutCopy" if tag.startswith("copy") else "ByteMixoutSplit"](dims, 16, bp)
        x = torch.randn(B, S, D).to(dtype).requires_grad_(True)
        y = mod(x)
        gout = torch.randn(y.shape).to(dtype)
        y.backward(gout)
        out[f"{tag}_x"] = np_(x)
        out[f"{tag}_y"] = np_(y)
        out[f"{tag}_gout"] = np_(gout)
        out[f"{tag}_gx"] = np_(x.grad)
        out[f"{tag}_bpt"] = np.asarray(bpt)
    np.savez_compressed(f"{HERE}/mixout.npz", **out)
    print("mixout:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    assert os.path.isdir(REF), "make_golden_r2.py runs only where /root/reference is mounted"
    golden_v3e()
    golden_mixout()
