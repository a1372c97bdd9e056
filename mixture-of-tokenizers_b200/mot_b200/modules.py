"""Drop-in nn.Modules with the reference's parameter names and forward signatures,
backed by the fused CUDA path (no PyTorch fallback).

Reference call sites:
  * modded-nanogpt runs:  `x_toks = embed_tokens(token_inputs)[None]; x_bytes = embed_bytes(byte_inputs)...;
    x = mixin_bytes(x_toks, x_bytes[, W])`  (runs/71:312-314, runs/7:317-319, ...)
  * scaled-pre-train:     `xt, xb = self.embed(tokens, bytes_padded, bytes_pulled); x = self.byte_mixin(xt, xb)`
    (spt/train_gpt.py:605-606)
  * mathblations:         `x = self.digit_mixin(self.wte(idx), self.dte(digits))` (mathblations/model.py:323-327)
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .ops import MixSpec, mot_embed

# variant name -> MixSpec kwargs (`slot_major` is the `.view(bpt,-1)` id layout of the sum runs)
RUN_VARIANTS = {
    "V0": dict(combine="tok_only"),
    "V3": dict(combine="add", slot_major=True),                                                  # runs/71
    "V3b": dict(combine="add", tok_norm=True, byte_norm=True, out_norm=False, slot_major=True),  # runs/73
    "V3c": dict(combine="add", tok_norm=True, byte_norm=True, out_norm=False, slot_major=True),  # runs/74 (+lambdas)
    "V3d": dict(combine="add", tok_norm=True, byte_norm=True, out_norm=True, slot_major=True),   # runs/71041 (+lambdas)
    "V4": dict(combine="concat"),                                                                # runs/711
    "V5": dict(combine="bytes_only"),                                                            # runs/4
}
_LAMBDA_VARIANTS = ("V3c", "V3d")


class MoTEmbedding(nn.Module):
    """The embedding front of the modded-nanogpt MoT runs: parameters `embed_tokens.weight`,
    `embed_bytes.weight` (and `lambdas` = the two trailing entries of the reference's `scalars`,
    [byte, token] order as in runs/74:259,314-315); forward(token_inputs [T] int32,
    byte_inputs int32 [bpt, T] or [1, T*bpt]) -> [1, T, model_dim].  With byte_inputs=None and a ttb
    table the byte ids are derived inside the kernel (no-pull path)."""

    def __init__(self, token_vocab_size: int, byte_vocab_size: int, token_dim: int, byte_dim: int,
                 bytes_per_token: int = 16, variant: str = "V3", ttb: Optional[torch.Tensor] = None):
        super().__init__()
        if variant not in RUN_VARIANTS:
            raise NotImplementedError(f"mot_b200: variant {variant!r} has no fused kernel")
        self.variant, self.bpt = variant, bytes_per_token
        self.spec = MixSpec(**RUN_VARIANTS[variant])
        self.embed_tokens = nn.Embedding(token_vocab_size, token_dim) if variant != "V5" else None
        self.embed_bytes = nn.Embedding(byte_vocab_size, byte_dim) if variant != "V0" else None
        self.lambdas = nn.Parameter(torch.tensor([0.5, 0.5])) if variant in _LAMBDA_VARIANTS else None
        self.register_buffer("ttb", ttb, persistent=False)

    def forward(self, token_inputs: torch.Tensor, byte_inputs: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert token_inputs.ndim == 1  # runs/71:300
        lam = None
        if self.lambdas is not None:  # kernel order (tok, byte); parameter order (byte, tok) like scalars[-2], scalars[-1]
            lam = self.lambdas.flip(0)
        spec = self.spec
        if byte_inputs is None and self.embed_bytes is not None:
            spec = MixSpec(**{**RUN_VARIANTS[self.variant], "slot_major": False,
                              "ttb_scramble": RUN_VARIANTS[self.variant].get("slot_major", False)})
        x = mot_embed(token_inputs if self.embed_tokens is not None else None, byte_inputs,
                      self.embed_tokens.weight if self.embed_tokens is not None else None,
                      self.embed_bytes.weight if self.embed_bytes is not None else None,
                      spec, bpt=self.bpt, lam=lam, ttb=self.ttb if byte_inputs is None else None,
                      seq_len=token_inputs.numel())
        return x[None]
