"""CPU oracle for the mixture-of-tokenizers byte-mix embedding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path
(`mixture-of-tokenizers_b200/`) may import this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs use it, and only as the checker / the CPU arm that is timed next to the
GPU number.

It is a plain restatement (numpy for the integer half, eager torch-CPU for the
floating-point half) of the reference's algorithm.  Every function cites the
reference lines it follows (paths relative to the reference checkout):

  * ttb build            modded-nanogpt/create_ttb.py:10-33
  * ttb -> table         scaled-pre-train/data_creation.py:43-58,
                         modded-nanogpt/runs/7_mot-in_toks-valemb.py:431-441
  * tokens_to_bytes      scaled-pre-train/data_creation.py:61-67
  * pull_from_left/right scaled-pre-train/data_creation.py:71-305
  * tokens_to_digits     mathblations/data.py:92-109
  * embedding + mixin    scaled-pre-train/train_gpt.py:172-186,327-379,430-443,
                         modded-nanogpt/runs/{7,71,72,73,74,711,71041,71042,71051,4_*}.py
                         (mixin_bytes + the three forward lines of GPT.forward),
                         mathblations/model.py:256-268,323-327,
                         inference/inference.py:267

Parity status
-------------
Integer half: PINNED by the reference's one checked-in table
(`modded-nanogpt/embeddings/ttb_8_left_pad.json`, committed here as
`tests/golden/ttb_8_left_pad.npz`) and by outputs of the reference's own
functions (`tests/golden/make_golden.py` imports them from /root/reference in
the build container and commits the vectors).
Floating half: the reference has no test that pins float results, so the pins
are outputs of the reference's own modules (AST-extracted and run eagerly on
CPU by `tests/golden/make_golden.py`), committed under `tests/golden/`.
"""
from __future__ import annotations

import dataclasses
import json
from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# constants (scaled-pre-train/train_gpt.py:149,677,828; runs/7:482-483,502-503)
# --------------------------------------------------------------------------
TOKEN_VOCAB = 50257
EOT_TOKEN = 50256
BYTE_VOCAB = 458
PAD_BYTE = 456
EOT_BYTE = 457
FP32_EPS = float(torch.finfo(torch.float32).eps)  # F.rms_norm default eps (fp32 math)


# ==========================================================================
# Integer half
# ==========================================================================
def create_ttb(decode: Callable[[int], str], byte_to_int: dict, n_tokens: int,
               bpt: int = 16, pad_position: str = "left") -> dict:
    """Restates modded-nanogpt/create_ttb.py:10-33 over an injected `decode`.

    The reference loops `range(encoding.max_token_value)` (:18) so the EOT id
    itself (50256) never gets a row; `n_tokens` plays that role here.  A token
    whose decoded string is "<|endoftext|>" maps to [endoftext]*bpt (:20-22).
    Characters (not UTF-8 bytes) are looked up in byte_to_int (:23), the FIRST
    bpt are kept (:24) and the row is padded left or right with "pad" (:25-28).
    """
    ttb = {}
    for index in range(n_tokens):
        token = decode(index)
        if token == "<|endoftext|>":
            ttb[index] = [byte_to_int["endoftext"]] * bpt
            continue
        b_seq = [byte_to_int[b] for b in token]
        b_seq = b_seq[:bpt]
        if pad_position == "left":
            b_seq = [byte_to_int["pad"]] * (bpt - len(b_seq)) + b_seq
        elif pad_position == "right":
            b_seq = b_seq + [byte_to_int["pad"]] * (bpt - len(b_seq))
        else:
            raise ValueError(f"Invalid pad_position: {pad_position}")
        ttb[index] = b_seq
    return ttb


def load_ttb_json(path: str) -> dict:
    """scaled-pre-train/data_creation.py:43-48."""
    with open(path, "r") as f:
        ttb = json.loads(f.read())
    return {int(k): [int(x) for x in v] for k, v in ttb.items()}


def ttb_dict_to_array(ttb: dict, vocab_size: int, bpt: int, eot_row: bool = True) -> np.ndarray:
    """Dense int16 [vocab_size, bpt] table.

    The reference copies the JSON rows into an nn.Embedding whose other rows
    keep their N(0,1) init (data_creation.py:51-58) -- i.e. garbage ids for any
    token without a row, in particular EOT 50256 for the checked-in table.  The
    scaled-pre-train pull functions only work if that row is [457]*bpt
    (create_ttb.py:20-22 shows the intent), so it is defined here explicitly;
    other missing rows are filled with PAD.
    """
    tab = np.full((vocab_size, bpt), PAD_BYTE, dtype=np.int16)
    for k, v in ttb.items():
        tab[k] = np.asarray(v, dtype=np.int16)
    if eot_row and EOT_TOKEN < vocab_size and EOT_TOKEN not in ttb:
        tab[EOT_TOKEN] = EOT_BYTE
    return tab


def ttb_repad(tab: np.ndarray, bpt_out: int, pad_position: str = "left",
              src_pad_position: str = "left") -> np.ndarray:
    """Re-derive a (bpt_out, pad_position) table from a table of known rows.

    Follows create_ttb.py:23-28 on the character lists recovered from the
    source table (its non-pad entries, in order).  Exact for every token whose
    string is fully contained in the source row; for tokens the source table
    already truncated (no pad left) the result is the truncation the source
    kept (the first chars), which is what create_ttb would also keep when
    bpt_out <= source bpt.
    """
    V, bpt_in = tab.shape
    out = np.full((V, bpt_out), PAD_BYTE, dtype=np.int16)
    for v in range(V):
        row = tab[v]
        if np.all(row == EOT_BYTE):
            out[v] = EOT_BYTE
            continue
        chars = row[row != PAD_BYTE][:bpt_out]
        n = len(chars)
        if pad_position == "left":
            out[v, bpt_out - n:] = chars
        elif pad_position == "right":
            out[v, :n] = chars
        else:
            raise ValueError(f"Invalid pad_position: {pad_position}")
    return out


def bf16_round_ids(ids: np.ndarray) -> np.ndarray:
    """The modded-nanogpt table quirk: the ttb table lives in an nn.Embedding
    cast to bf16 (runs/7:441) and is read back with `.to(int64)` (runs/7:446).
    bf16 keeps 8 significant bits, round-to-nearest-even: ids > 256 lose their
    last bit (457 -> 456, 455 -> 456, 257 -> 256, 259 -> 260 ...)."""
    t = torch.from_numpy(np.asarray(ids).astype(np.float32))
    return t.bfloat16().to(torch.int64).numpy().astype(np.asarray(ids).dtype)


def tokens_to_bytes(tokens: np.ndarray, table: np.ndarray) -> np.ndarray:
    """scaled-pre-train/data_creation.py:61-67: gather rows, int64, flatten to
    [B, T*bpt] for 2-D tokens or [1, T*bpt] for 1-D tokens."""
    tokens = np.asarray(tokens)
    rows = table[tokens.astype(np.int64)].astype(np.int64)
    if tokens.ndim == 2:
        return rows.reshape(tokens.shape[0], -1)
    return rows.reshape(1, -1)


def _pull_rows(byte_tensor: np.ndarray, bpt: int, pad_byte: int, eot_byte: int, left: bool) -> np.ndarray:
    byte_tensor = np.asarray(byte_tensor)
    B, T = byte_tensor.shape
    if T == 0:
        return byte_tensor.copy()
    assert T % bpt == 0, "T must be divisible by bytes_per_token"
    Tr = T // bpt
    view = byte_tensor.reshape(B, Tr, bpt)
    out = np.full_like(view, pad_byte)
    for b in range(B):
        is_eot = np.all(view[b] == eot_byte, axis=1)
        if left:
            # data_creation.py:179-305: bytes of every token after the previous EOT
            # up to and including t; keep the LAST min(bpt, n); right-align.
            pool: list = []
            for t in range(Tr):
                if is_eot[t]:
                    out[b, t] = view[b, t]
                    pool = []
                    continue
                row = view[b, t]
                pool.extend(row[row != pad_byte].tolist())
                pool = pool[-bpt:]
                n = len(pool)
                if n:
                    out[b, t, bpt - n:] = pool
        else:
            # data_creation.py:71-176: bytes of tokens t, t+1, ... up to (not incl.)
            # the next EOT / row end; keep the FIRST min(bpt, n); left-align.
            pool = []
            for t in range(Tr - 1, -1, -1):
                if is_eot[t]:
                    out[b, t] = view[b, t]
                    pool = []
                    continue
                row = view[b, t]
                pool = (row[row != pad_byte].tolist() + pool)[:bpt]
                n = len(pool)
                if n:
                    out[b, t, :n] = pool
    return out.reshape(B, T)


def pull_from_left(byte_tensor, bytes_per_token: int, pad_byte: int = PAD_BYTE, eot_byte: int = EOT_BYTE):
    """scaled-pre-train/data_creation.py:179-305 (== runs/7:351-428)."""
    return _pull_rows(byte_tensor, bytes_per_token, pad_byte, eot_byte, left=True)


def pull_from_right(byte_tensor, bytes_per_token: int, pad_byte: int = PAD_BYTE, eot_byte: int = EOT_BYTE):
    """scaled-pre-train/data_creation.py:71-176."""
    return _pull_rows(byte_tensor, bytes_per_token, pad_byte, eot_byte, left=False)


def scramble_view(flat_bytes: np.ndarray, bpt: int) -> np.ndarray:
    """runs/71:479: `.view(bpt, -1)` of the token-major (1, T*bpt) byte tensor.
    Row i, column s holds flat byte i*T + s."""
    flat = np.asarray(flat_bytes).reshape(-1)
    return flat.reshape(bpt, -1)


def tokens_to_digits(tokens: Sequence[int], max_digits_per_token: int,
                     op_token: int, eq_token: int, pad_token: int) -> np.ndarray:
    """mathblations/data.py:92-109: right-aligned decimal digits, 13 = pad,
    10 = op, 11 = eq, 12 = the pad token."""
    digits = []
    for token in [int(t) for t in tokens]:
        new_toks = [13] * max_digits_per_token
        if token == op_token:
            new_toks[-1] = 10
        elif token == eq_token:
            new_toks[-1] = 11
        elif token == pad_token:
            new_toks[-1] = 12
        else:
            for i, ch in enumerate(reversed(str(token))):
                new_toks[-i - 1] = int(ch)
        digits.extend(new_toks)
    return np.asarray(digits, dtype=np.int64)


# ==========================================================================
# Floating-point half
# ==========================================================================
def rms_norm(x: torch.Tensor, eps: Optional[float] = None) -> torch.Tensor:
    """`norm` of train_gpt.py:172-173 / runs/7:132-133: F.rms_norm over the last
    dim, no weight.  eps=None is torch's default (finfo(float32).eps with fp32
    math for both bf16 and fp32 inputs in torch 2.11); pass it explicitly when
    running the oracle in float64."""
    if eps is None:
        return F.rms_norm(x, (x.size(-1),))
    return F.rms_norm(x, (x.size(-1),), eps=eps)


@dataclasses.dataclass(frozen=True)
class MixSpec:
    """One member of the mixin-variant catalogue.

    combine:   "add" | "concat" | "tok_only" | "bytes_only" | "mean"
    tok_norm:  rms_norm the gathered token row before mixing
    byte_norm: rms_norm every gathered byte row (over byte_dim) before mixing
    out_norm:  rms_norm the mixed row
    proj:      None | "linear" -> F.linear(mix, W[, bias]) applied before out_norm
    bytes_first: concat order [bytes | tok] (mathblations) instead of [tok | bytes]
    byte_fc:   V3f: project the concatenated bytes with W before the add
    lam_over_sum: V3e: both lambdas are divided by their sum (runs/71042:311-313)
    """
    combine: str = "add"
    tok_norm: bool = False
    byte_norm: bool = False
    out_norm: bool = True
    proj: Optional[str] = None
    bytes_first: bool = False
    byte_fc: bool = False
    lam_over_sum: bool = False


# name -> (spec, reference citation)
VARIANTS = {
    "V0": (MixSpec(combine="tok_only"), "spt/train_gpt.py:342-348"),
    "V1": (MixSpec(combine="concat", tok_norm=True, byte_norm=True, proj="linear"), "runs/7:226-234,317-319; spt/train_gpt.py:361-369,439-443"),
    "V2": (MixSpec(combine="concat", proj="linear"), "runs/72:227-230,313-315"),
    "V3": (MixSpec(combine="add"), "runs/71:228-230,312-314"),
    "V3b": (MixSpec(combine="add", tok_norm=True, byte_norm=True, out_norm=False), "runs/73:229-231,313-315"),
    "V3c": (MixSpec(combine="add", tok_norm=True, byte_norm=True, out_norm=False), "runs/74:314-316 (+lambdas)"),
    "V3d": (MixSpec(combine="add", tok_norm=True, byte_norm=True, out_norm=True), "runs/71041:226-228,311-313 (+lambdas)"),
    "V3e": (MixSpec(combine="add", tok_norm=True, byte_norm=True, out_norm=True, lam_over_sum=True), "runs/71042:225-228,311-314 (+lambdas / their sum)"),
    "V3f": (MixSpec(combine="add", byte_fc=True), "runs/71051:226-229,312-314"),
    "V4": (MixSpec(combine="concat"), "runs/711:224-232,314-316"),
    "V5": (MixSpec(combine="bytes_only"), "runs/4_bytes-in_toks-valemb.py:226-232,313"),
    "V7": (MixSpec(combine="mean", out_norm=False), "inference/inference.py:267"),
    "V8": (MixSpec(combine="concat", out_norm=False, proj="linear", bytes_first=True), "mathblations/model.py:261-268"),
}


def gather_byte_rows(E_byte: torch.Tensor, byte_ids: torch.Tensor, n_pos: int, bpt: int,
                     slot_major: bool) -> torch.Tensor:
    """Return [n_pos, bpt, bd]: the byte-embedding row of every (position, slot).

    slot_major=False: byte_ids is token-major, flat index s*bpt + i
      (runs/7:229-231 `view(B,S,bpt,D)`; train_gpt.py:442 "B (S bpt) D -> B S (bpt D)").
    slot_major=True: byte_ids is the `[bpt, T]` tensor of the sum runs
      (runs/71:313 + :229 `torch.cat([b for b in byte_embs], dim=-1)`), flat index i*T + s.
    """
    ids = byte_ids.reshape(-1).long()
    if slot_major:
        ids = ids.view(bpt, n_pos).t()
    else:
        ids = ids.view(n_pos, bpt)
    return F.embedding(ids, E_byte)


def mot_embed_forward(spec: MixSpec, tokens: torch.Tensor, byte_ids: Optional[torch.Tensor],
                      E_tok: Optional[torch.Tensor], E_byte: Optional[torch.Tensor], *,
                      bpt: int = 16, slot_major: bool = False,
                      W: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                      lam_tok: Optional[torch.Tensor] = None, lam_byte: Optional[torch.Tensor] = None,
                      byte_ids2: Optional[torch.Tensor] = None,
                      eps: Optional[float] = None) -> torch.Tensor:
    """Forward of the whole family, fp32 (or fp64) math on whatever dtype the
    parameters were rounded to; returns [n_pos, out_dim] in the math dtype.

    byte_ids2: second id tensor for the `--add-padded-and-pulled` variant
    (train_gpt.py:371-379): rows are summed BEFORE the byte norm.
    lam_tok / lam_byte: the learnable scalars of runs/74:314-315 (applied after
    the per-input norms, before mixing).
    """
    tokens = tokens.reshape(-1).long()
    n_pos = tokens.numel()
    if spec.lam_over_sum:  # runs/71042:311: norm_scalrs_sum = scalars[-1] + scalars[-2]
        lam_sum = lam_tok + lam_byte
        lam_tok, lam_byte = lam_tok / lam_sum, lam_byte / lam_sum
    t = None
    if spec.combine != "bytes_only":
        t = F.embedding(tokens, E_tok)
        if spec.tok_norm:
            t = rms_norm(t, eps)
        if lam_tok is not None:
            t = t * lam_tok
    b = None
    if spec.combine != "tok_only":
        b = gather_byte_rows(E_byte, byte_ids, n_pos, bpt, slot_major)
        if byte_ids2 is not None:
            b = b + gather_byte_rows(E_byte, byte_ids2, n_pos, bpt, slot_major)
        if spec.byte_norm:
            b = rms_norm(b, eps)
        if lam_byte is not None:
            b = b * lam_byte
    if spec.combine == "tok_only":
        x = t
    elif spec.combine == "bytes_only":
        x = b.reshape(n_pos, -1)
    elif spec.combine == "mean":
        x = t + b.mean(dim=-2)
    elif spec.combine == "add":
        c = b.reshape(n_pos, -1)
        if spec.byte_fc:
            c = F.linear(c, W)
        x = t + c
    elif spec.combine == "concat":
        c = b.reshape(n_pos, -1)
        x = torch.cat([c, t], dim=-1) if spec.bytes_first else torch.cat([t, c], dim=-1)
    else:
        raise ValueError(spec.combine)
    if spec.proj == "linear":
        x = F.linear(x, W, bias)
    if spec.out_norm:
        x = rms_norm(x, eps)
    return x


def mot_embed_fwd_bwd(spec: MixSpec, tokens, byte_ids, E_tok, E_byte, grad_out, *,
                      math_dtype=torch.float32, **kw):
    """Forward + backward through torch autograd in `math_dtype`, parameters
    taken at whatever precision they were stored in (bf16 tables are upcast, so
    this is "fp32 math on bf16-rounded parameters": duplicates accumulate in
    fp32 and nothing is rounded until the caller casts).  Returns
    (out, dict of dense grads) all in math_dtype.
    """
    params = {}

    def leaf(x):
        if x is None:
            return None
        return x.detach().to(math_dtype).clone().requires_grad_(True)

    params["E_tok"] = leaf(E_tok)
    params["E_byte"] = leaf(E_byte)
    for name in ("W", "bias", "lam_tok", "lam_byte"):
        params[name] = leaf(kw.pop(name, None))
    eps = kw.pop("eps", None)
    if eps is None:
        eps = FP32_EPS
    out = mot_embed_forward(spec, tokens, byte_ids, params["E_tok"], params["E_byte"],
                            W=params["W"], bias=params["bias"], lam_tok=params["lam_tok"],
                            lam_byte=params["lam_byte"], eps=eps, **kw)
    out.backward(grad_out.detach().to(math_dtype).reshape(out.shape))
    grads = {k: (v.grad if v is not None else None) for k, v in params.items()}
    return out.detach(), grads


def value_embeds_fwd_bwd(tokens, tables, grad_outs, math_dtype=torch.float32):
    """The value embeddings gathered with the token ids (runs/7:252,308: `ve = [value_embed(token_inputs) for
    value_embed in self.value_embeds]`; spt/train_gpt.py:566,600) and their dense gradients
    (embedding_dense_backward: duplicates accumulate, rows never gathered are zero).  Returns (outs, grads)."""
    tokens = tokens.reshape(-1).long()
    leaves = [E.detach().to(math_dtype).clone().requires_grad_(True) for E in tables]
    outs = [F.embedding(tokens, E) for E in leaves]
    for o, g in zip(outs, grad_outs):
        o.backward(g.detach().to(math_dtype).reshape(o.shape))
    return [o.detach() for o in outs], [E.grad for E in leaves]


def split_residual_fwd_bwd(tokens, byte_ids, E_tok, E_byte, lam_tok, lam_byte, grads, *, bpt=16,
                           math_dtype=torch.float32, eps=None):
    """runs/71081:302-304,315 (V3g): x0t = norm(embed_tokens(tok)); x0b = cat of the per-byte-normed rows of the
    `[bpt, T]` id tensor; x = x0t * scalars[-1] + x0b * scalars[-2].  `grads` = upstream gradients of (x, x0t, x0b).
    Returns ((x, x0t, x0b), dict of dense grads)."""
    if eps is None:
        eps = FP32_EPS
    leaf = lambda t: t.detach().to(math_dtype).clone().requires_grad_(True)  # noqa: E731
    Et, Eb, lt, lb = leaf(E_tok), leaf(E_byte), leaf(lam_tok), leaf(lam_byte)
    tokens = tokens.reshape(-1).long()
    n = tokens.numel()
    x0t = rms_norm(F.embedding(tokens, Et), eps)
    x0b = rms_norm(gather_byte_rows(Eb, byte_ids, n, bpt, True), eps).reshape(n, -1)
    x = x0t * lt + x0b * lb
    torch.autograd.backward([x, x0t, x0b], [g.detach().to(math_dtype).reshape(n, -1) for g in grads])
    return (x.detach(), x0t.detach(), x0b.detach()), {"E_tok": Et.grad, "E_byte": Eb.grad, "lam_tok": lt.grad, "lam_byte": lb.grad}
