"""The ttb ("token to bytes") table layer: JSON / strings -> the int16 `[V, bpt]` device table the kernels read.

Reference: `create_ttb` (modded-nanogpt/create_ttb.py:10-33) builds `{token_id: [bpt character ids]}` offline and
writes it as JSON; `load_ttb` / `make_embedding` (scaled-pre-train/data_creation.py:43-58,
runs/7:431-441) load that JSON into a float `nn.Embedding(vocab, bpt)`.

Here the file format and the string handling stay on the host (JSON parsing, `decode(i)`, dictionary look-ups are
not device work); everything that produces or reshapes table ROWS is a kernel of libmot_b200 (`mot_ttb_build`,
`mot_ttb_repad`, `mot_ttb_expand`) and therefore needs a CUDA device: there is no CPU table builder.

Two things the reference leaves implicit are explicit here:
  * the end-of-text row.  `create_ttb` loops `range(encoding.max_token_value)` (:18), so token 50256 itself never
    gets a row and keeps `nn.Embedding`'s N(0,1) init in `make_embedding` (data_creation.py:53-56); `pull_from_*`
    only detect a document boundary when that row is `[457]*bpt` (:94,200; create_ttb.py:20-22 shows the intent).
    `table_from_rows(..., eot_token=50256)` / `with_eot_row` define it.
  * the container dtype.  The runs keep the table in a bf16 `nn.Embedding` (runs/7:441), which rounds 101 of the 458
    ids; `as_container(table, torch.bfloat16)` reproduces that container bit for bit, `from_container` reads either
    container back with the reference's truncation (`.to(int64)`, data_creation.py:63).
"""
from __future__ import annotations

import json
from typing import Callable, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

TOKEN_VOCAB = 50257   # spt/train_gpt.py:828, runs/7:502
EOT_TOKEN = 50256     # runs/7:267
PAD_BYTE = 456        # byte_to_int.json["pad"]
EOT_BYTE = 457        # byte_to_int.json["endoftext"]
EOT_STRING = "<|endoftext|>"   # create_ttb.py:20


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _need_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("mot_b200.ttb: table rows are built by CUDA kernels (no CPU fallback); pass a cuda device")
    return dev


# ---------------------------------------------------------------------------------------------------------------
# host side: file formats and strings
# ---------------------------------------------------------------------------------------------------------------
def load_json(path: str) -> Dict[int, List[int]]:
    """`load_ttb` (spt/data_creation.py:43-48): the JSON object `{"<token id>": [bpt ids]}` with integer keys."""
    with open(path, "r") as f:
        raw = json.load(f)
    return {int(k): [int(x) for x in row] for k, row in raw.items()}


def load_byte_to_int(path: str) -> Dict[str, int]:
    """`embeddings/byte_to_int.json` (create_ttb.py:12-14): character -> id, plus "pad" and "endoftext"."""
    with open(path, "r") as f:
        raw = json.load(f)
    return {ch: int(i) for ch, i in raw.items()}


def save_json(table: torch.Tensor, path: str, n_rows: Optional[int] = None) -> None:
    """Write a table in the reference's file format (create_ttb.py:32): keys are decimal token ids.  `n_rows` limits
    the rows written (the checked-in files stop before the end-of-text token, create_ttb.py:18)."""
    rows = table.detach().to("cpu", torch.int64).tolist()
    if n_rows is not None:
        rows = rows[:n_rows]
    with open(path, "w") as f:
        f.write(json.dumps({str(i): r for i, r in enumerate(rows)}))


def pack_rows(rows: Mapping[int, Sequence[int]], vocab_size: int, bpt: Optional[int] = None,
              pad_byte: int = PAD_BYTE) -> Tuple[np.ndarray, np.ndarray]:
    """Dense host image of a loaded JSON table: (`ids` int16 [vocab_size, bpt], `present` bool [vocab_size]).  Rows
    absent from the file are filled with `pad_byte` here (in the reference they keep a random float row,
    data_creation.py:53-56) and flagged in `present` so that the caller decides (see `with_eot_row`)."""
    if bpt is None:
        bpt = len(next(iter(rows.values())))
    ids = np.full((vocab_size, bpt), pad_byte, dtype=np.int16)
    present = np.zeros(vocab_size, dtype=bool)
    keys = np.fromiter(rows.keys(), dtype=np.int64, count=len(rows))
    if keys.size:
        if keys.min() < 0 or keys.max() >= vocab_size:
            raise ValueError("mot_b200.ttb: token id outside [0, vocab_size)")
        vals = np.asarray([rows[int(k)] for k in keys], dtype=np.int64)
        if vals.ndim != 2 or vals.shape[1] != bpt:
            raise ValueError(f"mot_b200.ttb: every row must hold {bpt} ids")
        if vals.min() < -32768 or vals.max() > 32767:
            raise ValueError("mot_b200.ttb: byte id does not fit int16")
        ids[keys] = vals.astype(np.int16)
        present[keys] = True
    return ids, present


def strings_to_chars(decode: Callable[[int], str], byte_to_int: Mapping[str, int],
                     n_tokens: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The host half of `create_ttb` (create_ttb.py:18-23): decode every token id, map each CHARACTER of the string
    (not UTF-8 byte: partial byte sequences decode to U+FFFD) through `byte_to_int`.  Returns the flat id stream
    `chars` int16, `offs` int32 [n_tokens + 1] and `is_eot` uint8 [n_tokens] (string == "<|endoftext|>", :20).
    Truncation to bpt and padding (:24-28) happen on the device (`build_table`)."""
    offs = np.zeros(n_tokens + 1, dtype=np.int32)
    is_eot = np.zeros(n_tokens, dtype=np.uint8)
    pieces: List[List[int]] = []
    total = 0
    for index in range(n_tokens):
        s = decode(index)
        if s == EOT_STRING:
            is_eot[index] = 1
        else:
            ids = [byte_to_int[ch] for ch in s]     # KeyError for an unknown character, like the reference
            pieces.append(ids)
            total += len(ids)
        offs[index + 1] = total
    chars = np.fromiter((i for p in pieces for i in p), dtype=np.int16, count=total)
    return chars, offs, is_eot


# ---------------------------------------------------------------------------------------------------------------
# device side: rows
# ---------------------------------------------------------------------------------------------------------------
def build_table(chars: np.ndarray, offs: np.ndarray, is_eot: Optional[np.ndarray], bpt: int = 16,
                pad_position: str = "left", *, device="cuda", pad_byte: int = PAD_BYTE,
                eot_byte: int = EOT_BYTE) -> torch.Tensor:
    """create_ttb.py:24-28 on the device (`mot_ttb_build`): keep the first bpt characters, pad left / right."""
    if pad_position not in ("left", "right"):
        raise ValueError(f"Invalid pad_position: {pad_position}")      # create_ttb.py:29-30
    dev = _need_cuda(device)
    n_rows = int(offs.shape[0]) - 1
    d_chars = torch.from_numpy(np.ascontiguousarray(chars, dtype=np.int16)).to(dev)
    d_offs = torch.from_numpy(np.ascontiguousarray(offs, dtype=np.int32)).to(dev)
    d_eot = torch.from_numpy(np.ascontiguousarray(is_eot, dtype=np.uint8)).to(dev) if is_eot is not None else None
    out = torch.empty((n_rows, bpt), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        rc = L.lib().mot_ttb_build(d_chars.data_ptr() if d_chars.numel() else None, d_offs.data_ptr(),
                                   d_eot.data_ptr() if d_eot is not None else None, n_rows, bpt,
                                   1 if pad_position == "left" else 0, pad_byte, eot_byte, out.data_ptr(), _stream(dev))
    L.check(rc, "mot_ttb_build")
    return out


def create_ttb(decode: Callable[[int], str], byte_to_int: Mapping[str, int], n_tokens: int, bpt: int = 16,
               pad_position: str = "left", *, device="cuda") -> torch.Tensor:
    """`create_ttb(bpt, pad_position)` (modded-nanogpt/create_ttb.py:10-33) over an injected `decode` (the reference
    calls tiktoken, which downloads the GPT-2 vocabulary at run time): int16 [n_tokens, bpt] on `device`."""
    chars, offs, is_eot = strings_to_chars(decode, byte_to_int, n_tokens)
    return build_table(chars, offs, is_eot, bpt, pad_position, device=device, pad_byte=byte_to_int["pad"],
                       eot_byte=byte_to_int["endoftext"])


def with_eot_row(table: torch.Tensor, vocab_size: int = TOKEN_VOCAB, eot_token: int = EOT_TOKEN,
                 eot_byte: int = EOT_BYTE, pad_byte: int = PAD_BYTE) -> torch.Tensor:
    """Grow `table` to `vocab_size` rows (new rows = pad) and set row `eot_token` to `[eot_byte]*bpt`
    (create_ttb.py:20-22; the row `pull_from_*` need, data_creation.py:94,200)."""
    rows, bpt = table.shape
    out = table
    if rows < vocab_size:
        out = torch.full((vocab_size, bpt), pad_byte, dtype=table.dtype, device=table.device)
        out[:rows] = table
    elif eot_token is not None:
        out = table.clone()
    if eot_token is not None:
        if not 0 <= eot_token < out.shape[0]:
            raise ValueError("mot_b200.ttb: eot_token outside the table")
        out[eot_token] = eot_byte
    return out


def table_from_rows(rows: Mapping[int, Sequence[int]], vocab_size: int = TOKEN_VOCAB, *, device="cuda",
                    eot_token: Optional[int] = EOT_TOKEN) -> torch.Tensor:
    """`make_embedding` (spt/data_creation.py:51-58) as an int16 device table: the JSON rows, plus the explicit
    end-of-text row when the file has none."""
    dev = _need_cuda(device)
    ids, present = pack_rows(rows, vocab_size)
    table = torch.from_numpy(ids).to(dev)
    if eot_token is not None and not present[eot_token]:
        table[eot_token] = EOT_BYTE
    return table


def from_json(path: str, vocab_size: int = TOKEN_VOCAB, *, device="cuda", eot_token: Optional[int] = EOT_TOKEN) -> torch.Tensor:
    """`make_embedding(filename, vocab_size)`: file -> int16 [vocab_size, bpt] device table."""
    return table_from_rows(load_json(path), vocab_size, device=device, eot_token=eot_token)


def repad(table: torch.Tensor, bpt_out: int, pad_position: str = "left", *, pad_byte: int = PAD_BYTE,
          eot_byte: int = EOT_BYTE) -> torch.Tensor:
    """The same strings under another (bpt, pad side) (`mot_ttb_repad`): what `create_ttb(bpt_out, pad_position)` gives
    for every token whose string the source row holds completely (source rows without a pad may already be truncated;
    they keep their first characters, which is also what create_ttb keeps when bpt_out <= source bpt)."""
    if pad_position not in ("left", "right"):
        raise ValueError(f"Invalid pad_position: {pad_position}")
    dev = _need_cuda(table.device)
    if table.dtype != torch.int16 or table.dim() != 2:
        raise NotImplementedError("mot_b200.ttb.repad: int16 [V, bpt] table expected (see from_container)")
    src = table.contiguous()
    out = torch.empty((src.shape[0], bpt_out), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        rc = L.lib().mot_ttb_repad(src.data_ptr(), src.shape[0], src.shape[1], bpt_out, 1 if pad_position == "left" else 0,
                                   pad_byte, eot_byte, out.data_ptr(), _stream(dev))
    L.check(rc, "mot_ttb_repad")
    return out


def as_container(table: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """The reference's float containers of the table: fp32 (`make_embedding`, spt/train_gpt.py:666) or the bf16 cast of
    the runs (`.cuda().bfloat16()`, runs/7:441), which rounds every id above 256 to 8 significant bits."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise NotImplementedError("mot_b200.ttb.as_container: float32 or bfloat16")
    return table.to(torch.float32).to(dtype)


def from_container(weight: torch.Tensor) -> torch.Tensor:
    """int16 table holding exactly the ids `tokens_to_bytes` reads out of a float container (`emb(tokens).to(int64)`,
    data_creation.py:63): one `mot_ttb_expand` over every row."""
    from .ops import ttb_expand
    dev = _need_cuda(weight.device)
    V, bpt = weight.shape
    ids = ttb_expand(torch.arange(V, dtype=torch.int32, device=dev), weight.detach(), out_dtype=torch.int32)
    return ids.view(V, bpt).to(torch.int16)
