"""Functional API over libmot_b200.so: ttb expansion and the fused byte-mix
embedding with autograd.  Tensors must live on a CUDA (B200) device; CPU tensors
raise -- there is no fallback path."""
from __future__ import annotations

import dataclasses
from typing import Optional

import torch

from . import _lib as L

FP32_EPS = float(torch.finfo(torch.float32).eps)  # F.rms_norm's default eps (train_gpt.py:172-173)

_COMBINE = {"add": L.ADD, "concat": L.CONCAT, "tok_only": L.TOK_ONLY, "bytes_only": L.BYTES_ONLY, "mean": L.MEAN}
_DTYPE = {torch.bfloat16: L.BF16, torch.float32: L.F32}
_TTB_DTYPE = {torch.int16: L.TTB_I16, torch.float32: L.TTB_F32, torch.bfloat16: L.TTB_BF16}


@dataclasses.dataclass(frozen=True)
class MixSpec:
    """Which member of the mixin catalogue the fused kernel computes."""
    combine: str = "add"          # add | concat | tok_only | bytes_only | mean
    tok_norm: bool = False        # rms_norm(E_tok[tok])            runs/7:317
    byte_norm: bool = False       # rms_norm(E_byte[id]) per byte   runs/7:318
    out_norm: bool = True         # rms_norm(mixed row)             runs/71:230
    bytes_first: bool = False     # concat order [bytes | tok]      mathblations/model.py:267
    slot_major: bool = False      # byte ids are [bpt, T]           runs/71:479
    ttb_scramble: bool = False    # ids from ttb with the `.view(bpt,-1)` index map
    eps: float = FP32_EPS


def _require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mot_b200: tensors must be on a CUDA device (no CPU fallback)")
        if dev is not None and t.device != dev:
            raise RuntimeError("mot_b200: tensors are on different devices")
        dev = t.device
    return dev


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_ABSENT = torch.empty(0)   # placeholder for "no tensor" in save_for_backward (one object, not one allocation per call)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)   # the cudaStream_t itself, no Stream object built


def _stream(dev) -> int:
    """cudaStream_t of torch's current stream on `dev` (called several times per step: torch.cuda.current_stream()
    costs ~15 us of host time per call, the raw getter well under one)."""
    if _raw_stream is not None and dev.index is not None:
        return _raw_stream(dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the common case costs one integer compare)."""
    __slots__ = ("dev", "prev")

    def __init__(self, dev):
        self.dev = dev

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.dev.index is not None and self.dev.index != self.prev:
            torch.cuda.set_device(self.dev)
        else:
            self.prev = -1

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _ttb_expand_impl(tok: torch.Tensor, ttb: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    dev = tok.device
    V, bpt = ttb.shape
    n = tok.numel()
    out = torch.empty((n, bpt), dtype=out_dtype, device=dev)
    with _on_device(dev):
        rc = L.lib().mot_ttb_expand(_ptr(tok), n, _ptr(ttb), V, bpt, _TTB_DTYPE[ttb.dtype], _ptr(out),
                                    1 if out_dtype == torch.int64 else 0, _stream(dev))
    L.check(rc, "mot_ttb_expand")
    return out


def ttb_expand(tokens: torch.Tensor, ttb: torch.Tensor, out_dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """tokens_to_bytes (spt/data_creation.py:61-67): [B,T] -> [B, T*bpt], [T] -> [1, T*bpt].

    `ttb` is the [V, bpt] table: int16 (native) or the reference's float containers (fp32
    nn.Embedding weight, or the bf16-cast weight of the runs with its id-rounding quirk)."""
    _require_cuda(tokens, ttb)
    if ttb.dtype not in _TTB_DTYPE or ttb.dim() != 2:
        raise NotImplementedError(f"mot_b200.ttb_expand: ttb must be a 2-D int16/float32/bfloat16 table, got {ttb.dtype}")
    if out_dtype not in (torch.int64, torch.int32):
        raise NotImplementedError("mot_b200.ttb_expand: out_dtype must be int64 or int32")
    tok = tokens.to(torch.int32).contiguous().reshape(-1)
    ttb = ttb.contiguous()
    if _custom_ops_wanted():
        out = torch.ops.mot_b200.ttb_expand(tok, ttb, out_dtype == torch.int64)
    else:
        out = _ttb_expand_impl(tok, ttb, out_dtype)
    if tokens.dim() == 2:
        return out.view(tokens.shape[0], -1)
    return out.view(1, -1)


def _tokens_to_digits_impl(t: torch.Tensor, dpt: int, op_token: int, eq_token: int, pad_token: int,
                           out_dtype: torch.dtype) -> torch.Tensor:
    dev = t.device
    out = torch.empty(t.numel() * dpt, dtype=out_dtype, device=dev)
    with _on_device(dev):
        rc = L.lib().mot_tokens_to_digits(_ptr(t), t.numel(), 1 if t.dtype == torch.int64 else 0, dpt,
                                          op_token, eq_token, pad_token, _ptr(out), 1 if out_dtype == torch.int64 else 0,
                                          _stream(dev))
    L.check(rc, "mot_tokens_to_digits")
    return out


def tokens_to_digits(tokens: torch.Tensor, max_digits_per_token: int, op_token: int, eq_token: int, pad_token: int,
                     out_dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """GenerateEquations.tokens_to_digits (mathblations/data.py:92-109) on the device: [n] -> [n * dpt]."""
    _require_cuda(tokens)
    if tokens.dtype not in (torch.int32, torch.int64) or out_dtype not in (torch.int32, torch.int64):
        raise NotImplementedError("mot_b200.tokens_to_digits: int32 / int64 only")
    t = tokens.reshape(-1).contiguous()
    if _custom_ops_wanted():
        return torch.ops.mot_b200.tokens_to_digits(t, max_digits_per_token, op_token, eq_token, pad_token,
                                                   out_dtype == torch.int64)
    return _tokens_to_digits_impl(t, max_digits_per_token, op_token, eq_token, pad_token, out_dtype)


def _pull_impl(x: torch.Tensor, bytes_per_token: int, pad_byte: int, eot_byte: int, from_right: bool) -> torch.Tensor:
    dev = x.device
    B, TB = x.shape
    out = torch.empty_like(x)
    T = TB // bytes_per_token
    if B * T == 0:
        return out
    ws = torch.empty(int(L.lib().mot_pull_workspace_bytes(B, T, bytes_per_token)), dtype=torch.uint8, device=dev)
    with _on_device(dev):
        rc = L.lib().mot_pull(_ptr(x), _ptr(out), B, T, bytes_per_token, 1 if x.dtype == torch.int64 else 0, pad_byte,
                              eot_byte, 1 if from_right else 0, _ptr(ws), ws.numel(), _stream(dev))
    L.check(rc, "mot_pull")
    return out


def _pull(byte_tensor: torch.Tensor, bytes_per_token: int, pad_byte: int, eot_byte: int, from_right: bool) -> torch.Tensor:
    _require_cuda(byte_tensor)
    if byte_tensor.dtype not in (torch.int32, torch.int64) or byte_tensor.dim() != 2:
        raise NotImplementedError("mot_b200.pull: byte_tensor must be a 2-D int32 / int64 tensor [B, T*bpt]")
    if byte_tensor.shape[1] % bytes_per_token:
        raise RuntimeError("T must be divisible by bytes_per_token")   # the reference's assert (data_creation.py:189)
    x = byte_tensor.contiguous()
    if _custom_ops_wanted():
        return torch.ops.mot_b200.pull(x, bytes_per_token, pad_byte, eot_byte, from_right)
    return _pull_impl(x, bytes_per_token, pad_byte, eot_byte, from_right)


def pull_from_left(byte_tensor: torch.Tensor, bytes_per_token: int, pad_byte: int = 456, eot_byte: int = 457) -> torch.Tensor:
    """spt/data_creation.py:179-305 (== runs/7:351-428), same signature: every non-EOT token receives the last
    bytes_per_token non-pad bytes of its segment up to and including itself, right-aligned.  No host syncs."""
    return _pull(byte_tensor, bytes_per_token, pad_byte, eot_byte, False)


def pull_from_right(byte_tensor: torch.Tensor, bytes_per_token: int, pad_byte: int = 456, eot_byte: int = 457) -> torch.Tensor:
    """spt/data_creation.py:71-176: the first bytes_per_token non-pad bytes of tokens t, t+1, ... before the next EOT."""
    return _pull(byte_tensor, bytes_per_token, pad_byte, eot_byte, True)


def make_desc(spec: MixSpec, n_tokens: int, E_tok, E_byte, bpt: int, *, ids: Optional[torch.Tensor],
              ttb: Optional[torch.Tensor], has_lam: bool, seq_len: int = 0, row_stride: int = 0,
              col_offset: int = 0, dp_slabs: int = 0) -> L.MotDesc:
    ref = E_tok if E_tok is not None else E_byte
    if ref.dtype not in _DTYPE:
        raise NotImplementedError(f"mot_b200: embedding dtype {ref.dtype} is not supported (bf16 / fp32 only)")
    if spec.combine not in _COMBINE:
        raise NotImplementedError(f"mot_b200: unknown combine mode {spec.combine!r}")
    has_tok, has_bytes = spec.combine != "bytes_only", spec.combine != "tok_only"
    Dt = E_tok.shape[1] if has_tok else 0
    bd = E_byte.shape[1] if has_bytes else 0
    if spec.combine in ("add", "tok_only", "mean"):
        Do = Dt
    elif spec.combine == "concat":
        Do = Dt + bpt * bd
    else:
        Do = bpt * bd
    flags = 0
    flags |= L.F_TOK_NORM if spec.tok_norm else 0
    flags |= L.F_BYTE_NORM if spec.byte_norm else 0
    flags |= L.F_OUT_NORM if spec.out_norm else 0
    flags |= L.F_BYTES_FIRST if spec.bytes_first else 0
    flags |= L.F_HAS_LAMBDAS if has_lam else 0
    ttb_dtype = 0
    if has_bytes:
        if ids is None:
            if ttb is None:
                raise RuntimeError("mot_b200: need byte ids or a ttb table")
            flags |= L.F_IDS_FROM_TTB
            flags |= L.F_TTB_SCRAMBLE if spec.ttb_scramble else 0
            ttb_dtype = _TTB_DTYPE[ttb.dtype]
        else:
            flags |= L.F_SLOT_MAJOR if spec.slot_major else 0
            flags |= L.F_IDS_I64 if ids.dtype == torch.int64 else 0
    return L.MotDesc(L.ABI_VERSION, _DTYPE[ref.dtype], n_tokens, seq_len,
                     E_tok.shape[0] if has_tok else 0, E_byte.shape[0] if has_bytes else 0, bpt if has_bytes else 0,
                     Dt, bd, Do, _COMBINE[spec.combine], flags, ttb_dtype, spec.eps, row_stride, col_offset, dp_slabs)


def embed_forward_out(desc: L.MotDesc, tok, ids, ttb, E_tok, E_byte, lam, out, stream: Optional[int] = None,
                      rstd: Optional[torch.Tensor] = None, addend: Optional[torch.Tensor] = None) -> None:
    """mot_embed_fwd on caller-allocated tensors (no allocation, no sync; CUDA-graph capturable).  `rstd` (fp32
    [n_tokens]) additionally keeps the reciprocal rms of every mixed row for the saved-output backward; `addend`
    ([n_tokens, out_dim]) is added to the mixed row before the output norm (mot_embed_fwd_ex)."""
    dev = out.device
    with _on_device(dev):
        if rstd is None and addend is None:
            rc = L.lib().mot_embed_fwd(desc, _ptr(tok), _ptr(ids), _ptr(ttb), _ptr(E_tok), _ptr(E_byte), _ptr(lam),
                                       _ptr(out), _stream(dev) if stream is None else stream)
        else:
            rc = L.lib().mot_embed_fwd_ex(desc, _ptr(tok), _ptr(ids), _ptr(ttb), _ptr(E_tok), _ptr(E_byte), _ptr(lam),
                                          _ptr(addend), _ptr(out), _ptr(rstd), _stream(dev) if stream is None else stream)
    L.check(rc, "mot_embed_fwd")


def embed_bwd_uses_saved(desc: L.MotDesc) -> bool:
    """Whether mot_embed_bwd_ex would read the kept forward result for this descriptor (MoT-sum, moderate N)."""
    return bool(L.lib().mot_embed_bwd_uses_saved(desc))


def embed_workspace_bytes(desc: L.MotDesc) -> int:
    return int(L.lib().mot_embed_workspace_bytes(desc))


def embed_workspace_init(desc: L.MotDesc, ws) -> None:
    """Zero the head of a fresh workspace once; afterwards every completed backward leaves it clean (MOT_WS_CLEAN)."""
    dev = ws.device
    with _on_device(dev):
        rc = L.lib().mot_embed_workspace_init(desc, _ptr(ws), ws.numel(), _stream(dev))
    L.check(rc, "mot_embed_workspace_init")


def embed_plan(desc: L.MotDesc, tok, ws, ws_clean: bool = False) -> None:
    dev = ws.device
    with _on_device(dev):
        rc = L.lib().mot_embed_plan(desc, _ptr(tok), _ptr(ws), ws.numel(), L.WS_CLEAN if ws_clean else 0, _stream(dev))
    L.check(rc, "mot_embed_plan")


def embed_backward_out(desc: L.MotDesc, tok, ids, ttb, E_tok, E_byte, lam, grad_out, gE_tok, gE_byte, g_lam, ws,
                       plan_ready: bool = False, ws_clean: bool = False, stream: Optional[int] = None,
                       out_saved: Optional[torch.Tensor] = None, rstd: Optional[torch.Tensor] = None,
                       addend: Optional[torch.Tensor] = None, d_addend: Optional[torch.Tensor] = None,
                       plan_joined: bool = False) -> None:
    """mot_embed_bwd on caller-allocated tensors; gE_tok / gE_byte are fully overwritten.  `ws_clean`: the caller
    vouches that the head of `ws` is zero (fresh from embed_workspace_init or left by a completed backward).
    `plan_joined`: the plan ran on the side stream and this stream waited for its event (MOT_WS_PLAN_JOINED).
    `out_saved` + `rstd` (what embed_forward_out(..., rstd=) produced) and `addend` / `d_addend`: mot_embed_bwd_ex."""
    dev = grad_out.device
    flags = (L.WS_PLAN_READY if plan_ready else 0) | (L.WS_CLEAN if ws_clean else 0) | \
        (L.WS_PLAN_JOINED if plan_ready and plan_joined else 0)
    with _on_device(dev):
        if (out_saved is None or rstd is None) and addend is None and d_addend is None:
            rc = L.lib().mot_embed_bwd(desc, _ptr(tok), _ptr(ids), _ptr(ttb), _ptr(E_tok), _ptr(E_byte), _ptr(lam),
                                       _ptr(grad_out), _ptr(gE_tok), _ptr(gE_byte), _ptr(g_lam), _ptr(ws), ws.numel(),
                                       flags, _stream(dev) if stream is None else stream)
        else:
            rc = L.lib().mot_embed_bwd_ex(desc, _ptr(tok), _ptr(ids), _ptr(ttb), _ptr(E_tok), _ptr(E_byte), _ptr(lam),
                                          _ptr(addend), _ptr(grad_out), _ptr(out_saved), _ptr(rstd), _ptr(gE_tok),
                                          _ptr(gE_byte), _ptr(g_lam), _ptr(d_addend), _ptr(ws), ws.numel(), flags,
                                          _stream(dev) if stream is None else stream)
    L.check(rc, "mot_embed_bwd")


def embed_backward_slab_out(desc: L.MotDesc, tok, ids, ttb, E_tok, E_byte, lam, grad_out, out_saved, rstd, gE_tok, gE_byte,
                            g_lam, ws, slab: int, n_slabs: int, *, reserve_sms: int = 0, plan_ready: bool = True,
                            ws_clean: bool = True, plan_joined: bool = False, stream: Optional[int] = None) -> None:
    """mot_embed_bwd_slab: slab `slab` of `n_slabs` of the backward (rows slab_rows(...) of gE_tok are final once its
    kernels have run; gE_byte / g_lam after the last slab).  desc must have been made with dp_slabs=n_slabs."""
    dev = grad_out.device
    flags = (L.WS_PLAN_READY if plan_ready else 0) | (L.WS_CLEAN if ws_clean else 0) | \
        (L.WS_PLAN_JOINED if plan_ready and plan_joined else 0)
    with _on_device(dev):
        rc = L.lib().mot_embed_bwd_slab(desc, _ptr(tok), _ptr(ids), _ptr(ttb), _ptr(E_tok), _ptr(E_byte), _ptr(lam),
                                        _ptr(grad_out), _ptr(out_saved), _ptr(rstd), _ptr(gE_tok), _ptr(gE_byte), _ptr(g_lam),
                                        _ptr(ws), ws.numel(), flags, slab, n_slabs, reserve_sms,
                                        _stream(dev) if stream is None else stream)
    L.check(rc, "mot_embed_bwd_slab")


def slab_rows(tok_vocab: int, slab: int, n_slabs: int):
    """Rows [lo, hi) of the token table that slab `slab` of `n_slabs` owns (mot_embed_slab_rows)."""
    import ctypes as C
    lo, hi = C.c_int32(), C.c_int32()
    L.check(L.lib().mot_embed_slab_rows(tok_vocab, slab, n_slabs, C.byref(lo), C.byref(hi)), "mot_embed_slab_rows")
    return lo.value, hi.value


class Workspace:
    """A backward workspace that remembers whether the library left it clean (so the steady-state step needs no
    memset), plus the fork / join events of the plan that runs on the side stream.  Instances are recycled through a
    pool per (device, table geometry)."""

    def __init__(self, key, dev):
        self.key = key
        self.buf: Optional[torch.Tensor] = None
        self.clean = False
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        cur = torch.cuda.current_stream(dev)
        self.ev_fork.record(cur)   # materialise the cudaEvent_t handles
        self.ev_join.record(cur)
        self.pending = False       # a plan was launched on the side stream and no backward has consumed it yet

    def reserve(self, desc: L.MotDesc, dev) -> torch.Tensor:
        need = embed_workspace_bytes(desc)
        if self.buf is None or self.buf.numel() < need or self.buf.device != dev:
            self.buf = torch.empty(need, dtype=torch.uint8, device=dev)
            self.clean = False
        return self.buf

    def __del__(self):   # dropped with a plan in flight (forward without backward): keep the allocator off the buffer
        try:
            if self.pending and self.buf is not None:
                self.buf.record_stream(side_stream(self.buf.device))
        except Exception:
            pass


_WS_POOL: dict = {}
_SIDE_STREAMS: dict = {}


def acquire_workspace(desc: L.MotDesc, dev) -> Workspace:
    # the zeroed head of the workspace (histogram | byte accumulators | one fp32 slot per stream chunk) is laid out by the
    # table geometry AND by (n_tokens, tok_dim): a buffer is only "clean" for the exact layout it was last used with
    key = (dev.index, desc.tok_vocab, desc.byte_vocab, desc.byte_dim, desc.combine, desc.n_tokens, desc.tok_dim, desc.dp_slabs)
    free = _WS_POOL.get(key)
    if free is None:
        free = _WS_POOL[key] = []
    ws = free.pop() if free else Workspace(key, dev)
    ws.reserve(desc, dev)
    return ws


def release_workspace(ws: Workspace) -> None:
    free = _WS_POOL.setdefault(ws.key, [])
    if len(free) < 4:
        free.append(ws)


def side_stream(dev) -> torch.cuda.Stream:
    """Per-device stream on which the backward plan (a counting sort of the token ids) runs beside the forward."""
    st = _SIDE_STREAMS.get(dev.index)
    if st is None:
        st = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return st


def embed_plan_async(desc: L.MotDesc, tok, ws: Workspace, dev, stream: Optional[int] = None) -> None:
    """mot_embed_plan_async: fork from the current stream, run the plan on the side stream, record ws.ev_join.  The
    backward waits for it with embed_plan_join()."""
    clean, ws.clean = ws.clean, False
    with _on_device(dev):
        rc = L.lib().mot_embed_plan_async(desc, _ptr(tok), _ptr(ws.buf), ws.buf.numel(), L.WS_CLEAN if clean else 0,
                                          _stream(dev) if stream is None else stream, side_stream(dev).cuda_stream,
                                          ws.ev_fork.cuda_event, ws.ev_join.cuda_event)
    L.check(rc, "mot_embed_plan_async")
    ws.pending = True


def embed_plan_join_if_capturing(ws: Optional[Workspace], dev, stream: Optional[int] = None) -> None:
    """Inside a CUDA-graph capture (torch.cuda.graph / make_graphed_callables) the side stream must rejoin before the
    capture of the forward ends; the graph keeps plan and forward concurrent.  Called by every forward that forked."""
    if ws is not None and ws.pending and torch.cuda.is_current_stream_capturing():
        embed_plan_join(ws, dev, stream)


def embed_plan_join(ws: Workspace, dev, stream: Optional[int] = None) -> None:
    rc = L.lib().mot_stream_wait_event(_stream(dev) if stream is None else stream, ws.ev_join.cuda_event)
    L.check(rc, "mot_stream_wait_event")
    ws.pending = False


class _MotEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: MixSpec, bpt: int, seq_len: int, tokens, byte_ids, ttb, E_tok, E_byte, lam, grad_bufs=None):
        dev = _require_cuda(tokens, byte_ids, ttb, E_tok, E_byte, lam)
        # optional ((param_tok, view_tok), (param_byte, view_byte)): views of a dp.GradBucket that become param.grad
        ctx.grad_bufs = grad_bufs
        tok = None
        if tokens is not None:
            tok = tokens.reshape(-1)
            tok = tok if tok.dtype == torch.int32 else tok.to(torch.int32)
            tok = tok.contiguous()
        ids = None
        if byte_ids is not None:
            if byte_ids.dtype not in (torch.int32, torch.int64):
                raise NotImplementedError("mot_b200: byte ids must be int32 or int64")
            ids = byte_ids.contiguous()
        n = tok.numel() if tok is not None else ids.numel() // bpt
        if ids is not None and ids.numel() != n * bpt:
            raise RuntimeError(f"mot_b200: byte ids have {ids.numel()} entries, expected {n}*{bpt}")
        E_tok_c = E_tok.contiguous() if E_tok is not None else None
        E_byte_c = E_byte.contiguous() if E_byte is not None else None
        if E_tok_c is not None and E_byte_c is not None and E_tok_c.dtype != E_byte_c.dtype:
            raise NotImplementedError("mot_b200: token and byte tables must share one dtype")
        lam_c = lam.detach().to(torch.float32).contiguous() if lam is not None else None
        desc = make_desc(spec, n, E_tok_c, E_byte_c, bpt, ids=ids, ttb=ttb, has_lam=lam is not None, seq_len=seq_len)
        # data-parallel pipeline: a bucket that holds exactly [token table | byte table] in symmetric memory lets the
        # backward run as vocabulary slabs with the exchange of slab k beside the backward of slab k + 1
        ctx.dp_bucket = None
        bucket = grad_bufs[2] if grad_bufs is not None and len(grad_bufs) > 2 else None
        if bucket is not None and bucket.pipelined and n > 0 and E_tok is not None and E_byte is not None \
                and len(bucket.params) == 2 and bucket.params[0] is grad_bufs[0][0] and bucket.params[1] is grad_bufs[1][0] \
                and bool(L.lib().mot_embed_bwd_uses_saved(desc)):
            desc = make_desc(spec, n, E_tok_c, E_byte_c, bpt, ids=ids, ttb=ttb, has_lam=lam is not None, seq_len=seq_len,
                             dp_slabs=bucket.n_slabs)
            ctx.dp_bucket = bucket
        ref = E_tok_c if E_tok_c is not None else E_byte_c
        out = torch.empty((n, desc.out_dim), dtype=ref.dtype, device=dev)
        # the backward needs the positions grouped by token id: start that sort now, beside the forward kernel
        ctx.ws = None
        st = _stream(dev)
        needs_grad = any(ctx.needs_input_grad[6:9])   # E_tok, E_byte, lam
        if needs_grad and tok is not None and n > 0:
            ctx.ws = acquire_workspace(desc, dev)
            embed_plan_async(desc, tok, ctx.ws, dev, st)
        # MoT-sum (runs/71): keep what rms_norm's autograd node keeps (its result and rstd); the backward then reads two
        # rows per occurrence instead of rebuilding the mixed row (mot_embed_bwd_ex)
        keep = needs_grad and n > 0 and bool(L.lib().mot_embed_bwd_uses_saved(desc))
        rstd = torch.empty(n, dtype=torch.float32, device=dev) if keep else None
        embed_forward_out(desc, tok, ids, ttb, E_tok_c, E_byte_c, lam_c, out, st, rstd=rstd)
        embed_plan_join_if_capturing(ctx.ws, dev, st)
        ctx.desc, ctx.dev = desc, dev
        saved = (tok, ids, ttb, E_tok_c, E_byte_c, lam_c, out if keep else None, rstd)
        ctx.save_for_backward(*[t if t is not None else _ABSENT for t in saved])
        ctx.present = [t is not None for t in saved]
        ctx.lam_dtype = lam.dtype if lam is not None else None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        saved = [t if ok else None for t, ok in zip(ctx.saved_tensors, ctx.present)]
        tok, ids, ttb, E_tok, E_byte, lam, out_saved, rstd = saved
        desc, dev = ctx.desc, ctx.dev
        g = grad_out.contiguous()
        if g.dtype != (E_tok if E_tok is not None else E_byte).dtype:
            g = g.to((E_tok if E_tok is not None else E_byte).dtype)
        direct = [None, None]
        if ctx.grad_bufs is not None:
            # the kernels overwrite every row, so they can write straight into the caller's flat bucket; the view is
            # installed as param.grad here (autograd would clone a view).  With a gradient already accumulated on the
            # parameter the normal path is taken and autograd adds to it.
            for i, (pv, E) in enumerate(zip(ctx.grad_bufs, (E_tok, E_byte))):
                if pv is not None and E is not None and pv[0].grad is None:
                    view = pv[1]
                    if view.dtype != E.dtype or view.shape != E.shape or not view.is_contiguous() or view.device != E.device:
                        raise TypeError("mot_b200: the gradient-bucket view of a table must match its dtype, shape and "
                                        f"device and be contiguous (table {tuple(E.shape)} {E.dtype}, view "
                                        f"{tuple(view.shape)} {view.dtype})")
                    direct[i] = pv
        gE_tok = direct[0][1] if direct[0] is not None else (torch.empty_like(E_tok) if E_tok is not None else None)
        gE_byte = direct[1][1] if direct[1] is not None else (torch.empty_like(E_byte) if E_byte is not None else None)
        g_lam = torch.empty(2, dtype=torch.float32, device=dev) if lam is not None else None
        st = _stream(dev)
        ws, planned = ctx.ws, ctx.ws is not None
        if ws is None:
            ws = acquire_workspace(desc, dev)
        if planned:
            if ws.pending:        # not yet joined (a captured forward joins at its own end)
                embed_plan_join(ws, dev, st)
            clean = True          # the plan ran on a clean (or freshly cleared) workspace and leaves it clean
        else:
            clean, ws.clean = ws.clean, False
        bucket = ctx.dp_bucket
        if bucket is not None and direct[0] is not None and direct[1] is not None and out_saved is not None \
                and not torch.cuda.is_current_stream_capturing():
            # slab k of the vocabulary, then its rows of the bucket go to the exchange stream while slab k + 1 is computed;
            # the byte table (finished by the last slab) travels with the last range.  bucket.all_reduce_avg() / wait() joins.
            K, V_, Dt_ = bucket.n_slabs, E_tok.shape[0], E_tok.shape[1]
            for k in range(K):
                embed_backward_slab_out(desc, tok, ids, ttb, E_tok, E_byte, lam, g, out_saved, rstd, gE_tok, gE_byte, g_lam,
                                        ws.buf, k, K, reserve_sms=bucket.reserve_sms if k > 0 else 0,
                                        plan_ready=planned or k > 0, ws_clean=clean or k > 0, plan_joined=planned, stream=st)
                lo, hi = slab_rows(V_, k, K)
                bucket.exchange_async(lo * Dt_, hi * Dt_ if k < K - 1 else bucket.flat.numel(), last=(k == K - 1))
        else:
            embed_backward_out(desc, tok, ids, ttb, E_tok, E_byte, lam, g, gE_tok, gE_byte, g_lam, ws.buf,
                               plan_ready=planned, ws_clean=clean, stream=st, out_saved=out_saved, rstd=rstd,
                               plan_joined=planned)
        ws.clean = True           # every completed backward leaves the head of the workspace zeroed
        if ctx.grad_bufs is not None and len(ctx.grad_bufs) > 2 and direct[0] is not None and tok is not None:
            bk = ctx.grad_bufs[2]
            if bk is not None and bk.sparse_rows and bk.params[0] is direct[0][0] and not bk._pending:
                bk.mark_rows(desc, ws.buf)      # the rows this batch gathered: the exchange moves only the union
        ctx.ws = None
        release_workspace(ws)
        if g_lam is not None:
            g_lam = g_lam.to(ctx.lam_dtype)
        if direct[0] is not None:
            direct[0][0].grad, gE_tok = gE_tok, None
        if direct[1] is not None:
            direct[1][0].grad, gE_byte = gE_byte, None
        return None, None, None, None, None, None, gE_tok, gE_byte, g_lam, None


def mot_embed(tokens: Optional[torch.Tensor], byte_ids: Optional[torch.Tensor], E_tok: Optional[torch.Tensor],
              E_byte: Optional[torch.Tensor], spec: MixSpec, *, bpt: int = 16, lam: Optional[torch.Tensor] = None,
              ttb: Optional[torch.Tensor] = None, seq_len: int = 0, grad_bufs=None) -> torch.Tensor:
    """Fused gather + pool + combine + norm.  Returns [n_tokens, out_dim] in the table dtype.

    tokens   int32/int64 [..] (flattened); byte_ids int32/int64 with n_tokens*bpt entries, token-major
    ([.., T*bpt]) or slot-major ([bpt, T], spec.slot_major); byte_ids=None derives the ids from `ttb`
    inside the kernel.  lam = float tensor [2] = (lam_tok, lam_byte) or None.  grad_bufs = optional
    ((param, view), (param, view)) pairs for the token / byte table: the backward writes the dense gradient into `view`
    (a slice of a dp.GradBucket) and installs it as `param.grad`."""
    if grad_bufs is None and _custom_ops_wanted():
        return _embed_via_custom_op(spec, bpt, seq_len, tokens, byte_ids, ttb, E_tok, E_byte, lam)
    return _MotEmbedFn.apply(spec, bpt, seq_len, tokens, byte_ids, ttb, E_tok, E_byte, lam, grad_bufs)


# ------------------------------------------------------------------------------------------------------------
# several tables gathered with the same token ids (the value embeddings, runs/7:252,308; spt/train_gpt.py:566,600)
# ------------------------------------------------------------------------------------------------------------
class _TokGatherFn(torch.autograd.Function):
    """outs[i] = tables[i][tokens]: plain gathers (no norm) of same-shaped tables.  The backward needs the positions
    grouped by token id once for all tables: one sort plan (started beside the first forward gather), one scatter
    launch per table reusing it (MOT_WS_PLAN_READY); the dense gradients need no token rows (R = 0)."""

    @staticmethod
    def forward(ctx, tokens, *tables):
        dev = _require_cuda(tokens, *tables)
        if not tables:
            raise RuntimeError("mot_b200.tok_gather: need at least one table")
        shape, dtype = tables[0].shape, tables[0].dtype
        for E in tables:
            if E.shape != shape or E.dtype != dtype or E.dim() != 2:
                raise NotImplementedError("mot_b200.tok_gather: tables must share one [V, D] shape and dtype")
        tok = tokens.reshape(-1)
        tok = (tok if tok.dtype == torch.int32 else tok.to(torch.int32)).contiguous()
        n = tok.numel()
        spec = MixSpec(combine="tok_only", out_norm=False)
        Es = [E.contiguous() for E in tables]
        desc = make_desc(spec, n, Es[0], None, 0, ids=None, ttb=None, has_lam=False)
        ctx.ws = None
        st = _stream(dev)
        if any(ctx.needs_input_grad[1:]) and n > 0:
            ctx.ws = acquire_workspace(desc, dev)
            embed_plan_async(desc, tok, ctx.ws, dev, st)
        outs = []
        for E in Es:
            out = torch.empty((n, shape[1]), dtype=dtype, device=dev)
            embed_forward_out(desc, tok, None, None, E, None, None, out, st)
            outs.append(out)
        embed_plan_join_if_capturing(ctx.ws, dev, st)
        ctx.desc, ctx.dev, ctx.n_tables = desc, dev, len(Es)
        ctx.save_for_backward(tok, *Es)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        tok, *Es = ctx.saved_tensors
        desc, dev = ctx.desc, ctx.dev
        st = _stream(dev)
        ws, planned = ctx.ws, ctx.ws is not None
        if ws is None:
            ws = acquire_workspace(desc, dev)
        if planned:
            if ws.pending:
                embed_plan_join(ws, dev, st)
            clean = True
        else:
            clean, ws.clean = ws.clean, False
        grads = []
        for i, (E, g) in enumerate(zip(Es, grad_outs)):
            if g is None or not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            g = g.reshape(-1, E.shape[1]).to(E.dtype).contiguous()
            gE = torch.empty_like(E)
            embed_backward_out(desc, tok, None, None, E, None, None, g, gE, None, None, ws.buf,
                               plan_ready=planned, ws_clean=clean, stream=st)
            planned, clean = True, True     # the first scatter leaves the plan in place and the accumulators zeroed
            grads.append(gE)
        ws.clean = True
        ctx.ws = None
        release_workspace(ws)
        return (None, *grads)


def tok_gather(tokens: torch.Tensor, *tables: torch.Tensor):
    """`[E(tokens) for E in tables]` (the value embeddings of runs/7:308 / spt/train_gpt.py:600): returns a tuple of
    [n_tokens, D] tensors; dense gradients, one token sort shared by every table."""
    if _custom_ops_wanted():
        return _tok_gather_via_custom_op(tokens, tables)
    return _TokGatherFn.apply(tokens, *tables)


# ------------------------------------------------------------------------------------------------------------
# concat + dense projection variants (V1 runs/7:226-234,317-319 and spt/train_gpt.py:439-443; V2 runs/72:227-230)
# ------------------------------------------------------------------------------------------------------------
def _gemm_dtype(*ts) -> int:
    """Operand dtype of the projection kernels: all bf16 (kind::f16) or all fp32 (read in place on the TF32 path)."""
    dts = {t.dtype for t in ts if t is not None}
    if dts == {torch.bfloat16}:
        return L.BF16
    if dts == {torch.float32}:
        return L.F32
    raise NotImplementedError(f"mot_b200: the tcgen05 projection kernels take all-bf16 or all-fp32 operands, got {sorted(map(str, dts))}")


def _linear_ws(n: int, in_dim: int, out_dim: int, dt: int, dev) -> Optional[torch.Tensor]:
    need = int(L.lib().mot_linear_workspace_bytes(n, in_dim, out_dim, dt))
    return torch.empty(need, dtype=torch.uint8, device=dev) if need else None


def linear_forward_out(x, w, y, bias=None) -> None:
    """y[n, Do] = x[n, K] . w[Do, K]^T (+ bias): mot_linear_fwd (tcgen05).  bf16 operands -> y bf16 or fp32; fp32
    operands (TF32 tensor cores) -> y fp32."""
    dt = _gemm_dtype(x, w)
    dev = _require_cuda(x, w, y, bias)
    n, K = x.shape
    if dt == L.F32 and y.dtype != torch.float32:
        raise NotImplementedError("mot_b200: fp32 operands produce an fp32 result")
    with _on_device(dev):
        rc = L.lib().mot_linear_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), n, K, w.shape[0], dt,
                                    1 if y.dtype == torch.float32 else 0, _stream(dev))
    L.check(rc, "mot_linear_fwd")


def linear_bwd_input_out(dy, w, dx) -> None:
    """dx[n, K] = dy[n, Do] . w[Do, K]: mot_linear_bwd_input (w read in place as an MN-major operand)."""
    dt = _gemm_dtype(dy, w, dx)
    dev = _require_cuda(dy, w, dx)
    ws = _linear_ws(dy.shape[0], w.shape[1], w.shape[0], dt, dev)
    with _on_device(dev):
        rc = L.lib().mot_linear_bwd_input(_ptr(dy), _ptr(w), _ptr(dx), dy.shape[0], w.shape[1], w.shape[0], dt,
                                          _ptr(ws), ws.numel() if ws is not None else 0, _stream(dev))
    L.check(rc, "mot_linear_bwd_input")


def linear_bwd_weight_out(dy, x, dw_f32, dw_bf16=None) -> None:
    """dw[Do, K] = dy^T . x reduced in fp32 (split over the tokens), optionally also cast to bf16."""
    dt = _gemm_dtype(dy, x)
    dev = _require_cuda(dy, x, dw_f32, dw_bf16)
    ws = _linear_ws(dy.shape[0], x.shape[1], dy.shape[1], dt, dev)
    with _on_device(dev):
        rc = L.lib().mot_linear_bwd_weight(_ptr(dy), _ptr(x), _ptr(dw_f32), _ptr(dw_bf16), dy.shape[0], x.shape[1],
                                           dy.shape[1], dt, _ptr(ws), ws.numel() if ws is not None else 0, _stream(dev))
    L.check(rc, "mot_linear_bwd_weight")


def rmsnorm_forward_out(y, out, eps: float = FP32_EPS) -> None:
    dev = _require_cuda(y, out)
    with _on_device(dev):
        rc = L.lib().mot_rmsnorm_fwd(_ptr(y), _ptr(out), y.shape[0], y.shape[1], _DTYPE[y.dtype], eps, _stream(dev))
    L.check(rc, "mot_rmsnorm_fwd")


def rmsnorm_backward_out(y, grad_out, dy, eps: float = FP32_EPS) -> None:
    dev = _require_cuda(y, grad_out, dy)
    with _on_device(dev):
        rc = L.lib().mot_rmsnorm_bwd(_ptr(y), _ptr(grad_out), _ptr(dy), y.shape[0], y.shape[1], _DTYPE[y.dtype], eps,
                                     _stream(dev))
    L.check(rc, "mot_rmsnorm_bwd")


def byte_pair_forward_out(ids_a, ids_b, bpt: int, E_byte, out, col_offset: int, eps: float = FP32_EPS) -> None:
    """mot_byte_pair_fwd: out[:, col_offset + k*bd : +bd] = rms_norm(E_byte[ids_a[:, k]] + E_byte[ids_b[:, k]])
    (spt/train_gpt.py:371-379), written into the byte columns of the [n, K] operand `out`."""
    dev = _require_cuda(ids_a, ids_b, E_byte, out)
    with _on_device(dev):
        rc = L.lib().mot_byte_pair_fwd(_ptr(ids_a), _ptr(ids_b), 1 if ids_a.dtype == torch.int64 else 0, out.shape[0], bpt,
                                       _ptr(E_byte), E_byte.shape[0], E_byte.shape[1], _DTYPE[E_byte.dtype], eps,
                                       _ptr(out), out.stride(0), col_offset, _stream(dev))
    L.check(rc, "mot_byte_pair_fwd")


def byte_pair_backward_out(ids_a, ids_b, bpt: int, E_byte, grad_rows, col_offset: int, gE_byte, eps: float = FP32_EPS) -> None:
    """mot_byte_pair_bwd: dense gE_byte from the byte columns of grad_rows [n, K]."""
    dev = _require_cuda(ids_a, ids_b, E_byte, grad_rows, gE_byte)
    need = int(L.lib().mot_byte_pair_workspace_bytes(E_byte.shape[0], E_byte.shape[1]))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    with _on_device(dev):
        rc = L.lib().mot_byte_pair_bwd(_ptr(ids_a), _ptr(ids_b), 1 if ids_a.dtype == torch.int64 else 0, grad_rows.shape[0],
                                       bpt, _ptr(E_byte), E_byte.shape[0], E_byte.shape[1], _DTYPE[E_byte.dtype], eps,
                                       _ptr(grad_rows), grad_rows.stride(0), col_offset, _ptr(gE_byte), _ptr(ws),
                                       ws.numel(), _stream(dev))
    L.check(rc, "mot_byte_pair_bwd")


def cast_out(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """CastedLinear's per-call `self.weight.type_as(x)` (spt/train_gpt.py:185-186) inside the library: fp32 -> bf16
    (mot_cast_f32_bf16).  Same dtype: returned as is."""
    w = w.detach()
    if w.dtype == dtype:
        return w.contiguous()
    if w.dtype != torch.float32 or dtype != torch.bfloat16:
        raise NotImplementedError(f"mot_b200: projection weight {w.dtype} with {dtype} tables is not supported "
                                  "(bf16 weight, or fp32 master weight cast to bf16 per call)")
    dev = _require_cuda(w)
    w = w.contiguous()
    out = torch.empty(w.shape, dtype=torch.bfloat16, device=dev)
    with _on_device(dev):
        rc = L.lib().mot_cast_f32_bf16(_ptr(w), _ptr(out), w.numel(), _stream(dev))
    L.check(rc, "mot_cast_f32_bf16")
    return out


def colsum_out(x: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of x [n, dim]: the bias gradient of F.linear (mot_colsum, fixed summation order)."""
    dev = _require_cuda(x)
    x = x.contiguous()
    out = torch.empty(x.shape[1], dtype=torch.float32, device=dev)
    with _on_device(dev):
        rc = L.lib().mot_colsum(_ptr(x), _ptr(out), x.shape[0], x.shape[1], _DTYPE[x.dtype], _stream(dev))
    L.check(rc, "mot_colsum")
    return out


def _mixout_copy_impl(x: torch.Tensor, bpt: int, backward: bool) -> torch.Tensor:
    dev = x.device
    rows, D = x.shape
    n = rows // bpt if backward else rows
    out = torch.empty((n if backward else n * bpt, D), dtype=x.dtype, device=dev)
    with _on_device(dev):
        fn = L.lib().mot_mixout_copy_bwd if backward else L.lib().mot_mixout_copy_fwd
        rc = fn(_ptr(x), _ptr(out), n, D, bpt, _DTYPE[x.dtype], _stream(dev))
    L.check(rc, "mot_mixout_copy_bwd" if backward else "mot_mixout_copy_fwd")
    return out


class _MixoutCopyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2d, bpt):
        ctx.bpt = bpt
        return _mixout_copy_impl(x2d, bpt, backward=False)

    @staticmethod
    def backward(ctx, g):
        return _mixout_copy_impl(g.contiguous(), ctx.bpt, backward=True), None


def mixout_copy(x: torch.Tensor, bytes_per_token: int) -> torch.Tensor:
    """ByteMixoutCopy's expand (spt/train_gpt.py:493): `einops.repeat(x, "... T D -> ... (T bpt) D")`; the backward sums
    the bpt copies of every row in fp32."""
    _require_cuda(x)
    if x.dtype not in _DTYPE:
        raise NotImplementedError(f"mot_b200.mixout_copy: dtype {x.dtype} is not supported (bf16 / fp32 only)")
    lead, T, D = x.shape[:-2], x.shape[-2], x.shape[-1]
    x2 = x.contiguous().view(-1, D)
    if _custom_ops_wanted():
        y = torch.ops.mot_b200.mixout_copy(x2, bytes_per_token)
    else:
        y = _MixoutCopyFn.apply(x2, bytes_per_token)
    return y.view(*lead, T * bytes_per_token, D)


def mixout_split(x: torch.Tensor, bytes_per_token: int) -> torch.Tensor:
    """ByteMixoutSplit's expand (spt/train_gpt.py:516): `rearrange(x, "... T (bpt D) -> ... (T bpt) D")` is a view of
    contiguous rows: no kernel, no copy."""
    lead, T, D = x.shape[:-2], x.shape[-2], x.shape[-1]
    if D % bytes_per_token:
        raise RuntimeError("model_dim must be divisible by bytes_per_token")   # spt/train_gpt.py:502
    return x.contiguous().view(*lead, T * bytes_per_token, D // bytes_per_token)


_PROJ_KEEP_OPERAND = not __import__("os").environ.get("MOT_PROJ_REGATHER")


class _MotEmbedProjFn(torch.autograd.Function):
    """out = f_out( [tok | bytes] . W^T + bias ): the fused gather builds the [n, K] operand (mot_embed_fwd, CONCAT),
    the projection runs on the tensor cores (mot_linear_fwd), the row norm after it is a separate HBM-bound pass over
    the bf16 product (the reference also rounds F.linear's output to bf16 before its rms_norm).  The [n, K] operand is
    kept for the backward (n*K*2 bytes: 268 MB at 64K x 2048, small on 180 GB; saves one gather pass, ~0.1 ms of a 1.2 ms
    step); MOT_PROJ_REGATHER=1 gathers it again instead (the round-1 behaviour, for memory-tight callers)."""

    @staticmethod
    def forward(ctx, spec: MixSpec, bpt: int, tokens, byte_ids, E_tok, E_byte, W, bias, byte_ids2=None):
        dev = _require_cuda(tokens, byte_ids, E_tok, E_byte, W, bias, byte_ids2)
        cdt = E_tok.dtype                      # compute dtype of the chain: bf16, or fp32 on the TF32 tensor-core path
        if cdt not in (torch.bfloat16, torch.float32) or E_byte.dtype != cdt:
            raise NotImplementedError("mot_b200: token and byte tables must both be bf16 or both fp32")
        tok = tokens.reshape(-1)
        tok = (tok if tok.dtype == torch.int32 else tok.to(torch.int32)).contiguous()
        if byte_ids.dtype not in (torch.int32, torch.int64):
            raise NotImplementedError("mot_b200: byte ids must be int32 or int64")
        ids = byte_ids.contiguous()
        n = tok.numel()
        if ids.numel() != n * bpt:
            raise RuntimeError(f"mot_b200: byte ids have {ids.numel()} entries, expected {n}*{bpt}")
        E_tok_c, E_byte_c = E_tok.contiguous(), E_byte.contiguous()
        ids2 = None
        if byte_ids2 is not None:
            # `--add-padded-and-pulled` (spt/train_gpt.py:371-379): two gathers summed before the per-byte norm.  The
            # token columns of the operand come from the fused kernel (tok-only, strided rows), the byte columns from
            # the pair kernels.
            if not spec.byte_norm or spec.slot_major or spec.bytes_first:
                raise NotImplementedError("mot_b200: the padded+pulled sum exists only with per-byte norms, token-major "
                                          "ids and [tok | bytes] order (spt/train_gpt.py:371-379,442-443)")
            ids2 = byte_ids2.contiguous()
            if ids2.dtype != ids.dtype or ids2.numel() != ids.numel():
                raise RuntimeError("mot_b200: the two byte-id tensors must share dtype and size")
            K = E_tok_c.shape[1] + bpt * E_byte_c.shape[1]
            desc = make_desc(MixSpec(combine="tok_only", tok_norm=spec.tok_norm, out_norm=False, eps=spec.eps), n, E_tok_c,
                             None, 0, ids=None, ttb=None, has_lam=False, row_stride=K, col_offset=0)
        else:
            a_spec = dataclasses.replace(spec, combine="concat", out_norm=False)
            desc = make_desc(a_spec, n, E_tok_c, E_byte_c, bpt, ids=ids, ttb=None, has_lam=False)
            K = desc.out_dim
        Do = W.shape[0]
        if W.shape[1] != K:
            raise RuntimeError(f"mot_b200: projection weight is {tuple(W.shape)}, expected [{Do}, {K}]")
        w16 = cast_out(W, cdt)                                 # CastedLinear: W.type_as(x) (spt/train_gpt.py:186)
        ctx.ws = None
        if any(ctx.needs_input_grad[4:6]) and n > 0:
            ctx.ws = acquire_workspace(desc, dev)
            embed_plan_async(desc, tok, ctx.ws, dev)
        A = torch.empty((n, K), dtype=cdt, device=dev)
        if ids2 is None:
            embed_forward_out(desc, tok, ids, None, E_tok_c, E_byte_c, None, A)
        elif n > 0:
            embed_forward_out(desc, tok, None, None, E_tok_c, None, None, A)
            byte_pair_forward_out(ids, ids2, bpt, E_byte_c, A, E_tok_c.shape[1], spec.eps)
        Y = torch.empty((n, Do), dtype=cdt, device=dev)
        if bias is not None and bias.dtype != torch.float32:
            raise NotImplementedError("mot_b200: the projection bias must be fp32 (mathblations/model.py:261)")
        b32 = bias.detach().contiguous() if bias is not None else None
        linear_forward_out(A, w16, Y, b32)
        keep_A = _PROJ_KEEP_OPERAND and any(ctx.needs_input_grad[4:8])
        if not keep_A:
            del A
        if spec.out_norm:
            out = torch.empty_like(Y)
            rmsnorm_forward_out(Y, out, spec.eps)
        else:
            out = Y
        embed_plan_join_if_capturing(ctx.ws, dev)
        ctx.desc, ctx.dev, ctx.spec = desc, dev, spec
        ctx.w_dtype, ctx.has_bias = W.dtype, bias is not None
        ctx.pair, ctx.bpt, ctx.K = ids2 is not None, bpt, K
        ctx.save_for_backward(tok, ids, E_tok_c, E_byte_c, w16, Y, ids2 if ids2 is not None else _ABSENT,
                              A if keep_A else _ABSENT)
        ctx.kept_A = keep_A
        return out

    @staticmethod
    def backward(ctx, grad_out):
        tok, ids, E_tok, E_byte, w16, Y, ids2, A_kept = ctx.saved_tensors
        desc, dev, spec = ctx.desc, ctx.dev, ctx.spec
        n, K, Do = tok.numel(), ctx.K, w16.shape[0]
        cdt = Y.dtype
        g = grad_out.reshape(n, Do).to(cdt).contiguous()
        if spec.out_norm:
            dY = torch.empty_like(Y)
            rmsnorm_backward_out(Y, g, dY, spec.eps)
        else:
            dY = g
        g_bias = colsum_out(dY) if ctx.has_bias else None
        if ctx.kept_A:
            A = A_kept                                                        # the forward's operand (n*K elements held)
            dA = torch.empty((n, K), dtype=cdt, device=dev)                   # a saved tensor is never overwritten
        else:
            A = torch.empty((n, K), dtype=cdt, device=dev)
            if not ctx.pair:
                embed_forward_out(desc, tok, ids, None, E_tok, E_byte, None, A)   # gathered again (MOT_PROJ_REGATHER=1)
            elif n > 0:
                embed_forward_out(desc, tok, None, None, E_tok, None, None, A)
                byte_pair_forward_out(ids, ids2, ctx.bpt, E_byte, A, E_tok.shape[1], spec.eps)
            dA = A                                                            # dX may overwrite the private copy
        dW32 = torch.empty((Do, K), dtype=torch.float32, device=dev)
        dW16 = torch.empty((Do, K), dtype=torch.bfloat16, device=dev) if ctx.w_dtype == torch.bfloat16 else None
        linear_bwd_weight_out(dY, A, dW32, dW16)
        linear_bwd_input_out(dY, w16, dA)
        gE_tok, gE_byte = torch.empty_like(E_tok), torch.empty_like(E_byte)
        ws, planned = ctx.ws, ctx.ws is not None
        if ws is None:
            ws = acquire_workspace(desc, dev)
        if planned:
            if ws.pending:
                embed_plan_join(ws, dev)
            clean = True
        else:
            clean, ws.clean = ws.clean, False
        if not ctx.pair:
            embed_backward_out(desc, tok, ids, None, E_tok, E_byte, None, dA, gE_tok, gE_byte, None, ws.buf,
                               plan_ready=planned, ws_clean=clean)
        else:   # token columns through the sorted-stream scatter, byte columns through the pair kernel
            embed_backward_out(desc, tok, None, None, E_tok, None, None, dA, gE_tok, None, None, ws.buf,
                               plan_ready=planned, ws_clean=clean)
            byte_pair_backward_out(ids, ids2, ctx.bpt, E_byte, dA, E_tok.shape[1], gE_byte, spec.eps)
        ws.clean = True
        ctx.ws = None
        release_workspace(ws)
        gW = dW16 if dW16 is not None else dW32          # cast_out admits bf16 or fp32 weights only
        return (None, None, None, None, gE_tok, gE_byte, gW, g_bias, None)


class _MotEmbedByteFcFn(torch.autograd.Function):
    """runs/71051:226-229,312-314 (V3f): out = norm(embed_tokens(tok) + F.linear(cat_k embed_bytes(ids_k), byte_fc)).
    Forward: bytes-only gather of the [n, D] operand (ids in the `.view(bpt,-1)` order), tcgen05 product, then the fused
    token gather + add + rms-norm (mot_embed_fwd_ex with the product as dense addend).  Backward: the fused kernel
    scatters d z into the token table and writes it densely as the gradient of the product; dW, dX on the tensor
    cores; bytes-only scatter.  The operand is gathered again in the backward, not kept."""

    @staticmethod
    def forward(ctx, spec_b: MixSpec, bpt: int, eps: float, tokens, byte_ids, E_tok, E_byte, W):
        dev = _require_cuda(tokens, byte_ids, E_tok, E_byte, W)
        cdt = E_tok.dtype
        if cdt not in (torch.bfloat16, torch.float32) or E_byte.dtype != cdt:
            raise NotImplementedError("mot_b200: token and byte tables must both be bf16 or both fp32")
        tok = tokens.reshape(-1)
        tok = (tok if tok.dtype == torch.int32 else tok.to(torch.int32)).contiguous()
        if byte_ids.dtype not in (torch.int32, torch.int64):
            raise NotImplementedError("mot_b200: byte ids must be int32 or int64")
        ids = byte_ids.contiguous()
        n = tok.numel()
        E_tok_c, E_byte_c = E_tok.contiguous(), E_byte.contiguous()
        D = E_tok_c.shape[1]
        if ids.numel() != n * bpt or bpt * E_byte_c.shape[1] != D or tuple(W.shape) != (D, D):
            raise RuntimeError("mot_b200: byte-FC mix needs bpt*byte_dim == token_dim and a square [D, D] weight")
        w16 = cast_out(W, cdt)
        desc_b = make_desc(spec_b, n, None, E_byte_c, bpt, ids=ids, ttb=None, has_lam=False)
        desc_t = make_desc(MixSpec(combine="tok_only", out_norm=True, eps=eps), n, E_tok_c, None, 0, ids=None, ttb=None,
                           has_lam=False)
        ctx.ws = None
        if ctx.needs_input_grad[5] and n > 0:
            ctx.ws = acquire_workspace(desc_t, dev)
            embed_plan_async(desc_t, tok, ctx.ws, dev)
        C = torch.empty((n, D), dtype=cdt, device=dev)
        embed_forward_out(desc_b, None, ids, None, None, E_byte_c, None, C)
        Y = torch.empty((n, D), dtype=cdt, device=dev)
        linear_forward_out(C, w16, Y)
        del C
        out = torch.empty((n, D), dtype=cdt, device=dev)
        if n > 0:
            embed_forward_out(desc_t, tok, None, None, E_tok_c, None, None, out, addend=Y)
        embed_plan_join_if_capturing(ctx.ws, dev)
        ctx.desc_b, ctx.desc_t, ctx.dev, ctx.w_dtype = desc_b, desc_t, dev, W.dtype
        ctx.save_for_backward(tok, ids, E_tok_c, E_byte_c, w16, Y)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        tok, ids, E_tok, E_byte, w16, Y = ctx.saved_tensors
        desc_b, desc_t, dev = ctx.desc_b, ctx.desc_t, ctx.dev
        n, D = Y.shape
        cdt = Y.dtype
        g = grad_out.reshape(n, D).to(cdt).contiguous()
        gE_tok, gE_byte = torch.empty_like(E_tok), torch.empty_like(E_byte)
        dY = torch.empty_like(Y)
        ws, planned = ctx.ws, ctx.ws is not None
        if ws is None:
            ws = acquire_workspace(desc_t, dev)
        if planned:
            if ws.pending:
                embed_plan_join(ws, dev)
            clean = True
        else:
            clean, ws.clean = ws.clean, False
        embed_backward_out(desc_t, tok, None, None, E_tok, None, None, g, gE_tok, None, None, ws.buf,
                           plan_ready=planned, ws_clean=clean, addend=Y, d_addend=dY)
        ws.clean = True
        ctx.ws = None
        release_workspace(ws)
        C = torch.empty((n, D), dtype=cdt, device=dev)
        embed_forward_out(desc_b, None, ids, None, None, E_byte, None, C)            # gathered again, not kept
        dW32 = torch.empty((D, D), dtype=torch.float32, device=dev)
        dW16 = torch.empty((D, D), dtype=torch.bfloat16, device=dev) if ctx.w_dtype == torch.bfloat16 else None
        linear_bwd_weight_out(dY, C, dW32, dW16)
        linear_bwd_input_out(dY, w16, C)                                             # dC overwrites the operand buffer
        wsb = acquire_workspace(desc_b, dev)
        clean_b, wsb.clean = wsb.clean, False
        embed_backward_out(desc_b, None, ids, None, None, E_byte, None, C, None, gE_byte, None, wsb.buf,
                           plan_ready=False, ws_clean=clean_b)
        wsb.clean = True
        release_workspace(wsb)
        gW = dW16 if dW16 is not None else dW32
        return None, None, None, None, None, gE_tok, gE_byte, gW


def mot_embed_byte_fc(tokens: torch.Tensor, byte_ids: torch.Tensor, E_tok: torch.Tensor, E_byte: torch.Tensor,
                      W_fc: torch.Tensor, *, bpt: int = 16, slot_major: bool = True, eps: float = FP32_EPS) -> torch.Tensor:
    """`norm(token_embs + F.linear(cat(byte_embs), byte_fc))` (runs/71051:226-229,312-314): [n_tokens, D]."""
    spec_b = MixSpec(combine="bytes_only", out_norm=False, slot_major=slot_major, eps=eps)
    if _custom_ops_wanted():
        return _byte_fc_via_custom_op(spec_b, bpt, eps, tokens, byte_ids, E_tok, E_byte, W_fc)
    return _MotEmbedByteFcFn.apply(spec_b, bpt, eps, tokens, byte_ids, E_tok, E_byte, W_fc)


def mot_embed_proj(tokens: torch.Tensor, byte_ids: torch.Tensor, E_tok: torch.Tensor, E_byte: torch.Tensor,
                   W: torch.Tensor, spec: MixSpec, *, bpt: int = 16, bias: Optional[torch.Tensor] = None,
                   byte_ids2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Concat + dense projection variants: `norm(F.linear(cat([f(tok), f(bytes)]), W))` (runs/7:226-234, runs/72,
    spt/train_gpt.py:439-443).  spec.tok_norm / byte_norm are the per-input norms, spec.out_norm the norm after the
    projection, spec.bytes_first the operand order.  Tables bf16 with W bf16 (runs) or an fp32 master (spt: cast per
    call, gradient returned in fp32); or everything fp32 (mathblations: TF32 tensor cores).  byte_ids2: the second id
    tensor of `--add-padded-and-pulled` (spt/train_gpt.py:371-379), rows summed before the per-byte norm.  Returns
    [n_tokens, W.shape[0]] in the table dtype."""
    if _custom_ops_wanted():
        return _proj_via_custom_op(spec, bpt, tokens, byte_ids, E_tok, E_byte, W, bias, byte_ids2)
    return _MotEmbedProjFn.apply(spec, bpt, tokens, byte_ids, E_tok, E_byte, W, bias, byte_ids2)


def launch_count() -> int:
    return int(L.lib().mot_launch_count())


def reset_launch_count() -> None:
    L.lib().mot_launch_count_reset()


# torch.library registration of the same entry points (torch.compile / compiled autograd see opaque operators)
from ._library import (custom_ops_wanted as _custom_ops_wanted, set_custom_ops,  # noqa: E402,F401
                       embed_via_custom_op as _embed_via_custom_op, tok_gather_via_custom_op as _tok_gather_via_custom_op,
                       proj_via_custom_op as _proj_via_custom_op, byte_fc_via_custom_op as _byte_fc_via_custom_op)
