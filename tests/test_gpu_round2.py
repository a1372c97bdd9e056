"""Round-2 GPU parity additions.

* element-wise bf16 bars: every element within one bf16 ulp of the fp32-math oracle (plus a small absolute
  floor for cancelled sums), beside the normalised max-abs bars of test_gpu_parity.py; full parity of the headline kernel
  (mot_bwd_sum_kernel) at the BASELINE shape 49152 x 768, uniform tokens;
* V3e (runs/71042:311-314) and the ByteMixout expands (spt/train_gpt.py:493,516) against outputs of the reference itself
  (tests/golden/runs_float_r2.npz, mixout.npz; generator: tests/golden/make_golden_r2.py);
* the widest reference concat operand (K = 3072), CUDA-graph capture of every module, workspace reuse across sizes,
  gradient-bucket view validation.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mot_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def nerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def bf16_elementwise_violations(got, want, ulps=1.0, floor_rel_rms=2.0 ** -12):
    """Number of elements with |got - want| > ulps * ulp_bf16(want) + floor, floor = floor_rel_rms * rms(want).
    ulp_bf16(x) = 2^(floor(log2|x|) - 7): the spacing of bf16 values at x (8 significant bits)."""
    g, w = got.detach().double().cpu().reshape(-1), want.detach().double().cpu().reshape(-1)
    ulp = torch.exp2(torch.floor(torch.log2(w.abs().clamp_min(1e-300))) - 7)
    floor = floor_rel_rms * float(w.pow(2).mean().sqrt())
    bad = (g - w).abs() > ulps * ulp + floor
    return int(bad.sum()), float(((g - w).abs() / (ulp + floor)).max())


# ------------------------------------------------------------------------------------------- element-wise bars
@pytest.mark.parametrize("N,V,zipf", [(2000, 1500, False), (6000, 900, True)])
def test_mot_sum_bf16_elementwise_vs_oracle(N, V, zipf):
    import mot_b200
    d = dev()
    g = torch.Generator().manual_seed(21)
    Dt, bd, bpt = 768, 48, 16
    toks = ((torch.rand(N, generator=g) ** 4 * V).long().clamp_(0, V - 1).int() if zipf
            else torch.randint(0, V, (N,), generator=g, dtype=torch.int32))
    ids = torch.randint(0, 458, (bpt, N), generator=g, dtype=torch.int32)
    E_tok, E_byte = torch.randn(V, Dt, generator=g).bfloat16(), torch.randn(458, bd, generator=g).bfloat16()
    gout = torch.randn(N, Dt, generator=g).bfloat16()
    want_out, want = O.mot_embed_fwd_bwd(O.VARIANTS["V3"][0], toks, ids, E_tok, E_byte, gout, bpt=bpt, slot_major=True)
    from mot_b200 import ops
    spec = mot_b200.MixSpec(combine="add", slot_major=True)
    tk, idd, Et, Eb, go = toks.to(d), ids.to(d), E_tok.to(d), E_byte.to(d), gout.to(d)
    desc = ops.make_desc(spec, N, Et, Eb, bpt, ids=idd, ttb=None, has_lam=False)
    out = torch.empty(N, Dt, dtype=torch.bfloat16, device=d)
    rstd = torch.empty(N, dtype=torch.float32, device=d)
    ops.embed_forward_out(desc, tk, idd, None, Et, Eb, None, out, rstd=rstd)
    # forward: fp32 math, one rounding -> within one bf16 ulp of the oracle everywhere
    bad, worst = bf16_elementwise_violations(out, want_out, ulps=1.0, floor_rel_rms=2.0 ** -20)
    assert bad == 0, f"out: {bad} elements beyond 1 ulp (worst {worst:.2f})"
    ws = torch.empty(ops.embed_workspace_bytes(desc), dtype=torch.uint8, device=d)
    # dense gradients: fp32 sums over the occurrences, one rounding.
    #  recompute kernel (rebuilds the mixed row in fp32): 1 ulp + 2^-12 of the rms (fp32 summation order);
    #  saved-output kernel: d z = r*g - o*(r*mean(g.o)) reads the bf16 forward result o, i.e. a 2^-9 relative error on the
    #  projection term only (F.rms_norm's own bf16 backward carries the same error on its saved INPUT): 1 ulp + 2^-9 rms.
    for name, kw, floor in (("recompute", {}, 2.0 ** -12), ("saved", dict(out_saved=out, rstd=rstd), 2.0 ** -9)):
        gt, gb = torch.empty_like(Et), torch.empty_like(Eb)
        ops.embed_backward_out(desc, tk, idd, None, Et, Eb, None, go, gt, gb, None, ws, plan_ready=False, ws_clean=False, **kw)
        for what, got, ref in (("gE_tok", gt, want["E_tok"]), ("gE_byte", gb, want["E_byte"])):
            bad, worst = bf16_elementwise_violations(got, ref, ulps=1.0, floor_rel_rms=floor)
            assert bad == 0, f"{name} {what}: {bad} elements beyond 1 ulp + floor (worst {worst:.2f})"


def test_headline_backward_full_parity_49152x768_uniform():
    """mot_bwd_sum_kernel at the BASELINE shape (config 3: 49152 tokens, V = 50257, 768 = 16 x 48, bf16, uniform ids):
    EVERY element of both dense gradients against the fp32 formulas restated with torch ops on the GPU (test-only; the
    CPU oracle needs minutes at this size), normalised and element-wise."""
    import mot_b200
    from mot_b200 import ops
    d = dev()
    g = torch.Generator(device="cuda").manual_seed(12345)
    N, V, Dt, bd, bpt = 49152, 50257, 768, 48, 16
    toks = torch.randint(0, V - 1, (N,), generator=g, device=d, dtype=torch.int32)
    ids = torch.randint(0, 458, (bpt, N), generator=g, device=d, dtype=torch.int32)
    Et = torch.randn(V, Dt, generator=g, device=d).bfloat16().requires_grad_(True)
    Eb = torch.randn(458, bd, generator=g, device=d).bfloat16().requires_grad_(True)
    gout = torch.randn(N, Dt, generator=g, device=d).bfloat16()
    spec = mot_b200.MixSpec(combine="add", slot_major=True)
    assert ops.embed_bwd_uses_saved(ops.make_desc(spec, N, Et, Eb, bpt, ids=ids, ttb=None, has_lam=False))
    out = mot_b200.mot_embed(toks, ids, Et, Eb, spec, bpt=bpt)
    out.backward(gout)
    z = Et.detach().float()[toks.long()] + Eb.detach().float()[ids.long().t()].reshape(N, -1)
    r = torch.rsqrt(z.pow(2).mean(-1, keepdim=True) + mot_b200.FP32_EPS)
    gf = gout.float()
    dz = r * gf - z * (r ** 3) * (gf * z).mean(-1, keepdim=True)
    want_t = torch.zeros(V, Dt, device=d).index_add_(0, toks.long(), dz)
    want_b = torch.zeros(458, bd, device=d).index_add_(0, ids.long().t().reshape(-1), dz.reshape(N * bpt, bd))
    assert nerr(out, z * r) <= 2.0 ** -8
    assert nerr(Et.grad, want_t) <= 2.0 ** -8 and nerr(Eb.grad, want_b) <= 2.0 ** -8
    bad, worst = bf16_elementwise_violations(out, z * r, floor_rel_rms=2.0 ** -20)
    assert bad == 0, f"out: {bad} (worst {worst:.2f})"
    # saved-output kernel: 1 ulp + 2^-9 of the rms (it reads the bf16 forward result in the projection term, see above)
    bad, worst = bf16_elementwise_violations(Et.grad, want_t, floor_rel_rms=2.0 ** -9)
    assert bad == 0, f"gE_tok: {bad} of {V * Dt} elements beyond 1 ulp + floor (worst {worst:.2f})"
    # ~1700 fp32 adds per byte-table element in atomic order vs index_add_'s own atomic order
    bad, worst = bf16_elementwise_violations(Eb.grad, want_b, floor_rel_rms=2.0 ** -9)
    assert bad == 0, f"gE_byte: {bad} (worst {worst:.2f})"
    untouched = torch.ones(V, dtype=torch.bool, device=d)
    untouched[toks.long()] = False
    assert float(Et.grad[untouched].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------- V3e, reference golden
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_v3e_reference_golden_through_cuda(golden_dir, dt):
    """runs/71042:311-314: both scalars divided by their sum before scaling the normalised inputs."""
    import mot_b200
    g = np.load(os.path.join(golden_dir, "runs_float_r2.npz"))
    d = dev()
    k = f"V3e_run71042_{dt}"
    dtype = torch.float32 if dt == "f32" else torch.bfloat16
    T_, Dt = g[f"{k}_tokens"].shape[0], g[f"{k}_E_tok"].shape[1]
    m = mot_b200.MoTEmbedding(g[f"{k}_E_tok"].shape[0], 458, Dt, g[f"{k}_E_byte"].shape[1], 16, variant="V3e").to(d)
    with torch.no_grad():
        m.embed_tokens.weight.data = torch.from_numpy(g[f"{k}_E_tok"]).to(d).to(dtype)
        m.embed_bytes.weight.data = torch.from_numpy(g[f"{k}_E_byte"]).to(d).to(dtype)
        m.lambdas.copy_(torch.from_numpy(g[f"{k}_scalars"][-2:]).to(d))          # [byte, token] = scalars[-2], scalars[-1]
    out = m(torch.from_numpy(g[f"{k}_tokens"]).to(d), torch.from_numpy(g[f"{k}_byte_inputs"]).to(d))
    out.backward(torch.from_numpy(g[f"{k}_gout"]).to(d).to(dtype).reshape(out.shape))
    # fp32: equal to round-off.  bf16: the eager reference rounds every intermediate to bf16, the kernel once.
    tol = 1e-5 if dt == "f32" else 3e-2
    assert nerr(out, torch.from_numpy(g[f"{k}_out"]).reshape(out.shape)) <= tol
    assert nerr(m.embed_tokens.weight.grad, torch.from_numpy(g[f"{k}_gE_tok"])) <= tol
    assert nerr(m.embed_bytes.weight.grad, torch.from_numpy(g[f"{k}_gE_byte"])) <= tol
    assert nerr(m.lambdas.grad, torch.from_numpy(g[f"{k}_gscalars"][-2:])) <= (1e-4 if dt == "f32" else 8e-2)
    # and against the oracle (fp32 math on the same parameters) at the kernel bars
    sc = torch.from_numpy(g[f"{k}_scalars"])
    want_out, want = O.mot_embed_fwd_bwd(O.VARIANTS["V3e"][0], torch.from_numpy(g[f"{k}_tokens"]),
                                         torch.from_numpy(g[f"{k}_byte_inputs"]), torch.from_numpy(g[f"{k}_E_tok"]).to(dtype),
                                         torch.from_numpy(g[f"{k}_E_byte"]).to(dtype), torch.from_numpy(g[f"{k}_gout"]).to(dtype),
                                         bpt=16, slot_major=True, lam_tok=sc[-1], lam_byte=sc[-2])
    kt = 1e-5 if dt == "f32" else 2.0 ** -8
    assert nerr(out, want_out.reshape(out.shape)) <= kt and nerr(m.embed_tokens.weight.grad, want["E_tok"]) <= kt
    assert nerr(m.lambdas.grad, torch.stack([want["lam_byte"], want["lam_tok"]])) <= (1e-4 if dt == "f32" else kt)


# ------------------------------------------------------------------------------------------- output-side expands
@pytest.mark.parametrize("tag", ["copy_f32", "copy_bf16", "split_f32", "split_bf16"])
def test_mixout_expands_match_reference_golden(golden_dir, tag):
    """ByteMixoutCopy / ByteMixoutSplit (spt/train_gpt.py:483-518) run with n_layer_out = 0: forward = the expand."""
    import mot_b200
    g = np.load(os.path.join(golden_dir, "mixout.npz"))
    d = dev()
    dtype = torch.float32 if tag.endswith("f32") else torch.bfloat16
    bpt = int(g[f"{tag}_bpt"])
    x = torch.from_numpy(g[f"{tag}_x"]).to(d).to(dtype).requires_grad_(True)
    fn = mot_b200.mixout_copy if tag.startswith("copy") else mot_b200.mixout_split
    y = fn(x, bpt)
    y.backward(torch.from_numpy(g[f"{tag}_gout"]).to(d).to(dtype))
    assert tuple(y.shape) == g[f"{tag}_y"].shape
    assert torch.equal(y.detach().float().cpu(), torch.from_numpy(g[f"{tag}_y"]).float())       # a copy: bit-exact
    assert nerr(x.grad, torch.from_numpy(g[f"{tag}_gx"])) <= (1e-6 if dtype == torch.float32 else 2.0 ** -8)


def test_mixout_copy_large_and_errors():
    import mot_b200
    d = dev()
    x = torch.randn(3, 1000, 768, device=d).bfloat16().requires_grad_(True)
    y = mot_b200.mixout_copy(x, 16)
    go = torch.randn_like(y)
    y.backward(go)
    assert torch.equal(y, x.detach().repeat_interleave(16, dim=1))
    want = go.float().view(3, 1000, 16, 768).sum(2)
    assert nerr(x.grad, want) <= 2.0 ** -8
    assert mot_b200.mixout_copy(x[:, :0], 16).shape == (3, 0, 768)
    with pytest.raises(RuntimeError):
        mot_b200.mixout_copy(x.detach().cpu(), 16)
    with pytest.raises(NotImplementedError):
        mot_b200.mixout_copy(x.detach().half(), 16)
    with pytest.raises(RuntimeError):
        mot_b200.mixout_split(x, 7)


# ------------------------------------------------------------------------------------------- ADVICE round 1
def test_widest_reference_concat_operand_k3072():
    """scaled-pre-train/experiments100_000steps.sh: token_dim 1024, byte_dim 128, bpt 16 -> K = 3072 (each half of the
    split concat fits one launch)."""
    import mot_b200
    d = dev()
    g = torch.Generator().manual_seed(5)
    N, V, bpt, Dt, bd, Do = 300, 400, 16, 1024, 128, 1024
    K = Dt + bpt * bd
    toks = torch.randint(0, V, (2, N // 2), generator=g, dtype=torch.int32)
    ids = torch.randint(0, 458, (2, N // 2 * bpt), generator=g, dtype=torch.int64)
    m = mot_b200.SptByteMixEmbedding(V, 458, Dt, bd, Do, bytes_per_token=bpt).to(d)
    m.embed.bfloat16()
    gout = torch.randn(2, N // 2, Do, generator=g).bfloat16()
    x = m(toks.to(d), None, ids.to(d))
    x.backward(gout.to(d))
    W = m.byte_mixin.mixin.mixin.weight
    want_out, want = O.mot_embed_fwd_bwd(O.VARIANTS["V1"][0], toks, ids.view(1, -1), m.embed.embed_tokens.weight.detach().cpu(),
                                         m.embed.embed_bytes.weight.detach().cpu(), gout, bpt=bpt, slot_major=False,
                                         W=W.detach().cpu().bfloat16())
    assert W.shape == (Do, K) and W.grad.dtype == torch.float32
    assert nerr(x, want_out.reshape(x.shape)) <= 2.0 ** -6
    assert nerr(m.embed.embed_tokens.weight.grad, want["E_tok"]) <= 2.0 ** -6
    assert nerr(m.embed.embed_bytes.weight.grad, want["E_byte"]) <= 2.0 ** -6
    assert nerr(W.grad, want["W"]) <= 2.0 ** -6


@pytest.mark.parametrize("kind", ["sum", "proj_runs", "proj_spt", "digits", "byte_fc", "value"])
def test_every_module_captures_into_cuda_graphs(kind):
    """torch.cuda.make_graphed_callables on each module family: the forward rejoins the side stream
    inside the capture; a replayed step reproduces the eager gradients."""
    import gc
    import mot_b200
    d = dev()
    # objects of earlier tests (graphs with private pools, tensors last used on a side stream) must not be finalised in the
    # middle of this test's capture: the caching allocator would query their events, which invalidates a global-mode capture
    gc.collect()
    torch.cuda.synchronize()
    torch.manual_seed(7)
    V, N, bpt = 600, 256, 16
    tok = torch.randint(0, V, (N,), device=d, dtype=torch.int32)
    if kind == "sum":
        m = mot_b200.MoTEmbedding(V, 458, 512, 32, bpt, variant="V3").to(d).bfloat16()
        args = (tok, torch.randint(0, 458, (bpt, N), device=d, dtype=torch.int32))
    elif kind == "proj_runs":
        m = mot_b200.MoTProjEmbedding(V, 458, 128, 16, 256, bpt, variant="V1").to(d).bfloat16()
        args = (tok, torch.randint(0, 458, (1, N * bpt), device=d, dtype=torch.int32))
    elif kind == "proj_spt":
        m = mot_b200.SptByteMixEmbedding(V, 458, 64, 16, 128, bytes_per_token=bpt, add_padded_and_pulled=True).to(d)
        m.embed.bfloat16()
        args = (tok.view(2, -1), torch.randint(0, 458, (2, N // 2 * bpt), device=d), torch.randint(0, 458, (2, N // 2 * bpt), device=d))
    elif kind == "digits":
        m = mot_b200.DigitMixinEmbedding(V, 64, 64, 4).to(d)
        args = (tok.view(2, -1).long(), torch.randint(0, 14, (2, N // 2 * 4), device=d))
    elif kind == "byte_fc":
        m = mot_b200.MoTByteFcEmbedding(V, 458, 512, 32, bpt).to(d).bfloat16()
        args = (tok, torch.randint(0, 458, (bpt, N), device=d, dtype=torch.int32))
    else:
        m = mot_b200.TokenValueEmbeddings(V, 256).to(d).bfloat16()
        args = (tok,)

    def run(mod):
        for p in m.parameters():
            p.grad = None
        out = mod(*args)
        outs = out if isinstance(out, (list, tuple)) else [out]
        torch.manual_seed(1)
        loss = sum((o.float() * torch.randn_like(o.float())).sum() for o in outs)
        loss.backward()
        return [p.grad.clone() for p in m.parameters()]

    want = run(m)
    graphed = torch.cuda.make_graphed_callables(m, tuple(a.clone() for a in args))
    for _ in range(2):
        got = run(graphed)
    for a, b in zip(got, want):
        assert nerr(a, b) <= 2.0 ** -7


def test_workspace_pool_survives_shrink_then_grow():
    """N1 -> N2 < N1 -> N1 on one table geometry: a pooled workspace is only reused as `clean` for the layout it was
    last used with (plan data of the smaller layout landed in the larger layout's zeroed region)."""
    import mot_b200
    d = dev()
    torch.manual_seed(11)
    V, Dt, bd, bpt = 3000, 512, 32, 16
    Et = torch.randn(V, Dt, device=d).bfloat16().requires_grad_(True)
    Eb = torch.randn(458, bd, device=d).bfloat16().requires_grad_(True)
    spec = mot_b200.MixSpec(combine="add", slot_major=True)

    def step(N, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        tok = (torch.rand(N, generator=g, device=d) ** 3 * V).long().clamp_(0, V - 1).int()
        ids = torch.randint(0, 458, (bpt, N), generator=g, device=d, dtype=torch.int32)
        go = torch.randn(N, Dt, generator=g, device=d).bfloat16()
        Et.grad = Eb.grad = None
        mot_b200.mot_embed(tok, ids, Et, Eb, spec, bpt=bpt).backward(go)
        return Et.grad.clone(), Eb.grad.clone()

    first = step(40000, 1)
    step(700, 2)
    step(5000, 3)
    again = step(40000, 1)
    assert nerr(again[0], first[0]) <= 2.0 ** -8 and nerr(again[1], first[1]) <= 2.0 ** -8


def test_grad_bucket_view_must_match_table():
    import mot_b200
    from mot_b200 import dp
    d = dev()
    m = mot_b200.MoTEmbedding(500, 458, 256, 16, 16, variant="V3").to(d).bfloat16()
    bad = dp.GradBucket([m.embed_tokens.weight, m.embed_bytes.weight], dtype=torch.float32)
    m.attach_grad_bucket(bad)
    tok = torch.randint(0, 500, (64,), device=d, dtype=torch.int32)
    ids = torch.randint(0, 458, (16, 64), device=d, dtype=torch.int32)
    with pytest.raises(TypeError):
        m(tok, ids).sum().backward()


def test_util_kernels_cast_and_colsum():
    from mot_b200 import ops
    d = dev()
    torch.manual_seed(3)
    w = torch.randn(1000, 1031, device=d)
    assert torch.equal(ops.cast_out(w, torch.bfloat16), w.bfloat16())          # round-to-nearest-even, like .to()
    assert ops.cast_out(w, torch.float32).data_ptr() == w.data_ptr()
    for dtype, n, dim in ((torch.bfloat16, 11264, 256), (torch.float32, 777, 1000), (torch.bfloat16, 1, 64)):
        x = torch.randn(n, dim, device=d).to(dtype)
        got = ops.colsum_out(x)
        assert got.dtype == torch.float32 and nerr(got, x.double().sum(0)) <= 1e-5
        assert torch.equal(got, ops.colsum_out(x))                              # fixed order: run-to-run identical


# ------------------------------------------------------------------------------------------- vocabulary slabs (dp pipeline)
@pytest.mark.parametrize("N,V,n_slabs,zipf,reserve", [(49152, 50257, 4, False, 8), (20000, 6000, 3, True, 16), (700, 50257, 8, False, 0),
                                                      (5, 64, 2, False, 0)])
def test_backward_as_vocabulary_slabs_equals_one_piece(N, V, n_slabs, zipf, reserve):
    """mot_embed_bwd_slab k = 0..n-1 (what the data-parallel pipeline runs beside the exchange) against mot_embed_bwd_ex
    in one piece: same dense gradients (another stream chunking: fp32 summation order of duplicates differs), and
    after slab k exactly the rows of slabs <= k have been written."""
    import mot_b200
    from mot_b200 import ops
    d = dev()
    g = torch.Generator(device="cuda").manual_seed(N + n_slabs)
    Dt, bd, bpt = 768, 48, 16
    toks = ((torch.rand(N, generator=g, device=d) ** 4 * V).long().clamp_(0, V - 1).int() if zipf
            else torch.randint(0, V, (N,), generator=g, device=d, dtype=torch.int32))
    ids = torch.randint(0, 458, (bpt, N), generator=g, device=d, dtype=torch.int32)
    Et = torch.randn(V, Dt, generator=g, device=d).bfloat16()
    Eb = torch.randn(458, bd, generator=g, device=d).bfloat16()
    go = torch.randn(N, Dt, generator=g, device=d).bfloat16()
    spec = mot_b200.MixSpec(combine="add", slot_major=True)
    desc1 = ops.make_desc(spec, N, Et, Eb, bpt, ids=ids, ttb=None, has_lam=False)
    descK = ops.make_desc(spec, N, Et, Eb, bpt, ids=ids, ttb=None, has_lam=False, dp_slabs=n_slabs)
    assert ops.embed_workspace_bytes(descK) >= ops.embed_workspace_bytes(desc1)
    out = torch.empty(N, Dt, dtype=torch.bfloat16, device=d)
    rstd = torch.empty(N, dtype=torch.float32, device=d)
    ops.embed_forward_out(desc1, toks, ids, None, Et, Eb, None, out, rstd=rstd)
    ws1 = torch.empty(ops.embed_workspace_bytes(desc1), dtype=torch.uint8, device=d)
    gt1, gb1 = torch.empty_like(Et), torch.empty_like(Eb)
    ops.embed_backward_out(desc1, toks, ids, None, Et, Eb, None, go, gt1, gb1, None, ws1, plan_ready=False, ws_clean=False,
                           out_saved=out, rstd=rstd)
    wsK = torch.empty(ops.embed_workspace_bytes(descK), dtype=torch.uint8, device=d)
    ops.embed_workspace_init(descK, wsK)
    ops.embed_plan(descK, toks, wsK, ws_clean=True)
    for rep in range(2):                       # twice: the slabs leave the workspace clean for the next step
        gt = torch.full((V, Dt), float("nan"), dtype=torch.bfloat16, device=d)
        gb = torch.full((458, bd), float("nan"), dtype=torch.bfloat16, device=d)
        if rep == 1:
            ops.embed_plan(descK, toks, wsK, ws_clean=True)
        for k in range(n_slabs):
            ops.embed_backward_slab_out(descK, toks, ids, None, Et, Eb, None, go, out, rstd, gt, gb, None, wsK, k, n_slabs,
                                        reserve_sms=reserve if k > 0 else 0)
            lo, hi = ops.slab_rows(V, k, n_slabs)
            torch.cuda.synchronize()
            assert not bool(torch.isnan(gt[:hi]).any()), f"slab {k}: a row below {hi} was not written"
            assert bool(torch.isnan(gt[hi:]).all()), f"slab {k}: wrote beyond its rows"
            assert bool(torch.isnan(gb).any()) == (k < n_slabs - 1)        # the byte table finishes with the last slab
        assert nerr(gt, gt1) <= 2.0 ** -8 and nerr(gb, gb1) <= 2.0 ** -8
        bad, worst = bf16_elementwise_violations(gt, gt1.float(), ulps=1.0, floor_rel_rms=2.0 ** -12)
        assert bad == 0, f"{bad} elements differ by more than one ulp (worst {worst:.2f})"
    with pytest.raises(RuntimeError):          # a descriptor that did not announce the slabs
        ops.embed_backward_slab_out(desc1, toks, ids, None, Et, Eb, None, go, out, rstd, gt, gb, None, ws1, 0, n_slabs)
