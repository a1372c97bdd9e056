"""GPU parity of the rows SURVEY.md 8(f)-2 / V6 / V3g add around the path: the value embeddings gathered with the
same token ids (plain gathers sharing one sort plan), the MoT value embeddings of runs/9, and the split residual of
runs/71081, each against the CPU oracle.  Bars as in test_gpu_parity.py / test_gpu_proj.py."""
import pytest
import torch

from oracle import mot_oracle as O

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 2.0 ** -8}


def dev():
    return torch.device("cuda:0")


def nerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("dtype,V,D,N,zipf", [
    (torch.float32, 300, 256, 1000, False),
    (torch.bfloat16, 2000, 1024, 5000, True),      # hot rows straddle stream chunks
    (torch.bfloat16, 50257, 768, 4096, False),     # mostly empty rows: zero fill
    (torch.float32, 64, 128, 1, False),
])
def test_value_embeddings_match_oracle(dtype, V, D, N, zipf):
    """ve = [value_embed(tokens) for value_embed in value_embeds] (runs/7:308) and its three dense gradients;
    one of the tables receives no gradient (frozen) to cover the partial case."""
    import mot_b200
    d = dev()
    g = torch.Generator().manual_seed(21)
    toks = ((torch.rand(N, generator=g) ** 4 * V).long().clamp_(0, V - 1).int() if zipf
            else torch.randint(0, V, (N,), generator=g, dtype=torch.int32))
    tables = [torch.randn(V, D, generator=g).to(dtype) for _ in range(3)]
    gouts = [torch.randn(N, D, generator=g).to(dtype) for _ in range(3)]
    want_out, want_g = O.value_embeds_fwd_bwd(toks, tables, gouts)
    mod = mot_b200.TokenValueEmbeddings(V, D).to(d).to(dtype)
    with torch.no_grad():
        for m, E in zip(mod.value_embeds, tables):
            m.weight.copy_(E)
    assert [n for n, _ in mod.named_parameters()] == [f"value_embeds.{i}.weight" for i in range(3)]
    mod.value_embeds[1].weight.requires_grad_(False)
    mot_b200.reset_launch_count()
    outs = mod(toks.to(d))
    torch.autograd.backward([outs[0], outs[2]], [gouts[0].to(d), gouts[2].to(d)])
    torch.cuda.synchronize()
    # 3 gathers + one plan (3 launches) + 2 x (scatter + finalize): the sort is not repeated per table
    assert mot_b200.launch_count() == 3 + 3 + 2 * 2
    for i in range(3):
        assert outs[i].dtype == dtype and tuple(outs[i].shape) == (N, D)
        assert torch.equal(outs[i].cpu(), want_out[i].to(dtype))            # a gather is exact
    assert mod.value_embeds[1].weight.grad is None
    for i in (0, 2):
        got = mod.value_embeds[i].weight.grad
        assert nerr(got, want_g[i]) <= TOL[dtype], f"table {i}: {nerr(got, want_g[i]):.3e}"
        untouched = torch.ones(V, dtype=torch.bool)
        untouched[toks.long()] = False
        if untouched.any():
            assert float(got[untouched.to(d)].abs().max()) == 0.0


def test_value_embeddings_full_size_vs_index_add():
    """65536 tokens over the GPT-2 vocabulary, 1024 columns (stream chunks of two batches): the dense gradient against
    torch's index_add_ in fp32 on the GPU (test-only)."""
    import mot_b200
    d = dev()
    g = torch.Generator(device="cuda").manual_seed(4)
    N, V, D = 65536, 50257, 1024
    toks = torch.randint(0, V, (N,), generator=g, device=d, dtype=torch.int32)
    E = torch.randn(V, D, generator=g, device=d).bfloat16().requires_grad_(True)
    gout = torch.randn(N, D, generator=g, device=d).bfloat16()
    (out,) = mot_b200.tok_gather(toks, E)
    out.backward(gout)
    assert torch.equal(out, E.detach()[toks.long()])
    want = torch.zeros(V, D, device=d).index_add_(0, toks.long(), gout.float())
    assert nerr(E.grad, want) <= 2.0 ** -8


def test_value_embeddings_2d_tokens_and_shape_errors():
    import mot_b200
    d = dev()
    mod = mot_b200.TokenValueEmbeddings(100, 64).to(d).bfloat16()
    toks = torch.randint(0, 100, (4, 33), device=d)
    outs = mod(toks)                                  # spt/train_gpt.py:600: toks_in is [B, S]
    assert len(outs) == 3 and tuple(outs[0].shape) == (4, 33, 64)
    assert torch.equal(outs[2], mod.value_embeds[2].weight[toks])
    with pytest.raises(NotImplementedError):
        mot_b200.tok_gather(toks, mod.value_embeds[0].weight, mod.value_embeds[1].weight[:, :32].contiguous())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mot_b200.tok_gather(toks.cpu(), mod.value_embeds[0].weight.cpu())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_split_residual_matches_oracle(dtype):
    """runs/71081:302-304,315: (x, x0t, x0b) and the gradients that flow back from all three uses."""
    import mot_b200
    d = dev()
    g = torch.Generator().manual_seed(5)
    V, Vb, bpt, bd, N = 400, 458, 16, 32, 777
    Dm = bpt * bd
    toks = torch.randint(0, V, (N,), generator=g, dtype=torch.int32)
    ids = torch.randint(0, Vb, (bpt, N), generator=g, dtype=torch.int32)
    E_tok = torch.randn(V, Dm, generator=g).to(dtype)
    E_byte = torch.randn(Vb, bd, generator=g).to(dtype)
    lam = torch.tensor([0.4, 0.7])            # [byte, token] like scalars[-2], scalars[-1]
    grads = [torch.randn(N, Dm, generator=g).to(dtype) for _ in range(3)]
    (wx, wt, wb), want = O.split_residual_fwd_bwd(toks, ids, E_tok, E_byte, lam[1], lam[0], grads, bpt=bpt)
    mod = mot_b200.MoTSplitResidualEmbedding(V, Vb, Dm, bd, bpt).to(d)
    mod.embed_tokens.to(dtype), mod.embed_bytes.to(dtype)
    with torch.no_grad():
        mod.embed_tokens.weight.copy_(E_tok); mod.embed_bytes.weight.copy_(E_byte); mod.lambdas.copy_(lam)
    x, x0t, x0b = mod(toks.to(d), ids.to(d))
    torch.autograd.backward([x, x0t, x0b], [gr.to(d).view(1, N, Dm) for gr in grads])
    torch.cuda.synchronize()
    tol = TOL[dtype]
    for got, ref, name in ((x, wx, "x"), (x0t, wt, "x0t"), (x0b, wb, "x0b")):
        assert tuple(got.shape) == (1, N, Dm)
        assert nerr(got[0], ref) <= tol, f"{name} {nerr(got[0], ref):.3e}"
    # three dense gradients rounded separately to the table dtype before autograd adds them (the reference's
    # autograd does the same with its bf16 grads): twice the single-rounding bar
    assert nerr(mod.embed_tokens.weight.grad, want["E_tok"]) <= 2 * tol
    assert nerr(mod.embed_bytes.weight.grad, want["E_byte"]) <= 2 * tol
    got_l = mod.lambdas.grad.cpu().double()
    ref_l = torch.stack([want["lam_byte"], want["lam_tok"]]).double()
    assert float((got_l - ref_l).abs().max() / ref_l.abs().max()) <= (1e-4 if dtype == torch.float32 else tol)


def test_mot_value_embeddings_match_oracle():
    """runs/9:252-254,311-313: ve_i = mixin_bytes(value_embeds_toks[i](tok), value_embeds_bytes[i](bytes), W_i) with the
    byte value tables sized by the token vocabulary."""
    import mot_b200
    d = dev()
    g = torch.Generator().manual_seed(9)
    V, Vb, bpt, Dt, bd, N = 600, 458, 16, 256, 32, 500
    toks = torch.randint(0, V, (N,), generator=g, dtype=torch.int32)
    ids = torch.randint(0, Vb, (1, N * bpt), generator=g, dtype=torch.int32)
    mod = mot_b200.MoTValueEmbeddings(V, Vb, Dt, bd, bpt).to(d).bfloat16()
    names = [n for n, _ in mod.named_parameters()]
    assert "value_embeds_toks.0.weight" in names and "value_embeds_bytes.2.weight" in names \
        and "value_byte_mixin_weights.1" in names
    assert tuple(mod.value_embeds_bytes[0].weight.shape) == (V, bd)
    gouts = [torch.randn(1, N, Dt, generator=g).bfloat16() for _ in range(3)]
    outs = mod(toks.to(d), ids.to(d))
    torch.autograd.backward(outs, [go.to(d) for go in gouts])
    torch.cuda.synchronize()
    spec = O.VARIANTS["V1"][0]
    for i in range(3):
        Et = mod.value_embeds_toks[i].weight.detach().cpu()
        Eb = mod.value_embeds_bytes[i].weight.detach().cpu()
        W = mod.value_byte_mixin_weights[i].detach().cpu()
        want_out, want = O.mot_embed_fwd_bwd(spec, toks, ids, Et, Eb, gouts[i], bpt=bpt, W=W)
        assert nerr(outs[i][0], want_out) <= 2.0 ** -6
        assert nerr(mod.value_embeds_toks[i].weight.grad, want["E_tok"]) <= 2.0 ** -6
        gb = mod.value_embeds_bytes[i].weight.grad
        assert tuple(gb.shape) == (V, bd) and float(gb[Vb:].abs().max()) == 0.0     # rows no byte id can reach
        assert nerr(gb, want["E_byte"]) <= 2.0 ** -6
        assert nerr(mod.value_byte_mixin_weights[i].grad, want["W"]) <= 2.0 ** -6
