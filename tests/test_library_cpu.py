"""torch.library registration of the operators (mot_b200/_library.py), checked without a GPU: schemas, fake (meta)
implementations and Dynamo traceability.  The reference wraps its models in torch.compile(model, dynamic=False)
(scaled-pre-train/train_gpt.py:1195; runs/7:623): a module swap must not introduce graph breaks.  Values and the
autograd formulas are checked on the GPU (tests/test_gpu_compile.py): the autograd engine needs a CUDA context even
for fake CUDA tensors."""
import pytest
import torch
from torch import nn
from torch._subclasses.fake_tensor import FakeTensorMode

import mot_b200


@pytest.fixture
def force_ops():
    mot_b200.set_custom_ops(True)
    yield
    mot_b200.set_custom_ops(None)


def test_every_operator_is_registered_with_a_fake_impl():
    import torch._library.utils as lu  # noqa: F401
    names = ["embed", "embed_bwd", "tok_gather", "tok_gather_bwd", "embed_proj", "embed_proj_bwd", "embed_byte_fc",
             "embed_byte_fc_bwd", "ttb_expand", "pull", "tokens_to_digits", "mixout_copy", "mixout_copy_bwd"]
    for n in names:
        op = getattr(torch.ops.mot_b200, n).default
        assert op._schema.name == f"mot_b200::{n}"
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "Meta") or \
            torch._library.simple_registry.singleton.find(op._name).fake_impl.kernel is not None


def test_forward_shapes_through_the_operators_in_fake_mode(force_ops):
    with FakeTensorMode():
        d = "cuda"
        N, V, bpt, bd = 256, 1000, 16, 32
        Dt = bpt * bd
        tok = torch.empty(N, dtype=torch.int32, device=d)
        ids = torch.empty(bpt, N, dtype=torch.int32, device=d)
        Et = torch.empty(V, Dt, dtype=torch.bfloat16, device=d, requires_grad=True)
        Eb = torch.empty(458, bd, dtype=torch.bfloat16, device=d, requires_grad=True)
        out = mot_b200.mot_embed(tok, ids, Et, Eb, mot_b200.MixSpec(combine="add", slot_major=True), bpt=bpt)
        assert out.shape == (N, Dt) and out.dtype == torch.bfloat16 and "mot_b200_embed" in type(out.grad_fn).__name__
        o3 = torch.ops.mot_b200.embed(tok, ids, None, Et, Eb, None, mot_b200._library.pack_spec(
            mot_b200.MixSpec(combine="add", slot_major=True)), bpt, 0, mot_b200.FP32_EPS, True)
        assert o3[1].shape == (N,) and o3[1].dtype == torch.float32      # rstd kept for the saved-output backward
        assert o3[2].dtype == torch.uint8 and o3[2].numel() > N * 8      # the workspace travels as a tensor
        with torch.no_grad():
            out = mot_b200.mot_embed(tok, ids, Et, Eb, mot_b200.MixSpec(combine="add", slot_major=True), bpt=bpt)
            assert out.grad_fn is None
        W = torch.empty(128, Dt + bpt * bd, dtype=torch.float32, device=d, requires_grad=True)
        idt = torch.empty(1, N * bpt, dtype=torch.int64, device=d)
        spec = mot_b200.MixSpec(combine="concat", tok_norm=True, byte_norm=True)
        o = mot_b200.mot_embed_proj(tok, idt, Et, Eb, W, spec, bpt=bpt)
        assert o.shape == (N, 128) and o.dtype == torch.bfloat16
        o = mot_b200.mot_embed_proj(tok, idt, Et, Eb, W, spec, bpt=bpt, byte_ids2=idt)
        assert o.shape == (N, 128)
        Wf = torch.empty(Dt, Dt, dtype=torch.bfloat16, device=d, requires_grad=True)
        assert mot_b200.mot_embed_byte_fc(tok, ids, Et, Eb, Wf, bpt=bpt).shape == (N, Dt)
        T3 = [torch.empty(V, 64, dtype=torch.bfloat16, device=d, requires_grad=True) for _ in range(3)]
        outs = mot_b200.tok_gather(tok, *T3)
        assert len(outs) == 3 and all(o.shape == (N, 64) for o in outs)
        tab = torch.empty(V, bpt, dtype=torch.int16, device=d)
        b = mot_b200.ttb_expand(tok, tab)
        assert b.shape == (1, N * bpt) and b.dtype == torch.int64
        assert mot_b200.pull_from_left(b, bpt).shape == b.shape and mot_b200.pull_from_right(b, bpt).dtype == torch.int64
        x = torch.empty(2, 5, 64, dtype=torch.bfloat16, device=d, requires_grad=True)
        assert mot_b200.mixout_copy(x, 4).shape == (2, 20, 64) and mot_b200.mixout_split(x, 4).shape == (2, 20, 16)
        assert mot_b200.tokens_to_digits(tok, 4, 10000, 10001, 10002).shape == (N * 4,)


class _Tiny(nn.Module):
    """A model front like the reference's GPT.forward: the embedding modules, then something trainable on top."""

    def __init__(self):
        super().__init__()
        self.emb = mot_b200.MoTEmbedding(1000, 458, 512, 32, 16, variant="V3")
        self.spt = mot_b200.SptByteMixEmbedding(1000, 458, 64, 16, 128, bytes_per_token=16)
        self.head = nn.Linear(512, 8)

    def forward(self, tok, ids, tok2, ids2):
        x = self.emb(tok, ids)
        y = self.spt(tok2, None, ids2)
        return self.head(x.float()).sum() + y.float().sum()


def test_dynamo_traces_the_modules_without_graph_breaks():
    """Default dispatch rule: while Dynamo traces, the public functions route through torch.ops.mot_b200.* (opaque
    nodes); torch._dynamo.export raises on any graph break."""
    with FakeTensorMode():
        with torch.device("cuda"):
            m = _Tiny()
        for mod in (m.emb.embed_tokens, m.emb.embed_bytes, m.spt.embed.embed_tokens, m.spt.embed.embed_bytes):
            mod.weight = nn.Parameter(torch.empty(mod.weight.shape, dtype=torch.bfloat16, device="cuda"))
        tok = torch.empty(256, dtype=torch.int32, device="cuda")
        ids = torch.empty(16, 256, dtype=torch.int32, device="cuda")
        tok2 = torch.empty(2, 128, dtype=torch.int32, device="cuda")
        ids2 = torch.empty(2, 128 * 16, dtype=torch.int64, device="cuda")
        exp = torch._dynamo.export(m, aten_graph=False)(tok, ids, tok2, ids2)
    targets = [str(n.target) for n in exp.graph_module.graph.nodes if n.op == "call_function"]
    assert any("mot_b200.embed_proj" in t for t in targets) and any(t.endswith("mot_b200.embed')>") or "mot_b200.embed" in t for t in targets)


def test_spec_round_trip():
    from mot_b200._library import pack_spec, unpack_spec
    for kw in mot_b200.RUN_VARIANTS.values():
        s = mot_b200.MixSpec(**kw)
        assert unpack_spec(pack_spec(s), s.eps) == s
    s = mot_b200.MixSpec(combine="concat", bytes_first=True, out_norm=False, ttb_scramble=True, eps=1e-3)
    assert unpack_spec(pack_spec(s), 1e-3) == s
