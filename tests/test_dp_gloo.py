"""Host-side data-parallel logic of the path on CPU: world_size 2 over gloo (the N > 1 path of bench.py and of the
modules uses the same GradBucket / broadcast / sharding code with NCCL)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mot_b200 import dp
        torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT tables
        E_tok = torch.nn.Parameter(torch.randn(50, 16))
        E_byte = torch.nn.Parameter(torch.randn(458, 8))
        dp.broadcast_params([E_tok, E_byte], src=0)        # C1: replicate from rank 0
        ref = torch.Generator().manual_seed(100)
        want_tok = torch.randn(50, 16, generator=ref)
        assert torch.equal(E_tok.detach(), want_tok)
        # sharding: rank r owns its slice of the global batch (runs/7:468-474)
        stream = torch.arange(1000)
        mine = dp.shard_for_rank(stream, pos=100, local=64, rank=rank)
        assert mine[0].item() == 100 + rank * 64 and mine.numel() == 64
        # one flat bucket, per-parameter views, one collective
        # "auto": the own exchange kernels need CUDA + symmetric memory; on CPU / gloo the bucket is ordinary memory
        assert dp.pick_algo(2, False) == "p2p" and dp.pick_algo(8, True) == "nvls" and dp.pick_algo(4, True) == "nvls"
        assert dp.pick_algo(4, False) == "p2p" and dp.pick_algo(3, False) == "nccl" and dp.pick_algo(6, True) == "nvls"
        bucket = dp.GradBucket([E_tok, E_byte], symmetric="auto")
        assert bucket._symm is None and bucket.algo == "nccl" and not bucket.pipelined
        v_tok, v_byte = bucket.views()
        assert v_tok.shape == E_tok.shape and v_byte.shape == E_byte.shape
        assert v_tok.data_ptr() == bucket.flat.data_ptr() and bucket.offsets[1] % 4 == 0
        v_tok.fill_(float(rank + 1))                       # what the backward kernels would have written
        v_byte.copy_(torch.full_like(v_byte, 10.0 * (rank + 1)))
        bucket.attach()
        assert E_tok.grad.data_ptr() == v_tok.data_ptr()
        bucket.all_reduce_avg()
        bucket.wait()                                      # nothing pending: a no-op
        with pytest.raises(RuntimeError):
            bucket.exchange_async(0, 16, True)             # the pipelined exchange exists only over symmetric memory
        assert torch.allclose(E_tok.grad, torch.full_like(E_tok, 1.5))      # mean of 1, 2
        assert torch.allclose(E_byte.grad, torch.full_like(E_byte, 15.0))   # mean of 10, 20
        # a gradient produced outside the bucket is copied in by attach()
        E_tok.grad = torch.full_like(E_tok, 7.0)
        bucket.attach()
        assert float(bucket.flat[0]) == 7.0 and E_tok.grad.data_ptr() == v_tok.data_ptr()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_grad_bucket_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_grad_bucket_single_process_layout():
    sys.path.insert(0, os.path.join(ROOT, "mixture-of-tokenizers_b200"))
    from mot_b200 import dp
    a = torch.nn.Parameter(torch.zeros(3, 5, dtype=torch.bfloat16))     # 15 elements -> padded to 16
    b = torch.nn.Parameter(torch.zeros(458, 8, dtype=torch.bfloat16))
    bk = dp.GradBucket([a, b])
    assert bk.offsets == [0, 16] and bk.flat.numel() == 16 + 458 * 8
    assert bk.view_of(b).data_ptr() == bk.flat.data_ptr() + 32
    assert bk.all_reduce_avg() is None      # no process group: nothing to do
    c = torch.nn.Parameter(torch.zeros(4, 4))                           # fp32 master weight in a bf16 bucket: refused
    bk2 = dp.GradBucket([c], dtype=torch.bfloat16)
    c.grad = torch.ones(4, 4)
    with pytest.raises(TypeError):
        bk2.attach()


def test_pick_algo_policy_and_overrides(monkeypatch):
    """Which exchange runs the bucket (mot_b200/dp.py:pick_algo; measured in profiles/r2_dp.md): peer-to-peer two-shot at 2
    ranks, the in-switch reduction from 3 ranks where the switch offers multicast, peer-to-peer at 4 / 8 without it, NCCL
    for everything else; MOT_DP_ALGO / MOT_DP_NCCL force one and never select an exchange the box cannot run."""
    from mot_b200 import dp
    for k in ("MOT_DP_ALGO", "MOT_DP_NCCL"):
        monkeypatch.delenv(k, raising=False)
    assert dp.pick_algo(2, True) == "p2p" and dp.pick_algo(2, False) == "p2p"
    assert [dp.pick_algo(n, True) for n in (3, 4, 8, 16)] == ["nvls"] * 4
    assert dp.pick_algo(4, False) == "p2p" and dp.pick_algo(8, False) == "p2p"
    assert dp.pick_algo(3, False) == "nccl" and dp.pick_algo(6, False) == "nccl"
    monkeypatch.setenv("MOT_DP_ALGO", "nvls")
    assert dp.pick_algo(8, True) == "nvls" and dp.pick_algo(8, False) == "nccl"     # no multicast: NCCL, not a broken path
    monkeypatch.setenv("MOT_DP_ALGO", "p2p")
    assert dp.pick_algo(4, True) == "p2p" and dp.pick_algo(3, True) == "nccl"       # P2P kernels exist for 2 / 4 / 8 ranks
    monkeypatch.setenv("MOT_DP_NCCL", "1")
    assert dp.pick_algo(8, True) == "nccl"
    assert dp.default_slabs(2) == 1 and dp.default_slabs(8) == 1                   # the slab pipeline is opt-in


def test_grad_bucket_without_process_group_is_plain_memory():
    """No initialised process group: the bucket is ordinary memory, the exchange a no-op, the pipeline refused."""
    import torch
    from mot_b200 import dp
    ps = [torch.nn.Parameter(torch.zeros(10, 8)), torch.nn.Parameter(torch.zeros(3, 8))]
    b = dp.GradBucket(ps, symmetric="auto")
    assert b.algo == "nccl" and not b.pipelined and not b.sparse_rows and b.n_slabs == 1
    assert b.all_reduce_avg() is None
    with pytest.raises(RuntimeError):
        b.exchange_async(0, 8, last=True)
    b.wait()                                                                       # nothing pending: returns
