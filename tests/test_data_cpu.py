"""CPU checks of the data side (mot_b200.data): the .bin shard format of modded-nanogpt/data/fineweb.py:28-50 /
spt/train_gpt.py:628-648 (256 x int32 header, magic 20240520, version 1, uint16 tokens), the reference's assertions on
bad files, and the window arithmetic of distributed_data_generator (:788-806) without touching a device."""
import numpy as np
import pytest
import torch

import mot_b200
from mot_b200 import data as D


def test_shard_round_trip_and_header_checks(tmp_path):
    toks = np.random.default_rng(0).integers(0, 50257, size=1000).astype(np.uint16)
    p = tmp_path / "fineweb_train_000001.bin"
    D.write_data_shard(str(p), toks)
    raw = np.fromfile(p, dtype=np.int32, count=256)
    assert raw[0] == 20240520 and raw[1] == 1 and raw[2] == 1000 and p.stat().st_size == 256 * 4 + 2 * 1000
    got = D.load_data_shard(str(p), pin_memory=False)
    assert got.dtype == torch.uint16 and np.array_equal(got.numpy(), toks)
    # `bytes/` shards hold int32 (spt/train_gpt.py:645)
    (tmp_path / "bytes").mkdir()
    pb = tmp_path / "bytes" / "x.bin"
    D.write_data_shard(str(pb), np.arange(70, dtype=np.int32))
    gb = D.load_data_shard(str(pb), pin_memory=False)
    assert gb.dtype == torch.int32 and gb.tolist() == list(range(70))
    # the reference's assertions
    bad = tmp_path / "bad_magic.bin"
    h = np.zeros(256, dtype=np.int32); h[0], h[1], h[2] = 123, 1, 4
    bad.write_bytes(h.tobytes() + np.zeros(4, dtype=np.uint16).tobytes())
    with pytest.raises(AssertionError, match="magic number mismatch"):
        D.load_data_shard(str(bad), pin_memory=False)
    h[0], h[1] = 20240520, 2
    bad.write_bytes(h.tobytes() + np.zeros(4, dtype=np.uint16).tobytes())
    with pytest.raises(AssertionError, match="unsupported version"):
        D.load_data_shard(str(bad), pin_memory=False)
    h[1], h[2] = 1, 10                       # header claims more tokens than the file holds
    bad.write_bytes(h.tobytes() + np.zeros(4, dtype=np.uint16).tobytes())
    with pytest.raises(AssertionError, match="does not match header"):
        D.load_data_shard(str(bad), pin_memory=False)


def test_generator_refuses_cpu_and_invalid_combinations(tmp_path):
    D.write_data_shard(str(tmp_path / "a.bin"), np.arange(5000, dtype=np.uint16))
    gen = D.distributed_data_generator(str(tmp_path / "*.bin"), seq_len=15, batch_size=4, rank=0, world_size=2,
                                       bytes_per_token=8, byte_in=False, pull_in=False, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        next(gen)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        D.create_batch(torch.zeros(2, 5, dtype=torch.int32), None, None, 8, byte_in=False, pull_in=False)
    assert mot_b200.data.SHARD_MAGIC == 20240520
