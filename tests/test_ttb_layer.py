"""The product-side ttb table layer (mot_b200.ttb): host file/strings logic on CPU, table rows on the GPU.

Reference: modded-nanogpt/create_ttb.py:10-33, scaled-pre-train/data_creation.py:43-58.  The pin is the reference's one
checked-in table (tests/golden/ttb_8_left_pad.npz = embeddings/ttb_8_left_pad.json); mot_b200.ttb never imports the
oracle, the GPU tests use it only as the checker."""
import json
import os

import numpy as np
import pytest
import torch

from mot_b200 import ttb as T


def _golden(golden_dir):
    tab = np.load(os.path.join(golden_dir, "ttb_8_left_pad.npz"))["table"]
    b2i = json.load(open(os.path.join(golden_dir, "byte_to_int.json")))
    i2b = {v: k for k, v in b2i.items()}
    strings = ["".join(i2b[int(c)] for c in row[row != 456]) for row in tab]
    return tab, b2i, strings


# ---------------------------------------------------------------------------------------- host logic (CPU)
def test_json_round_trip_and_pack_rows(golden_dir, tmp_path):
    tab, _, _ = _golden(golden_dir)
    path = tmp_path / "ttb_8_left_pad.json"
    # the reference's file format: {"<id>": [ids]} (create_ttb.py:32)
    path.write_text(json.dumps({str(i): [int(x) for x in r] for i, r in enumerate(tab[:3000])}))
    rows = T.load_json(str(path))
    assert sorted(rows) == list(range(3000)) and all(isinstance(k, int) for k in rows)
    ids, present = T.pack_rows(rows, 3001)
    assert ids.dtype == np.int16 and ids.shape == (3001, 8)
    assert np.array_equal(ids[:3000], tab[:3000])
    assert present[:3000].all() and not present[3000] and np.all(ids[3000] == T.PAD_BYTE)
    with pytest.raises(ValueError):
        T.pack_rows({5: [1] * 8}, 4)
    with pytest.raises(ValueError):
        T.pack_rows({0: [1] * 8, 1: [1] * 7}, 4, bpt=8)
    assert T.load_byte_to_int(os.path.join(golden_dir, "byte_to_int.json"))["pad"] == T.PAD_BYTE


def test_strings_to_chars(golden_dir):
    tab, b2i, strings = _golden(golden_dir)
    n = 5000
    vocab = strings[:n] + ["<|endoftext|>"]
    chars, offs, is_eot = T.strings_to_chars(lambda i: vocab[i], b2i, n + 1)
    assert chars.dtype == np.int16 and offs.dtype == np.int32 and offs.shape == (n + 2,)
    assert is_eot.tolist() == [0] * n + [1] and offs[-1] == offs[-2] == len(chars)
    for v in (0, 1, 17, 4999):
        row = tab[v]
        assert chars[offs[v]:offs[v + 1]].tolist() == row[row != 456].tolist()
    with pytest.raises(KeyError):
        T.strings_to_chars(lambda i: "\x00\x01not-a-known-char\U0001F600", b2i, 1)


def test_refuses_cpu():
    with pytest.raises(RuntimeError):
        T.build_table(np.zeros(1, np.int16), np.array([0, 1], np.int32), None, 8, "left", device="cpu")
    with pytest.raises(RuntimeError):
        T.repad(torch.zeros(4, 8, dtype=torch.int16), 16)
    with pytest.raises(ValueError):
        T.build_table(np.zeros(1, np.int16), np.array([0, 1], np.int32), None, 8, "middle", device="cpu")


# ---------------------------------------------------------------------------------------- table rows (GPU)
@pytest.mark.gpu
def test_create_ttb_reproduces_checked_in_table(golden_dir):
    tab, b2i, strings = _golden(golden_dir)
    got = T.create_ttb(lambda i: strings[i], b2i, 50256, bpt=8, pad_position="left")
    assert got.dtype == torch.int16 and got.is_cuda
    assert np.array_equal(got.cpu().numpy(), tab)                       # bit-exact vs the reference's golden file
    full = T.with_eot_row(got)
    assert full.shape == (50257, 8) and bool((full[50256] == 457).all()) and np.array_equal(full[:50256].cpu().numpy(), tab)
    eot = T.create_ttb(lambda i: "<|endoftext|>" if i == 1 else "ab", b2i, 3, bpt=4)
    assert eot.cpu().tolist() == [[456, 456, b2i["a"], b2i["b"]], [457] * 4, [456, 456, b2i["a"], b2i["b"]]]


@pytest.mark.gpu
@pytest.mark.parametrize("bpt,side", [(8, "right"), (4, "left"), (4, "right"), (16, "left"), (16, "right"), (18, "left"),
                                      (20, "right"), (32, "left"), (1, "left")])
def test_build_and_repad_match_oracle(golden_dir, bpt, side):
    from oracle import mot_oracle as O
    tab, b2i, strings = _golden(golden_dir)
    want = O.ttb_dict_to_array(O.create_ttb(lambda i: strings[i], b2i, 50256, bpt=bpt, pad_position=side), 50257, bpt)
    built = T.with_eot_row(T.create_ttb(lambda i: strings[i], b2i, 50256, bpt=bpt, pad_position=side))
    assert np.array_equal(built.cpu().numpy(), want)
    src = T.with_eot_row(torch.from_numpy(tab).cuda())
    got = T.repad(src, bpt, side)
    assert np.array_equal(got.cpu().numpy(), O.ttb_repad(src.cpu().numpy(), bpt, side))
    assert np.array_equal(got.cpu().numpy(), want)                      # same strings: same table
    # left -> right -> left is the identity on rows that were not truncated
    back = T.repad(T.repad(src, 8, "right"), 8, "left")
    assert torch.equal(back, src)


@pytest.mark.gpu
def test_from_json_and_containers(golden_dir, tmp_path):
    from oracle import mot_oracle as O
    tab, _, _ = _golden(golden_dir)
    path = tmp_path / "ttb_8_left_pad.json"
    path.write_text(json.dumps({str(i): [int(x) for x in r] for i, r in enumerate(tab)}))
    table = T.from_json(str(path))
    assert table.shape == (50257, 8) and table.dtype == torch.int16
    assert np.array_equal(table[:50256].cpu().numpy(), tab) and bool((table[50256] == 457).all())
    T.save_json(table, str(tmp_path / "again.json"), n_rows=50256)
    assert T.load_json(str(tmp_path / "again.json")) == T.load_json(str(path))
    # float containers of the reference: fp32 is exact, bf16 rounds 101 ids (runs/7:441)
    f32 = T.as_container(table, torch.float32)
    assert torch.equal(T.from_container(f32), table)
    bf = T.as_container(table, torch.bfloat16)
    assert np.array_equal(T.from_container(bf).cpu().numpy(), O.bf16_round_ids(table.cpu().numpy()))
    assert int((T.from_container(bf)[50256] == 456).all()) == 1         # EOT 457 -> 456: the quirk
