"""CPU tests that pin the oracle (oracle/text_encoder.py).

* against the committed HuggingFace M2M100Encoder golden (independent implementation
  the reference's notebook uses as the SONAR encoder; tests/golden/make_m2m100_golden.py);
* against the reference's pooling known-answer tests
  (/root/reference/tests/unit_tests/test_sonar_pooling.py:16-68, values restated here).
"""

import os

import pytest
import torch
from torch.testing import assert_close

from oracle.text_encoder import (OracleEncoderConfig, OracleTextEncoder, encoder_flops, make_synthetic_state_dict,
                                 sinusoidal_table, static_pooling)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "m2m100_small.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=True)


def test_oracle_matches_m2m100_golden(golden):
    cfg = OracleEncoderConfig(**golden["config"])
    enc = OracleTextEncoder(cfg, golden["state_dict"])
    emb, x = enc(golden["ids"], golden["seq_lens"])
    s = golden["ids"].shape[1]
    valid = (torch.arange(s)[None] < golden["seq_lens"][:, None])[:, :, None]
    assert_close(emb, golden["sentence_embeddings"], rtol=1e-5, atol=1e-5)
    assert_close(x * valid, golden["encoded_seqs"] * valid, rtol=1e-5, atol=1e-5)


def test_oracle_pad_value_is_irrelevant(golden):
    cfg = OracleEncoderConfig(**golden["config"])
    enc = OracleTextEncoder(cfg, golden["state_dict"])
    ids2 = golden["ids"].clone()
    s = ids2.shape[1]
    pad = ~(torch.arange(s)[None] < golden["seq_lens"][:, None])
    ids2[pad] = 7
    a, _ = enc(golden["ids"], golden["seq_lens"])
    b, _ = enc(ids2, golden["seq_lens"])
    assert_close(a, b, rtol=0, atol=0)


def test_oracle_batch_composition_invariance(golden):
    """Property the reference pins at tests/integration_tests/test_text_sonar.py:120-161."""
    cfg = OracleEncoderConfig(**golden["config"])
    enc = OracleTextEncoder(cfg, golden["state_dict"])
    full, _ = enc(golden["ids"], golden["seq_lens"])
    for i, n in enumerate(golden["seq_lens"].tolist()):
        one, _ = enc(golden["ids"][i : i + 1, :n], None)
        assert_close(one[0], full[i], rtol=1.3e-6, atol=1e-5)


def test_position_table_offset():
    t = sinusoidal_table(6, 8, legacy_pad_idx=1)
    # row 0 encodes position index 2 (pad_idx + 1), [sin | cos] halves
    assert_close(t[0, 0], torch.sin(torch.tensor(2.0)))
    assert_close(t[0, 4], torch.cos(torch.tensor(2.0)))
    assert_close(t[3, 0], torch.sin(torch.tensor(5.0)))


# ---- reference pooling KATs (test_sonar_pooling.py) ----
SEQS = torch.tensor([[[7, 2], [3, 4], [10, 20]], [[-1, -2], [100, 1000], [-10, -20]]], dtype=torch.float32)
LENS = torch.tensor([2, 1])


@pytest.mark.parametrize("mode,expected", [
    ("max", [[7.0, 4.0], [-1.0, -2.0]]),
    ("mean", [[5.0, 3.0], [-1.0, -2.0]]),
    ("last", [[3.0, 4.0], [-1.0, -2.0]]),
])
def test_pooling_kat_with_mask(mode, expected):
    exp = torch.tensor(expected)
    assert_close(static_pooling(SEQS, LENS, mode), exp)
    assert_close(static_pooling(SEQS.unsqueeze(3), LENS, mode), exp.unsqueeze(2))


def test_pooling_kat_no_mask():
    seqs = torch.tensor([[[7, 2], [3, 2], [2, 20]], [[-1, -3], [-4, 2], [-7, -2]]], dtype=torch.float32)
    assert_close(static_pooling(seqs, None, "last"), torch.tensor([[2.0, 20], [-7, -2]]))
    assert_close(static_pooling(seqs, None, "max"), torch.tensor([[7.0, 20], [-1, 2]]))
    assert_close(static_pooling(seqs, None, "mean"), torch.tensor([[4.0, 8], [-4, -1]]))


def test_flop_model():
    assert abs(encoder_flops(128) - 130.46e9) / 130.46e9 < 1e-3
    assert abs(encoder_flops(64) - 64.83e9) / 64.83e9 < 1e-3


def test_synthetic_weights_deterministic():
    cfg = OracleEncoderConfig(model_dim=64, vocab_size=50, num_layers=1, num_heads=1, ffn_inner_dim=128)
    a = make_synthetic_state_dict(cfg, seed=1)
    b = make_synthetic_state_dict(cfg, seed=1)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert set(a) >= {"encoder_frontend.embed.weight", "encoder.layers.0.self_attn.q_proj.weight", "layer_norm.bias"}
