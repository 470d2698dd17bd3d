"""CPU: the speech-frontend oracle (oracle/speech_frontend.py) against the committed torchaudio Kaldi-fbank golden
(tests/golden/fbank_golden.pt, make_fbank_golden.py), and the host-built filterbank tables of the CUDA library."""

import os

import torch

from oracle.speech_frontend import (collate_fbank, fbank, mel_banks, num_frames, povey_window, standardize,
                                    waveform_to_fbank)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "fbank_golden.pt")


def test_oracle_fbank_matches_torchaudio_golden():
    g = torch.load(GOLDEN, weights_only=True)
    for w, raw, std in zip(g["waveforms"], g["fbank_raw"], g["fbank_standardized"]):
        f = fbank(w)
        assert f.shape == raw.shape == (num_frames(w.numel()), 80)
        torch.testing.assert_close(f, raw, rtol=0, atol=5e-4)          # log-mel, fp32 FFT round-off
        torch.testing.assert_close(standardize(f), std, rtol=0, atol=5e-4)


def test_frame_count_and_collate():
    assert num_frames(160000) == 998 and num_frames(400) == 1 and num_frames(399) == 0  # 10 s -> 998
    feats = [waveform_to_fbank(torch.randn(3000) * 0.1), waveform_to_fbank(torch.randn(1760) * 0.1)]
    batch, lens = collate_fbank(feats)
    assert lens == [17, 9] and batch.shape == (2, 18, 80)  # pad_to_multiple=2
    assert float(batch[0, 17].abs().max()) == 0.0 and float(batch[1, 9:].abs().max()) == 0.0


def test_native_tables_match_oracle(native_lib):
    n = native_lib.sb_fbank_tables_bytes()
    buf = torch.empty(n, dtype=torch.uint8)
    assert native_lib.sb_fbank_build_tables(buf.data_ptr()) == 0
    f, ints = buf.view(torch.float32), buf.view(torch.int32)
    torch.testing.assert_close(f[:400], povey_window(), rtol=0, atol=1e-7)
    off = 400 + 2 * 128 + 2 * 256
    lo, ln = ints[off : off + 80], ints[off + 80 : off + 160]
    w = f[off + 160 : off + 160 + 80 * 64].view(80, 64)
    mb = mel_banks()
    for m in range(80):
        dense = torch.zeros(257)
        dense[lo[m] : lo[m] + ln[m]] = w[m, : ln[m]]
        torch.testing.assert_close(dense, mb[m], rtol=0, atol=1e-7)


def test_oracle_conformer_block_matches_hf_golden():
    from oracle.speech_encoder import OracleSpeechConfig, OracleSpeechEncoder

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "conformer_layer_small.pt"), weights_only=True)
    enc = OracleSpeechEncoder(OracleSpeechConfig(**g["config"]), g["state_dict"])
    s = g["x"].shape[1]
    ok = torch.arange(s)[None, :] < g["lens"][:, None]
    out = enc.conformer_block(0, g["x"], ok)
    torch.testing.assert_close(out * ok[:, :, None], g["out"] * ok[:, :, None], rtol=1e-5, atol=1e-5)


def test_oracle_frontend_and_block_stack_match_hf_conformer_encoder_golden():
    """The whole stack before the pooler -- 2-frame stacking, LayerNorm(160) + projection, two Conformer blocks, the final
    LayerNorm -- against HuggingFace's SeamlessM4T (w2v-BERT) feature projection + Conformer encoder on shared random
    weights (golden made by tests/golden/make_conformer_encoder_golden.py), ragged batch, valid positions compared."""
    from oracle.speech_encoder import OracleSpeechConfig, OracleSpeechEncoder

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "conformer_encoder_small.pt"), weights_only=True)
    enc = OracleSpeechEncoder(OracleSpeechConfig(**g["config"]), g["state_dict"])
    out, ok, lens = enc.encode(g["fbank"], g["frame_lens"])
    assert lens == [n // 2 for n in g["frame_lens"]]
    torch.testing.assert_close(out * ok[:, :, None], g["out"] * ok[:, :, None], rtol=1e-5, atol=2e-5)


def test_oracle_speech_padding_invariance():
    """An utterance's embedding must not depend on its batch neighbours / padding (key masks, zeroed conv inputs)."""
    from oracle.speech_encoder import OracleSpeechConfig, OracleSpeechEncoder, make_synthetic_speech_state_dict

    cfg = OracleSpeechConfig(model_dim=64, num_layers=2, num_heads=4, ffn_inner_dim=128, pooler_layers=2, pooler_heads=4,
                             pooler_ffn_inner_dim=128, pooler_vocab=64)
    enc = OracleSpeechEncoder(cfg, make_synthetic_speech_state_dict(cfg, seed=1, std=0.1))
    fb = torch.randn(2, 24, 80)
    fb[1, 14:] = 0
    both, _, lens = enc(fb, [24, 14])
    alone, _, _ = enc(fb[1:, :14], [14])
    assert lens == [12, 7]
    torch.testing.assert_close(both[1], alone[0], rtol=1e-5, atol=1e-5)


def test_oracle_pooler_layers_match_hf_bart_post_ln_golden():
    """The pooler's POST-LN decoder layers against an independent implementation (HuggingFace BartDecoderLayer; golden made
    by tests/golden/make_pooler_golden.py): self-attention over the single query, cross-attention with a key-padding mask
    (one utterance keeps a single valid key), ReLU FFN."""
    from oracle.speech_encoder import OracleSpeechConfig, OracleSpeechEncoder

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pooler_layers_small.pt"), weights_only=True)
    enc = OracleSpeechEncoder(OracleSpeechConfig(**g["config"]), g["state_dict"])
    s = g["enc"].shape[1]
    key_ok = torch.arange(s)[None, :] < g["lens"][:, None]
    out = enc.pooler_layers(g["x0"], g["enc"], key_ok)
    torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=1e-5)
