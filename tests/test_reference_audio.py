"""The speech frontend on the REFERENCE'S OWN audio fixtures: the two FLEURS clips and the manifest its
speech tests read (/root/reference/tests/integration_tests/data/audio_files/audio_{1,2}.wav, audio_ref.tsv, used at
tests/integration_tests/test_sonar_speech_encoder.py:36-44 and test_sonar_speech_pipeline_models.py:19-22), copied to
tests/golden/reference_audio/.  Golden = torchaudio's Kaldi fbank of those clips (make_reference_audio_golden.py).
CPU tests pin the oracle and the WAV reader; GPU tests pin sb_fbank and drive the TSV-manifest pipeline end to end."""

import os
from pathlib import Path

import pytest
import torch

from oracle.speech_frontend import fbank, num_frames, standardize

DIR = Path(os.path.dirname(__file__)) / "golden" / "reference_audio"


def _golden():
    return torch.load(DIR / "reference_audio_fbank.pt", weights_only=True)


def test_wav_reader_and_oracle_fbank_on_the_reference_clips():
    from sonar_b200.inference_pipelines.speech import _read_wav, read_tsv_audio_paths

    g = _golden()
    paths = list(read_tsv_audio_paths(DIR / "audio_ref.tsv", 1))  # the reference passes audio_path_index=1
    assert paths == g["files"] == ["audio_1.wav", "audio_2.wav"]
    for name, n, raw in zip(g["files"], g["num_samples"], g["fbank_raw"]):
        w = _read_wav(DIR / name)
        assert w.shape == (1, n) and w.dtype == torch.float32 and float(w.abs().max()) <= 1.0  # [C=1, T] in [-1, 1]
        f = fbank(w[0])
        assert f.shape == raw.shape == (num_frames(n), 80)
        torch.testing.assert_close(f, raw, rtol=0, atol=1e-3)  # log-mel of real speech (values span ~[-2, 20]), fp32 FFT round-off
        s, m = torch.std_mean(raw, dim=0)
        torch.testing.assert_close(standardize(f), (raw - m) / s, rtol=0, atol=1e-3)


@pytest.mark.gpu
def test_fbank_kernel_on_the_reference_clips(native_lib, cuda_device):
    from sonar_b200.inference_pipelines.speech import _read_wav
    from sonar_b200.speech_frontend import WaveformToFbank

    g = _golden()
    waves = [_read_wav(DIR / n) for n in g["files"]]
    out, frames = WaveformToFbank(cuda_device)(waves)
    assert frames == [502, 478] and out.shape == (2, 502, 80)
    for i, raw in enumerate(g["fbank_raw"]):
        s, m = torch.std_mean(raw, dim=0)
        torch.testing.assert_close(out[i, : raw.shape[0]].cpu(), (raw - m) / s, rtol=0, atol=2e-3)
    assert float(out[1, 478:].abs().max()) == 0.0  # Collater(pad_value=0, pad_to_multiple=2)


@pytest.mark.gpu
def test_tsv_manifest_pipeline_on_the_reference_manifest(native_lib, cuda_device):
    """`SpeechToEmbeddingPipeline.build_pipeline(SpeechInferenceParams(...))` exactly as the reference test builds it
    (test_sonar_speech_encoder.py:36-46,68-78) on a 2-layer synthetic model: one batch of two embeddings, equal to the oracle
    speech encoder fed the torchaudio features of the same clips, and to `SpeechToEmbeddingModelPipeline.predict` on paths."""
    from oracle.speech_encoder import OracleSpeechConfig, OracleSpeechEncoder, make_synthetic_speech_state_dict
    from oracle.speech_frontend import collate_fbank
    from sonar_b200 import B200SpeechEncoderModel, sonar_speech_encoder_config
    from sonar_b200.inference_pipelines import (SpeechInferenceParams, SpeechToEmbeddingModelPipeline,
                                                SpeechToEmbeddingPipeline)

    ocfg = OracleSpeechConfig(num_layers=2, pooler_layers=2)
    sd = make_synthetic_speech_state_dict(ocfg, seed=3)
    model = B200SpeechEncoderModel(sonar_speech_encoder_config("english", num_encoder_layers=2, num_decoder_layers=2), sd,
                                   cuda_device)
    params = SpeechInferenceParams(data_file=DIR / "audio_ref.tsv", audio_root_dir=DIR, audio_path_index=1,
                                   target_lang="fra_Latn", batch_size=4, pad_idx=0, device=cuda_device,
                                   fbank_dtype=torch.float32, n_parallel=1)
    batches = list(SpeechToEmbeddingPipeline(model).build_pipeline(params))
    assert len(batches) == 1
    emb = batches[0].sentence_embeddings
    assert emb.shape == (2, 1024)
    g = _golden()
    feats = []
    for raw in g["fbank_raw"]:
        s, m = torch.std_mean(raw, dim=0)
        feats.append((raw - m) / s)
    fb, lens = collate_fbank(feats)
    ref, _, _ = OracleSpeechEncoder(ocfg, sd)(fb, lens)
    cos = torch.nn.functional.cosine_similarity(emb.cpu().double(), ref.double(), dim=1)
    rel = (emb.cpu().double() - ref.double()).norm(dim=1) / ref.double().norm(dim=1)
    assert float((1 - cos).max()) <= 1e-3 and float(rel.max()) <= 2e-2, (cos, rel)
    direct = SpeechToEmbeddingModelPipeline(model, device=cuda_device).predict(
        [str(DIR / "audio_1.wav"), str(DIR / "audio_2.wav")], batch_size=4)
    torch.testing.assert_close(direct, emb, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.skipif(not os.environ.get("SONAR_B200_CHECKPOINT_DIR"), reason="needs the real sonar_speech_encoder_eng checkpoint")
def test_real_checkpoint_reproduces_the_reference_speech_golden(native_lib, cuda_device):
    """Env-gated (SONAR_B200_CHECKPOINT_DIR/sonar_speech_encoder_eng.pt): the reference's own golden
    (tests/integration_tests/test_sonar_speech_encoder.py:68-78, default assert_close tolerances are fp32 1.3e-6/1e-5; the
    bf16 engine is held to 1 - cos <= 1e-3 per BASELINE.json north_star)."""
    from sonar_b200 import B200SpeechEncoderModel, sonar_speech_encoder_config
    from sonar_b200.inference_pipelines import SpeechInferenceParams, SpeechToEmbeddingPipeline

    ckpt = Path(os.environ["SONAR_B200_CHECKPOINT_DIR"]) / "sonar_speech_encoder_eng.pt"
    if not ckpt.exists():
        pytest.skip(f"{ckpt} not found")
    model = B200SpeechEncoderModel.from_checkpoint(ckpt, sonar_speech_encoder_config("english"), cuda_device)
    params = SpeechInferenceParams(data_file=DIR / "audio_ref.tsv", audio_root_dir=DIR, audio_path_index=1, batch_size=4,
                                   device=cuda_device)
    emb = next(iter(SpeechToEmbeddingPipeline(model).build_pipeline(params))).sentence_embeddings.cpu()
    want = torch.load(DIR / "speech_embedding.pt", weights_only=False).detach()
    cos = torch.nn.functional.cosine_similarity(emb.double(), want.double(), dim=1)
    assert float((1 - cos).max()) <= 1e-3, cos
