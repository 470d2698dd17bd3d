"""GPU parity of the speech path against the CPU oracle (BASELINE.json config 3)."""

import os

import pytest
import torch

from oracle.speech_frontend import collate_fbank, waveform_to_fbank

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fbank_golden.pt")


def test_fbank_kernel_matches_oracle_and_golden(native_lib, cuda_device):
    from sonar_b200.speech_frontend import WaveformToFbank

    conv = WaveformToFbank(cuda_device)
    g = torch.load(GOLDEN, weights_only=True)
    gen = torch.Generator().manual_seed(3)
    waves = list(g["waveforms"]) + [(torch.randn(160000, generator=gen) * 0.05).clamp(-1, 1),   # config-3 shape: 10 s
                                    (torch.randn(16000 * 3 + 77, generator=gen) * 0.3).clamp(-1, 1),
                                    torch.sin(torch.arange(400 + 160 * 5) * 0.05) * 0.5]
    out, frames = conv([w.to(cuda_device) for w in waves])
    torch.cuda.synchronize()
    ref, ref_lens = collate_fbank([waveform_to_fbank(w) for w in waves])
    assert frames == ref_lens and out.shape == ref.shape
    # standardised log-mel, fp32 FFT/mel round-off on both sides (the oracle itself sits 1e-4 from torchaudio)
    torch.testing.assert_close(out.cpu(), ref, rtol=0, atol=2e-3)
    for i, std in enumerate(g["fbank_standardized"]):  # and against the torchaudio golden directly
        torch.testing.assert_close(out[i, : std.shape[0]].cpu(), std, rtol=0, atol=2e-3)
    assert frames[2] == 998


def test_fbank_rejects_bad_input(native_lib, cuda_device):
    from sonar_b200.speech_frontend import WaveformToFbank

    conv = WaveformToFbank(cuda_device)
    with pytest.raises(ValueError):
        conv([torch.zeros(100, device=cuda_device)])
    with pytest.raises(ValueError):
        conv([torch.zeros((2, 1000), device=cuda_device)])
    with pytest.raises(RuntimeError):
        WaveformToFbank("cpu")
