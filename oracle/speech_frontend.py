"""CPU oracle for the speech path's feature frontend (BASELINE.json config 3).
TEST INFRASTRUCTURE ONLY.

Restates fairseq2n ``WaveformToFbankConverter(num_mel_bins=80, waveform_scale=2**15, channel_last=True,
standardize=True)`` as the SONAR speech pipelines configure it (``sonar/inference_pipelines/speech.py:120-127,
283-290``) = kaldi-native-fbank defaults [fs2]: 25 ms / 10 ms frames, snip_edges, DC
removal, pre-emphasis 0.97, Povey window, 512-point FFT, power spectrum, 80 Kaldi-mel triangles 20 Hz..Nyquist,
log with FLT_EPSILON floor, no dither; then per-utterance standardisation over time with the unbiased std; then
``Collater(pad_value=0, pad_to_multiple=2)`` (``speech.py:139,384,444``).

Pinned against ``torchaudio.compliance.kaldi.fbank`` (an independent implementation of the Kaldi recipe, present
in this image) through the committed fixture ``tests/golden/fbank_golden.pt``
(``tests/golden/make_fbank_golden.py``).  Against fairseq2n itself: parity unpinned (not installable; the reference's
golden ``tests/integration_tests/data/speech_embedding.pt`` needs the downloaded speech checkpoint).
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

SAMPLE_RATE = 16000
FRAME_LEN = 400      # 25 ms
FRAME_SHIFT = 160    # 10 ms
NFFT = 512
NUM_MEL = 80
LOW_FREQ = 20.0
PREEMPH = 0.97
WAVEFORM_SCALE = 2.0 ** 15
EPS = 1.1920928955078125e-07  # FLT_EPSILON


def num_frames(num_samples: int) -> int:
    return 0 if num_samples < FRAME_LEN else 1 + (num_samples - FRAME_LEN) // FRAME_SHIFT


def povey_window() -> Tensor:
    n = torch.arange(FRAME_LEN, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2 * math.pi * n / (FRAME_LEN - 1))).pow(0.85).float()


def mel_banks() -> Tensor:
    """[80, 257] triangular weights on the Kaldi mel scale (last FFT bin gets weight 0 like Kaldi)."""
    def mel(f):
        return 1127.0 * torch.log(1.0 + f / 700.0)

    nyq = SAMPLE_RATE / 2
    lo, hi = mel(torch.tensor(LOW_FREQ, dtype=torch.float64)), mel(torch.tensor(nyq, dtype=torch.float64))
    delta = (hi - lo) / (NUM_MEL + 1)
    b = torch.arange(NUM_MEL, dtype=torch.float64)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
    m = mel(torch.arange(NFFT // 2, dtype=torch.float64) * (SAMPLE_RATE / NFFT))[None, :]
    w = torch.clamp(torch.minimum((m - left) / (center - left), (right - m) / (right - center)), min=0.0)
    return torch.nn.functional.pad(w, (0, 1)).float()


def fbank(waveform: Tensor) -> Tensor:
    """waveform float [T] in [-1, 1] at 16 kHz -> log-mel [frames, 80] (not yet standardised)."""
    x = waveform.float() * WAVEFORM_SCALE
    m = num_frames(x.numel())
    if m == 0:
        return torch.empty((0, NUM_MEL))
    frames = x.unfold(0, FRAME_LEN, FRAME_SHIFT)[:m].clone()
    frames = frames - frames.mean(dim=1, keepdim=True)
    prev = torch.cat([frames[:, :1], frames[:, :-1]], dim=1)
    frames = (frames - PREEMPH * prev) * povey_window()[None]
    spec = torch.fft.rfft(torch.nn.functional.pad(frames, (0, NFFT - FRAME_LEN))).abs().pow(2.0)
    return torch.clamp(spec @ mel_banks().T, min=EPS).log()


def standardize(feat: Tensor) -> Tensor:
    """Per-utterance, per-bin (f - mean) / std over time, unbiased std (App. B.1)."""
    std, mean = torch.std_mean(feat, dim=0)
    return (feat - mean) / std


def waveform_to_fbank(waveform: Tensor) -> Tensor:
    return standardize(fbank(waveform))


def collate_fbank(feats: Sequence[Tensor], pad_to_multiple: int = 2) -> Tuple[Tensor, List[int]]:
    """``Collater(pad_value=0, pad_to_multiple=2)``: [B, Tmax, 80] zero-padded, true frame counts."""
    lens = [int(f.shape[0]) for f in feats]
    tmax = max(lens)
    tmax = (tmax + pad_to_multiple - 1) // pad_to_multiple * pad_to_multiple
    out = torch.zeros((len(feats), tmax, feats[0].shape[1]))
    for i, f in enumerate(feats):
        out[i, : lens[i]] = f
    return out, lens
