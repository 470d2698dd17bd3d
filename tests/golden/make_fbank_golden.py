"""Generate tests/golden/fbank_golden.pt with torchaudio's Kaldi-compatible fbank (independent implementation):

    python tests/golden/make_fbank_golden.py

Two synthetic 16 kHz utterances (noise + tones; 0.53 s and 0.31 s), features = kaldi.fbank(w * 2**15,
num_mel_bins=80) with torchaudio defaults (dither 0, povey, 25/10 ms, snip_edges) + std_mean standardisation,
i.e. what fairseq2n's WaveformToFbankConverter(standardize=True) is documented to produce."""

import os

import torch
import torchaudio.compliance.kaldi as kaldi


def main() -> None:
    g = torch.Generator().manual_seed(99)
    waves = []
    for n, f0 in ((8480, 220.0), (5000, 1333.0)):
        t = torch.arange(n) / 16000.0
        w = 0.05 * torch.randn(n, generator=g) + 0.2 * torch.sin(2 * torch.pi * f0 * t) + 0.1 * torch.sin(2 * torch.pi * 3.1 * f0 * t)
        waves.append(w.clamp(-1, 1))
    raw, std = [], []
    for w in waves:
        f = kaldi.fbank(w[None] * 2 ** 15, num_mel_bins=80, sample_frequency=16000.0)
        raw.append(f)
        s, m = torch.std_mean(f, dim=0)
        std.append((f - m) / s)
    here = os.path.dirname(os.path.abspath(__file__))
    torch.save({"waveforms": waves, "fbank_raw": raw, "fbank_standardized": std,
                "generator": "torchaudio.compliance.kaldi.fbank"}, os.path.join(here, "fbank_golden.pt"))
    print("wrote fbank_golden.pt", [f.shape for f in raw])


if __name__ == "__main__":
    main()
