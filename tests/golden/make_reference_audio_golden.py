"""Generate tests/golden/reference_audio/reference_audio_fbank.pt from the reference's own audio fixtures:

    python tests/golden/make_reference_audio_golden.py

For each of audio_1.wav / audio_2.wav (16 kHz mono PCM-16, read with the standard library):
features = torchaudio.compliance.kaldi.fbank(w * 2**15, num_mel_bins=80) (dither 0, povey window, 25/10 ms, snip_edges), i.e.
what fairseq2n's WaveformToFbankConverter(num_mel_bins=80, waveform_scale=2**15, standardize=True) computes before its
per-utterance standardisation (reference sonar/inference_pipelines/speech.py:283-290)."""

import os
import wave

import torch
import torchaudio.compliance.kaldi as kaldi

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_audio")


def read_wav(path):
    with wave.open(path, "rb") as f:
        assert f.getsampwidth() == 2 and f.getnchannels() == 1 and f.getframerate() == 16000
        pcm = torch.frombuffer(bytearray(f.readframes(f.getnframes())), dtype=torch.int16)
    return pcm.float() / 32768.0


def main() -> None:
    out = {"files": [], "num_samples": [], "fbank_raw": []}
    for name in ("audio_1.wav", "audio_2.wav"):
        w = read_wav(os.path.join(HERE, name))
        out["files"].append(name)
        out["num_samples"].append(int(w.numel()))
        out["fbank_raw"].append(kaldi.fbank(w[None] * 2 ** 15, num_mel_bins=80, sample_frequency=16000.0))
    out["generator"] = "torchaudio.compliance.kaldi.fbank"
    torch.save(out, os.path.join(HERE, "reference_audio_fbank.pt"))
    print("wrote reference_audio_fbank.pt", [tuple(f.shape) for f in out["fbank_raw"]])


if __name__ == "__main__":
    main()
