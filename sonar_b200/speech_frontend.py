"""Speech feature frontend on the GPU: waveform -> standardised 80-bin Kaldi fbank batch.

Mirror of what ``SpeechToEmbeddingModelPipeline`` builds from fairseq2n operators
(``sonar/inference_pipelines/speech.py:283-290,439-451``): ``WaveformToFbankConverter(num_mel_bins=80,
waveform_scale=2**15, channel_last=True, standardize=True)`` per utterance, then
``Collater(pad_value=0, pad_to_multiple=2)`` -> ``{"fbank": {"seqs": [B,T,80], "seq_lens": [B]}}``.
All arithmetic runs in ``sb_fbank`` (``csrc/fbank.cu``).
"""

from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

FRAME_LEN, FRAME_SHIFT, NUM_MEL, SAMPLE_RATE = 400, 160, 80, 16000


def num_frames(num_samples: int) -> int:
    return 0 if num_samples < FRAME_LEN else 1 + (num_samples - FRAME_LEN) // FRAME_SHIFT


class WaveformToFbank:
    """Callable: list of mono 16 kHz waveforms (1-D, or [C,T] / [T,C] with one channel) ->
    (fbank fp32 CUDA [B, Tpad, 80], frame counts)."""

    def __init__(self, device, pad_to_multiple: int = 2) -> None:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("WaveformToFbank runs on a CUDA device only (no CPU path)")
        self.device = dev
        self.pad_to_multiple = pad_to_multiple
        lib = _lib.load()
        n = lib.sb_fbank_tables_bytes()
        host = torch.empty(n, dtype=torch.uint8)
        _lib.check(lib.sb_fbank_build_tables(host.data_ptr()), "sb_fbank_build_tables")
        self._tables = host.to(dev)
        self._lib = lib

    @torch.inference_mode()
    def __call__(self, waveforms: Sequence[Tensor]) -> Tuple[Tensor, List[int]]:
        flat = []
        for w in waveforms:
            if w.dim() == 2:
                if 1 not in w.shape:
                    raise ValueError("only mono waveforms are supported")
                w = w.reshape(-1)
            if w.dim() != 1:
                raise ValueError("waveform must be 1-D or [1, T]")
            flat.append(w.to(dtype=torch.float32))
        lens = [int(w.numel()) for w in flat]
        frames = [num_frames(n) for n in lens]
        if min(frames) < 1:
            raise ValueError("waveform shorter than one 25 ms frame (400 samples at 16 kHz)")
        dev = self.device
        packed = torch.cat([w.to(dev, non_blocking=True) for w in flat])
        woff = torch.zeros(len(flat) + 1, dtype=torch.int64)
        woff[1:] = torch.cumsum(torch.tensor(lens), 0)
        foff = torch.zeros(len(flat) + 1, dtype=torch.int32)
        foff[1:] = torch.cumsum(torch.tensor(frames), 0).to(torch.int32)
        total = int(foff[-1])
        m = self.pad_to_multiple
        tpad = (max(frames) + m - 1) // m * m
        raw = torch.empty((total, NUM_MEL), dtype=torch.float32, device=dev)
        out = torch.empty((len(flat), tpad, NUM_MEL), dtype=torch.float32, device=dev)
        woff_d, foff_d = woff.to(dev), foff.to(dev)
        with torch.cuda.device(dev):
            rc = self._lib.sb_fbank(packed.data_ptr(), woff_d.data_ptr(), foff_d.data_ptr(), len(flat), total,
                                    self._tables.data_ptr(), raw.data_ptr(), out.data_ptr(), tpad,
                                    torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "sb_fbank")
        return out, frames
