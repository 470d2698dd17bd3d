"""Look-alikes of the fairseq2 containers the SONAR pipelines pass around.

fairseq2 is not a dependency of this package, so it ships its own minimal
``PaddingMask`` / ``SequenceBatch`` (same attribute names as fairseq2 0.4:
``seqs``, ``padding_mask.seq_lens``, ``.materialize()``) and the reference's
``SonarEncoderOutput`` (``sonar/models/encoder_model.py:17-38``).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
from torch import Tensor


class PaddingMask:
    """Sequence lengths of a right-padded batch (fairseq2 ``PaddingMask`` look-alike).

    ``seq_lens_host`` caches the lengths as Python ints so the CUDA engine never
    needs a device->host sync to size its packed-token buffers.
    """

    def __init__(self, seq_lens: Tensor, batch_seq_len: int,
                 seq_lens_host: Optional[Sequence[int]] = None) -> None:
        self.seq_lens = seq_lens
        self.batch_seq_len = int(batch_seq_len)
        self._host: Optional[List[int]] = list(map(int, seq_lens_host)) if seq_lens_host is not None else None

    @property
    def seq_lens_host(self) -> List[int]:
        if self._host is None:
            self._host = [int(v) for v in self.seq_lens.detach().cpu().tolist()]
        return self._host

    def materialize(self) -> Tensor:
        """Boolean mask [N, S], True at real positions (fairseq2 semantics)."""
        idx = torch.arange(self.batch_seq_len, device=self.seq_lens.device)
        return idx[None, :] < self.seq_lens[:, None]

    def to(self, device) -> "PaddingMask":
        return PaddingMask(self.seq_lens.to(device), self.batch_seq_len, self._host)


@dataclass
class SequenceBatch:
    """``seqs`` int64 [N, S] token ids (or [N, S, *] features); ``padding_mask`` or None."""

    seqs: Tensor
    padding_mask: Optional[PaddingMask] = None

    @property
    def batch_size(self) -> int:
        return self.seqs.size(0)


@dataclass
class SonarEncoderOutput:
    """Same fields as the reference dataclass (``sonar/models/encoder_model.py:17-38``)."""

    encoded_seqs: Optional[Tensor]
    """[N, S, M] final-LayerNormed states, or None when not requested (the text pipeline
    only consumes ``sentence_embeddings``, ``sonar/inference_pipelines/text.py:245``)."""

    sentence_embeddings: Tensor
    """[N, M] pooled representation."""

    padding_mask: Optional[PaddingMask]
