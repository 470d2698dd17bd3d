"""Mirror of ``sonar/inference_pipelines/utils.py`` (``extract_sequence_batch`` :18-21,
``add_progress_bar`` :24-46)."""

from __future__ import annotations

import math
from pathlib import Path
from typing import Iterable, Optional, Union

from ..batching import to_sequence_batch
from ..sequence import SequenceBatch


def extract_sequence_batch(x: dict, device) -> SequenceBatch:
    """``x`` is collater output ``{"seqs", "seq_lens", "is_ragged"}`` (fairseq2 ``SequenceData``)."""
    lens = [int(v) for v in x["seq_lens"]]
    return to_sequence_batch(x["seqs"], lens, bool(x["is_ragged"]), device)


def add_progress_bar(sequence: Iterable, inputs: Optional[Union[Iterable, str, Path]] = None,
                     batch_size: Optional[int] = 1, **kwargs) -> Iterable:
    """Wrap the input into a tqdm progress bar (total = ceil(len(inputs) / batch_size) when known)."""
    from tqdm.auto import tqdm

    total = None
    if inputs is None:
        inputs = sequence
    if batch_size is not None:
        if hasattr(inputs, "__len__") and not isinstance(inputs, (str, Path)):
            total = math.ceil(len(inputs) / batch_size)  # type: ignore
    return tqdm(sequence, total=total, **kwargs)
