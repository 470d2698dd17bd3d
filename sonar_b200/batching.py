"""Host-side batcher: the fairseq2n data-pipeline operators the text pipeline chains at
``sonar/inference_pipelines/text.py:231-247``, restated as plain Python generators.

* ``dynamic_bucket`` -- fairseq2 ``DataPipelineBuilder.dynamic_bucket`` semantics [fs2]
  (SURVEY App. A.1 step 3 / F5): examples are appended until
  ``count >= max_num_examples`` or (``sum(cost) >= threshold`` and ``count >= min``);
  the example that crosses the threshold is included; the last partial bucket is kept
  unless ``drop_remainder``.
* ``collate`` -- ``Collater(pad_value)`` [fs2]: right-pad ragged int64 vectors to the
  batch maximum; ``is_ragged`` False => the pipeline passes ``padding_mask=None``
  (``sonar/inference_pipelines/utils.py:18-21``).
* ``prefetch`` -- ``.prefetch(n)``: everything upstream runs in a background thread.
"""

from __future__ import annotations

import queue
import threading
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple, TypeVar

import torch
from torch import Tensor

from .sequence import PaddingMask, SequenceBatch

T = TypeVar("T")


def dynamic_bucket(items: Iterable[T], threshold: float, cost_fn: Callable[[T], float], *,
                   min_num_examples: int = 1, max_num_examples: Optional[int] = None,
                   drop_remainder: bool = False) -> Iterator[List[T]]:
    bucket: List[T] = []
    cost = 0.0
    for ex in items:
        bucket.append(ex)
        cost += cost_fn(ex)
        full = max_num_examples is not None and len(bucket) >= max_num_examples
        if full or (cost >= threshold and len(bucket) >= min_num_examples):
            yield bucket
            bucket, cost = [], 0.0
    if bucket and not drop_remainder:
        yield bucket


def bucket(items: Iterable[T], size: int, drop_remainder: bool = False) -> Iterator[List[T]]:
    """``.bucket(batch_size)``: fixed-count batches (used by the decoder/speech pipelines)."""
    cur: List[T] = []
    for ex in items:
        cur.append(ex)
        if len(cur) == size:
            yield cur
            cur = []
    if cur and not drop_remainder:
        yield cur


def collate(seqs: Sequence[Tensor], pad_value: int, *, pin_memory: bool = False) -> Tuple[Tensor, List[int], bool]:
    """-> (ids int64 [N, Smax] right-padded, seq_lens, is_ragged)."""
    lens = [int(t.shape[0]) for t in seqs]
    smax = max(lens) if lens else 0
    ragged = any(n != smax for n in lens)
    pin = bool(pin_memory and torch.cuda.is_available())
    if not ragged and seqs:  # one stack instead of a Python loop over rows
        out = torch.empty((len(seqs), smax), dtype=torch.int64, pin_memory=pin)
        torch.stack(list(seqs), out=out)
        return out, lens, False
    out = torch.full((len(seqs), smax), int(pad_value), dtype=torch.int64, pin_memory=pin)
    if seqs:  # scatter the concatenated tokens to (row, position) in one indexed write
        lt = torch.tensor(lens, dtype=torch.int64)
        row = torch.repeat_interleave(torch.arange(len(seqs)), lt)
        starts = torch.cumsum(lt, 0) - lt
        pos = torch.arange(int(lt.sum())) - torch.repeat_interleave(starts, lt)
        out[row, pos] = torch.cat(list(seqs))
    return out, lens, ragged


def to_sequence_batch(ids: Tensor, lens: List[int], is_ragged: bool, device) -> SequenceBatch:
    """``extract_sequence_batch`` (``utils.py:18-21``): H2D copy + PaddingMask-or-None."""
    dev = torch.device(device)
    seqs = ids.to(dev, non_blocking=True) if dev.type == "cuda" else ids
    mask = None
    if is_ragged:
        mask = PaddingMask(torch.tensor(lens, dtype=torch.int64), ids.shape[1], seq_lens_host=lens)
    return SequenceBatch(seqs, mask)


_SENTINEL = object()


def prefetch(it: Iterable[T], depth: int = 2) -> Iterator[T]:
    """Run ``it`` in a daemon thread, keeping up to ``depth`` results queued; exceptions
    raised upstream are re-raised in the consumer thread."""
    if depth <= 0:
        yield from it
        return
    q: "queue.Queue" = queue.Queue(maxsize=depth)
    stop = threading.Event()

    def worker() -> None:
        try:
            for x in it:
                while not stop.is_set():
                    try:
                        q.put((x, None), timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put((_SENTINEL, None))
        except BaseException as e:  # noqa: BLE001 - forwarded to the consumer
            q.put((_SENTINEL, e))

    th = threading.Thread(target=worker, daemon=True, name="sonar_b200-prefetch")
    th.start()
    try:
        while True:
            x, err = q.get()
            if err is not None:
                raise err
            if x is _SENTINEL:
                return
            yield x
    finally:
        stop.set()
