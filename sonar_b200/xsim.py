"""xsim cosine-margin mining over sentence embeddings on the B200 (BASELINE.json config 5).

``knn`` / ``xsim`` run entirely in ``libsonar_b200.so`` (``sb_xsim_knn``: tcgen05 GEMM with a fused
running top-k, exact fp64 re-rank; ``sb_xsim_margin_predict``).  ``xsim_distributed`` shards the
query rows over the ranks of a ``torch.distributed`` group: one all-gather assembles the embedding
matrices on every rank (the single exchange step of the path), each rank mines its own
row block in both directions, and a second, tiny all-gather merges the per-row neighbour averages.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

_MARGINS = {"absolute": 0, "ratio": 1, "distance": 2}


def _need_cuda_f32(t: Tensor) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError("sonar_b200.xsim runs on CUDA tensors only (no CPU fallback exists)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def knn(x: Tensor, y: Tensor, k: int = 4) -> Tuple[Tensor, Tensor]:
    """Exact-cosine k nearest rows of ``y`` for every row of ``x`` -> (cos fp64 [n,k], idx int32 [n,k])."""
    x, y = _need_cuda_f32(x), _need_cuda_f32(y)
    n, d = x.shape
    m = y.shape[0]
    assert y.shape[1] == d
    lib = _lib.load()
    need = C.c_size_t()
    _lib.check(lib.sb_xsim_workspace_bytes(n, m, d, C.byref(need)), "sb_xsim_workspace_bytes")
    ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
    val = torch.empty((n, k), dtype=torch.float64, device=x.device)
    idx = torch.empty((n, k), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.sb_xsim_knn(x.data_ptr(), y.data_ptr(), n, m, d, k, val.data_ptr(), idx.data_ptr(), ws.data_ptr(),
                             ws.numel(), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "sb_xsim_knn")
    return val, idx


def margin_predict(val_xy: Tensor, idx_xy: Tensor, val_yx: Optional[Tensor], m: int, margin: str = "ratio") -> Tensor:
    n, k = val_xy.shape
    pred = torch.empty((n,), dtype=torch.int32, device=val_xy.device)
    lib = _lib.load()
    with torch.cuda.device(val_xy.device):
        rc = lib.sb_xsim_margin_predict(val_xy.data_ptr(), idx_xy.data_ptr(),
                                        val_yx.data_ptr() if val_yx is not None else None, n, m, k, _MARGINS[margin],
                                        pred.data_ptr(), torch.cuda.current_stream(val_xy.device).cuda_stream)
    _lib.check(rc, "sb_xsim_margin_predict")
    return pred


def xsim(x: Tensor, y: Tensor, margin: str = "ratio", k: int = 4) -> Tuple[int, int, Tensor]:
    """LASER-style xsim: row i of ``x`` should retrieve row i of ``y``.  -> (errors, n, predictions int32 [n])."""
    if margin not in _MARGINS:
        raise ValueError(f"margin must be one of {sorted(_MARGINS)}")
    val_xy, idx_xy = knn(x, y, k)
    val_yx = None
    if margin != "absolute":
        val_yx, _ = knn(y, x, k)
    pred = margin_predict(val_xy, idx_xy, val_yx, y.shape[0], margin)
    n = x.shape[0]
    err = int((pred.long() != torch.arange(n, device=pred.device)).sum().item())
    return err, n, pred


def _xsim_distributed_impl(x_shard: Tensor, y_shard: Tensor, margin: str, k: int, group, knn_fn, margin_fn):
    """Collective plumbing of ``xsim_distributed`` with the two compute steps passed in as callables (the gloo test in
    ``tests/test_distributed_gloo.py`` drives it with a CPU checker; the public function below binds the CUDA kernels)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ns, d = x_shard.shape
    x_all = torch.empty((world * ns, d), dtype=torch.float32, device=x_shard.device)
    y_all = torch.empty((world * ns, d), dtype=torch.float32, device=x_shard.device)
    dist.all_gather_into_tensor(x_all, x_shard, group=group)  # the exchange step: [N,1024] on every rank
    dist.all_gather_into_tensor(y_all, y_shard, group=group)
    val_xy, idx_xy = knn_fn(x_shard, y_all, k)  # this rank's query rows against all of y
    val_yx_all = None
    if margin != "absolute":
        val_yx, _ = knn_fn(y_shard, x_all, k)  # reverse direction for this rank's y rows
        val_yx_all = torch.empty((world * ns, k), dtype=torch.float64, device=x_shard.device)
        dist.all_gather_into_tensor(val_yx_all, val_yx, group=group)  # tiny: [N,k] fp64
    pred = margin_fn(val_xy, idx_xy, val_yx_all, world * ns, margin)
    target = torch.arange(rank * ns, (rank + 1) * ns, device=pred.device)
    err = (pred.long() != target).sum()
    dist.all_reduce(err, group=group)
    return int(err.item()), world * ns, pred


def xsim_distributed(x_shard: Tensor, y_shard: Tensor, margin: str = "ratio", k: int = 4, group=None):
    """Every rank holds the same number of rows of x and y (its batch shard of the encoded sentences).
    -> (global errors, global n, predictions for this rank's rows as GLOBAL y indices).  CUDA tensors only."""
    if margin not in _MARGINS:
        raise ValueError(f"margin must be one of {sorted(_MARGINS)}")
    x_shard, y_shard = _need_cuda_f32(x_shard), _need_cuda_f32(y_shard)
    return _xsim_distributed_impl(x_shard, y_shard, margin, k, group, knn, margin_predict)
