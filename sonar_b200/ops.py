"""Thin torch-tensor wrappers over the kernel-level C-ABI entry points.

Used by the parity tests, the micro-benchmarks and ``static_pooling``.  Every
function requires CUDA tensors and launches on torch's current stream; there is
no CPU implementation behind any of them.
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib

POOLING_MODES = {"max": _lib.SB_POOL_MAX, "mean": _lib.SB_POOL_MEAN, "last": _lib.SB_POOL_LAST}


def _need_cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sonar_b200 ops run on CUDA tensors only (no CPU fallback exists)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def gemm_bf16(a: Tensor, w: Tensor, bias: Tensor, *, epilogue: str = "bias", residual: Optional[Tensor] = None,
              out_dtype: torch.dtype = torch.bfloat16, out: Optional[Tensor] = None, cta_group: int = 2) -> Tensor:
    """out[M,N] = epi(a[M,K] @ w[N,K]^T + bias[N]); epilogue in {bias, relu, residual}."""
    _need_cuda(a, w, bias, residual, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    m, k = a.shape
    n = w.shape[0]
    assert w.shape[1] == k and bias.numel() == n
    epi = {"bias": _lib.SB_EPI_BIAS, "relu": _lib.SB_EPI_BIAS_RELU, "residual": _lib.SB_EPI_BIAS_RESIDUAL, "silu": 5}[epilogue]
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    assert out.dtype in (torch.bfloat16, torch.float32) and out.stride(1) == 1
    if epilogue == "residual":
        assert residual is not None and residual.dtype == out.dtype and residual.stride(1) == 1
    rc = _lib.load().sb_gemm_bf16(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0),
        1 if out.dtype == torch.float32 else 0, bias.data_ptr(), _ptr(residual),
        residual.stride(0) if residual is not None else 0, m, n, k, epi, cta_group, _stream())
    _lib.check(rc, "sb_gemm_bf16")
    return out


def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    """bf16 LayerNorm(x fp32 [T,D])."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().sb_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(),
                                  x.shape[0], x.shape[1], _stream())
    _lib.check(rc, "sb_layernorm")
    return y


def cu_seqlens_of(seq_lens) -> Tensor:
    lens = torch.as_tensor(seq_lens, dtype=torch.int64).cpu()
    cu = torch.zeros(lens.numel() + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0).to(torch.int32)
    return cu


ATTENTION_IMPLS = {"auto": 0, "mma_sync": 1, "tcgen05": 2}


def fold_layernorm(w: Tensor, bias: Tensor, gamma: Tensor, beta: Tensor):
    """LayerNorm folding, weight side: -> (Wf bf16 [N,K], colsum fp32 [N], bias_f fp32 [N])."""
    _need_cuda(w, bias, gamma, beta)
    assert w.dtype == torch.bfloat16 and w.is_contiguous()
    n, k = w.shape
    wf = torch.empty_like(w)
    colsum = torch.empty((n,), dtype=torch.float32, device=w.device)
    bias_f = torch.empty((n,), dtype=torch.float32, device=w.device)
    rc = _lib.load().sb_fold_layernorm(w.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(), n, k, wf.data_ptr(),
                                       colsum.data_ptr(), bias_f.data_ptr(), _stream())
    _lib.check(rc, "sb_fold_layernorm")
    return wf, colsum, bias_f


def gemm_residual_stats(a: Tensor, w: Tensor, bias: Tensor, x: Tensor):
    """x += a . w^T + bias in place (fp32); -> (h = bf16(x) [M,N], stats fp32 [M, N/128, 2]): partial 2*t + g holds (mean, M2)
    of the 128 columns {256 t + 32 c + j : c % 2 == g, j < 32} of the new row (epilogue warpgroup g's share of tile t)."""
    _need_cuda(a, w, bias, x)
    m, k = a.shape
    n = w.shape[0]
    assert x.shape == (m, n) and x.dtype == torch.float32 and x.is_contiguous() and n % 256 == 0
    h = torch.empty((m, n), dtype=torch.bfloat16, device=a.device)
    stats = torch.empty((m, n // 128, 2), dtype=torch.float32, device=a.device)
    rc = _lib.load().sb_gemm_residual_stats(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), x.data_ptr(), n,
                                            bias.data_ptr(), h.data_ptr(), n, stats.data_ptr(), m, n, k, _stream())
    _lib.check(rc, "sb_gemm_residual_stats")
    return h, stats


def gemm_residual_splitk(a: Tensor, w: Tensor, bias: Tensor, x: Tensor, counters: Tensor) -> Tensor:
    """x += a . w^T + bias in place (fp32) with the decoder step's ordered split-K (``sb_gemm_residual_splitk``);
    ``counters``: zero-initialised int32 device tensor with >= 4 * (M/256 * N/256) entries, left zero."""
    _need_cuda(a, w, bias, x, counters)
    m, k = a.shape
    n = w.shape[0]
    assert x.shape == (m, n) and x.dtype == torch.float32 and x.is_contiguous() and counters.dtype == torch.int32
    rc = _lib.load().sb_gemm_residual_splitk(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), x.data_ptr(), n,
                                             bias.data_ptr(), m, n, k, counters.data_ptr(), counters.numel(), _stream())
    _lib.check(rc, "sb_gemm_residual_splitk")
    return x


def gemm_ln_consumer(a: Tensor, wf: Tensor, bias_f: Tensor, colsum: Tensor, stats: Tensor, eps: float = 1e-5,
                     relu: bool = False) -> Tensor:
    """bf16 [M,N] = [relu](rstd * (a . wf^T - mean * colsum) + bias_f), (mean, rstd) merged from stats [M, K/128, 2]
    (any partition of the row into 128-column subsets)."""
    _need_cuda(a, wf, bias_f, colsum, stats)
    m, k = a.shape
    n = wf.shape[0]
    out = torch.empty((m, n), dtype=torch.bfloat16, device=a.device)
    rc = _lib.load().sb_gemm_ln_consumer(a.data_ptr(), a.stride(0), wf.data_ptr(), wf.stride(0), out.data_ptr(), n,
                                         bias_f.data_ptr(), colsum.data_ptr(), stats.data_ptr(), eps, m, n, k,
                                         1 if relu else 0, _stream())
    _lib.check(rc, "sb_gemm_ln_consumer")
    return out


def attention(qkv: Tensor, cu_seqlens: Tensor, max_len: int, num_heads: int, impl: str = "auto") -> Tensor:
    """Packed bidirectional MHA: qkv bf16 [T, 3*64*H] -> bf16 [T, 64*H]."""
    _need_cuda(qkv, cu_seqlens)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and cu_seqlens.dtype == torch.int32
    t = qkv.shape[0]
    d = 64 * num_heads
    assert qkv.shape[1] == 3 * d
    out = torch.empty((t, d), dtype=torch.bfloat16, device=qkv.device)
    rc = _lib.load().sb_attention(qkv.data_ptr(), cu_seqlens.data_ptr(), cu_seqlens.numel() - 1, max_len,
                                  num_heads, t, ATTENTION_IMPLS[impl], out.data_ptr(), _stream())
    _lib.check(rc, "sb_attention")
    return out


def embed(ids: Tensor, cu_seqlens: Tensor, table: Tensor, pos_table: Tensor, scale: float, total_tokens: int) -> Tensor:
    """x[cu[b]+t] = table[ids[b,t]] * scale + pos_table[t]  (fp32 [T, D])."""
    _need_cuda(ids, cu_seqlens, table, pos_table)
    assert ids.dtype == torch.int64 and ids.stride(1) == 1 and table.dtype == torch.bfloat16
    b, s = ids.shape
    d = table.shape[1]
    x = torch.empty((total_tokens, d), dtype=torch.float32, device=ids.device)
    err = torch.zeros(1, dtype=torch.int32, device=ids.device)
    rc = _lib.load().sb_embed(ids.data_ptr(), ids.stride(0), cu_seqlens.data_ptr(), b, s, table.data_ptr(),
                              table.shape[0], pos_table.data_ptr(), pos_table.shape[0], d, scale, x.data_ptr(),
                              err.data_ptr(), _stream())
    _lib.check(rc, "sb_embed")
    if int(err.item()) != 0:
        raise ValueError("sb_embed: token id outside [0, vocab_size)")
    return x


def pool_packed(x: Tensor, cu_seqlens: Tensor, pooling: str, *, gamma: Optional[Tensor] = None,
                beta: Optional[Tensor] = None, eps: float = 1e-5, encoded_seq_len: int = 0):
    """(optional LayerNorm +) pooling of packed rows fp32 [T,D] -> fp32 [B,D]
    (and, if ``encoded_seq_len`` > 0, the padded [B,S,D] states)."""
    _need_cuda(x, cu_seqlens, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous()
    b = cu_seqlens.numel() - 1
    d = x.shape[1]
    out = torch.empty((b, d), dtype=torch.float32, device=x.device)
    enc = torch.empty((b, encoded_seq_len, d), dtype=torch.float32, device=x.device) if encoded_seq_len > 0 else None
    rc = _lib.load().sb_pool(x.data_ptr(), cu_seqlens.data_ptr(), b, d, _ptr(gamma), _ptr(beta), eps,
                             1 if gamma is not None else 0, POOLING_MODES[pooling.lower()], out.data_ptr(),
                             _ptr(enc), encoded_seq_len, _stream())
    _lib.check(rc, "sb_pool")
    return (out, enc) if enc is not None else out
