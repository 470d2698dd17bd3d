"""CPU oracle for xsim cosine-margin mining.  TEST INFRASTRUCTURE ONLY (see oracle/text_encoder.py).

PARITY UNPINNED against the reference: the repository contains no xsim code, fixture or golden -- only the word
in README.md:5 and result tables (README.md:19-22).  Pinned instead against an independent implementation
(tests/test_oracle_xsim.py: scikit-learn's brute-force cosine k-NN for the search, a dense evaluation of the margin
formula for the scoring).  This file restates the public LASER
``xsim.py`` algorithm in float64 NumPy: L2-normalise, cosine = inner
product, k-NN in both directions, ratio/distance margin over the forward k candidates, prediction
= candidate with the best margin score, ties broken by the lowest candidate rank (first maximum),
candidates ordered by (score descending, index ascending).  The only reference call sites of the
underlying maths are ``normalize`` + ``matmul`` in tests/integration_tests/test_text_sonar.py:42,51.
"""

from __future__ import annotations

import numpy as np


def l2_normalize(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-30)


def topk_desc(sim: np.ndarray, k: int):
    """Row-wise top-k of a dense similarity block: (values, indices), sorted by (value desc, index asc)."""
    n, m = sim.shape
    k = min(k, m)
    if m <= 4 * k:
        # stable argsort on -sim gives value-descending with index-ascending tie-break
        idx = np.argsort(-sim, axis=1, kind="stable")[:, :k]
        return np.take_along_axis(sim, idx, axis=1), idx
    # same result without sorting whole rows: every entry >= the row's k-th largest value survives (ties included), then the
    # survivors are ordered by (value desc, index asc)
    kth = np.partition(sim, m - k, axis=1)[:, m - k]
    idx = np.empty((n, k), dtype=np.int64)
    for r in range(n):
        cand = np.flatnonzero(sim[r] >= kth[r])  # ascending indices
        order = np.argsort(-sim[r, cand], kind="stable")[:k]
        idx[r] = cand[order]
    val = np.take_along_axis(sim, idx, axis=1)
    return val, idx


def knn(x: np.ndarray, y: np.ndarray, k: int, block: int = 4096):
    """Exact float64 k nearest neighbours (cosine) of every row of x among the rows of y."""
    xn, yn = l2_normalize(x), l2_normalize(y)
    vals, idxs = [], []
    for s in range(0, xn.shape[0], block):
        v, i = topk_desc(xn[s : s + block] @ yn.T, k)
        vals.append(v)
        idxs.append(i)
    return np.concatenate(vals), np.concatenate(idxs)


def xsim(x: np.ndarray, y: np.ndarray, margin: str = "ratio", k: int = 4):
    """-> (error_count, n, predictions[n]) with the i-th row of x expected to match the i-th row of y."""
    n = x.shape[0]
    cos_xy, idx_xy = knn(x, y, k)
    if margin == "absolute":
        pred = idx_xy[:, 0]
    else:
        cos_yx, _ = knn(y, x, k)
        avg_x = cos_xy.mean(axis=1)  # [n]
        avg_y = cos_yx.mean(axis=1)  # [m]
        denom = (avg_x[:, None] + avg_y[idx_xy]) / 2.0
        if margin == "ratio":
            score = cos_xy / denom
        elif margin == "distance":
            score = cos_xy - denom
        else:
            raise ValueError(margin)
        pred = idx_xy[np.arange(n), np.argmax(score, axis=1)]  # argmax = first maximum
    err = int((pred != np.arange(n)).sum())
    return err, n, pred
