"""Shared parity metrics for the CUDA-vs-oracle tests."""

import torch


def parity_metrics(got: torch.Tensor, ref: torch.Tensor) -> dict:
    """Per-sentence 1-cos, mean-centred cosine, relative L2 of [N,D] embeddings (fp64 maths)."""
    g, r = got.double().cpu(), ref.double().cpu()
    cos = torch.nn.functional.cosine_similarity(g, r, dim=1)
    gc, rc = g - g.mean(0, keepdim=True), r - r.mean(0, keepdim=True)
    ccos = torch.nn.functional.cosine_similarity(gc, rc, dim=1)
    rel = (g - r).norm(dim=1) / r.norm(dim=1)
    rn = torch.nn.functional.normalize(r, dim=1)
    off = (rn @ rn.T)[~torch.eye(len(r), dtype=torch.bool)] if len(r) > 1 else torch.zeros(1)
    return {
        "one_minus_cos_max": float((1 - cos).max()),
        "centred_cos_min": float(ccos.min()) if len(r) > 2 else 1.0,
        "rel_l2_max": float(rel.max()),
        "max_abs": float((g - r).abs().max()),
        "ref_offdiag_cos_mean": float(off.mean()),
    }


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    g, r = got.double(), ref.double()
    return float((g - r).norm() / r.norm().clamp_min(1e-30))
