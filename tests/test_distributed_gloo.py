"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: shard bounds, the all-gather that assembles the
embedding matrix, batch-sharded encoding, and the collective plumbing of xsim_distributed (with the NumPy oracle
injected as the k-NN checker -- the product path always runs the CUDA kernels)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sonar_b200.distributed import encode_sharded, gather_rows, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- gather_rows with unequal shards ---
        n, d = 7, 5
        full = torch.arange(n * d, dtype=torch.float32).view(n, d)
        s, e = shard_bounds(n, world, rank)
        got = gather_rows(full[s:e].clone(), n)
        assert torch.equal(got, full)

        # --- batch-sharded encoding with a deterministic stand-in for pipeline.predict ---
        sents = [f"sentence number {i}" for i in range(11)]

        def fake_predict(batch):
            return torch.tensor([[float(len(t)), float(sum(map(ord, t)) % 97)] for t in batch])

        emb = encode_sharded(fake_predict, sents)
        assert torch.equal(emb, fake_predict(sents))

        # --- xsim_distributed plumbing, oracle injected as the k-NN / margin checker ---
        from oracle import xsim as ox
        from sonar_b200.xsim import _xsim_distributed_impl, xsim_distributed

        g = torch.Generator().manual_seed(3)
        y = torch.randn((64, 32), generator=g)
        x = y + 0.9 * torch.randn((64, 32), generator=g)

        def knn_cpu(a, b, k):
            v, i = ox.knn(a.numpy(), b.numpy(), k)
            return torch.from_numpy(v), torch.from_numpy(i.astype(np.int32))

        def knn_bidir_cpu(a, b, k):
            v1, i1 = knn_cpu(a, b, k)
            v2, i2 = knn_cpu(b, a, k)
            return v1, i1, v2, i2

        def margin_cpu(val_xy, idx_xy, val_yx, m, margin):
            v, i = val_xy.numpy(), idx_xy.numpy()
            if margin == "absolute":
                return torch.from_numpy(i[:, 0].copy())
            denom = (v.mean(1)[:, None] + val_yx.numpy().mean(1)[i]) / 2.0
            score = v / denom if margin == "ratio" else v - denom
            return torch.from_numpy(i[np.arange(len(i)), score.argmax(1)].copy())

        ns = 64 // world
        sl = slice(rank * ns, (rank + 1) * ns)
        for margin in ("ratio", "distance", "absolute"):
            err, n_tot, pred = _xsim_distributed_impl(x[sl], y[sl], margin, 4, None, knn_cpu, margin_cpu, knn_bidir_cpu)
            ref_err, ref_n, ref_pred = ox.xsim(x.numpy(), y.numpy(), margin=margin, k=4)
            assert n_tot == ref_n and err == ref_err, (margin, err, ref_err)
            assert np.array_equal(pred.numpy(), ref_pred[sl])
        with pytest.raises(RuntimeError, match="CUDA"):  # the public function has no CPU path and no injection seam
            xsim_distributed(x[sl], y[sl])
        # --- more ranks than sentences: the empty shard must still take part in every collective ---
        one = encode_sharded(fake_predict, sents[:1])
        assert torch.equal(one, fake_predict(sents[:1]))
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}
