"""NCCL tests of the multi-GPU path: batch-sharded encoding + all-gather, and distributed xsim mining, against the
single-process results / the NumPy oracle.  World size = min(2, visible GPUs): on a single-GPU box the same code runs
as a 1-rank NCCL group (all-gather = copy), with 2 GPUs (`gpurun --gpus 2`) the collectives cross NVLink."""

import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

VOCAB = 4096


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import xsim as ox
        from oracle.text_encoder import OracleEncoderConfig, make_synthetic_state_dict
        from sonar_b200 import B200TextEncoderModel, VocabularyInfo, sonar_text_encoder_config
        from sonar_b200.distributed import encode_sharded
        from sonar_b200.inference_pipelines import TextToEmbeddingModelPipeline
        from sonar_b200.tokenizer import SyntheticTokenizer
        from sonar_b200.xsim import xsim_distributed

        # ---- distributed xsim vs the float64 oracle ----
        g = torch.Generator().manual_seed(5)
        y = torch.randn((2048, 1024), generator=g)
        x = y + 1.0 * torch.randn((2048, 1024), generator=g) * y.norm(dim=1, keepdim=True) / 32.0
        ns = 2048 // world
        sl = slice(rank * ns, (rank + 1) * ns)
        for margin in ("ratio", "absolute"):
            err, n, pred = xsim_distributed(x[sl].to(dev), y[sl].to(dev), margin=margin, k=4)
            ref_err, ref_n, ref_pred = ox.xsim(x.numpy(), y.numpy(), margin=margin, k=4)
            assert (err, n) == (ref_err, ref_n), (margin, err, ref_err)
            assert np.array_equal(pred.cpu().numpy(), ref_pred[sl])

        # ---- batch-sharded encode + all-gather == single-process encode ----
        ocfg = OracleEncoderConfig(vocab_size=VOCAB, num_layers=2)
        sd = make_synthetic_state_dict(ocfg, seed=1)
        cfg = sonar_text_encoder_config("basic", num_encoder_layers=2,
                                        vocab_info=VocabularyInfo(size=VOCAB, unk_idx=1, bos_idx=2, eos_idx=3, pad_idx=1))
        pipe = TextToEmbeddingModelPipeline(B200TextEncoderModel(cfg, sd, dev), SyntheticTokenizer(vocab_size=VOCAB),
                                            device=dev)
        sents = [" ".join(f"w{(i * 7 + j) % 50}" for j in range(1 + i % 13)) for i in range(37)]
        emb = encode_sharded(lambda s: pipe.predict(s, "eng_Latn", batch_size=8), sents)
        full = pipe.predict(sents, "eng_Latn", batch_size=8)
        assert emb.shape == (37, 1024)
        torch.testing.assert_close(emb, full, rtol=1.3e-6, atol=1e-5)
        # more ranks than sentences: the rank with an empty shard still joins every collective on its CUDA device
        one = encode_sharded(lambda s: pipe.predict(s, "eng_Latn", batch_size=8), sents[:1])
        torch.testing.assert_close(one, full[:1], rtol=1.3e-6, atol=1e-5)
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_encode_and_xsim(native_lib):
    import torch.multiprocessing as mp

    world = min(2, torch.cuda.device_count())
    assert world >= 1
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {r: "ok" for r in range(world)}
