"""BASELINE.json config 5: batch-sharded encode -> ONE all-gather -> distributed xsim, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_config5.py [--per-gpu 125000] [--seq-len 128]

Every rank encodes its shard of `per_gpu` synthetic sentences (batches of 4096 x seq_len through the engine, no collective),
the [N,1024] fp32 matrix is assembled with one NCCL all-gather, a second set y_i = x_i + 0.1 N(0,1) |x_i| / sqrt(d) is built
per shard (so the true neighbour of x_i is y_i), and `xsim_distributed` (ratio margin, k=4) scores X against Y.  Rank 0
prints one JSON line: encode sentences/s (whole job), all-gather GB/s (bytes received per rank / time), xsim pairs/s and
the error count.  All times are CUDA-event times on the device, max over ranks.
"""

import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import D, VOCAB, dist_env, flops_per_sentence, load_peaks, synthetic_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=125000)
    ap.add_argument("--seq-len", type=int, default=128)
    ap.add_argument("--batch", type=int, default=4096)
    args = ap.parse_args()
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:  # a 1-rank group still exercises the same code (gather = copy)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    import __graft_entry__

    if rank == 0:
        __graft_entry__.build()
    dist.barrier()
    from sonar_b200 import B200TextEncoderModel, SequenceBatch, sonar_text_encoder_config
    from sonar_b200.xsim import xsim_distributed

    model = B200TextEncoderModel(sonar_text_encoder_config("basic"), synthetic_state_dict(dev), dev)
    torch.cuda.empty_cache()
    ns, S, B = args.per_gpu, args.seq_len, args.batch
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    x_shard = torch.empty((ns, D), dtype=torch.float32, device=dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def max_ms(e0, e1):
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def encode():
        for s in range(0, ns, B):
            n = min(B, ns - s)
            ids = torch.randint(4, VOCAB, (n, S), generator=g, device=dev, dtype=torch.int64)
            x_shard[s:s + n] = model(SequenceBatch(ids, None)).sentence_embeddings

    # warm-up: one batch through the engine, one small collective
    model(SequenceBatch(torch.randint(4, VOCAB, (B, S), generator=g, device=dev, dtype=torch.int64), None))
    dist.all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    encode()
    e1.record()
    torch.cuda.synchronize()
    enc_ms = max_ms(e0, e1)

    x_all = torch.empty((world * ns, D), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(x_all, x_shard)  # warm the communicator at this size
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    dist.all_gather_into_tensor(x_all, x_shard)
    e1.record()
    torch.cuda.synchronize()
    ag_ms = max_ms(e0, e1)
    del x_all

    y_shard = x_shard + 0.1 * torch.randn((ns, D), generator=g, device=dev) * x_shard.norm(dim=1, keepdim=True) / math.sqrt(D)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    err, n, _ = xsim_distributed(x_shard, y_shard, margin="ratio", k=4)
    e1.record()
    torch.cuda.synchronize()
    xs_ms = max_ms(e0, e1)
    if rank == 0:
        peaks, src = load_peaks()
        n_total = world * ns
        sent_s = n_total / enc_ms * 1e3
        # both k-NN directions are scored: 2 x (N/world x N) per rank
        pairs = 2.0 * n_total * n_total
        print(json.dumps({
            "config": f"config 5: {n_total} synthetic sentences x {S} tokens sharded {ns}/GPU over {world} GPU(s); one fp32 "
                      f"all-gather; xsim ratio margin k=4 of [{n_total},1024] vs noisy copies (both k-NN directions)",
            "n_gpus": world,
            "encode": {"value": sent_s, "unit": "sentences/s", "ms": enc_ms,
                       "roofline_frac": sent_s * flops_per_sentence(S) / 1e12 / (world * float(peaks["bf16_tflops_sustained"]))},
            "all_gather": {"bytes_received_per_rank": (world - 1) * ns * D * 4, "ms": ag_ms,
                           "value": (world - 1) * ns * D * 4 / ag_ms / 1e6 if world > 1 else None, "unit": "GB/s per rank (receive)"},
            "xsim": {"value": pairs / xs_ms * 1e3, "unit": "pairs/s", "ms": xs_ms, "errors": err, "n": n,
                     "includes": "the two [N,1024] all-gathers inside xsim_distributed, L2 normalisation, bf16 GEMM + top-16, "
                                 "fp64 re-rank, margin scoring, error all-reduce",
                     "tensor_tflops": 2.0 * pairs * D / xs_ms / 1e9,
                     "roofline_frac": 2.0 * pairs * D / xs_ms / 1e9 / (world * float(peaks["bf16_tflops_sustained"]))},
            "peaks_source": src}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
