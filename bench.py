"""Benchmark of the SONAR text-embedding hot path (BASELINE.json metric: sentences/sec -> 1024-d).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement of the reference path

Workload (BASELINE.json configs[1]): text_sonar_basic_encoder architecture (24 layers, d=1024,
16 heads, FFN 8192, vocab 256206, random-init weights), batch 4096 sentences x 128 tokens of
synthetic ids per GPU per step.  A "step" is one pass of the hot path over one batch:
embed -> 24 encoder layers -> final LN -> mean-pool -> [4096,1024] fp32.

* `value`  : whole-job sentences/s with the ids already resident in HBM (CUDA events, max over ranks)
* `e2e`    : same metric through the reference-facing model call with HOST (pinned) ids in and
             HOST embeddings out, copies inside the timed region
* `roofline`: the dominant kernel (tcgen05 GEMM, FFN inner-projection instantiation) timed alone
             with CUDA events on its launch stream, against MEASURED_PEAKS.json
* `cpu_baseline`: the fp32 PyTorch restatement of the fairseq2 op sequence (oracle/, "port") on the
             host cores, on a bounded sample of the same workload (rank 0, N=1 only)
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH, SEQ, D, FFN, LAYERS, HEADS, VOCAB = 4096, 128, 1024, 8192, 24, 16, 256206
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def flops_per_sentence(s: int) -> float:
    """F(S) = L*S*(2*(4d^2 + 2df) + 4*S*d)."""
    return LAYERS * s * (2.0 * (4 * D * D + 2 * D * FFN) + 4.0 * s * D)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


def synthetic_state_dict(device, layers=LAYERS, vocab=VOCAB, seed=1, std=0.02):
    """Seeded random-init weights of the `basic` architecture under the fairseq2 state-dict names
    (SURVEY §8(d) distributions), generated directly on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)

    def rn(*shape, s=std):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * s

    sd = {"encoder_frontend.embed.weight": rn(vocab, D, s=D ** -0.5)}
    for i in range(layers):
        p = f"encoder.layers.{i}."
        for name in ("q_proj", "k_proj", "v_proj", "output_proj"):
            sd[p + f"self_attn.{name}.weight"] = rn(D, D)
            sd[p + f"self_attn.{name}.bias"] = rn(D)
        sd[p + "self_attn_layer_norm.weight"] = 1.0 + rn(D)
        sd[p + "self_attn_layer_norm.bias"] = rn(D)
        sd[p + "ffn.inner_proj.weight"] = rn(FFN, D)
        sd[p + "ffn.inner_proj.bias"] = rn(FFN)
        sd[p + "ffn.output_proj.weight"] = rn(D, FFN)
        sd[p + "ffn.output_proj.bias"] = rn(D)
        sd[p + "ffn_layer_norm.weight"] = 1.0 + rn(D)
        sd[p + "ffn_layer_norm.bias"] = rn(D)
    sd["layer_norm.weight"] = 1.0 + rn(D)
    sd["layer_norm.bias"] = rn(D)
    return sd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,power.limit")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw, plim = [], None, set(), [], None
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            try:  # board power next to the clocks: the step runs under sw_power_cap, this says how close to the limit
                pw.append(float(f[3]))
                if len(f) > 8:
                    plim = float(f[8])
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None, "power_limit_w": plim}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def calibrate_cpu_threads() -> int:
    """Pick the torch intra-op thread count that runs an encoder-layer-shaped fp32 GEMM fastest on this host
    (more threads than physical cores / NUMA-local memory can be slower); a few seconds."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (cores, cores // 2, cores // 4, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    a = torch.randn(8192, 1024)
    w = torch.randn(8192, 1024)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.linear(a, w)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.linear(a, w)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_port_throughput(state_dict_cpu, target_seconds: float, sentences_per_step: int = 64):
    """Time the CPU restatement of the reference path (oracle) on a bounded sample of the workload."""
    from oracle.text_encoder import OracleEncoderConfig, OracleTextEncoder

    torch.set_float32_matmul_precision("high")  # reference precision_context for fp32 (text.py:36-54,255)
    cfg = OracleEncoderConfig(vocab_size=state_dict_cpu["encoder_frontend.embed.weight"].shape[0],
                              num_layers=LAYERS)
    enc = OracleTextEncoder(cfg, state_dict_cpu)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(4, cfg.vocab_size, (sentences_per_step, SEQ), generator=g)
    enc(ids[:2], None)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        emb, _ = enc(ids, None)
        n += sentences_per_step
        dt = time.perf_counter() - t0
        if dt >= target_seconds:
            break
    return n / dt, n, dt, emb, ids


def run_reference(args):
    """`--impl reference`: the reference's own (CPU, fp32) path -- fairseq2 cannot be installed here, so
    this is the oracle port of its op sequence -- on all host threads.  Rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    calibrate_cpu_threads()
    sd = synthetic_state_dict("cpu", vocab=VOCAB)
    from oracle.text_encoder import OracleEncoderConfig, OracleTextEncoder

    torch.set_float32_matmul_precision("high")
    enc = OracleTextEncoder(OracleEncoderConfig(vocab_size=VOCAB, num_layers=LAYERS), sd)
    per_step = 16
    ids = torch.randint(4, VOCAB, (per_step, SEQ), generator=torch.Generator().manual_seed(0))
    for _ in range(max(args.warmup, 1) if args.warmup < 3 else 3):
        enc(ids, None)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        enc(ids, None)
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "sentences/sec->1024-d", "value": val, "unit": "sentences/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"text_sonar_basic_encoder arch, batch {BATCH} x seq_len {SEQ} per GPU "
                               f"(each CPU step = a bounded sample of {per_step} sentences x {SEQ} tokens)"},
        "cpu_baseline": {"value": val, "unit": "sentences/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{per_step} sentences x {SEQ} tokens x {LAYERS} layers per step, fp32, "
                                   f"host cpu_count={cores}"},
        "e2e": {"value": val, "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--seq-len", type=int, default=SEQ)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cta-group", type=int, default=2)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__

    if rank == 0:
        __graft_entry__.build()
    if dist is not None:
        dist.barrier()
    from sonar_b200 import B200TextEncoderModel, SequenceBatch, ops, sonar_text_encoder_config

    B, S = args.batch, args.seq_len
    sd = synthetic_state_dict(dev)
    model = B200TextEncoderModel(sonar_text_encoder_config("basic"), sd, dev, cta_group=args.cta_group)
    sd_cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        sd_cpu = {k: v.cpu() for k, v in sd.items()}
    del sd
    torch.cuda.empty_cache()

    g = torch.Generator().manual_seed(1000 + rank)
    ids_host = torch.randint(4, VOCAB, (B, S), generator=g, dtype=torch.int64).pin_memory()
    ids_dev = ids_host.to(dev)
    batch_dev = SequenceBatch(ids_dev, None)  # all rows full length -> padding_mask None (utils.py:18-21)
    gather_buf = torch.empty((world * B, D), dtype=torch.float32, device=dev) if world > 1 else None

    def step_resident():
        out = model(batch_dev).sentence_embeddings
        if dist is not None:  # the one exchange step of the path: assemble [N,1024] on every rank
            dist.all_gather_into_tensor(gather_buf, out)
        return out

    out_host = torch.empty((B, D), dtype=torch.float32).pin_memory()

    def step_e2e():
        out = model(SequenceBatch(ids_host, None)).sentence_embeddings  # H2D of the ids inside
        if dist is not None:
            dist.all_gather_into_tensor(gather_buf, out)
        out_host.copy_(out, non_blocking=True)  # D2H of the result
        return out

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    # dominant kernel timed INSIDE the real steps: events recorded by the engine around the middle layer's FFN1 GEMM
    k_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    k_ev[0].record(); k_ev[1].record()  # materialise the handles
    model.profile_ffn1(*k_ev)
    total_ms, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    in_step_kernel_ms = k_ev[0].elapsed_time(k_ev[1])  # the last timed step's launch
    model.profile_ffn1(None, None)
    e2e_ms, _ = timed(step_e2e, args.steps, 1)
    model.check_inputs()
    value = world * B * args.steps / (total_ms / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    # ---- dominant kernel alone: tcgen05 GEMM, FFN inner-projection instantiation (bias+ReLU, bf16 out) ----
    peaks, peak_kind = load_peaks()
    roofline = None
    if rank == 0:
        T = B * S
        a = torch.randn((T, D), device=dev, dtype=torch.float32).to(torch.bfloat16)
        w = model._layer_bufs[0]["w1"]
        b1 = model._layer_bufs[0]["b1"]
        f = torch.empty((T, FFN), device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm_bf16(a, w, b1, epilogue="relu", out=f, cta_group=args.cta_group)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        k0.record()
        for _ in range(reps):
            ops.gemm_bf16(a, w, b1, epilogue="relu", out=f, cta_group=args.cta_group)
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / reps
        flops = 2.0 * T * FFN * D  # algorithmic FLOPs of one launch
        alone_tflops = flops / (kms / 1e3) / 1e12
        achieved = flops / (in_step_kernel_ms / 1e3) / 1e12  # the launch inside the last timed step
        peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
        burst = float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"]))
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_gemm_ffn1.json")
        if os.path.exists(tp) and (B, S) == (BATCH, SEQ):
            with open(tp) as fh:
                tj = json.load(fh)
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]  # per launch, from the committed ncu capture
        roofline = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel<cta_group,EPI_BIAS_RELU,bf16> "
                    f"M={T} N={FFN} K={D}", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "ms_per_launch_in_step": in_step_kernel_ms,
                    "timed_alone": {"achieved": alone_tflops, "peak": burst, "frac": alone_tflops / burst,
                                    "peak_source": f"{peak_kind} bf16_tflops (burst)", "ms_per_launch": kms},
                    "algorithmic_bytes": 2.0 * T * D + 2.0 * FFN * D + 4.0 * FFN + 2.0 * T * FFN,
                    "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside the long step)",
                    "whole_step_frac": (value / world) * flops_per_sentence(S) / 1e12 / peak}
        del a, f

    # ---- ragged variant: lengths U{16..128}; the engine packs tokens, the reference would pad to 128 ----
    ragged = None
    if rank == 0 and world == 1 and (B, S) == (BATCH, SEQ):
        from sonar_b200 import PaddingMask

        gl = torch.Generator().manual_seed(7)
        lens = torch.randint(16, 129, (B,), generator=gl)
        lens_list = lens.tolist()
        rb = SequenceBatch(ids_dev, PaddingMask(lens, S, lens_list))
        for _ in range(2):
            model(rb)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(3):
            model(rb)
        r1.record()
        torch.cuda.synchronize()
        rms = r0.elapsed_time(r1) / 3
        rflops = sum(flops_per_sentence(n) for n in lens_list)
        peak_s = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
        ragged = {"lengths": "U{16..128}, seed 7", "tokens": int(lens.sum()), "padded_tokens": B * S,
                  "value": B / rms * 1e3, "unit": "sentences/s", "ms_per_step": rms,
                  "roofline_frac_of_real_flops": rflops / (rms / 1e3) / 1e12 / peak_s}

    cpu_baseline = None
    if sd_cpu is not None:
        calibrate_cpu_threads()
        v, n, dt, emb_cpu, ids_cpu = cpu_port_throughput(sd_cpu, args.cpu_seconds)
        got = model(SequenceBatch(ids_cpu.to(dev), None)).sentence_embeddings.cpu().double()
        ref = emb_cpu.double()
        cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
        rel = (got - ref).norm(dim=1) / ref.norm(dim=1)
        cpu_baseline = {"value": v, "unit": "sentences/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} sentences x {S} tokens x {LAYERS} layers in {dt:.1f}s, fp32 torch CPU "
                                  f"restatement of the fairseq2 op sequence (oracle/text_encoder.py), "
                                  f"host cpu_count={os.cpu_count()}",
                        "parity_vs_gpu": {"one_minus_cos_max": float((1 - cos).max()),
                                          "rel_l2_max": float(rel.max()), "sentences": int(len(cos))}}

    if rank == 0:
        launches_per_step = 1 + LAYERS * 7 + 1
        line = {
            "metric": "sentences/sec->1024-d", "value": value, "unit": "sentences/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"text_sonar_basic_encoder arch (24L, d=1024, 16 heads, FFN 8192, vocab {VOCAB}), "
                                   f"batch {B} x seq_len {S} per GPU, random-init weights, synthetic ids",
                       "l2": "inputs larger than L2 (per-step activations ~15 GB vs 126 MB L2)",
                       "parallelism": f"dp{world}" + (" + all_gather of embeddings" if world > 1 else ""),
                       "cta_group": args.cta_group},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "sentences/s", "h2d_bytes_per_step": B * S * 8,
                    "d2h_bytes_per_step": B * D * 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "ragged": ragged,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
