"""Measurements for the secondary configurations of BASELINE.json (configs 3, 4, 5) on ONE B200 -- the headline line
stays `bench.py`.  Prints one JSON line per configuration with the device-timed metric, the roofline arithmetic of
BASELINE.md §3 and the CPU oracle timed beside it on a bounded sample.

    python scripts/bench_secondary.py [speech] [decoder] [xsim]
"""

import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import calibrate_cpu_threads, load_peaks  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_speech(peaks):
    from oracle.speech_encoder import OracleSpeechConfig, OracleSpeechEncoder, make_synthetic_speech_state_dict
    from oracle.speech_frontend import collate_fbank, waveform_to_fbank
    from sonar_b200 import B200SpeechEncoderModel, PaddingMask, SequenceBatch, sonar_speech_encoder_config
    from sonar_b200.speech_frontend import WaveformToFbank

    n = 256
    ocfg = OracleSpeechConfig()
    sd = make_synthetic_speech_state_dict(ocfg, seed=3)
    model = B200SpeechEncoderModel(sonar_speech_encoder_config("english"), sd, DEV)
    g = torch.Generator().manual_seed(0)
    waves = [(torch.randn(160000, generator=g) * 0.05).clamp(-1, 1) for _ in range(n)]
    wd = [w.to(DEV) for w in waves]
    conv = WaveformToFbank(DEV)

    def run():
        fb, fr = conv(wd)
        return model(SequenceBatch(fb, PaddingMask(torch.tensor(fr), fb.shape[1], fr))).sentence_embeddings

    ms = timed(run)
    flop_per_utt = 499 * 24 * 52.38e6 + 7e9
    peak = float(peaks["bf16_tflops_sustained"])
    val = n / ms * 1e3
    # CPU oracle on 2 utterances
    calibrate_cpu_threads()
    oracle = OracleSpeechEncoder(ocfg, sd)
    t0 = time.perf_counter()
    fb, fl = collate_fbank([waveform_to_fbank(w) for w in waves[:2]])
    ref, _, _ = oracle(fb, fl)
    dt = time.perf_counter() - t0
    got = run()[:2].cpu().double()
    cos = torch.nn.functional.cosine_similarity(got, ref.double(), dim=1)
    return {"config": "sonar_speech_encoder_eng arch: 256 x 10 s 16 kHz synthetic waveforms, fbank + 24 Conformer + 3 pooler layers",
            "metric": "utterances/sec->1024-d", "value": val, "unit": "utterances/s", "ms_per_step": ms,
            "roofline": {"bound": "tensor", "achieved": val * flop_per_utt / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": val * flop_per_utt / 1e12 / peak},
            "cpu_baseline": {"value": 2 / dt, "unit": "utterances/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"2 utterances in {dt:.1f}s (fp32 torch oracle incl. fbank)",
                             "parity_vs_gpu": {"one_minus_cos_max": float((1 - cos).max())}}}


def bench_decoder(peaks):
    from sonar_b200 import B200TextDecoderModel, sonar_text_decoder_config
    from sonar_b200.generation import BeamSearchSeq2SeqGenerator

    g = torch.Generator(device=DEV).manual_seed(3)
    sd = {}

    def rn(*shape, s=0.02):
        return torch.randn(*shape, generator=g, device=DEV) * s

    sd["decoder_frontend.embed.weight"] = rn(256206, 1024, s=1 / 32)
    for i in range(24):
        p = f"decoder.layers.{i}."
        for a in ("self_attn", "encoder_decoder_attn"):
            for nme in ("q_proj", "k_proj", "v_proj", "output_proj"):
                sd[p + f"{a}.{nme}.weight"], sd[p + f"{a}.{nme}.bias"] = rn(1024, 1024), rn(1024)
            sd[p + f"{a}_layer_norm.weight"], sd[p + f"{a}_layer_norm.bias"] = 1 + rn(1024), rn(1024)
        sd[p + "ffn.inner_proj.weight"], sd[p + "ffn.inner_proj.bias"] = rn(8192, 1024), rn(8192)
        sd[p + "ffn.output_proj.weight"], sd[p + "ffn.output_proj.bias"] = rn(1024, 8192), rn(1024)
        sd[p + "ffn_layer_norm.weight"], sd[p + "ffn_layer_norm.bias"] = 1 + rn(1024), rn(1024)
    sd["decoder.layer_norm.weight"], sd["decoder.layer_norm.bias"] = 1 + rn(1024), rn(1024)
    model = B200TextDecoderModel(sonar_text_decoder_config("basic"), sd, DEV)
    del sd
    n, beam, max_seq_len = 512, 5, 128
    emb = torch.randn((n, 1024), device=DEV) * 0.25 / math.sqrt(1024) * 32
    prompt = torch.tensor([3, 256100])
    runs = {}
    for label, flag in (("eager", False), ("cuda_graphs", True)):
        gen = BeamSearchSeq2SeqGenerator(model, beam_size=beam, max_seq_len=max_seq_len, pad_idx=0, cuda_graphs=flag)
        walls = []
        for _ in range(3):  # the first call allocates the 32 GB KV cache / records the graphs
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = gen(emb, None, prompt, None)
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)
        runs[label] = walls
    dt = min(runs["eager"][1:] + runs["cuda_graphs"][1:])
    steps = max(len(h[0].seq) for h in out.hypotheses if h)
    # the pipelines' default batch (5 sentences x beam 5 = 25 hypothesis rows): launch-bound unless the step is a CUDA graph
    small = {}
    emb5 = emb[:5].contiguous()
    for label, flag, fused in (("eager_torch_beam_ops", False, False), ("eager", False, True), ("cuda_graphs", True, True)):
        gsm = BeamSearchSeq2SeqGenerator(model, beam_size=beam, max_seq_len=max_seq_len, pad_idx=0, cuda_graphs=flag,
                                         fused_beam_step=fused)
        gsm(emb5, None, prompt, None)  # warm-up (records the graphs)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        o5 = gsm(emb5, None, prompt, None)
        torch.cuda.synchronize()
        d5 = time.perf_counter() - t1
        st5 = max(len(h[0].seq) for h in o5.hypotheses if h)
        small[label] = {"wall_s": d5, "steps": st5, "ms_per_step": d5 / st5 * 1e3}
    hyp_tokens = n * beam * steps
    peak = float(peaks["bf16_tflops_sustained"])
    return {"config": f"text_sonar_basic_decoder arch: {n} embeddings, beam {beam}, max_seq_len {max_seq_len}, random weights "
                      f"(random-weight hypotheses rarely emit EOS early: {steps} steps ran)",
            "metric": "sentences/sec decoded", "value": n / dt, "unit": "sentences/s", "wall_s": dt, "steps": steps,
            "ms_per_step": dt / steps * 1e3, "hypothesis_tokens_per_s": hyp_tokens / dt,
            "wall_s_by_mode_3_calls_each": runs, "batch5_beam5": small,
            "roofline": {"bound": "tensor", "achieved": hyp_tokens * 1.63e9 / dt / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": hyp_tokens * 1.63e9 / dt / 1e12 / peak,
                         "note": "1.63 GFLOP per hypothesis-token (BASELINE.md §3); KV-cache reads (4 KB x t per hypothesis-layer) "
                                 "are the second bound"}}


def bench_xsim(peaks):
    import numpy as np

    from oracle import xsim as ox
    from sonar_b200 import xsim

    n = m = 262144
    g = torch.Generator(device=DEV).manual_seed(0)
    y = torch.randn((m, 1024), generator=g, device=DEV)
    x = y + 0.1 * torch.randn((n, 1024), generator=g, device=DEV) * y.norm(dim=1, keepdim=True) / 32.0  # §8(d) config 5
    ms = timed(lambda: xsim.knn(x, y, 4), iters=2, warm=1)
    err, _, pred = xsim.xsim(x[:65536], y[:65536], margin="ratio", k=4)
    peak = float(peaks["bf16_tflops_sustained"])
    pairs = n * m / ms * 1e3
    # CPU oracle (fp64 numpy) on a 4096 x 16384 slice, and index parity on it
    xs, ys = x[:4096].cpu().numpy(), y[:16384].cpu().numpy()
    t0 = time.perf_counter()
    rv, ri = ox.knn(xs, ys, 4)
    dt = time.perf_counter() - t0
    gv, gi = xsim.knn(x[:4096], y[:16384], 4)
    return {"config": f"xsim k-NN (k=4) of [{n},1024] x [{m},1024] noisy copies on 1 GPU (one direction)",
            "metric": "xsim pairs/sec", "value": pairs, "unit": "pairs/s", "ms_per_step": ms,
            "xsim_error_64k_ratio_margin": err,
            "roofline": {"bound": "tensor", "achieved": 2.0 * n * m * 1024 / ms / 1e9, "peak": peak, "unit": "TFLOP/s",
                         "frac": 2.0 * n * m * 1024 / ms / 1e9 / peak},
            "cpu_baseline": {"value": 4096 * 16384 / dt, "unit": "pairs/s", "kind": "port",
                             "sample": f"4096 x 16384 float64 numpy oracle in {dt:.1f}s",
                             "top4_indices_identical": bool(np.array_equal(gi.cpu().numpy(), ri))}}


if __name__ == "__main__":
    which = sys.argv[1:] or ["speech", "decoder", "xsim"]
    peaks, _ = load_peaks()
    for name in which:
        fn = {"speech": bench_speech, "decoder": bench_decoder, "xsim": bench_xsim}[name]
        print(json.dumps({"name": name, **fn(peaks)}), flush=True)
        torch.cuda.empty_cache()
