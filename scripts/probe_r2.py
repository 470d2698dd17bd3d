"""Round-2 micro-benchmarks (CUDA events, kernels launched back to back, inputs >> L2):
    python scripts/probe_r2.py [attention] [lnfold] [predict]
attention: tcgen05 pipelined kernel vs the mma.sync kernel at 4096 x 128, a ragged batch and long sentences.
lnfold   : (residual GEMM + LayerNorm kernel + consumer GEMM) vs (residual+stats GEMM + folded consumer GEMM) at the
           encoder's shapes.
predict  : where the wall time of TextToEmbeddingModelPipeline.predict goes (2-layer model so the host side dominates)."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sonar_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1:] or ["attention", "lnfold"]
g = torch.Generator(device=dev).manual_seed(0)
HBM = 6572.2


def timed(fn, iters=5, warm=2):
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


if "attention" in which:
    D = 1024
    for name, lens in (("4096 x 128", [128] * 4096),
                       ("ragged U{16..128} x 4096", torch.randint(16, 129, (4096,), generator=torch.Generator().manual_seed(7)).tolist()),
                       ("256 x 514", [514] * 256), ("1024 x 200", [200] * 1024)):
        T = sum(lens)
        qkv = torch.randn((T, 3 * D), device=dev, generator=g).to(torch.bfloat16)
        cu = ops.cu_seqlens_of(lens).to(dev)
        mx = max(lens)
        t_tc = timed(lambda: ops.attention(qkv, cu, mx, 16, impl="tcgen05"))
        t_ms = timed(lambda: ops.attention(qkv, cu, mx, 16, impl="mma_sync"))
        byts = T * D * 2 * 4  # q, k, v read + out written
        print(f"attention {name:28s} T={T:7d}: tcgen05 {t_tc:7.3f} ms ({byts / t_tc / 1e6:6.0f} GB/s = {byts / t_tc / 1e6 / HBM:.2f} of HBM) | "
              f"mma.sync {t_ms:7.3f} ms", flush=True)
        del qkv

if "lnfold" in which:
    T, D, F = 4096 * 128, 1024, 8192
    x = torch.randn((T, D), device=dev, generator=g)
    gamma = 1 + 0.02 * torch.randn(D, device=dev, generator=g)
    beta = 0.02 * torch.randn(D, device=dev, generator=g)
    a1 = torch.randn((T, D), device=dev, generator=g).to(torch.bfloat16)
    for name, K in (("out-proj K=1024", D), ("FFN2 K=8192", F)):
        a = a1 if K == D else torch.randn((T, K), device=dev, generator=g).to(torch.bfloat16)
        w = (torch.randn((D, K), device=dev, generator=g) / math.sqrt(K)).to(torch.bfloat16)
        b = torch.randn(D, device=dev, generator=g) * 0.02
        t_acc = timed(lambda: ops.gemm_bf16(a, w, b, epilogue="residual", residual=x, out=x))
        t_ln = timed(lambda: ops.layernorm(x, gamma, beta))
        t_st = timed(lambda: ops.gemm_residual_stats(a, w, b, x))
        fl = 2.0 * T * D * K
        print(f"{name:16s}: reduce-add GEMM {t_acc:6.3f} ms ({fl / t_acc / 1e9:5.0f} TF/s) + LayerNorm {t_ln:6.3f} ms = {t_acc + t_ln:6.3f} | "
              f"residual+stats GEMM {t_st:6.3f} ms ({fl / t_st / 1e9:5.0f} TF/s)", flush=True)
        del a, w
    h, stats = ops.gemm_residual_stats(a1, (torch.randn((D, D), device=dev, generator=g) / 32).to(torch.bfloat16),
                                       torch.zeros(D, device=dev), x)
    for name, N, relu in (("QKV N=3072", 3 * D, False), ("FFN1 N=8192", F, True)):
        w = (torch.randn((N, D), device=dev, generator=g) / 32).to(torch.bfloat16)
        b = torch.randn(N, device=dev, generator=g) * 0.02
        wf, cs, bf = ops.fold_layernorm(w, b, gamma, beta)
        out = torch.empty((T, N), device=dev, dtype=torch.bfloat16)
        t_pl = timed(lambda: ops.gemm_bf16(h, w, b, epilogue="relu" if relu else "bias", out=out))
        del out
        t_fd = timed(lambda: ops.gemm_ln_consumer(h, wf, bf, cs, stats, 1e-5, relu=relu))
        fl = 2.0 * T * N * D
        print(f"{name:16s}: plain GEMM {t_pl:6.3f} ms ({fl / t_pl / 1e9:5.0f} TF/s) | LayerNorm-folded GEMM {t_fd:6.3f} ms "
              f"({fl / t_fd / 1e9:5.0f} TF/s)", flush=True)
        del w, wf

if "predict" in which:
    import bench
    from sonar_b200 import B200TextEncoderModel, VocabularyInfo, sonar_text_encoder_config
    from sonar_b200.inference_pipelines import TextToEmbeddingModelPipeline
    from sonar_b200.tokenizer import SyntheticTokenizer

    for layers in (2, 24):
        sd = bench.synthetic_state_dict(dev, layers=layers)
        model = B200TextEncoderModel(sonar_text_encoder_config("basic", num_encoder_layers=layers), sd, dev)
        pipe = TextToEmbeddingModelPipeline(model, SyntheticTokenizer(vocab_size=bench.VOCAB), device=dev)
        sents = bench.synthetic_sentences(32768, 126)
        pipe.predict(sents[:8192], "eng_Latn", batch_size=4096, target_device="cpu")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.predict(sents, "eng_Latn", batch_size=4096, target_device="cpu")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ids = torch.randint(4, bench.VOCAB, (4096, 128), device=dev)
        from sonar_b200 import SequenceBatch
        t_gpu = timed(lambda: model(SequenceBatch(ids, None)), iters=3, warm=1)
        print(f"predict {layers:2d} layers: 32768 sentences in {dt:.2f} s = {32768 / dt:.0f} sent/s; GPU alone {t_gpu:.1f} ms per 4096 "
              f"-> {8 * t_gpu / 1e3:.2f} s for 8 batches", flush=True)
        del model, pipe, sd
        torch.cuda.empty_cache()

if "decstep" in which:
    # decoder step time against the position, with every hypothesis on its own KV-cache rows (identity ancestry) and with
    # the five hypotheses of a sentence sharing one ancestor chain (what beam search converges to): how much of the step
    # is KV-cache traffic, and how much of it L2 already removes when the rows are shared
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from bench import synthetic_decoder_state_dict  # noqa: E402
    from sonar_b200 import B200TextDecoderModel, sonar_text_decoder_config  # noqa: E402
    model = B200TextDecoderModel(sonar_text_decoder_config("basic"), synthetic_decoder_state_dict(dev), dev)
    n, beam, tmax = 512, 5, 130
    model.begin(torch.randn((n, 1024), device=dev) * 0.25, beam, tmax)
    r = n * beam
    ident = torch.arange(r, dtype=torch.int32, device=dev)[:, None].expand(r, tmax).contiguous()
    shared = ((torch.arange(r, dtype=torch.int32, device=dev) // beam) * beam)[:, None].expand(r, tmax).contiguous()
    tk = torch.randint(4, 256000, (r,), device=dev)
    for t in (8, 64, 120):
        a = timed(lambda: model.step(tk, ident, t), iters=8, warm=3)
        b = timed(lambda: model.step(tk, shared, t), iters=8, warm=3)
        print(f"decoder step 2560 rows, t={t:3d}: own rows {a:.3f} ms   shared ancestors {b:.3f} ms")
