"""Bring-up / diagnosis probe for the CUDA kernels (run under gpurun, one stage per process so a
trapped kernel cannot poison later stages):

    python scripts/gpu_probe.py elementwise | gemm1 | gemm2 | perf | encoder
"""

import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sonar_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")


def rnd(shape, scale, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(DEV, dtype)


def err_map(out, ref, rb=32, cb=64):
    """coarse map of where the result is wrong: fraction of bad elements per (row block, col block)."""
    bad = ((out - ref).abs() > (ref.abs() * 0.02 + 0.05)).float()
    m, n = bad.shape
    mm, nn = (m // rb) * rb, (n // cb) * cb
    if mm == 0 or nn == 0:
        return bad.mean().item()
    blk = bad[:mm, :nn].view(mm // rb, rb, nn // cb, cb).mean(dim=(1, 3))
    return blk


IMPL = "tcgen05"


def stage_attn_tc():
    """first contact with the tcgen05 attention kernel: small cases, printed errors"""
    h, d = 16, 1024
    for lens in ([128], [128] * 3, [64], [1, 17, 64, 65, 100, 128], [128] * 600):
        t = sum(lens)
        qkv = rnd((t, 3 * d), 1.0, 5, torch.bfloat16)
        cu = ops.cu_seqlens_of(lens).to(DEV)
        o = ops.attention(qkv, cu, max(lens), h, impl="tcgen05")
        r = ops.attention(qkv, cu, max(lens), h, impl="mma_sync")
        torch.cuda.synchronize()
        e = (o.float() - r.float()).abs()
        print(f"attn_tc lens={lens[:6]}{'...' if len(lens) > 6 else ''} max diff vs mma_sync {e.max().item():.4g} "
              f"nan={bool(torch.isnan(o.float()).any())}", flush=True)
        if e.max().item() > 0.05:
            rows = (e.max(1).values > 0.05).nonzero().flatten()[:10].tolist()
            cols = (e.max(0).values > 0.05).nonzero().flatten()[:16].tolist()
            print("  bad rows", rows, "bad cols", cols, "\n  got", o[rows[0], :8].float().tolist(), "\n  ref", r[rows[0], :8].float().tolist())


def stage_elementwise():
    d = 1024
    x = rnd((300, d), 3.0, 1) + 0.5
    g, b = 1 + rnd((d,), 0.1, 2), rnd((d,), 0.1, 3)
    y = ops.layernorm(x, g, b)
    ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-5)
    print("layernorm max err", (y.float() - ref).abs().max().item())
    lens = [5, 128, 1, 77]
    xx = rnd((sum(lens), d), 2.0, 4)
    cu = ops.cu_seqlens_of(lens).to(DEV)
    out = ops.pool_packed(xx, cu, "mean", gamma=g, beta=b)
    yy = torch.nn.functional.layer_norm(xx, (d,), g, b, 1e-5)
    refp = torch.stack([yy[int(cu[i]):int(cu[i + 1])].mean(0) for i in range(len(lens))])
    print("ln_pool max err", (out - refp).abs().max().item())
    # attention
    h = 16
    for lens in ([128] * 3, [1, 17, 64, 65, 130, 514]):
        t = sum(lens)
        qkv = rnd((t, 3 * d), 1.0, 5, torch.bfloat16)
        cu = ops.cu_seqlens_of(lens).to(DEV)
        o = ops.attention(qkv, cu, max(lens), h, impl=IMPL if max(lens) <= 128 else "auto")
        torch.cuda.synchronize()
        s0, worst = 0, 0.0
        for n in lens:
            blk = qkv[s0:s0 + n].float()
            q, k, v = (blk[:, i * d:(i + 1) * d].view(n, h, 64).transpose(0, 1) for i in range(3))
            r = torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1).reshape(n, d)
            worst = max(worst, (o[s0:s0 + n].float() - r).abs().max().item())
            s0 += n
        print("attention lens", lens, "max err", worst)


def stage_gemm(cg):
    shapes = [(128, 256, 64), (128, 256, 128), (256, 256, 64), (256, 512, 256), (300, 512, 192),
              (1000, 1024, 1024), (4096, 3072, 1024), (777, 1024, 8192), (8192, 8192, 1024)]
    for (m, n, k) in shapes:
        a = rnd((m, k), 1.0, 1, torch.bfloat16)
        w = rnd((n, k), 1 / math.sqrt(k), 2, torch.bfloat16)
        bias = rnd((n,), 0.5, 3)
        ref = a.float() @ w.float().T + bias
        for od in (torch.bfloat16, torch.float32):
            out = ops.gemm_bf16(a, w, bias, epilogue="bias", out_dtype=od, cta_group=cg)
            torch.cuda.synchronize()
            e = (out.float() - ref).abs().max().item()
            ok = e < 0.1
            print(f"gemm cg={cg} {m}x{n}x{k} out={str(od)[6:]} max_err={e:.4g} {'OK' if ok else 'WRONG'}", flush=True)
            if not ok:
                em = err_map(out.float(), ref)
                torch.set_printoptions(linewidth=200, precision=2)
                print("bad-fraction map (rows=32-row blocks, cols=64-col blocks):\n", em)
                print("out[0,:8]", out[0, :8].float().tolist(), "\nref[0,:8]", ref[0, :8].tolist())
                return
    # epilogues
    m, n, k = 1500, 1024, 1024
    a = rnd((m, k), 1.0, 4, torch.bfloat16)
    w = rnd((n, k), 1 / math.sqrt(k), 5, torch.bfloat16)
    bias = rnd((n,), 0.5, 6)
    out = ops.gemm_bf16(a, w, bias, epilogue="relu", cta_group=cg)
    print("relu max err", (out.float() - torch.relu(a.float() @ w.float().T + bias)).abs().max().item())
    x = rnd((m, n), 2.0, 7)
    ref = x + a.float() @ w.float().T + bias
    ops.gemm_bf16(a, w, bias, epilogue="residual", residual=x, out=x, cta_group=cg)
    print("residual fp32 in-place max err", (x - ref).abs().max().item())


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2], ts[0]


def stage_perf():
    m = 131072
    for (n, k, epi) in [(3072, 1024, "bias"), (1024, 1024, "residual"), (8192, 1024, "relu"), (1024, 8192, "residual")]:
        a = rnd((m, k), 1.0, 1, torch.bfloat16)
        w = rnd((n, k), 1 / math.sqrt(k), 2, torch.bfloat16)
        bias = rnd((n,), 0.5, 3)
        fl = 2.0 * m * n * k
        od = torch.float32 if epi == "residual" else torch.bfloat16
        out = torch.empty((m, n), dtype=od, device=DEV)
        res = out if epi == "residual" else None
        for cg in (1, 2):
            med, best = timeit(lambda: ops.gemm_bf16(a, w, bias, epilogue=epi, residual=res, out=out, cta_group=cg))
            print(f"gemm cg={cg} M={m} N={n} K={k} {epi}: median {med:.3f} ms  {fl / med / 1e9:.1f} TFLOP/s (best {fl / best / 1e9:.1f})", flush=True)
        med, best = timeit(lambda: torch.matmul(a, w.T))
        print(f"cuBLAS          M={m} N={n} K={k}: median {med:.3f} ms  {fl / med / 1e9:.1f} TFLOP/s (best {fl / best / 1e9:.1f})", flush=True)
        del a, w, out
    t, d, h = 131072, 1024, 16
    x = rnd((t, d), 1.0, 1)
    g, b = 1 + rnd((d,), 0.1, 2), rnd((d,), 0.1, 3)
    med, _ = timeit(lambda: ops.layernorm(x, g, b))
    print(f"layernorm T={t}: {med:.3f} ms  {t * d * 6 / med / 1e6:.0f} GB/s")
    qkv = rnd((t, 3 * d), 1.0, 5, torch.bfloat16)
    cu = ops.cu_seqlens_of([128] * (t // 128)).to(DEV)
    for impl in ("mma_sync", "tcgen05"):
        med, _ = timeit(lambda: ops.attention(qkv, cu, 128, h, impl=impl))
        print(f"attention[{impl}] T={t} S=128: {med:.3f} ms  {t * d * 8 / med / 1e6:.0f} GB/s  {4.0 * t * 128 * d / med / 1e9:.1f} TFLOP/s")
    lens = torch.randint(16, 129, (t // 72,), generator=torch.Generator().manual_seed(0)).tolist()
    tt = sum(lens)
    qkv2 = rnd((tt, 3 * d), 1.0, 6, torch.bfloat16)
    cu2 = ops.cu_seqlens_of(lens).to(DEV)
    for impl in ("mma_sync", "tcgen05"):
        med, _ = timeit(lambda: ops.attention(qkv2, cu2, 128, h, impl=impl))
        print(f"attention[{impl}] ragged U(16..128) T={tt}: {med:.3f} ms  {tt * d * 8 / med / 1e6:.0f} GB/s")


def stage_encoder():
    from oracle.text_encoder import OracleEncoderConfig, OracleTextEncoder, make_synthetic_state_dict
    from sonar_b200 import B200TextEncoderModel, PaddingMask, SequenceBatch, VocabularyInfo, sonar_text_encoder_config
    from tests.helpers import parity_metrics

    V = 4096
    for layers in (1, 4):
        ocfg = OracleEncoderConfig(vocab_size=V, num_layers=layers)
        sd = make_synthetic_state_dict(ocfg, seed=1)
        cfg = sonar_text_encoder_config("basic", num_encoder_layers=layers,
                                        vocab_info=VocabularyInfo(size=V, unk_idx=1, bos_idx=2, eos_idx=3, pad_idx=1))
        model = B200TextEncoderModel(cfg, sd, DEV)
        oracle = OracleTextEncoder(ocfg, sd)
        lens = [64, 1, 17, 33, 48, 5, 63, 31]
        g = torch.Generator().manual_seed(0)
        ids = torch.zeros((len(lens), 64), dtype=torch.int64)
        for i, n in enumerate(lens):
            ids[i, :n] = torch.randint(4, V, (n,), generator=g)
        ref, _ = oracle(ids, torch.tensor(lens))
        out = model(SequenceBatch(ids.to(DEV), PaddingMask(torch.tensor(lens), 64, lens))).sentence_embeddings
        torch.cuda.synchronize()
        print(f"encoder {layers} layers:", parity_metrics(out, ref), flush=True)


def stage_xsim():
    import numpy as np
    from oracle import xsim as ox
    from sonar_b200 import xsim
    g = torch.Generator().manual_seed(0)
    for (n, m, d) in [(300, 1000, 1024), (1000, 777, 1024), (4096, 8192, 1024)]:
        y = torch.randn((m, d), generator=g)
        x = torch.randn((n, d), generator=g)
        val, idx = xsim.knn(x.to(DEV), y.to(DEV), 4)
        torch.cuda.synchronize()
        rv, ri = ox.knn(x.numpy(), y.numpy(), 4)
        same = np.array_equal(idx.cpu().numpy(), ri)
        print(f"xsim knn {n}x{m}: idx equal={same} max val err={np.abs(val.cpu().numpy() - rv).max():.3e}", flush=True)
        if not same:
            bad = (idx.cpu().numpy() != ri).any(1)
            print("  bad rows", bad.sum(), "first", np.nonzero(bad)[0][:5], idx.cpu().numpy()[bad][:3], ri[bad][:3])
    n = m = 131072
    x = torch.randn((n, 1024), device=DEV)
    y = torch.randn((m, 1024), device=DEV)
    med, best = timeit(lambda: xsim.knn(x, y, 4), iters=3, warm=1)
    print(f"xsim knn {n}x{m}: {med:.2f} ms  {n * m / med / 1e6:.2f} Gpairs/s  GEMM-equivalent {2.0 * n * m * 1024 / med / 1e9:.0f} TFLOP/s")


def stage_decoder():
    from oracle.text_decoder import OracleDecoderConfig, OracleTextDecoder, make_synthetic_decoder_state_dict
    from sonar_b200 import B200TextDecoderModel, VocabularyInfo, sonar_text_decoder_config
    V = 4096
    ocfg = OracleDecoderConfig(vocab_size=V, num_layers=2, max_seq_len=64)
    sd = make_synthetic_decoder_state_dict(ocfg, seed=2)
    cfg = sonar_text_decoder_config("basic", num_decoder_layers=2, max_seq_len=64,
                                    vocab_info=VocabularyInfo(size=V, unk_idx=1, bos_idx=2, eos_idx=3, pad_idx=1))
    model = B200TextDecoderModel(cfg, sd, DEV)
    oracle = OracleTextDecoder(ocfg, sd)
    n, beam, steps = 3, 2, 5
    emb = torch.randn((n, 1024), generator=torch.Generator().manual_seed(0)) * 0.25
    toks = torch.randint(4, V, (n * beam, steps), generator=torch.Generator().manual_seed(1))
    model.begin(emb.to(DEV), beam, 16)
    r = n * beam
    table = torch.arange(r, dtype=torch.int32, device=DEV)[:, None].expand(r, 16).contiguous()
    enc_rows = emb[:, None, :].repeat_interleave(beam, 0)
    for t in range(steps):
        lp, tok, eos_lp = model.step(toks[:, t].contiguous().to(DEV), table, t)
        torch.cuda.synchronize()
        ref = oracle.step_lprobs(toks[:, :t + 1], enc_rows)
        d = (lp.cpu() - torch.gather(ref, 1, tok.cpu().long())).abs().max().item()
        top1 = (tok.cpu()[:, 0].long() == ref.argmax(1)).float().mean().item()
        print(f"decoder step {t}: max |lprob - oracle| = {d:.4f}  eos err {(eos_lp.cpu() - ref[:, 3]).abs().max().item():.4f} top1 match {top1:.2f}", flush=True)
    # full-size step timing (config 4: 512 sentences x beam 5, 24 layers, vocab 256206)
    from bench import synthetic_state_dict  # noqa
    import time
    full = sonar_text_decoder_config("basic")
    g = torch.Generator(device=DEV).manual_seed(3)
    sdf = {}
    def rn(*shape, s=0.02):
        return torch.randn(*shape, generator=g, device=DEV) * s
    sdf["decoder_frontend.embed.weight"] = rn(256206, 1024, s=1 / 32)
    for i in range(24):
        p = f"decoder.layers.{i}."
        for a in ("self_attn", "encoder_decoder_attn"):
            for nme in ("q_proj", "k_proj", "v_proj", "output_proj"):
                sdf[p + f"{a}.{nme}.weight"] = rn(1024, 1024); sdf[p + f"{a}.{nme}.bias"] = rn(1024)
            sdf[p + f"{a}_layer_norm.weight"] = 1 + rn(1024); sdf[p + f"{a}_layer_norm.bias"] = rn(1024)
        sdf[p + "ffn.inner_proj.weight"] = rn(8192, 1024); sdf[p + "ffn.inner_proj.bias"] = rn(8192)
        sdf[p + "ffn.output_proj.weight"] = rn(1024, 8192); sdf[p + "ffn.output_proj.bias"] = rn(1024)
        sdf[p + "ffn_layer_norm.weight"] = 1 + rn(1024); sdf[p + "ffn_layer_norm.bias"] = rn(1024)
    sdf["decoder.layer_norm.weight"] = 1 + rn(1024); sdf["decoder.layer_norm.bias"] = rn(1024)
    big = B200TextDecoderModel(full, sdf, DEV)
    del sdf
    n, beam, tmax = 512, 5, 130
    big.begin(torch.randn((n, 1024), device=DEV) * 0.25, beam, tmax)
    r = n * beam
    table = torch.arange(r, dtype=torch.int32, device=DEV)[:, None].expand(r, tmax).contiguous()
    tk = torch.randint(4, 256000, (r,), device=DEV)
    for t in (0, 1, 2):
        big.step(tk, table, t)
    torch.cuda.synchronize()
    for t0 in (3, 64, 120):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(t0, t0 + 4):
            big.step(tk, table, t)
        e1.record()
        torch.cuda.synchronize()
        print(f"full-size decoder step at t~{t0}: {e0.elapsed_time(e1) / 4:.2f} ms/step for {r} hypotheses "
              f"({r / (e0.elapsed_time(e1) / 4) * 1e3:.0f} hyp-tokens/s)", flush=True)


def stage_speech():
    from oracle.speech_encoder import OracleSpeechConfig, OracleSpeechEncoder, make_synthetic_speech_state_dict
    from oracle.speech_frontend import collate_fbank, waveform_to_fbank
    from sonar_b200 import B200SpeechEncoderModel, PaddingMask, SequenceBatch, sonar_speech_encoder_config
    from sonar_b200.speech_frontend import WaveformToFbank
    from tests.helpers import parity_metrics
    g = torch.Generator().manual_seed(3)
    waves = [(torch.randn(n, generator=g) * 0.1).clamp(-1, 1) for n in (16000, 9000, 48000)]
    out, frames = WaveformToFbank(DEV)([w.to(DEV) for w in waves])
    ref, rl = collate_fbank([waveform_to_fbank(w) for w in waves])
    print("fbank frames", frames, rl, "max err", (out.cpu() - ref).abs().max().item(), flush=True)
    frames = [300, 131, 64, 2]
    fb = torch.zeros((len(frames), 300, 80))
    for i, n in enumerate(frames):
        fb[i, :n] = torch.randn((n, 80), generator=g)
    for (nl, npool) in ((0, 0), (1, 0), (1, 1), (2, 2)):
        ocfg = OracleSpeechConfig(num_layers=nl, pooler_layers=npool)
        sd = make_synthetic_speech_state_dict(ocfg, seed=3)
        model = B200SpeechEncoderModel(sonar_speech_encoder_config("english", num_encoder_layers=nl, num_decoder_layers=npool), sd, DEV)
        oracle = OracleSpeechEncoder(ocfg, sd)
        remb, renc, lens = oracle(fb, frames)
        model.return_encoded_seqs = True
        o = model(SequenceBatch(fb.to(DEV), PaddingMask(torch.tensor(frames), 300, frames)))
        torch.cuda.synchronize()
        st, worst = 0, 0.0
        for i, n in enumerate(lens):
            got, exp = o.encoded_seqs[st:st + n].cpu(), renc[i, :n]
            worst = max(worst, float((got - exp).norm() / exp.norm()))
            st += n
        print(f"speech layers={nl} pooler={npool}: encoded rel-L2 max {worst:.4g}; emb", parity_metrics(o.sentence_embeddings, remb), flush=True)
    # timing at config-3 scale: 64 x 10 s utterances, full 24+3 layers
    ocfg = OracleSpeechConfig()
    sd = make_synthetic_speech_state_dict(ocfg, seed=3)
    model = B200SpeechEncoderModel(sonar_speech_encoder_config("english"), sd, DEV)
    wv = [(torch.randn(160000, generator=g) * 0.05).clamp(-1, 1).to(DEV) for _ in range(64)]
    conv = WaveformToFbank(DEV)
    def run():
        fbk, fr = conv(wv)
        return model(SequenceBatch(fbk, PaddingMask(torch.tensor(fr), fbk.shape[1], fr))).sentence_embeddings
    med, best = timeit(run, iters=3, warm=1)
    print(f"speech encoder 64 x 10 s (fbank + 24 Conformer + 3 pooler layers): {med:.1f} ms -> {64 / med * 1e3:.0f} utt/s", flush=True)


if __name__ == "__main__":
    stage = sys.argv[1]
    t0 = time.time()
    print(f"== stage {stage} on {torch.cuda.get_device_name(0)}", flush=True)
    {"elementwise": stage_elementwise, "gemm1": lambda: stage_gemm(1), "gemm2": lambda: stage_gemm(2),
     "perf": stage_perf, "encoder": stage_encoder, "xsim": stage_xsim, "attn_tc": stage_attn_tc, "decoder": stage_decoder, "speech": stage_speech}[stage]()
    torch.cuda.synchronize()
    print(f"== stage {stage} done in {time.time() - t0:.1f}s", flush=True)
