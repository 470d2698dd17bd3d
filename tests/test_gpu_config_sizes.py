"""Parity at the BASELINE.json configuration SIZES: the kernels take different tile schedules at
524 288 tokens / 2 560 hypothesis rows than at the toy sizes of the other GPU tests, so the full-size launches are checked
against the fp32 CPU oracle on row subsets the oracle finishes in well under a minute ("parity ... on a
256-sentence subset").  Tolerances are the ones the small tests use."""

import math

import pytest
import torch

from tests.helpers import parity_metrics

pytestmark = pytest.mark.gpu


def test_config2_encoder_4096x128x24_layers_vs_oracle_on_256_rows(native_lib, cuda_device):
    """BASELINE config 2: B=4096, S=128, 24 layers, vocabulary 256 206 -- ONE full-size forward; 256 of its rows (every
    16th sentence) against the oracle run on those sentences alone (the engine is batch-composition invariant bit for bit,
    which the last assertion re-checks at this size)."""
    import bench
    from oracle.text_encoder import OracleEncoderConfig, OracleTextEncoder
    from sonar_b200 import B200TextEncoderModel, SequenceBatch, sonar_text_encoder_config

    sd = bench.synthetic_state_dict(cuda_device)
    model = B200TextEncoderModel(sonar_text_encoder_config("basic"), sd, cuda_device)
    ids = torch.randint(4, bench.VOCAB, (4096, 128), generator=torch.Generator().manual_seed(1000), dtype=torch.int64)
    out = model(SequenceBatch(ids.to(cuda_device), None)).sentence_embeddings
    model.check_inputs()
    rows = torch.arange(0, 4096, 16)
    bench.calibrate_cpu_threads()
    oracle = OracleTextEncoder(OracleEncoderConfig(vocab_size=bench.VOCAB, num_layers=24), {k: v.cpu() for k, v in sd.items()})
    ref = torch.cat([oracle(ids[rows[i:i + 64]], None)[0] for i in range(0, 256, 64)])
    m = parity_metrics(out[rows.to(cuda_device)], ref)
    print("config-2 size, 256 rows:", m)
    assert m["one_minus_cos_max"] <= 1e-3 and m["centred_cos_min"] >= 0.999 and m["rel_l2_max"] <= 1e-2, m
    sub = model(SequenceBatch(ids[rows].to(cuda_device), None)).sentence_embeddings
    assert torch.equal(sub, out[rows.to(cuda_device)])  # 256-row batch == rows of the 4096-row batch, bitwise


def test_config4_decoder_step_2560_rows_vs_oracle_on_64_rows(native_lib, cuda_device):
    """BASELINE config 4: 512 sentences x beam 5 = 2 560 hypothesis rows, 24 layers, vocabulary 256 206 -- teacher-forced
    steps; 64 rows against the oracle's full-recompute log-softmax (2e-2 + 2e-3*|lprob|, as tests/test_gpu_decoder.py)."""
    import bench
    from oracle.text_decoder import OracleDecoderConfig, OracleTextDecoder
    from sonar_b200 import B200TextDecoderModel, sonar_text_decoder_config

    sd = bench.synthetic_decoder_state_dict(cuda_device)
    model = B200TextDecoderModel(sonar_text_decoder_config("basic"), sd, cuda_device)
    n, beam, tmax, steps = 512, 5, 8, 4
    R = n * beam
    emb = torch.randn((n, 1024), generator=torch.Generator().manual_seed(5)) * 0.25 / math.sqrt(1024) * 32
    model.begin(emb.to(cuda_device), beam, tmax)
    table = torch.arange(R, dtype=torch.int32, device=cuda_device)[:, None].expand(R, tmax).contiguous()
    toks = torch.randint(4, 256000, (R, steps), generator=torch.Generator().manual_seed(9))
    rows = torch.arange(0, R, R // 64)[:64]
    sd_cpu = {k: v.float().cpu() for k, v in sd.items()}
    sd_cpu["final_proj.weight"] = sd_cpu["decoder_frontend.embed.weight"]
    oracle = OracleTextDecoder(OracleDecoderConfig(), sd_cpu)
    enc_rows = emb[rows // beam][:, None, :]
    for t in range(steps):
        probe = toks[:, (t + 1) % steps].contiguous()
        lp, tk, eos_lp, probe_lp = model.step(toks[:, t].contiguous().to(cuda_device), table, t, probe.to(cuda_device))
        ref = oracle.step_lprobs(toks[rows, : t + 1], enc_rows)
        lp, tk = lp.cpu()[rows], tk.cpu().long()[rows]
        torch.testing.assert_close(lp, torch.gather(ref, 1, tk), rtol=2e-3, atol=2e-2)
        torch.testing.assert_close(eos_lp.cpu()[rows], ref[:, 3], rtol=2e-3, atol=2e-2)
        torch.testing.assert_close(probe_lp.cpu()[rows], torch.gather(ref, 1, probe[rows][:, None])[:, 0], rtol=2e-3, atol=2e-2)
        assert bool((lp[:, :-1] >= lp[:, 1:]).all())
        assert bool((tk == ref.argmax(1, keepdim=True)).any(1).all())
        kth = ref.topk(16, dim=1).values[:, -1:]
        assert bool((lp[:, -1:] >= kth - 0.25).all())
    model.check_inputs()
