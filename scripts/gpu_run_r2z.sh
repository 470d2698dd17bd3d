#!/bin/bash
# round-2 GPU call Z: sampling generator on the CUDA decoder; is the mma.sync rel-pos attention batch-invariant bit for bit?
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu > gpurun_out/pytest_r2z.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_r2z.log
timeout 600 python - > gpurun_out/probe_r2z.log 2>&1 <<'PY'
import torch
from oracle.speech_encoder import OracleSpeechConfig, make_synthetic_speech_state_dict
from sonar_b200 import B200SpeechEncoderModel, PaddingMask, SequenceBatch, sonar_speech_encoder_config
dev = torch.device("cuda:0")
ocfg = OracleSpeechConfig(num_layers=2, pooler_layers=2)
sd = make_synthetic_speech_state_dict(ocfg, seed=11)
cfg = sonar_speech_encoder_config("english", num_encoder_layers=2, num_decoder_layers=2)
g = torch.Generator().manual_seed(12)
frames = [998, 258, 256, 2, 514, 770, 254, 600]
fb = torch.zeros((len(frames), 998, 80))
for i, n in enumerate(frames):
    fb[i, :n] = torch.randn((n, 80), generator=g)
for impl in ("mma_sync", "tcgen05"):
    m = B200SpeechEncoderModel(cfg, sd, dev, attn_impl=impl)
    a = m(SequenceBatch(fb.to(dev), PaddingMask(torch.tensor(frames), 998, frames))).sentence_embeddings
    for i in (1, 4, 7):
        n = frames[i]
        alone = m(SequenceBatch(fb[i:i + 1, :n].contiguous().to(dev), None)).sentence_embeddings
        print(impl, i, "bitwise equal:", bool(torch.equal(alone[0], a[i])), "max abs diff:", float((alone[0] - a[i]).abs().max()))
PY
cat gpurun_out/probe_r2z.log | tail -8
