"""Small driver for ncu captures: runs each hot kernel a few times at the bench shapes.
    ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 2 -o gpurun_out/gemm python scripts/profile_kernels.py gemm
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sonar_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
T, D, F = 4096 * 128, 1024, 8192
if len(sys.argv) > 2:
    T = int(sys.argv[2])
g = torch.Generator(device=dev).manual_seed(0)
if which == "gemm":      # FFN inner projection (bias + ReLU, bf16 out)
    a = torch.randn((T, D), device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn((F, D), device=dev, generator=g) / math.sqrt(D)).to(torch.bfloat16)
    b = torch.randn((F,), device=dev, generator=g)
    out = torch.empty((T, F), device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.gemm_bf16(a, w, b, epilogue="relu", out=out)
elif which == "gemm_res":  # out-proj (bias + fp32 residual in place)
    a = torch.randn((T, D), device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn((D, D), device=dev, generator=g) / math.sqrt(D)).to(torch.bfloat16)
    b = torch.randn((D,), device=dev, generator=g)
    x = torch.randn((T, D), device=dev, generator=g)
    for _ in range(4):
        ops.gemm_bf16(a, w, b, epilogue="residual", residual=x, out=x)
elif which == "attention":
    qkv = torch.randn((T, 3 * D), device=dev, generator=g).to(torch.bfloat16)
    cu = ops.cu_seqlens_of([128] * (T // 128)).to(dev)
    for _ in range(4):
        ops.attention(qkv, cu, 128, 16)
elif which == "layernorm":
    x = torch.randn((T, D), device=dev, generator=g)
    gg, bb = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    for _ in range(4):
        ops.layernorm(x, gg, bb)
torch.cuda.synchronize()
print("done", which)
