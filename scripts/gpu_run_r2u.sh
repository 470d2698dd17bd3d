#!/bin/bash
# round-2 GPU call U: ncu --set full of the FFN1 GEMM as the text encoder runs it (one-warpgroup epilogue): the 3rd GEMM of
# layer 12 of the second forward (GEMM order per layer: QKV, out-proj, FFN1, FFN2 -> skip 96 + 12 * 4 + 2 launches)
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:gemm_bf16_tcgen05 -s 146 -c 1 -o gpurun_out/gemm_ffn1_instep_r2u python scripts/profile_kernels.py text_step > gpurun_out/prof_r2u.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/prof_r2u.log
ls -la gpurun_out/*r2u*
