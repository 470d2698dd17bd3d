#!/bin/bash
# round-2 GPU call AD: ncu --set full of the kernels that still had no tracked summary
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 200 $NCU -k regex:attention_relpos_kernel -s 24 -c 1 -o gpurun_out/relpos_mma_r2ad python scripts/profile_kernels.py speech > /dev/null 2>&1
timeout 200 $NCU -k regex:glu_dwconv -s 24 -c 1 -o gpurun_out/glu_dwconv_r2ad python scripts/profile_kernels.py speech > /dev/null 2>&1
timeout 200 $NCU -k regex:layernorm_bf16 -s 100 -c 1 -o gpurun_out/layernorm_speech_r2ad python scripts/profile_kernels.py speech > /dev/null 2>&1
timeout 200 $NCU -k regex:pool_attention -s 3 -c 1 -o gpurun_out/pool_attention_r2ad python scripts/profile_kernels.py speech > /dev/null 2>&1
timeout 200 $NCU -k regex:gemm_skinny -s 40 -c 1 -o gpurun_out/gemm_skinny_r2ad python scripts/profile_kernels.py decoder_small > /dev/null 2>&1
timeout 200 $NCU -k regex:vocab_merge -s 2 -c 1 -o gpurun_out/vocab_merge_r2ad python scripts/profile_kernels.py decoder > /dev/null 2>&1
ls -la gpurun_out/*r2ad*
