#!/bin/bash
# round-2 GPU call T: end-of-round tree: full GPU suite, full bench, ncu --set full of the FFN1 GEMM as the text encoder
# runs it (one-warpgroup epilogue), taken inside a real forward
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2t.log
tail -5 gpurun_out/pytest_r2t.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r2t.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k "regex:gemm_bf16_tcgen05_kernel<2, 1," -s 36 -c 1 -o gpurun_out/gemm_ffn1_instep_r2t python scripts/profile_kernels.py text_step > gpurun_out/prof_r2t.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*r2t* | tail -8
