#!/bin/bash
# round-2 GPU call R: evidence run of the end-of-round tree: full GPU suite, full bench (+ reference arm), launch list of
# one text step, ncu --set full of the dominant kernel (FFN1 GEMM) and of the kernels this round added
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2r.log
tail -5 gpurun_out/pytest_r2r.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r2r.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2r_ref.json 2> gpurun_out/bench_r2r_ref.err; echo "ref rc=$?"
KREG='regex:gemm_bf16_tcgen05|attention_tc|layernorm|embed_kernel|ln_pool'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 170 -c 170 --csv --log-file gpurun_out/launches_r2r.csv python scripts/profile_kernels.py text_step > gpurun_out/prof_r2r.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/gemm_ffn1_r2r python scripts/profile_kernels.py gemm > /dev/null 2>&1
timeout 400 $NCU -k regex:gemm_bf16_tcgen05 -s 1 -c 1 -o gpurun_out/xsim_sweep_filter_r2r python scripts/profile_kernels.py xsim_bidir 131072 > /dev/null 2>&1
timeout 400 $NCU -k regex:col_rerank -s 1 -c 1 -o gpurun_out/col_rerank_r2r python scripts/profile_kernels.py xsim_bidir 131072 > /dev/null 2>&1
timeout 400 $NCU -k regex:glu_dwconv -s 24 -c 1 -o gpurun_out/glu_dwconv_r2r python scripts/profile_kernels.py speech > /dev/null 2>&1
ls -la gpurun_out/*r2r* | tail -14
