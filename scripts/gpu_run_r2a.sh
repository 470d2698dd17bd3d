#!/bin/bash
# round-2 GPU call A: full GPU test-suite, bench (all blocks), ncu captures for the kernels that lacked a tracked summary
set -x
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2a.log
tail -5 gpurun_out/pytest_r2a.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r2a.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:fbank_kernel -s 1 -c 1 -o gpurun_out/fbank_r2a python scripts/profile_kernels.py fbank > /dev/null 2>&1
timeout 300 $NCU -k regex:decode_attention -s 2 -c 1 -o gpurun_out/decattn_r2a python scripts/profile_kernels.py decoder > /dev/null 2>&1
timeout 300 $NCU -k regex:gemm_bf16_tcgen05 -s 1 -c 1 -o gpurun_out/xsim_topk_r2a python scripts/profile_kernels.py xsim > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
