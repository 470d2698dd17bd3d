#!/bin/bash
# round-2 GPU call C: attention with P in tensor memory + split rings; GEMM with two epilogue warpgroups
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_r2c_kernels.log 2>&1; echo "kernels rc=$?"
tail -12 gpurun_out/pytest_r2c_kernels.log
timeout 600 python scripts/probe_r2.py attention lnfold > gpurun_out/probe_r2c.log 2>&1; echo "probe rc=$?"
cat gpurun_out/probe_r2c.log | tail -12
timeout 600 python scripts/gpu_probe.py perf > gpurun_out/probe_perf_r2c.log 2>&1; tail -25 gpurun_out/probe_perf_r2c.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_r2c.log
timeout 900 python bench.py --steps 5 --warmup 3 --only predict > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2c.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attention_tc -s 2 -c 1 -o gpurun_out/attn_tc_r2c python scripts/profile_kernels.py attention > /dev/null 2>&1
timeout 300 $NCU -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/gemm_res_r2c python scripts/profile_kernels.py gemm_res > /dev/null 2>&1
timeout 300 $NCU -k regex:gemm_bf16_tcgen05 -s 2 -c 1 -o gpurun_out/gemm_resstats_r2c python scripts/profile_kernels.py gemm_resstats > /dev/null 2>&1
ls -la gpurun_out/*r2c* | tail -12
