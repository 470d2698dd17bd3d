#!/bin/bash
# round-2 GPU call Q: ordered split-K of the decoder's residual GEMMs
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "splitk or gemm" > gpurun_out/pytest_r2q.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r2q.log
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_config_sizes.py -x -q -m gpu > gpurun_out/pytest_r2q2.log 2>&1; echo "pytest2 rc=$?"
tail -4 gpurun_out/pytest_r2q2.log
timeout 600 python scripts/probe_r2.py decstep > gpurun_out/probe_r2q.log 2>&1; tail -3 gpurun_out/probe_r2q.log
timeout 900 python bench.py --steps 3 --warmup 3 --only decoder > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2q.err
