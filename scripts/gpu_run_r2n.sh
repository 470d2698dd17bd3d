#!/bin/bash
# round-2 GPU call N: one-pass bidirectional xsim (column filter in the sweep epilogue)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_xsim.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_r2n.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_r2n.log
timeout 900 python bench.py --steps 3 --warmup 3 --only xsim > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2n.err
ls -la gpurun_out/*r2n*
