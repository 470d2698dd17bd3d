#!/bin/bash
# round-2 GPU call W: bidirectional xsim with order-statistic column thresholds (16th best of the 1/8 sample)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_xsim.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/pytest_r2w.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_r2w.log
timeout 900 python bench.py --steps 3 --warmup 3 --only xsim > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2w.err
