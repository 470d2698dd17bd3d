#!/bin/bash
# round-2 GPU call O: dwconv taps through shared memory (speech), decoder step time against position / ancestry sharing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_speech.py -x -q -m gpu > gpurun_out/pytest_r2o.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r2o.log
timeout 600 python scripts/probe_r2.py decstep > gpurun_out/probe_r2o.log 2>&1; tail -5 gpurun_out/probe_r2o.log
timeout 900 python bench.py --steps 3 --warmup 3 --only speech > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2o.err
