#!/bin/bash
# round-2 GPU call S: predict() with the pinned staging ring + preallocated result; ln_fold default epilogue groups
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py -x -q -m gpu > gpurun_out/pytest_r2s.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r2s.log
for i in 1 2; do
timeout 900 python bench.py --steps 3 --warmup 3 --only predict > gpurun_out/bench_r2s_$i.json 2> gpurun_out/bench_r2s.err; echo "bench rc=$?"
done
tail -c 300 gpurun_out/bench_r2s.err
