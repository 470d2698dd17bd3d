#!/bin/bash
# round-2 GPU call D: top-k sweep split over both epilogue warpgroups; in-step A/B of the schedule variants
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_r2d.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2d.err
ls -la gpurun_out/*r2d* | tail
