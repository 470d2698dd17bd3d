#!/bin/bash
# round-2 GPU call AA: conformer GLU + depthwise conv with packed FFMA2 over channel pairs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_speech.py -x -q -m gpu > gpurun_out/pytest_r2aa.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r2aa.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:glu_dwconv -c 48 --csv --log-file gpurun_out/launches_dwconv_r2aa.csv python scripts/profile_kernels.py speech > /dev/null 2>&1
grep -c glu_dwconv gpurun_out/launches_dwconv_r2aa.csv
timeout 900 python bench.py --steps 3 --warmup 3 --only speech > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err; echo "bench rc=$?"
