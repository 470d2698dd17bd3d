#!/bin/bash
# round-2 GPU call I: rel-pos tcgen05 attention after the event-loop / masking fixes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_speech.py tests/test_gpu_kernels.py -x -q -m gpu -k "speech or relpos or attention" > gpurun_out/pytest_r2i.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r2i.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_speech_r2i.csv python scripts/profile_kernels.py speech > /dev/null 2>&1
timeout 600 python scripts/probe_r2.py attention > gpurun_out/probe_r2i.log 2>&1; tail -4 gpurun_out/probe_r2i.log
ls -la gpurun_out/*r2i*
