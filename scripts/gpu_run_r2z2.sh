#!/bin/bash
# round-2 GPU call Z2: sampling generator test (flatter model), speech batch invariance bit for bit
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_speech.py -x -q -m gpu > gpurun_out/pytest_r2z2.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_r2z2.log
