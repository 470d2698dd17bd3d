#!/bin/bash
# round-2 GPU call J: rel-pos tcgen05 attention, single-pass softmax with a lazily raised reference
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_speech.py tests/test_reference_audio.py -x -q -m gpu > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_r2j.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_speech_r2j.csv python scripts/profile_kernels.py speech > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_relpos_tc -s 2 -c 1 -o gpurun_out/relpos_tc_r2j python scripts/profile_kernels.py speech > /dev/null 2>&1
ls -la gpurun_out/*r2j*
