#!/bin/bash
# round-2 GPU call F: launch lists (per-kernel device time) of the decoder step and the speech forward
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decoder_r2f.csv python scripts/profile_kernels.py decoder > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_speech_r2f.csv python scripts/profile_kernels.py speech > /dev/null 2>&1
ls -la gpurun_out/*r2f*
