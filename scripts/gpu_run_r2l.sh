#!/bin/bash
# round-2 GPU call L: launch list of the one-pass bidirectional xsim at the bench size
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_xsim_bidir_r2l.csv python scripts/profile_kernels.py xsim_bidir > gpurun_out/prof_r2l.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/prof_r2l.log
