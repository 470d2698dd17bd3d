#!/bin/bash
# round-2 GPU call H: relative-position attention on tcgen05 (speech)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_speech.py tests/test_reference_audio.py -x -q -m gpu > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_r2h.log
timeout 600 python bench.py --steps 3 --warmup 3 --only speech > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2h.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_speech_r2h.csv python scripts/profile_kernels.py speech > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_relpos_tc -s 2 -c 1 -o gpurun_out/relpos_tc_r2h python scripts/profile_kernels.py speech > /dev/null 2>&1
ls -la gpurun_out/*r2h*
