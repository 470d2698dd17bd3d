#!/bin/bash
# round-2 GPU call G: queue-based top-k sweep epilogue (decoder vocabulary projection, xsim)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_xsim.py tests/test_gpu_config_sizes.py -x -q -m gpu > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r2g.log
timeout 900 python bench.py --steps 3 --warmup 3 --only decoder,xsim > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_decoder_r2g.csv python scripts/profile_kernels.py decoder > /dev/null 2>&1
ls -la gpurun_out/*r2g*
