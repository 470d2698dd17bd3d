#!/bin/bash
# round-2 GPU call B: new attention kernel + LayerNorm folding: kernel tests first (bounded), probes, full suite, bench, ncu
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/pytest_r2b_kernels.log 2>&1; echo "kernels rc=$?"
tail -15 gpurun_out/pytest_r2b_kernels.log
timeout 600 python scripts/probe_r2.py attention lnfold > gpurun_out/probe_r2b.log 2>&1; echo "probe rc=$?"
cat gpurun_out/probe_r2b.log | tail -20
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_r2b.log
timeout 600 python scripts/probe_r2.py predict > gpurun_out/probe_r2b_predict.log 2>&1; tail -4 gpurun_out/probe_r2b_predict.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r2b.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:attention_tc -s 2 -c 1 -o gpurun_out/attn_tc_r2b python scripts/profile_kernels.py attention > /dev/null 2>&1
ls -la gpurun_out/*r2b* | tail -12
