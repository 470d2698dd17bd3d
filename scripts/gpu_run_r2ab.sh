#!/bin/bash
# round-2 GPU call AB: end-of-round tree: smoke(), full GPU suite, full bench; dwconv launch list after the occupancy change
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2ab.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2ab.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2ab.log
tail -5 gpurun_out/pytest_r2ab.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2ab.json 2> gpurun_out/bench_r2ab.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r2ab.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:glu_dwconv -c 48 --csv --log-file gpurun_out/launches_dwconv_r2ab.csv python scripts/profile_kernels.py speech > /dev/null 2>&1
