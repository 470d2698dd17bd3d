#!/bin/bash
# round-2 GPU call X (2 GPUs): NCCL tests, bench.py --gpus 2 (incl. the config-5 block), reference arm under torchrun
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_r2x.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_r2x.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2x_n2.json 2> gpurun_out/bench_r2x_n2.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r2x_n2.err
