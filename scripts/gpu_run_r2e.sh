#!/bin/bash
# round-2 GPU call E (2 GPUs): NCCL tests on 2 ranks + bench --gpus 2 with the config-5 block (reduced shard so the call stays short)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_decoder.py -x -q -m gpu > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r2e.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --config5-per-gpu 65536 > gpurun_out/bench_r2e_n2.json 2> gpurun_out/bench_r2e_n2.err; echo "bench n2 rc=$?"
tail -c 600 gpurun_out/bench_r2e_n2.err
timeout 600 python bench.py --steps 3 --warmup 3 --only decoder > gpurun_out/bench_r2e_dec.json 2> gpurun_out/bench_r2e_dec.err; echo "bench dec rc=$?"
ls -la gpurun_out/*r2e*
