#!/bin/bash
# round-2 GPU call P: decoder self-attention with one warp per (sentence, head): explicit sharing vs L1 sharing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r2p.log
timeout 600 python scripts/probe_r2.py decstep > gpurun_out/probe_r2p_dedup.log 2>&1; tail -3 gpurun_out/probe_r2p_dedup.log
SB_DECODE_ATTN_L1=1 timeout 600 python scripts/probe_r2.py decstep > gpurun_out/probe_r2p_l1.log 2>&1; tail -3 gpurun_out/probe_r2p_l1.log
