#!/usr/bin/env bash
# compute-sanitizer over the kernels round 2 added or rewrote (small cases of the GPU tests):
#   attention_tc_kernel (P in tensor memory, TS-form P.V, split rings), attention_relpos_tc_kernel, the GEMM's two-warpgroup
#   epilogue / LayerNorm-folding epilogues / ordered split-K, the xsim column filter + col_rerank + merge_lists kernels,
#   the probe-token path of vocab_merge_kernel, beam_step_kernel with the fairseq2 EOS rules.
#     bash scripts/sanitize_r2.sh > gpurun_out/sanitize_r2.log 2>&1
set -u
cd "$(dirname "$0")/.."
echo "=== memcheck: GEMM epilogues (bias / residual / statistics / LN consumer / split-K), attention ==="
timeout 360 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -x -q -m gpu tests/test_gpu_kernels.py \
  -k "(splitk and (700 or 2500)) or (residual_stats and (300 or 77)) or (ln_consumer and 130) or attention or (test_gemm_residual_fp32 and 515)" 2>&1 | tail -8
echo "=== memcheck: xsim (one direction, bidirectional, overflow, narrow spread), speech rel-pos attention on tcgen05, decoder ==="
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -x -q -m gpu tests/test_gpu_xsim.py \
  tests/test_gpu_speech.py::test_relpos_attention_tcgen05_agrees_with_mma_sync_and_is_batch_invariant \
  tests/test_gpu_decoder.py -k "not large_slice" 2>&1 | tail -8
echo "=== racecheck: attention (tcgen05), rel-pos attention (tcgen05), xsim bidirectional small cases ==="
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_kernels.py tests/test_gpu_xsim.py \
  tests/test_gpu_speech.py::test_relpos_attention_tcgen05_agrees_with_mma_sync_and_is_batch_invariant \
  -k "attention_vs_sdpa or (bidir_matches and (64 or 5-3)) or relpos_attention" 2>&1 | tail -8
echo "=== synccheck: the same ==="
timeout 300 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_kernels.py tests/test_gpu_xsim.py \
  -k "attention_vs_sdpa or (bidir_matches and (64 or 5-3))" 2>&1 | tail -8
