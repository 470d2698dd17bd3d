#!/usr/bin/env bash
# compute-sanitizer passes over the small GPU paths (SURVEY §5 "race detection / sanitizers": the reference has none; the
# mbarrier / TMEM / TMA code here is where misuse would hide).  Run on a B200 box:
#     bash scripts/sanitize.sh > gpurun_out/sanitize.log 2>&1
# Each tool runs the smoke() encoder forward (tcgen05 GEMM + attention, LayerNorm, embed, pool) and the small kernel
# tests that cover the xsim, decoder and speech kernels.  A clean run prints "ERROR SUMMARY: 0 errors" per tool.
set -u
cd "$(dirname "$0")/.."
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool : smoke() ==="
  timeout 900 compute-sanitizer --tool "$tool" --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
done
echo "=== compute-sanitizer --tool memcheck : pytest small decoder / speech / xsim cases ==="
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_decoder.py::test_teacher_forced_steps_match_oracle \
  tests/test_gpu_speech.py::test_fbank_kernel_matches_oracle_and_golden \
  tests/test_gpu_speech.py::test_speech_encoder_vs_oracle tests/test_gpu_xsim.py::test_xsim_matches_oracle 2>&1 | tail -15
echo "=== compute-sanitizer --tool racecheck : speech encoder (rel-pos attention smem ring) + decoder step ==="
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_speech.py::test_speech_encoder_vs_oracle \
  tests/test_gpu_decoder.py::test_teacher_forced_steps_match_oracle 2>&1 | tail -15
echo "=== compute-sanitizer --tool memcheck : skinny GEMM, fused-LayerNorm GEMM, graph-replayed beam search ==="
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest -x -q -m gpu \
  tests/test_gpu_kernels.py -k "skinny" \
  "tests/test_gpu_encoder.py::test_fused_layernorm_is_bitwise_the_separate_kernel" \
  tests/test_gpu_decoder.py::test_cuda_graph_replay_equals_eager_generation 2>&1 | tail -15
