#!/bin/bash
# compute-sanitizer evidence for one learner step of every kernel family.
# Run on the GPU box:  bash tools/sanitize.sh   -> gpurun_out/sanitizer_*.log (copy the summaries to profiles/)
# racecheck covers shared-memory hazards inside a block; the tcgen05 / TMA (async proxy) traffic is invisible to it, so
# the mbarrier pipelines are additionally covered by the parity tests (bit-exact fused-vs-unfused, 30-run bias-sum stress).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" | tee gpurun_out/sanitizer_$tool.log
  timeout 900 $CS --tool $tool --print-limit 20 python tests/diag/sanitize_step.py ${MODES:-bf16 fused split three lstm ops} >> gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit code|total_loss|ops: ok" gpurun_out/sanitizer_$tool.log | tail -14
done
