#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / TMA protocols: memcheck on the stage
# tests of both precision modes and on one small whole forward, racecheck (shared-memory hazards) on the attention and
# mask-einsum stages.  Run on the GPU box:   gpurun --timeout 1500 -- tools/sanitize.sh
# Logs land in gpurun_out/sanitize_*.log; the summaries are committed under profiles/.
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
PYT="python -m pytest -q -x -p no:cacheprovider"
run() {  # name, tool, pytest selection
  echo "== $1 ($2)"
  eval "timeout 900 $CS --tool $2 --error-exitcode 99 --launch-timeout 0 $PYT $3" > $OUT/sanitize_$1.log 2>&1
  echo "exit $?" >> $OUT/sanitize_$1.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" $OUT/sanitize_$1.log | tail -4
}
run memcheck_attention memcheck "tests/test_gpu_bf16.py -k attention_stage"
run memcheck_einsum memcheck "tests/test_gpu_bf16.py -k batched_einsum"
run memcheck_fp32 memcheck "tests/test_gpu_parity.py -k \"mask_bits_stage or masked_attention_stage or tiny\""
run memcheck_train memcheck "tests/test_gpu_train.py -k dispatch"
run racecheck_attention racecheck "tests/test_gpu_bf16.py -k \"attention_stage and 64-0.5\""
run racecheck_einsum racecheck "tests/test_gpu_bf16.py -k batched_einsum"
# round 2c: the tf32 GEMM (all operand-major pairs, split-K, two-level batches), the attention row kernels, the matching
# and caption kernels
run memcheck_tf32 memcheck "tests/test_gpu_tf32.py -k \"True or two_level or 24-20\""
run memcheck_matching memcheck "tests/test_gpu_matching.py -k \"point_sample or cost or 1-gts0 or fixture\""
run memcheck_caption memcheck "tests/test_gpu_caption.py -k \"fp32\""
run racecheck_tf32 racecheck "tests/test_gpu_tf32.py -k \"case0 and True\""
