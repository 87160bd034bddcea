#!/bin/bash
# compute-sanitizer memcheck over the kernels touched after tools/sanitize_r02e.sh ran: the training kernels (parallel
# LayerNorm-backward reduce, float4 ReLU-backward / add-rows / batch-sum), the deformable-attention backward (3-D grid) and
# the whole pixel-decoder gradient test.  gpurun --timeout 900 -- tools/sanitize_r02f.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
PYT="python -m pytest -q -x -p no:cacheprovider"
run() {
  echo "== $1 ($2)"
  eval "timeout 400 $CS --tool $2 --error-exitcode 99 --launch-timeout 0 $PYT $3" > $OUT/sanitize_$1.log 2>&1
  echo "exit $?" >> $OUT/sanitize_$1.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" $OUT/sanitize_$1.log | tail -3
}
run memcheck_train_f memcheck "tests/test_gpu_train.py -k \"24-2-128-160-49 and fp32\""
run memcheck_pixdec_grad memcheck "tests/test_gpu_pixel_decoder.py -k \"ms_deform or gradients_match or golden\""
run racecheck_lnbwd racecheck "tests/test_gpu_pixel_decoder.py -k \"gradients_match\""
