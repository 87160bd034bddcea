#!/bin/bash
# compute-sanitizer over what round 2e added or changed: the tf32 GEMM's new epilogues (shared-memory staging, bias row,
# conv taps with TMA zero fill, shared-weight batches) and the pixel-decoder kernels.  gpurun --timeout 1200 -- tools/sanitize_r02e.sh
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
PYT="python -m pytest -q -x -p no:cacheprovider"
run() {  # name, tool, pytest selection
  echo "== $1 ($2)"
  eval "timeout 600 $CS --tool $2 --error-exitcode 99 --launch-timeout 0 $PYT $3" > $OUT/sanitize_$1.log 2>&1
  echo "exit $?" >> $OUT/sanitize_$1.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" $OUT/sanitize_$1.log | tail -4
}
run memcheck_tf32_e memcheck "tests/test_gpu_tf32.py -k \"True or two_level or 24-20\""
run memcheck_pixdec_stages memcheck "tests/test_gpu_pixel_decoder.py -k \"ms_deform or group_norm or conv3x3 or upsample or conv1x1\""
run memcheck_pixdec_whole memcheck "tests/test_gpu_pixel_decoder.py -k \"forward_matches and size1\""
run racecheck_tf32_e racecheck "tests/test_gpu_tf32.py -k \"(case0 or case2 or case10) and True\""
run racecheck_pixdec racecheck "tests/test_gpu_pixel_decoder.py -k \"ms_deform or (group_norm and shape0) or (conv3x3 and geom0)\""
