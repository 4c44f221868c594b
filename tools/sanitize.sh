#!/bin/bash
# compute-sanitizer over the GPU parity, animation and fuzz tests (under gpurun); summaries land in gpurun_out/<tag>_sanitizer_*.log
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh r02'
tag=${1:-run}; out=gpurun_out; mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
run() { # tool, log suffix, pytest args...
  tool=$1; sfx=$2; shift 2
  timeout 900 $S --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > $out/${tag}_sanitizer_$sfx.log 2>&1
  echo "$tool $sfx rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $out/${tag}_sanitizer_$sfx.log | tr '\n' ' ')"
}
run memcheck memcheck_parity tests/test_gpu_parity.py -k "not 8k and not division and not filter and not partial_readback"
run memcheck memcheck_fuzz_anim tests/test_fuzz_gpu.py tests/test_animation.py tests/test_dropin_gpu.py
run racecheck racecheck_fuzz tests/test_fuzz_gpu.py -k "not layers"
run initcheck initcheck_fuzz tests/test_fuzz_gpu.py tests/test_animation.py
run synccheck synccheck_fuzz tests/test_fuzz_gpu.py -k "not layers"
