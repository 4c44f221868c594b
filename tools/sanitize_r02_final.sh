#!/bin/bash
# compute-sanitizer over what the last session of round 2 added: the staged read-back with three frames in flight
# (k_stage_rect), the cost classes of the busy list, the 128-thread span kernel (SWEGL_B200_SHARED_GPU=1 makes every
# context take it), plus a seeded GPU fuzz with that kernel.   gpurun --timeout 900 -- 'bash tools/sanitize_r02_final.sh r02f'
tag=${1:-run}; out=gpurun_out; mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
run() { # tool, log suffix, pytest args...
  tool=$1; sfx=$2; shift 2
  timeout 400 $S --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > $out/${tag}_sanitizer_$sfx.log 2>&1
  echo "$tool $sfx rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $out/${tag}_sanitizer_$sfx.log | tr '\n' ' ')"
}
run memcheck memcheck_async tests/test_gpu_parity.py -k "three_frames and 1080 or shared_gpu or pipelined_readback or partial_readback and 1080"
SWEGL_B200_SHARED_GPU=1 run memcheck memcheck_narrow_fuzz tests/test_fuzz_gpu.py -k "not layers"
SWEGL_B200_SHARED_GPU=1 run racecheck racecheck_narrow_fuzz tests/test_fuzz_gpu.py -k "not layers"
run initcheck initcheck_async tests/test_gpu_parity.py -k "three_frames and 1080 or shared_gpu"
SWEGL_B200_SHARED_GPU=1 timeout 200 python tools/fuzz_gpu.py 4000 4300 $out/${tag}_fuzz_gpu_narrow.json > $out/${tag}_fuzz_narrow.log 2>&1; tail -1 $out/${tag}_fuzz_narrow.log
