#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (run on the GPU box through gpurun):
#   tools/sanitize.sh [memcheck|racecheck|synccheck]
tool=${1:-memcheck}
out=gpurun_out/sanitize_$tool.log
: > $out
run() {
  echo "=== $*" >> $out
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 "$@" >> $out 2>&1
  echo "=== exit $?" >> $out
}
run python -m pytest -x -q -m gpu tests/test_gpu_parity.py -k "test_golden_log_spec_and_energy and (A or S512 or S256 or W512 or B-) and mel"
run python -m pytest -x -q -m gpu tests/test_gpu_parity.py -k "test_ragged_batch_matches_oracle and (S512 or W256 or ST2 or N400) and mel-librosa"
run python -m pytest -x -q -m gpu tests/test_gpu_frontend.py -k "loudness_batch_matches_oracle or loudness_scan or gate_decisions"
run python -m pytest -x -q -m gpu tests/test_gpu_backward.py -k "hifigan or styletts2 or deterministic or packed_job"
grep -E "^=== |ERROR SUMMARY|passed|failed|Error|error" $out | head -60
