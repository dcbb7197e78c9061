#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (SURVEY.md §5 race-detection row): memcheck on every family,
# racecheck (shared-memory hazards) on the kernels that communicate through shared memory / hand-rolled barriers.
# Usage (GPU box): bash scripts/sanitizer.sh gpurun_out/sanitizer      -> <prefix>_<tool>_<name>.log + <prefix>_summary.txt
prefix=${1:-gpurun_out/sanitizer}
export STEMSEG_SANITIZER=1
run() {
  tool=$1; name=$2; shift 2
  log=${prefix}_${tool}_${name}.log
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest -x -q -p no:cacheprovider "$@" > $log 2>&1
  rc=$?
  errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
  echo "$tool $name rc=$rc :: $errs :: $(grep -E 'passed|failed' $log | tail -1)" | tee -a ${prefix}_summary.txt
}
: > ${prefix}_summary.txt
run memcheck decoder tests/test_decoder_gpu.py -k "golden and (emb_xyff_t8 or semseg_42_t8 or emb_xytff_t16 or maxpool_t8)"
run memcheck cluster tests/test_cluster_gpu.py -k "golden"
run memcheck cluster_stream tests/test_cluster_gpu.py -k "large and 414720"
run memcheck foreground tests/test_foreground_gpu.py
run memcheck stitch tests/test_chaining_gpu.py -k "stitcher"
run memcheck pipeline tests/test_decoder_gpu.py -k "whole_step_graph or submit_result"
run memcheck train tests/test_training_gpu.py -k "trainer_step_matches_oracles"
run memcheck loss tests/test_loss_gpu.py -k "fixture and (three or single or empty_first)"
run racecheck cluster tests/test_cluster_gpu.py -k "golden"
run racecheck foreground tests/test_foreground_gpu.py
run racecheck stitch tests/test_chaining_gpu.py -k "stitcher_matches_reference"
run racecheck decoder tests/test_decoder_gpu.py -k "golden and emb_xyff_t8"
cat ${prefix}_summary.txt
