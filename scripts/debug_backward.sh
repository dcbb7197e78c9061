#!/bin/bash
# runs every backward op check in its own process; failing ops are re-run under compute-sanitizer
if [ -n "$PROBE" ]; then echo "=== tma probe"
for swz in 0 1; do for c0 in 0 8 -8 3 -1 -71 489 505; do timeout 60 scripts/tma_probe $swz $c0; done; done; fi
for op in ${OPS:-upsample pool gn chsum head dgrad dgrad1 wgrad1 wgrad}; do
  echo "=== $op"
  if ! CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/debug_backward.py $op 2>&1 | tail -12; then echo "(pipeline fail)"; fi
  if [ "${PIPESTATUS[0]}" != "0" ]; then
    echo "--- sanitizer $op"
    timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python scripts/debug_backward.py $op 2>&1 | grep -v "^=========     Host Frame\|^=========         in\|^=========$" | head -40
  fi
done
