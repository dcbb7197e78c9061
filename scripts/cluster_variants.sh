#!/bin/bash
# A/B of the streaming clustering kernel variants (one process each: the variant is read once per process)
out=${1:-gpurun_out/cluster_variants.jsonl}
: > $out
for shape in "6635520 8 0" "3317760 4 2"; do
  for v in legacy 256x2 256x4 256x8 512x4 512x8; do
    STEMSEG_CLUSTER_STREAM=$v python scripts/profile_cluster.py $shape 2>/dev/null | tail -1 >> $out
  done
done
cat $out | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print('%-8s %-40s %8.3f ms  frac %.3f  checksum %d assigned %d' % (r['variant'], r['kernel'], r['launch_ms'], r['frac'], r['labels_checksum'], r['assigned']))
"
