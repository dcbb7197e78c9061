#!/bin/bash
# ncu evidence for round 2 (one GPU; numbers printed under ncu are never bench values).
# Outputs under gpurun_out/: r02_launches_bench.csv (every launch of 2 default steps with its device time),
# r02_conv_fp32.ncu-rep / r02_conv_bf16.ncu-rep (--set full of the tcgen05 conv launches), r02_cluster_e8.ncu-rep
# (--set full of the streaming clustering kernel on the HBM-resident shape).
set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
# the merged gpurun_out/ must stay below 64 MiB: the conv reports (22 launches each) are reduced to their raw CSV page
for mode in fp32 bf16; do
  timeout 600 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 44 -c 22 -o gpurun_out/r02_conv_$mode \
      python scripts/ncu_one_step.py $mode > gpurun_out/r02_conv_$mode.log 2>&1
  ncu -i gpurun_out/r02_conv_$mode.ncu-rep --page raw --csv > gpurun_out/r02_conv_${mode}_raw.csv 2>/dev/null
  rm -f gpurun_out/r02_conv_$mode.ncu-rep
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:seq_cluster -s 1 -c 1 -o gpurun_out/r02_cluster_e8 \
    python scripts/profile_cluster.py > gpurun_out/r02_cluster_e8.log 2>&1
ls -la gpurun_out | grep r02
