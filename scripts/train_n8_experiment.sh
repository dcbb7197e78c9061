#!/bin/bash
# 8-GPU data-parallel training step: does the gradient all-reduce really overlap the late backward?  Variants:
#   default                      persistent conv CTAs use ~200 KB of shared memory (NCCL CTAs cannot co-reside)
#   STEMSEG_CONV_SMEM_KB=150     leaves ~75 KB per SM for NCCL's CTAs
#   NCCL_MAX_NCHANNELS=4         fewer NCCL CTAs
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
export PROBE_SHORT=1
$TR scripts/train_overlap_probe.py timeline > gpurun_out/train_n${N}_default.txt 2>&1
STEMSEG_CONV_SMEM_KB=150 $TR scripts/train_overlap_probe.py > gpurun_out/train_n${N}_smem150.txt 2>&1
NCCL_MAX_NCHANNELS=4 $TR scripts/train_overlap_probe.py > gpurun_out/train_n${N}_nch4.txt 2>&1
grep -h "ms/step" gpurun_out/train_n${N}_*.txt
