#!/bin/bash
# round 2, GPU call 4: v5.2 (split slow path, Estrin exp, L2 prefetch option) parity + A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests4.log
{
bash scripts/ab_run.sh base u2p0 u1p1 u2p1
for wt in 25 24 23; do echo "wt=$wt"; PB_REFL_WT=$wt bash scripts/ab_run.sh base u2p1; done
} > gpurun_out/r2_ab4.log 2>&1
python scripts/kernel_times.py --only refl --reps 30 > gpurun_out/r2_kt_refl4.jsonl 2>&1
