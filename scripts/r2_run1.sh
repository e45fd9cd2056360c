#!/bin/bash
# round 2, GPU call 1: parity of the v5 reflected kernel, machine numbers, A/B of its build variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests1.log
python scripts/microbench.py > gpurun_out/r2_microbench.jsonl 2>&1
{
PB_REFL_KERNEL=4 bash scripts/ab_run.sh base | sed 's/^base/v4/'
bash scripts/ab_run.sh base r128u2 r112u1 r128u1
for wt in 32 24 23 20 16; do echo "wt=$wt"; PB_REFL_WT=$wt bash scripts/ab_run.sh base r128u1; done
} > gpurun_out/r2_ab1.log 2>&1
python scripts/kernel_times.py --only refl --reps 30 > gpurun_out/r2_kt_refl.jsonl 2>&1
