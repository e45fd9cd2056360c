#!/bin/bash
# round 2, GPU call 2: parity of v5 (factorised denominators), A/B of build variants and tile widths, ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests2.log
{
PB_REFL_KERNEL=4 bash scripts/ab_run.sh base | sed 's/^base/v4/'
bash scripts/ab_run.sh base r128u2 r112u1 r128u1
for wt in 32 24 23 20 16; do echo "wt=$wt"; PB_REFL_WT=$wt bash scripts/ab_run.sh base r128u1; done
} > gpurun_out/r2_ab2.log 2>&1
python scripts/kernel_times.py --only refl --reps 30 > gpurun_out/r2_kt_refl.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:refl_toa_kernel5 -s 30 -c 2 -o gpurun_out/r2_refl_v5 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/r2_ncu_v5.log 2>&1
