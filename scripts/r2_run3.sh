#!/bin/bash
# round 2, GPU call 3: v5.1 (single barrier per chunk, branch-free exp) parity, A/B, ncu capture with source
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests3.log
{
bash scripts/ab_run.sh base r128u2 r136u1 r136u2
for wt in 32 24 22; do echo "wt=$wt"; PB_REFL_WT=$wt bash scripts/ab_run.sh base r136u2; done
} > gpurun_out/r2_ab3.log 2>&1
python scripts/kernel_times.py --only refl --reps 30 > gpurun_out/r2_kt_refl3.jsonl 2>&1
ncu --set full --clock-control none --import-source on -k regex:refl_toa_kernel5 -s 30 -c 2 -o gpurun_out/r2_refl_v51 python bench.py --no-cpu-baseline --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/r2_ncu_v51.log 2>&1
