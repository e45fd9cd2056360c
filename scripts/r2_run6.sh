#!/bin/bash
# round 2, GPU call 6: SH4 tile kernel parity + timing, full-size tests, cfg4/cfg5 bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests6.log
for c in cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
done
PB_SH_TILE=0 python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg3_perangle.json 2>> gpurun_out/r2_bench_cfg3.err
python scripts/kernel_times.py --only sh --reps 20 > gpurun_out/r2_kt_sh6.jsonl 2>&1
