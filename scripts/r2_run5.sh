#!/bin/bash
# round 2, GPU call 5: bench.py --config for every BASELINE configuration (N = 1), headline bench, ncu captures of the
# cfg2 / cfg3 / cfg4 kernels (the round-1 captures of SH4 and transit were of other kernels)
mkdir -p gpurun_out
for c in cfg2 cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
  tail -c 600 gpurun_out/r2_bench_$c.err
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_headline.json 2> gpurun_out/r2_bench_headline.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_headline_ref.json 2>> gpurun_out/r2_bench_headline.err
ncu --set full --clock-control none --import-source on -k regex:sh_reflected_kernel -s 4 -c 1 -o gpurun_out/r2_sh4_cfg3 python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_ncu_sh4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:therm_toa -s 6 -c 1 -o gpurun_out/r2_therm_cfg2 python bench.py --config cfg2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_ncu_therm.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:transit_kernel|opacity_layer_kernel" -s 8 -c 2 -o gpurun_out/r2_cfg4 python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_ncu_cfg4.log 2>&1
