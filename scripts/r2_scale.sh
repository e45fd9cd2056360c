#!/bin/bash
# usage: scripts/r2_scale.sh N  - multi-GPU lines (torchrun, one rank per GPU): headline (weak), cfg3 (strong), cfg5 (strong)
N=$1
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5 > gpurun_out/r2_scale_headline_n$N.json 2> gpurun_out/r2_scale_headline_n$N.err
run --config cfg3 --steps 10 --warmup 3 > gpurun_out/r2_scale_cfg3_n$N.json 2> gpurun_out/r2_scale_cfg3_n$N.err
run --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2_scale_cfg5_n$N.json 2> gpurun_out/r2_scale_cfg5_n$N.err
for c in headline cfg3 cfg5; do echo "== $c N=$N"; cut -c1-330 gpurun_out/r2_scale_${c}_n$N.json; tail -c 500 gpurun_out/r2_scale_${c}_n$N.err | grep -v "^$" | tail -4; done
