#!/bin/bash
# A/B the reflected TOA kernel variants built by scripts/ab_build.py (run on the GPU box)
for t in "$@"; do
  lib=picaso_b200/_build/libpb_$t.so
  [ "$t" = base ] && lib=picaso_b200/_build/libpicaso_b200.so
  PICASO_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$t', round(1e3*d['ms_per_step'],2), 'us', d['config']['parity_albedo_max_rel_err'])"
done
