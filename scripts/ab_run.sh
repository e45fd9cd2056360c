#!/bin/bash
# A/B the reflected TOA kernel variants built by scripts/ab_build.py (run on the GPU box)
for t in "$@"; do
  lib=picaso_b200/_build/libpb_$t.so
  [ "$t" = base ] && lib=picaso_b200/_build/libpicaso_b200.so
  out=$(PICASO_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1)
  echo "$out" | python -c "import sys,json
s=sys.stdin.read()
try:
    d=json.loads(s); print('$t', round(1e3*d['ms_per_step'],2), 'us', d['config']['parity_albedo_max_rel_err'])
except Exception: print('$t FAILED:', s[-300:])"
done
