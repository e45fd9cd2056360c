#!/bin/bash
# A/B of transit kernel variants built by scripts/ab_build.py: scripts/ab_transit.sh base b48u4 pf32u4 ...
for t in "$@"; do
  lib=picaso_b200/_build/libpb_$t.so
  [ "$t" = base ] && lib=picaso_b200/_build/libpicaso_b200.so
  PICASO_B200_LIB=$PWD/$lib python scripts/kernel_times.py --only transit --reps 200 2>&1 | grep transit | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$t  %-40s %9.1f us'%(d['config'][:40], 1e3*d['ms_per_launch']))"
done
