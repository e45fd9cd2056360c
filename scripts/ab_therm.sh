#!/bin/bash
for t in "$@"; do
  lib=picaso_b200/_build/libpb_$t.so
  [ "$t" = base ] && lib=picaso_b200/_build/libpicaso_b200.so
  echo "== $t"
  PICASO_B200_LIB=$PWD/$lib PB_THERM_KERNEL=wave python scripts/kernel_times.py --only batch,thermal --reps 100 2>&1 | grep "thermal_toon" | grep -v "levels=1" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  %-80s %9.1f us'%(d['config'][:80], 1e3*d['ms_per_launch']))"
done
