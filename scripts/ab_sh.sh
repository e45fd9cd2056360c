#!/bin/bash
for t in "$@"; do
  lib=picaso_b200/_build/libpb_$t.so
  [ "$t" = base ] && lib=picaso_b200/_build/libpicaso_b200.so
  echo "== $t"
  PICASO_B200_LIB=$PWD/$lib python scripts/kernel_times.py --only sh --reps 50 2>&1 | grep "reflected_SH" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  %-70s %9.1f us'%(d['config'][:70], 1e3*d['ms_per_launch']))"
done
