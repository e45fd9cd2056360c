"""Build an A/B variant of the library: python scripts/ab_build.py <tag> [-DNAME[=V] ...]
-> picaso_b200/_build/libpb_<tag>.so ; run with PICASO_B200_LIB=<that path>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from picaso_b200 import build as b
tag = sys.argv[1]
defs = [a[2:] for a in sys.argv[2:] if a.startswith("-D")]
print(b.build(force=True, defines=defs, lib=os.path.join(b.OUT, "libpb_%s.so" % tag)))
