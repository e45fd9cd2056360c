"""cProfile of the host side of the spectrum-level chain (run on the GPU box)."""
import cProfile, pstats, sys, os, importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
import picaso_b200 as pb
ctx = pb.Context(0)
db, ray, atms, ducks = b._spectrum_setup()
opa, one = b.spectrum_gpu_factory(pb, ctx, db, ray, ducks)
for i in range(5):
    one(i)
pr = cProfile.Profile(); pr.enable()
for i in range(200):
    one(i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
