"""one warm + 2 timed pb.get_fluxes calls with device-resident opacities (for an ncu launch list)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import picaso_b200 as pb
from picaso_b200 import synth
from picaso_b200.optics import DeviceArray
ctx = pb.Context(0)
d = synth.climate_inputs(L=90, W=661, K=8, seed=77, ng=1)
args = [d["Atmosphere"], d["OpacityWEd"], d["OpacityNoEd"], d["ScatteringPhase"], d["Disco"], d["Opagrid"], d["F0PI"], True, True]
args[1] = type(d["OpacityWEd"])(*[DeviceArray.from_numpy(ctx, a) for a in d["OpacityWEd"]])
args[2] = type(d["OpacityNoEd"])(*[DeviceArray.from_numpy(ctx, a) for a in d["OpacityNoEd"]])
for _ in range(3):
    pb.get_fluxes(*args, ctx=ctx)
