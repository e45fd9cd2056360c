"""fp64-pipe machine numbers of the GPU box (python scripts/microbench.py) - one JSON line each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import picaso_b200 as pb
ctx = pb.Context(0)
nsm = ctx.sm_count()
names = ["dfma_throughput", "dfma_latency", "dfma_half_warp_throughput", "lds64_throughput", "rcp_latency",
         "dfma_3warps_per_smsp_ilp2"]
for w, n in enumerate(names):
    r = ctx.microbench(w, 8192)
    r["name"] = n
    r["sm"] = nsm
    if "throughput" in n or "3warps" in n:
        r["warp_instr_per_clk_per_sm_at_1965MHz"] = r["gops"] * 1e9 / 32 / nsm / 1.965e9
    if w in (0, 2, 5):
        r["tflops"] = 2 * r["gops"] / 1e3
    print(json.dumps(r))
