"""CPU: world_size-2 (and 3) gloo runs of the wavelength-sharding host logic.  The per-rank
compute is the CPU oracle here (test infrastructure); on the GPU box the same code path runs the
CUDA kernels (tests/test_gpu_parity.py::test_sharded_matches_unsharded)."""
import os
import socket

import numpy as np
import pytest

import cases as C
import oracle
from picaso_b200 import sharded, synth

KW = dict(single_phase=3, multi_phase=0, toon_coefficients=0, get_lvl_flux=0)


def test_partition_properties():
    for n, w in [(10, 3), (0, 2), (7, 8), (196000, 8), (10000, 1)]:
        p = sharded.partition(n, w)
        assert len(p) == w and p[0][0] == 0 and p[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(p, p[1:]))
        sizes = [e - s for s, e in p]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, W, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synth.reflected_inputs(L=12, W=W, seed=42)

    def compute(sd):
        x, _ = oracle.get_reflected_1d(*C.reflected_args(sd, KW))
        return oracle.compress_disco(sd["nwno"], sd["cos_theta"], x, sd["gweight"], sd["tweight"], sd["F0PI"])

    full = sharded.run_sharded(compute, d, W)
    xs = sharded.run_sharded(lambda sd: oracle.get_reflected_1d(*C.reflected_args(sd, KW))[0], d, W)
    q.put((rank, full, xs))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W", [(2, 37), (3, 2)])  # (3, 2): the last rank owns no wavelength
def test_gloo_sharded_equals_unsharded(world, W):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = synth.reflected_inputs(L=12, W=W, seed=42)
    x, _ = oracle.get_reflected_1d(*C.reflected_args(d, KW))
    alb = oracle.compress_disco(W, d["cos_theta"], x, d["gweight"], d["tweight"], d["F0PI"])
    for rank, full, xs in res:
        assert np.array_equal(full, alb), "rank %d albedo" % rank
        assert np.array_equal(xs, x), "rank %d xint" % rank
