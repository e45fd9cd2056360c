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


def _gloo_exchange(dist, world):
    """exchange(obj) -> every rank's obj, over the test's own gloo group (the product takes any such callable)"""
    def exchange(obj):
        parts = [None] * world
        dist.all_gather_object(parts, obj)
        return parts
    return exchange


def _worker(rank, world, port, W, q, transport="gloo"):
    if transport == "gloo":
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        exchange = _gloo_exchange(dist, world)
    else:
        exchange = sharded.TcpExchange(rank, world, port=port)
    d = synth.reflected_inputs(L=12, W=W, seed=42)

    def compute(sd):
        x, _ = oracle.get_reflected_1d(*C.reflected_args(sd, KW))
        return oracle.compress_disco(sd["nwno"], sd["cos_theta"], x, sd["gweight"], sd["tweight"], sd["F0PI"])

    full = sharded.run_sharded(compute, d, W, rank, world, exchange)
    xs = sharded.run_sharded(lambda sd: oracle.get_reflected_1d(*C.reflected_args(sd, KW))[0], d, W, rank, world, exchange)
    q.put((rank, full, xs))
    if transport == "gloo":
        dist.destroy_process_group()
    else:
        exchange.close()


@pytest.mark.parametrize("transport", ["gloo", "tcp"])
@pytest.mark.parametrize("world,W", [(2, 37), (3, 2)])  # (3, 2): the last rank owns no wavelength
def test_gloo_sharded_equals_unsharded(world, W, transport):
    """world_size 2 / 3 over torch.distributed gloo, and over the package's own socket rendezvous (TcpExchange)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, q, transport)) for r in range(world)]
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


# ---- atmosphere sharding of the batched retrieval model (BASELINE cfg5; picaso_b200/batch.py: shard) ----
def _batch_inputs(B, L, W):
    ds = [synth.thermal_inputs(L=L, W=W, seed=700 + b, t_range=(300.0 + 15 * b, 1400.0 + 30 * b)) for b in range(B)]
    d0 = ds[0]
    return dict(wno=d0["wno"], tlevel=np.array([d["tlevel"] for d in ds]), plevel=np.array([d["plevel"] for d in ds]),
                dtau=np.array([d["dtau"] for d in ds]), w0=np.array([d["w0"] for d in ds]),
                cosb=np.array([d["cosb"] for d in ds]), ubar1=d0["ubar1"], gweight=d0["gweight"], tweight=d0["tweight"])


def _batch_worker(rank, world, port, B, q):
    import torch
    import torch.distributed as dist
    from oracle import regrid as oreg
    from picaso_b200.batch import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = _batch_inputs(B, 9, 120)
    newx = np.linspace(500.0, 9000.0, 17)
    lo, hi = shard(B, rank, world)
    sub = dict(kw, **{k: kw[k][lo:hi] for k in ("tlevel", "plevel", "dtau", "w0", "cosb")})
    _, y = (oreg.thermal_batch(**sub, newx=newx, scale=2.0) if hi > lo else (None, np.zeros((0, 17))))
    # no exchange on the data path; gathering the per-rank rows (ragged) only assembles the result
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, y))
    full = np.zeros((B, 17))
    for a, b, rows in parts:
        full[a:b] = rows
    q.put((rank, full))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 5), (3, 2)])  # (3, 2): the last rank owns no atmosphere
def test_gloo_atmosphere_shards_equal_whole_batch(world, B):
    import torch.multiprocessing as mp
    from oracle import regrid as oreg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, want = oreg.thermal_batch(**_batch_inputs(B, 9, 120), newx=np.linspace(500.0, 9000.0, 17), scale=2.0)
    for rank, full in res:
        assert np.array_equal(full, want, equal_nan=True), "rank %d" % rank


def test_product_has_no_torch_import():
    """north_star: Python host code over ctypes, no PyTorch - the package never imports torch"""
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "picaso_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import torch|from torch)", src, flags=re.M), fn
