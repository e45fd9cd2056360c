"""Wavelength sharding across ranks (one process per GPU).

Wavelength bins are independent through the whole hot path: every
reference loop is elementwise over wavelengths or recurrent over layers.  So the multi-GPU
scheme is: contiguous wave slabs per rank, no data-path collective, and one all-gather of
the final [nwno] vector(s) (albedo / thermal flux / transit depth).  torch.distributed is
used for the plumbing only (NCCL on GPUs, gloo in the CPU tests) and imported lazily.
"""
import numpy as np

__all__ = ["partition", "wave_slice", "shard_inputs", "allgather_waves", "run_sharded"]


def partition(nwno, world):
    """Balanced contiguous slabs: list of (start, stop), sizes differ by at most one."""
    base, rem = divmod(int(nwno), int(world))
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((s, s + n))
        s += n
    return out


def wave_slice(nwno, rank, world):
    s, e = partition(nwno, world)[rank]
    return slice(s, e)


# arguments of the reference flux functions that carry a wavelength axis (last axis)
WAVE_KEYS = frozenset(["dtau", "tau", "w0", "cosb", "gcos2", "ftau_cld", "ftau_ray", "dtau_og",
                       "tau_og", "w0_og", "cosb_og", "f_deltaM", "surf_reflect", "F0PI", "b_top",
                       "wno", "dwno", "DTAU", "w0_no_raman"])


def shard_inputs(inputs, nwno, rank, world, wave_keys=None):
    """Slice the wavelength (last) axis of the named arrays (default: WAVE_KEYS, the reference's
    argument names); everything else (geometry, scalars, per-level profiles) is replicated."""
    sl = wave_slice(nwno, rank, world)
    keys = WAVE_KEYS if wave_keys is None else wave_keys
    out = {}
    for k, v in inputs.items():
        is_wave = k in keys and isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[-1] == nwno
        out[k] = np.ascontiguousarray(v[..., sl]) if is_wave else v
    if "nwno" in out:
        out["nwno"] = sl.stop - sl.start
    return out


def allgather_waves(local, nwno, group=None):
    """All-gather per-rank results along the wavelength (last) axis -> full [..., nwno] array on
    every rank.  `local` is this rank's numpy slab (or a torch tensor on the rank's device)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = partition(nwno, world)
    nmax = max(e - s for s, e in parts)
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        t = t.cuda()
    lead = tuple(t.shape[:-1])
    pad = torch.zeros(lead + (nmax,), dtype=t.dtype, device=t.device)
    pad[..., : t.shape[-1]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    full = torch.cat([b[..., : e - s] for b, (s, e) in zip(bufs, parts)], dim=-1)
    return full.cpu().numpy() if is_np else full


def run_sharded(compute, inputs, nwno, group=None, wave_keys=None):
    """compute(shard_dict) -> ndarray [..., n_local]; returns the gathered [..., nwno] result."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = compute(shard_inputs(inputs, nwno, rank, world, wave_keys))
    return allgather_waves(np.asarray(local), nwno, group)
