"""Wavelength sharding across ranks (one process per GPU).

Wavelength bins are independent through the whole hot path: every
reference loop is elementwise over wavelengths or recurrent over layers.  So the multi-GPU
scheme is: contiguous wave slabs per rank, no data-path collective, and one all-gather of
the final [nwno] vector(s) (albedo / thermal flux / transit depth).

No PyTorch in here.  The host-side helpers take an `exchange` callable - exchange(obj) returns the list of
every rank's obj - which is all the plumbing they need: `TcpExchange` (plain sockets, below) for
stand-alone use, or a one-line wrapper around whatever the caller already runs (torch.distributed
all_gather_object in bench.py and in the gloo tests, mpi4py allgather under the retrieval driver).
`PeerAllGather` is the device-side collective: the producing kernel (or a side-stream copy kernel)
stores each rank's slab into every rank's buffer over NVLink peer memory; `exchange` is used once, to
swap the CUDA IPC handles.
"""
import pickle
import socket
import struct
import time

import numpy as np

__all__ = ["partition", "wave_slice", "shard_inputs", "allgather_waves", "run_sharded", "PeerAllGather", "TcpExchange"]


class TcpExchange:
    """exchange(obj) -> [obj of rank 0, ..., obj of rank world-1] over plain TCP (rank 0 listens on addr:port;
    the connections stay open between calls).  Host-side rendezvous only - nothing on the data path."""

    def __init__(self, rank, world, addr="127.0.0.1", port=29533, timeout=120.0):
        self.rank, self.world = int(rank), int(world)
        self.peers, self.sock = [], None
        if self.world == 1:
            return
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(self.world)
            srv.settimeout(timeout)
            conns = {}
            while len(conns) < self.world - 1:
                c, _ = srv.accept()
                c.settimeout(timeout)
                conns[struct.unpack("<i", self._recvn(c, 4))[0]] = c
            srv.close()
            self.peers = [conns[r] for r in range(1, self.world)]
        else:
            t0 = time.time()
            while True:
                try:
                    self.sock = socket.create_connection((addr, port), timeout=timeout)
                    break
                except OSError:
                    if time.time() - t0 > timeout:
                        raise
                    time.sleep(0.05)
            self.sock.sendall(struct.pack("<i", self.rank))

    @staticmethod
    def _recvn(c, n):
        buf = b""
        while len(buf) < n:
            chunk = c.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("peer closed the exchange socket")
            buf += chunk
        return buf

    @classmethod
    def _send(cls, c, obj):
        data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        c.sendall(struct.pack("<q", len(data)) + data)

    @classmethod
    def _recv(cls, c):
        n = struct.unpack("<q", cls._recvn(c, 8))[0]
        return pickle.loads(cls._recvn(c, n))

    def __call__(self, obj):
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            parts = [obj] + [self._recv(c) for c in self.peers]
            for c in self.peers:
                self._send(c, parts)
            return parts
        self._send(self.sock, obj)
        return self._recv(self.sock)

    def close(self):
        for c in self.peers + ([self.sock] if self.sock else []):
            try:
                c.close()
            except OSError:
                pass
        self.peers, self.sock = [], None


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


def allgather_waves(local, nwno, rank, world, exchange):
    """All-gather per-rank results along the wavelength (last) axis -> full [..., nwno] numpy array on every
    rank.  `local` is this rank's slab (ragged slabs are fine, a rank may own no wavelength at all);
    exchange(obj) returns every rank's obj."""
    parts = partition(nwno, world)
    mine = np.ascontiguousarray(local)
    s, e = parts[rank]
    if mine.shape[-1] != e - s:
        raise ValueError("rank %d holds %d wavelengths, its slab has %d" % (rank, mine.shape[-1], e - s))
    slabs = exchange(mine)
    return np.concatenate([np.asarray(x) for x in slabs], axis=-1)


def run_sharded(compute, inputs, nwno, rank, world, exchange, wave_keys=None):
    """compute(shard_dict) -> ndarray [..., n_local]; returns the gathered [..., nwno] result."""
    local = compute(shard_inputs(inputs, nwno, rank, world, wave_keys))
    return allgather_waves(np.asarray(local), nwno, rank, world, exchange)


def ctypes_addr(obj):
    import ctypes
    return ctypes.addressof(obj)


class PeerAllGather:
    """The all-gather of the final [nwno] slab over NVLink peer memory (include/picaso_b200.h: pb_peer_gather).

    One instance per rank (one process per GPU).  Every rank owns `nbuf` rotating gathered buffers
    [world][nwno] and an arrival-flag array [world]; `exchange(obj)` must return the list of every rank's
    `obj` (TcpExchange, or a wrapper of the caller's own collective) and is used once, to swap the CUDA IPC handles.
    Per step: ``a.gather = g.next()`` on the ReflectedArgs of a PB_DEVICE `pb_reflected_toon_1d` call; the
    slab of step s lands in row `rank` of buffer s % nbuf on every rank.  ``g.wait()`` enqueues a
    stream-ordered wait for the last step of every rank; ``g.gathered()`` reads the local buffer.
    push=True (default): a side-stream copy kernel pushes the slab while the next launch computes;
    push=False: the solver kernel's epilogue stores to the peers itself and publishes the flags before it retires;
    push="lazy": the epilogue only stores - the flags of step s are published by the FIRST CTA of launch s + 1
    (stores of a finished grid are performed system-wide, so no fence / NVLink round trip sits on any launch's
    critical path); `wait()` publishes the last step.
    push="deferred": the solver CTAs only fill the local row; the launch of step s + 1 carries one extra CTA that
    pushes the slab of step s to the peers and publishes it while the solver CTAs compute (no peer store in any
    solver CTA, nothing at the grid's tail); `wait()` delivers the last step (pb_peer_flush).  `barrier()` is a device-side barrier over the same peer
    mappings (stream-ordered): ranks leave it within a microsecond of each other whatever the host skew."""

    def __init__(self, ctx, rank, world, nwno, nbuf=3, push=True, exchange=None, _peers=None):
        import ctypes
        from ._lib import PeerGather
        if not 1 <= world <= 8:
            raise ValueError("PeerAllGather: 1 <= world <= 8")
        if not 3 <= int(nbuf) <= 8:
            raise ValueError("PeerAllGather: nbuf must be 3..8 (ranks may run nbuf - 2 steps apart; pb_peer_gather has 8 slots)")
        self.ctx, self.rank, self.world, self.nwno, self.nbuf = ctx, rank, world, nwno, int(nbuf)
        self.push = 2 if push == "lazy" else 3 if push == "deferred" else int(bool(push))
        self.step, self.epoch, self._flushed = 0, 0, 0
        gbytes = self.nbuf * world * nwno * 8
        self.d_gath, self.d_flags, self.d_done = ctx.dev_alloc(gbytes), ctx.dev_alloc(256), ctx.dev_alloc(256)
        for ptr, nb in ((self.d_gath, gbytes), (self.d_flags, 256), (self.d_done, 256)):
            ctx.check(ctx.lib.pb_memset(ctx.h, ptr, 0, nb))
        ctx.sync()
        self._opened = []
        if _peers is not None:            # ranks living in one process (tests): plain device pointers
            self._setup = lambda: self._bind([p.d_gath for p in _peers], [p.d_flags for p in _peers])
            return
        hg, hf = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        ctx.check(ctx.lib.pb_ipc_export(ctx.h, self.d_gath, hg))
        ctx.check(ctx.lib.pb_ipc_export(ctx.h, self.d_flags, hf))
        if world == 1:
            handles = [(hg.raw, hf.raw)]
        elif exchange is not None:
            handles = exchange((hg.raw, hf.raw))
        else:
            raise ValueError("PeerAllGather over several processes needs an `exchange` callable (sharded.TcpExchange "
                             "or a wrapper of the caller's all-gather) to swap the CUDA IPC handles")
        pg, pf = [], []
        for r, (rg, rf) in enumerate(handles):
            if r == rank:
                pg.append(self.d_gath)
                pf.append(self.d_flags)
            else:
                a, b = ctypes.c_void_p(), ctypes.c_void_p()
                ctx.check(ctx.lib.pb_ipc_open(ctx.h, rg, ctypes.byref(a)))
                ctx.check(ctx.lib.pb_ipc_open(ctx.h, rf, ctypes.byref(b)))
                pg.append(a.value)
                pf.append(b.value)
                self._opened += [a.value, b.value]
        self._bind(pg, pf)

    def _bind(self, peer_g, peer_f):
        import ctypes
        from ._lib import PeerGather
        world, W = self.world, self.nwno
        self._flag_ptrs = (ctypes.c_void_p * world)(*peer_f)
        self._alb_ptrs = [(ctypes.c_void_p * world)(*[g + b * world * W * 8 for g in peer_g]) for b in range(self.nbuf)]
        self._structs = []
        for b in range(self.nbuf):
            s = PeerGather()
            s.nranks, s.rank = world, self.rank
            s.albedo, s.flags = ctypes.addressof(self._alb_ptrs[b]), ctypes.addressof(self._flag_ptrs)
            s.done_counter = self.d_done
            self._structs.append(s)

    @classmethod
    def local_group(cls, ctxs, nwno, nbuf=3, push=True):
        """`len(ctxs)` ranks inside one process (one Context = one stream each), e.g. on a single GPU"""
        world = len(ctxs)
        group = []
        for r, c in enumerate(ctxs):
            group.append(cls(c, r, world, nwno, nbuf=nbuf, push=push, _peers=group))
        for g in group:
            g._setup()
        return group

    def next(self):
        """address of the pb_peer_gather struct of the next step (valid until the call after next)"""
        import ctypes
        self.step += 1
        s = self._structs[self.step % self.nbuf]
        s.step, s.wait_step = self.step, max(0, self.step - (self.nbuf - 1))
        s.push, s.slot = self.push, self.step % self.nbuf
        s.albedo_prev = ctypes.addressof(self._alb_ptrs[(self.step - 1) % self.nbuf]) if self.step > 1 else None
        return ctypes.addressof(s)

    def wait(self):
        if self.step > 0:
            if self.push == 2:  # lazy flags: nobody has published the last step yet
                self.ctx.check(self.ctx.lib.pb_peer_signal(self.ctx.h, ctypes_addr(self._flag_ptrs), self.world, self.rank, 0, self.step))
            elif self.push == 3:  # deferred: the last step's slab has no next launch to carry it
                last = self._structs[self.step % self.nbuf]
                if self._flushed != self.step:
                    self.ctx.check(self.ctx.lib.pb_peer_flush(self.ctx.h, ctypes_addr(last), self.nwno))
                    self._flushed = self.step
            self.ctx.check(self.ctx.lib.pb_gather_wait(self.ctx.h, self.d_flags, self.world, self.step, self.d_done + 8))

    BARRIER_OFFSET = 16   # flag words 16 .. 16 + world of the 256-byte flag block count barrier epochs

    def barrier(self):
        """device-side barrier on the context's stream: signal this rank's arrival on every rank, wait for all"""
        self.epoch += 1
        self.ctx.check(self.ctx.lib.pb_peer_signal(self.ctx.h, ctypes_addr(self._flag_ptrs), self.world, self.rank,
                                                   self.BARRIER_OFFSET, self.epoch))
        self.ctx.check(self.ctx.lib.pb_gather_wait(self.ctx.h, self.d_flags + 8 * self.BARRIER_OFFSET, self.world, self.epoch,
                                                   self.d_done + 8))

    def gathered(self, step=None):
        step = self.step if step is None else step
        return self.ctx.from_device(self.d_gath + (step % self.nbuf) * self.world * self.nwno * 8, (self.world, self.nwno))

    def courier_us(self):
        """push="deferred": how long the last courier CTA took (guard wait + copies + fence), microseconds"""
        import struct
        w = struct.unpack("<8I", self.ctx.from_device(self.d_done, (4,)).tobytes())
        return w[5] / 1e3

    def timed_out(self):
        import struct
        w = struct.unpack("<4I", self.ctx.from_device(self.d_done, (2,)).tobytes())
        return bool(w[1] or w[2])

    def close(self):
        for p in self._opened:
            self.ctx.lib.pb_ipc_close(self.ctx.h, p)
        self._opened = []
        for p in (self.d_gath, self.d_flags, self.d_done):
            if p is not None:
                self.ctx.dev_free(p)
        self.d_gath = self.d_flags = self.d_done = None
