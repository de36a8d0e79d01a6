"""Multi-GPU parallel tempering: the temperature ladder sharded across ranks (one process per GPU).

Each ensemble's ladder of `n_rungs` temperatures is cut into `world_size` contiguous blocks; rank r owns
rungs [r*L, (r+1)*L) of EVERY ensemble (L = n_rungs / world_size). Between exchange rounds the chains are
independent, so no collective is on the data path. One exchange round (MAIN:594-603, 962-1011, 1059-1065):

  1. all_gather of 4 doubles per chain (eta, stored likelihood, adapttemp, SSE of `w`)   -- <= 32 B / chain
  2. every rank runs the SAME deterministic sequential sweep (pairs 0-1, 1-2, ... per ensemble) with the
     same Philox swap uniforms -> identical permutation everywhere, no second collective for agreement
  3. configurations whose destination is on another rank travel once (send/recv of the fp32 vector);
     local moves are device copies
  4. every chain's `w` becomes a detached dict (MAIN:602)

The sweep arithmetic restates `run_swap_decide` of csrc/bae_kernels.cu on the host in float64.
"""
from __future__ import annotations

import math

import numpy as np

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c, k):
    """Philox4x32-10; c: 4 ints, k: 2 ints -> 4 ints (same generator as csrc/bae_common.cuh)."""
    c = [int(x) & 0xFFFFFFFF for x in c]
    k = [int(x) & 0xFFFFFFFF for x in k]
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF,
             p0 & 0xFFFFFFFF]
        k = [(k[0] + _W0) & 0xFFFFFFFF, (k[1] + _W1) & 0xFFFFFFFF]
    return c


def swap_uniform(seed, ensemble_global, rnd, pair):
    """The uniform `swap_procedure` draws for (ensemble, round, pair): bit-identical to the device stream."""
    r = philox4x32_10([pair, 0xFE | (3 << 8), rnd, ensemble_global], [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
    return float(np.float32((np.float32(r[0] >> 8) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)))


def swap_uniforms(seed, ensembles, rnd, n_pairs):
    """All swap uniforms of one exchange round, [len(ensembles), n_pairs] (vectorised Philox, same values as swap_uniform)."""
    E = len(ensembles)
    c = [np.tile(np.arange(n_pairs, dtype=np.uint64), (E, 1)), np.full((E, n_pairs), 0xFE | (3 << 8), np.uint64),
         np.full((E, n_pairs), rnd, np.uint64), np.repeat(np.asarray(ensembles, np.uint64)[:, None], n_pairs, 1)]
    k = [np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)]
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(_M0) * c[0], np.uint64(_M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k[0]) & m32, p1 & m32, ((p0 >> np.uint64(32)) ^ c[3] ^ k[1]) & m32, p0 & m32]
        k = [(k[0] + np.uint64(_W0)) & m32, (k[1] + np.uint64(_W1)) & m32]
    return ((c[0] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def loglik(sse, n, tau_sq, temp):
    return (-(n * 0.5) * math.log(2.0 * math.pi * tau_sq) - 0.5 * tau_sq * sse) / temp


def sweep(table, n_elems, uniforms):
    """Sequential adjacent-pair sweep over ONE ensemble (MAIN:1059-1065 with the message-slot quirk of
    MAIN:999-1005). table: [n_rungs, 4] = (eta, stored likelihood, adapttemp, SSE of w) by ladder position.
    Returns (src, n_swapped): src[p] = position whose configuration ends at p."""
    n = table.shape[0]
    src = list(range(n))
    cfg_prev, lh_prev, T_prev = 0, table[0, 1], table[0, 2]
    swapped = 0
    for p in range(n - 1):
        cfg1, cfg2 = cfg_prev, p + 1
        lhood1, T1 = lh_prev, T_prev
        lhood2, T2 = table[cfg2, 1], table[cfg2, 2]
        lhood12 = loglik(table[cfg1, 3], n_elems, math.exp(table[cfg1, 0]), T2)
        lhood21 = loglik(table[cfg2, 3], n_elems, math.exp(table[cfg2, 0]), T1)
        ex = (lhood12 - lhood1) + (lhood21 - lhood2)
        try:
            ee = math.exp(ex)
        except OverflowError:
            ee = math.inf
        prob = ee if ee < 1.0 else 1.0     # MAIN:988 min(1, exp(.)): a NaN exponent gives 1, as there and on the device
        if uniforms[p] < prob:
            swapped += 1
            src[p] = cfg2
            cfg_prev, lh_prev, T_prev = cfg1, lhood12, T1
        else:
            src[p] = cfg1
            cfg_prev, lh_prev, T_prev = cfg2, lhood2, T2
    src[n - 1] = cfg_prev
    return src, swapped


def sweep_native(table, n_elems, uniforms):
    """`sweep` through the library's host function `bae_host_sweep` (same arithmetic in C; None if the library is missing)."""
    try:
        import ctypes as C
        from . import _native as N
        lib = N.load()
    except Exception:
        return None
    tab = np.ascontiguousarray(table, dtype=np.float64)
    u = np.ascontiguousarray(uniforms, dtype=np.float32)
    src = np.zeros(tab.shape[0], np.int32)
    nsw = C.c_int(0)
    N.check(lib.bae_host_sweep(tab.ctypes.data, int(tab.shape[0]), float(n_elems), u.ctypes.data, src.ctypes.data, C.byref(nsw)))
    return [int(x) for x in src], int(nsw.value)


def plan_exchange(src, world, rank):
    """The transfers `rank` posts for the permutation `src` [E, R] (library host function `bae_plan_exchange`):
    list of (kind 'send' / 'recv', peer, ensemble, rung)."""
    import ctypes as C
    from . import _native as N
    lib = N.load()
    src = np.ascontiguousarray(src, dtype=np.int32)
    E, R = src.shape
    out = np.zeros((2 * E * R, 4), np.int32)
    n = lib.bae_plan_exchange(src.ctypes.data, E, R, int(world), int(rank), out.ctypes.data, 2 * E * R)
    if n < 0:
        N.check(n)
    return [("send" if k == 0 else "recv", int(peer), int(e), int(r)) for k, peer, e, r in out[:n]]


class NcclComm:
    """An ncclComm_t over the ranks of a torch.distributed job, created with the libnccl.so.2 that torch has loaded
    (the unique id travels through torch.distributed). The library never creates communicators itself:
    this object is what a caller passes to `DeviceSampler.swap_exchange` / `bae_swap_exchange`."""

    def __init__(self, dist, rank, world, device):
        import ctypes as C
        import torch

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]

        self._lib = C.CDLL("libnccl.so.2")
        self._lib.ncclGetErrorString.restype = C.c_char_p
        uid = UniqueId()
        if rank == 0:
            self._check(self._lib.ncclGetUniqueId(C.byref(uid)))
        backend = dist.get_backend()
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=(torch.device("cuda", device) if backend == "nccl" else "cpu"))
        dist.broadcast(t, 0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        torch.cuda.set_device(device)
        comm = C.c_void_p()
        self._lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        self._check(self._lib.ncclCommInitRank(C.byref(comm), int(world), uid, int(rank)))
        self.handle = comm
        self.rank, self.world = rank, world
        self.open_all_pairs()

    def open_all_pairs(self):
        """NCCL opens point-to-point connections lazily (tens of milliseconds each, on first use). A configuration can travel
        to any rank in a sweep, so every pair is opened here with a one-float exchange instead of inside a timed round."""
        import ctypes as C
        import torch
        if self.world < 2:
            return
        a = torch.zeros(self.world, dtype=torch.float32, device="cuda")
        b = torch.zeros(self.world, dtype=torch.float32, device="cuda")
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L = self._lib
        L.ncclSend.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ncclRecv.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self._check(L.ncclGroupStart())
        for peer in range(self.world):
            if peer == self.rank:
                continue
            self._check(L.ncclSend(C.c_void_p(a.data_ptr() + 4 * peer), 1, 7, peer, self.handle, stream))   # 7 = ncclFloat32
            self._check(L.ncclRecv(C.c_void_p(b.data_ptr() + 4 * peer), 1, 7, peer, self.handle, stream))
        self._check(L.ncclGroupEnd())
        torch.cuda.synchronize()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"NCCL error {rc}: {self._lib.ncclGetErrorString(rc).decode()}")

    def close(self):
        if self.handle:
            self._lib.ncclCommDestroy.argtypes = [__import__("ctypes").c_void_p]
            self._lib.ncclCommDestroy(self.handle)
            self.handle = None


class ShardedLadder:
    """Exchange rounds for a ladder sharded over `world_size` ranks. `transport` supplies the chain state:

      transport.scalars(local_chain) -> (eta, likelihood, adapttemp, last_sse)
      transport.pack(local_chain)    -> torch tensor [w_size] (fp32, on the rank's device) of the posted `w`
      transport.unpack(local_chain, tensor, eta, last_sse)
      transport.finish()             -> detach every chain's `w` (MAIN:602)

    Local chain index = ensemble * L + (rung - rank * L).
    """

    def __init__(self, dist, rank, world_size, n_ensembles, n_rungs, n_elems, seed, transport):
        if n_rungs % world_size:
            raise ValueError("n_rungs must be a multiple of world_size")
        self.dist, self.rank, self.world = dist, rank, world_size
        self.E, self.R, self.L = n_ensembles, n_rungs, n_rungs // world_size
        self.n_elems, self.seed, self.t = n_elems, seed, transport
        self.native_sweep = hasattr(transport, "apply_local")   # device transports use the library's host sweep
        self.round = 0
        self.num_swap = 0
        self.total_swap_proposals = 0
        self.vectors_sent = 0

    def owner(self, rung):
        return rung // self.L

    def local_index(self, ensemble, rung):
        return ensemble * self.L + (rung - self.rank * self.L)

    def gather_table(self):
        import torch
        if hasattr(self.t, "table"):
            loc = torch.from_numpy(np.ascontiguousarray(self.t.table(), dtype=np.float64))                  # [E*L, 4], one read
        else:
            loc = torch.tensor([self.t.scalars(c) for c in range(self.E * self.L)], dtype=torch.float64)   # [E*L, 4]
        dev = self.t.collective_device()
        loc = loc.to(dev)
        out = [torch.empty_like(loc) for _ in range(self.world)]
        self.dist.all_gather(out, loc)
        tab = np.zeros((self.E, self.R, 4))
        for r, part in enumerate(out):
            tab[:, r * self.L:(r + 1) * self.L, :] = part.cpu().numpy().reshape(self.E, self.L, 4)
        return tab

    def exchange(self):
        """One exchange round. Returns the per-ensemble permutations (identical on every rank)."""
        import contextlib
        with (self.t.ordering() if hasattr(self.t, "ordering") else contextlib.nullcontext()):
            return self._exchange()

    def _exchange(self):
        import torch
        tab = self.gather_table()
        perms = []
        uni = swap_uniforms(self.seed, list(range(self.E)), self.round, self.R - 1)
        for e in range(self.E):
            res = sweep_native(tab[e], self.n_elems, uni[e]) if self.native_sweep else None
            if res is None:
                res = sweep(tab[e], self.n_elems, [float(x) for x in uni[e]])
            src, nsw = res
            perms.append(src)
            self.num_swap += nsw
            self.total_swap_proposals += self.R - 1
        # data movement: each configuration moves once, from position src[p] to position p
        sends, recvs, local_moves = [], [], []
        for e in range(self.E):
            for p, s in enumerate(perms[e]):
                if s == p:
                    continue
                po, so = self.owner(p), self.owner(s)
                if po == self.rank and so == self.rank:
                    local_moves.append((e, s, p))
                elif so == self.rank:
                    sends.append((e, s, p, po))
                elif po == self.rank:
                    recvs.append((e, s, p, so))
        batched = hasattr(self.t, "apply_local")
        staged = {}
        if not batched:
            for e, s, p in local_moves:
                staged[(e, p)] = self.t.pack(self.local_index(e, s)).clone()
        ops, bufs = [], {}
        for e, s, p, dst in sends:       # packed BEFORE any local move can overwrite the source chain
            buf = self.t.pack(self.local_index(e, s)).clone()
            bufs[("s", e, p)] = buf
            ops.append(self.dist.P2POp(self.dist.isend, buf, dst))
            self.vectors_sent += 1
        for e, s, p, frm in recvs:
            buf = torch.empty_like(self.t.pack(self.local_index(e, p))) if not batched else self.t.empty_vector()
            bufs[("r", e, p)] = buf
            ops.append(self.dist.P2POp(self.dist.irecv, buf, frm))
        if batched:
            # every same-GPU move (and the detach of all chains, MAIN:602) in one gather + one scatter pass on the device
            src = np.arange(self.E * self.L, dtype=np.int32)
            for e, s, p in local_moves:
                src[self.local_index(e, p)] = self.local_index(e, s)
            self.t.apply_local(src)
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        if not batched:
            for e, s, p in local_moves:
                self.t.unpack(self.local_index(e, p), staged[(e, p)], tab[e, s, 0], tab[e, s, 3])
        for e, s, p, frm in recvs:
            self.t.unpack(self.local_index(e, p), bufs[("r", e, p)], tab[e, s, 0], tab[e, s, 3])
        if not batched:
            self.t.finish()
        self.round += 1
        return perms


class DeviceTransport:
    """`ShardedLadder` transport over a `DeviceSampler` (CUDA tensors, NCCL)."""

    def __init__(self, sampler, device):
        import torch
        self.s = sampler
        self.dev = torch.device("cuda", device)
        self._buf = torch.empty(sampler.w_size, dtype=torch.float32, device=self.dev)
        self._sc = np.zeros(4, np.float64)
        # the library's stream as torch's current stream during an exchange: packs, clones, NCCL sends and unpacks are
        # then ordered on the device without host synchronisation
        self._stream = torch.cuda.ExternalStream(sampler.lib.bae_stream(sampler._h), device=self.dev)

    def ordering(self):
        import torch
        return torch.cuda.stream(self._stream)

    def collective_device(self):
        return self.dev

    def scalars(self, chain):
        st = self.s.get_state(chain, vectors=False)
        return (st["eta"], st["likelihood"], st["adapttemp"], st["last_sse_train"])

    def table(self):
        from . import _native as N
        out = np.zeros((self.s.n_chains, 4), np.float64)
        N.check(self.s.lib.bae_exchange_table(self.s._h, out.ctypes.data))
        return out

    def pack(self, chain):
        from . import _native as N
        N.check(self.s.lib.bae_exchange_pack(self.s._h, int(chain), self._buf.data_ptr(), None))
        return self._buf

    def unpack(self, chain, tensor, eta, last_sse):
        from . import _native as N
        sc = np.array([eta, last_sse], np.float64)
        N.check(self.s.lib.bae_exchange_unpack(self.s._h, int(chain), tensor.data_ptr(), sc.ctypes.data))

    def finish(self):
        from . import _native as N
        N.check(self.s.lib.bae_exchange_finish(self.s._h))

    def empty_vector(self):
        import torch
        return torch.empty(self.s.w_size, dtype=torch.float32, device=self.dev)

    def apply_local(self, src):
        from . import _native as N
        src = np.ascontiguousarray(src, dtype=np.int32)
        N.check(self.s.lib.bae_exchange_local(self.s._h, src.ctypes.data))
