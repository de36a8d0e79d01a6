"""Chain-batched device sampler: thin Python object over the C-ABI handle.

One `DeviceSampler` drives the chains resident on one GPU. All arithmetic of the reference's
`ptReplica.run` loop body (MAIN:432-592) and of `ParallelTempering.swap_procedure` (MAIN:962-1011)
happens inside the library's CUDA step program; this class only moves host buffers in and out.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _native as N

SORTED_KEYS = [
    "decode.0.bias", "decode.0.num_batches_tracked", "decode.0.running_mean", "decode.0.running_var",
    "decode.0.weight", "decode.1.bias", "decode.1.weight", "decode.3.bias", "decode.3.weight",
    "decode.5.bias", "decode.5.weight", "encode.0.bias", "encode.0.weight", "encode.2.bias",
    "encode.2.weight", "encode.4.bias", "encode.4.weight",
]


def param_shapes(dims4):
    """Shapes of the state_dict tensors in `getparameters` order (MAIN:151-176, 228-241)."""
    d0, d1, d2, d3 = [int(x) for x in dims4]
    return {
        "decode.0.bias": (d3,), "decode.0.num_batches_tracked": (), "decode.0.running_mean": (d3,),
        "decode.0.running_var": (d3,), "decode.0.weight": (d3,),
        "decode.1.bias": (d2,), "decode.1.weight": (d2, d3), "decode.3.bias": (d1,), "decode.3.weight": (d1, d2),
        "decode.5.bias": (d0,), "decode.5.weight": (d0, d1), "encode.0.bias": (d1,), "encode.0.weight": (d1, d0),
        "encode.2.bias": (d2,), "encode.2.weight": (d2, d1), "encode.4.bias": (d3,), "encode.4.weight": (d3, d2),
    }


def param_offsets(dims4):
    off, out = 0, {}
    for k in SORTED_KEYS:
        shp = param_shapes(dims4)[k]
        n = int(np.prod(shp)) if shp else 1
        out[k] = (off, n, shp)
        off += n
    return out, off


def default_theta_init(dims4, rng: np.random.Generator) -> np.ndarray:
    """A torch-default-style initial state_dict as a flat fp32 vector: Linear weights and biases
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)); BatchNorm gamma 1, beta 0, running_mean 0, running_var 1,
    num_batches_tracked 0 (what `Model()` holds before the chain starts, MAIN:270)."""
    offs, total = param_offsets(dims4)
    th = np.zeros(total, dtype=np.float32)
    shapes = param_shapes(dims4)
    for name in ("encode.0", "encode.2", "encode.4", "decode.1", "decode.3", "decode.5"):
        o, n, shp = offs[name + ".weight"]
        bound = 1.0 / math.sqrt(shp[1])
        th[o:o + n] = rng.uniform(-bound, bound, n).astype(np.float32)
        ob, nb, _ = offs[name + ".bias"]
        th[ob:ob + nb] = rng.uniform(-bound, bound, nb).astype(np.float32)
    o, n, _ = offs["decode.0.weight"]
    th[o:o + n] = 1.0
    o, n, _ = offs["decode.0.running_var"]
    th[o:o + n] = 1.0
    return th


@dataclass
class Hyper:
    lr: float = 0.01          # MAIN:80,89,101
    step_w: float = 0.005     # MAIN:81,88,102
    step_eta: float = 0.2     # MAIN:387
    l_prob: float = 0.9       # MAIN:286
    sigma_squared: float = 25.0
    nu_1: float = 0.0
    nu_2: float = 3.0
    use_langevin: bool = True
    pt_samples: float = 0.5   # MAIN:65
    drift_mode: int = 0       # 0 Adam only (experiment_type 2, MAIN:72); 1 Adam then SGD after pt (experiment_type 1); 2 MALA


TRACE_DTYPE = np.dtype([
    ("likelihood_proposal", "f8"), ("likelihood", "f8"), ("sum_value", "f8"), ("log_u", "f8"),
    ("prior_proposal", "f8"), ("diff_prop", "f8"), ("mse_train", "f8"), ("mse_test", "f8"),
    ("mse_train_proposal", "f8"), ("mse_test_proposal", "f8"), ("eta", "f8"), ("weights", "f4", (13,)),
    ("accepted", "i4"), ("langevin", "i4"), ("step", "i4"),
], align=True)
assert TRACE_DTYPE.itemsize == C.sizeof(N.Trace), (TRACE_DTYPE.itemsize, C.sizeof(N.Trace))


class DeviceSampler:
    def __init__(self, dims4, train, test, n_chains, n_rungs, samples, temperatures, swap_interval=5,
                 seed=0xBAE5EED, device=0, chain_id_base=0, hyper: Hyper = None, gemm_path=0, act_group=0):
        self.lib = N.load()
        self.hyper = hyper or Hyper()
        h = self.hyper
        train = np.ascontiguousarray(train, dtype=np.float32)
        test = np.ascontiguousarray(test, dtype=np.float32)
        if train.ndim != 2 or test.ndim != 2 or train.shape[1] != dims4[0] or test.shape[1] != dims4[0]:
            raise ValueError("train/test must be [rows, in_shape] matrices")
        cfg = N.Config()
        for i in range(4):
            cfg.dims[i] = int(dims4[i])
        cfg.n_train, cfg.n_test = train.shape[0], test.shape[0]
        cfg.n_chains, cfg.n_rungs, cfg.chain_id_base = int(n_chains), int(n_rungs), int(chain_id_base)
        cfg.samples = int(samples)
        cfg.pt_switch_step = int(h.pt_samples * samples)     # MAIN:429
        cfg.swap_interval = int(swap_interval)
        cfg.use_langevin = 1 if h.use_langevin else 0
        cfg.lr, cfg.step_w, cfg.step_eta, cfg.l_prob = h.lr, h.step_w, h.step_eta, h.l_prob
        cfg.sigma_squared, cfg.nu_1, cfg.nu_2 = h.sigma_squared, h.nu_1, h.nu_2
        cfg.seed = int(seed)
        cfg.gemm_path = int(gemm_path)
        cfg.drift_mode = int(h.drift_mode)
        cfg.act_group = int(act_group)
        self.cfg = cfg
        self.dims4 = tuple(int(x) for x in dims4)
        self.n_chains, self.n_rungs, self.samples = int(n_chains), int(n_rungs), int(samples)
        self.swap_interval = int(swap_interval)
        self._h = C.c_void_p()
        N.check(self.lib.bae_create(C.byref(cfg), int(device), C.byref(self._h)))
        self.w_size = self.lib.bae_w_size(self._h)
        self.upload_data(train, test)
        T = np.ascontiguousarray(temperatures, dtype=np.float64)
        if T.shape != (n_chains,):
            raise ValueError("need one temperature per chain")
        N.check(self.lib.bae_set_ladder(self._h, T.ctypes.data))
        self.temperatures = T

    # -- lifecycle ----------------------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.bae_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data / state -------------------------------------------------------------------------
    def upload_data(self, train, test):
        train = np.ascontiguousarray(train, dtype=np.float32)
        test = np.ascontiguousarray(test, dtype=np.float32)
        self._train, self._test = train, test   # keep host buffers alive for async copies
        N.check(self.lib.bae_upload_data(self._h, train.ctypes.data, test.ctypes.data))

    def init_chain(self, chain, theta_init, w0):
        th = np.ascontiguousarray(theta_init, dtype=np.float32)
        w0 = np.ascontiguousarray(w0, dtype=np.float64)
        assert th.shape == (self.w_size,) and w0.shape == (self.w_size,)
        N.check(self.lib.bae_init_chain(self._h, int(chain), th.ctypes.data, w0.ctypes.data))

    def get_state(self, chain, vectors=True):
        sc = N.ChainScalars()
        if vectors:
            th, w, m, v = (np.zeros(self.w_size, np.float32) for _ in range(4))
            N.check(self.lib.bae_get_state(self._h, int(chain), th.ctypes.data, w.ctypes.data, m.ctypes.data,
                                           v.ctypes.data, C.byref(sc)))
        else:
            th = w = m = v = None
            N.check(self.lib.bae_get_state(self._h, int(chain), None, None, None, None, C.byref(sc)))
        out = {k: getattr(sc, k) for k, _ in N.ChainScalars._fields_}
        out.update(theta=th, w=w, m=m, v=v)
        return out

    def set_state(self, chain, theta=None, w=None, m=None, v=None, scalars: dict = None):
        def p(a):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dtype=np.float32)
            assert a.shape == (self.w_size,)
            return a, a.ctypes.data
        th, pth = p(theta)
        ww, pw = p(w)
        mm, pm = p(m)
        vv, pv = p(v)
        sc = None
        if scalars is not None:
            sc = N.ChainScalars()
            for k, val in scalars.items():
                if k in ("theta", "w", "m", "v"):
                    continue
                setattr(sc, k, val)
        N.check(self.lib.bae_set_state(self._h, int(chain), pth, pw, pm, pv, C.byref(sc) if sc is not None else None))

    # -- sampling -----------------------------------------------------------------------------
    def step(self, n=1):
        N.check(self.lib.bae_step(self._h, int(n)))

    def swap(self):
        N.check(self.lib.bae_swap_local(self._h))

    def run(self, n):
        N.check(self.lib.bae_run(self._h, int(n)))

    def synchronize(self):
        N.check(self.lib.bae_synchronize(self._h))

    def last_ms(self):
        ms = C.c_float()
        N.check(self.lib.bae_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.bae_launch_count(self._h))

    def gemm_path(self):
        return int(self.lib.bae_gemm_path(self._h))

    def step_kernel(self):
        """1 = persistent FP32 tile kernel, 2 = tcgen05 phase program, 3 = narrow-width cluster kernel."""
        return int(self.lib.bae_step_kernel(self._h))

    def mma_kind(self):
        """0 = FFMA tiles, 1 = tcgen05 3xTF32 split, 2 = tcgen05 3xFP16 split (the default on the tensor-core path)."""
        return int(self.lib.bae_mma_kind(self._h))

    def traces(self, chain, first=0, n=None):
        n = self.samples - first if n is None else n
        out = np.zeros(n, dtype=TRACE_DTYPE)
        N.check(self.lib.bae_read_traces(self._h, int(chain), int(first), int(n), out.ctypes.data))
        return out

    def traces_all(self, first=0, n=None):
        """Trace records [n_chains][n] of every resident chain in one device -> host copy."""
        n = self.samples - first if n is None else n
        out = np.zeros((self.n_chains, n), dtype=TRACE_DTYPE)
        N.check(self.lib.bae_read_traces_all(self._h, int(first), int(n), out.ctypes.data))
        return out

    def swap_counts(self):
        a, b = C.c_int64(), C.c_int64()
        N.check(self.lib.bae_swap_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def forward_all(self, which=0, features=True, recon=True):
        """Batched forward of the train (0) / test (1) matrix through every resident chain's live model:
        dict(features [C, rows, d3] = `encode(x)`, recon [C, rows, d0] = `forward(x)`, sse [C]); MAIN:709-818."""
        rows = self.cfg.n_train if which == 0 else self.cfg.n_test
        f = np.zeros((self.n_chains, rows, self.dims4[3]), np.float32) if features else None
        r = np.zeros((self.n_chains, rows, self.dims4[0]), np.float32) if recon else None
        sse = np.zeros(self.n_chains, np.float64)
        N.check(self.lib.bae_forward_all(self._h, int(which), f.ctypes.data if features else None,
                                         r.ctypes.data if recon else None, sse.ctypes.data))
        return dict(features=f, recon=r, sse=sse)

    # -- ladder sharded over GPUs (one process per GPU) ---------------------------------------------------
    def shard(self, rank, world):
        """Declare this handle's chains (E ensembles x n_rungs local rungs) to be rank `rank`'s block of ladders of
        n_rungs * world rungs (`bae_shard_config`); create every rank's sampler with the SAME chain_id_base."""
        N.check(self.lib.bae_shard_config(self._h, int(rank), int(world)))
        self.shard_rank, self.shard_world = int(rank), int(world)

    def swap_exchange(self, nccl_comm):
        """One exchange round across the ranks (`bae_swap_exchange`); `nccl_comm`: distributed.NcclComm or a raw ncclComm_t."""
        N.check(self.lib.bae_swap_exchange(self._h, getattr(nccl_comm, "handle", nccl_comm)))

    def vectors_sent(self):
        return int(self.lib.bae_exchange_vectors_sent(self._h))

    def last_sweep(self):
        """src[p] = chain whose configuration ended at position p in the most recent exchange sweep."""
        src = np.zeros(self.n_chains, np.int32)
        N.check(self.lib.bae_last_sweep(self._h, src.ctypes.data))
        return src

    # -- probes -------------------------------------------------------------------------------
    def eval(self, chain, which=0, w=None, grads=False, out=False):
        sse, mse = C.c_double(), C.c_double()
        pw = None
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float32)
            pw = w.ctypes.data
        g = np.zeros(self.w_size, np.float32) if grads else None
        rows = self.cfg.n_train if which == 0 else self.cfg.n_test
        o = np.zeros((rows, self.dims4[0]), np.float32) if out else None
        N.check(self.lib.bae_eval(self._h, int(chain), int(which), pw, C.byref(sse), C.byref(mse),
                                  g.ctypes.data if grads else None, o.ctypes.data if out else None))
        return dict(sse=sse.value, mse=mse.value, grads=g, out=o)

    def eval_bn_stats(self, which=0):
        """Batch mean / biased variance of the BatchNorm layer in the most recent `eval` call."""
        mu, var = np.zeros(self.dims4[3], np.float32), np.zeros(self.dims4[3], np.float32)
        N.check(self.lib.bae_eval_bn_stats(self._h, int(which), mu.ctypes.data, var.ctypes.data))
        return mu, var

    def debug_draws(self, chain, step):
        noise = np.zeros(self.w_size, np.float32)
        sc = np.zeros(3, np.float64)
        N.check(self.lib.bae_debug_draws(self._h, int(chain), int(step), noise.ctypes.data, sc.ctypes.data))
        return dict(noise=noise, lx=sc[0], eta_noise=sc[1], u=sc[2])

    def debug_buffer(self, name, chain, rows, lo_part=False):
        """Raw [rows, ld] view of a per-chain device array (debug probe)."""
        ld = C.c_int()
        probe = np.zeros(4, np.float32)
        N.check(self.lib.bae_debug_buffer(self._h, name.encode(), int(chain), probe.ctypes.data, 4, C.byref(ld), int(lo_part)))
        out = np.zeros((rows, ld.value), np.float32)
        N.check(self.lib.bae_debug_buffer(self._h, name.encode(), int(chain), out.ctypes.data, out.size, C.byref(ld), int(lo_part)))
        return out

    def debug_swap_uniform(self, ensemble, rnd, pair):
        u = C.c_double()
        N.check(self.lib.bae_debug_swap_uniform(self._h, int(ensemble), int(rnd), int(pair), C.byref(u)))
        return u.value
