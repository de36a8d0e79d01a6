"""Drive the UNMODIFIED reference with injected randomness (TEST INFRASTRUCTURE ONLY; build container only).

Runs `ptReplica.run()` in-process (never `.start()`), one Python thread per replica, with
`queue.Queue` / `threading.Event` stand-ins for the multiprocessing primitives (MAIN:849-852) and
the real `ParallelTempering.swap_procedure` (MAIN:962-1011) executed by the calling thread exactly
like `run_chains` does (MAIN:1035-1072). All random draws the loop consumes are replaced by
caller-supplied values:

  np.random.uniform  MAIN:447 (lx), MAIN:994 (swap u)      np.random.normal  MAIN:500 (eta noise)
  np.random.randn    MAIN:389 (start vector)               random.uniform    MAIN:526 (MH u)
  torch.Tensor.normal_  MAIN:260 (weight noise; routed per tensor NAME because addnoiseandcopy
                                  iterates in the incoming dict's key order)

The reference method bodies are not altered; thin wrappers only record arguments/results.
"""
from __future__ import annotations

import queue
import threading
import time
from unittest import mock

import numpy as np
import torch

from . import ref_import
from .bae_oracle import Layout, SORTED_KEYS


class _Tls(threading.local):
    ctx = None


_tls = _Tls()


class ReplicaCtx:
    """Per-replica injected randomness and captured internals."""

    def __init__(self, layout: Layout, w0_randn, lx, noise, eta_noise, u):
        self.L = layout
        self.w0_randn = np.asarray(w0_randn, dtype=np.float64)
        self.lx = list(lx)
        self.noise = np.asarray(noise, dtype=np.float32)      # [steps, w_size], sorted-key flat order
        self.eta_noise = list(eta_noise)
        self.u = list(u)
        self.step = -1            # advanced when lx is drawn (first draw of an iteration)
        self.pending_keys = []
        self.grads = []           # flat (sorted-key) gradient after each langevin_gradient call
        self.lik_calls = []       # (n_rows, tau_sq, temp, loss_over_temp, mse)
        self.saved = {}           # np.savetxt captures, full precision
        self.theta_init = None


def _flat_grad(ref_model, L: Layout):
    g = np.zeros(L.w_size)
    named = dict(ref_model.named_parameters())
    for k in SORTED_KEYS:
        if k in named and named[k].grad is not None:
            g[L.slice(k)] = named[k].grad.detach().numpy().reshape(-1)
    return g


def _install_patches(ref):
    """Context manager stack replacing the global RNG entry points with per-thread injected streams."""
    orig_uniform = np.random.uniform
    orig_normal = np.random.normal
    orig_randn = np.random.randn
    orig_tnormal = torch.Tensor.normal_
    orig_addnoise = ref.Model.addnoiseandcopy
    orig_langevin = ref.Model.langevin_gradient
    orig_lik = ref.ptReplica.likelihood_func
    orig_savetxt = np.savetxt

    def uniform(*a, **k):
        c = _tls.ctx
        if c is None:
            return orig_uniform(*a, **k)
        if len(a) == 3:                 # np.random.uniform(0, 1, 1)  MAIN:447
            c.step += 1
            return np.array([c.lx[c.step]])
        return c.swap_u.pop(0)          # np.random.uniform(0, 1)    MAIN:994 (main thread ctx)

    def normal(*a, **k):
        c = _tls.ctx
        if c is None:
            return orig_normal(*a, **k)
        return np.array([c.eta_noise[c.step]])

    def randn(*a, **k):
        c = _tls.ctx
        if c is None:
            return orig_randn(*a, **k)
        return c.w0_randn.copy()

    def py_uniform(a, b):
        c = _tls.ctx
        return c.u[c.step]

    def tnormal(self, mean=0, std=1, *, generator=None):
        c = _tls.ctx
        if c is None or not c.pending_keys:
            return orig_tnormal(self, mean, std, generator=generator)
        key = c.pending_keys.pop(0)
        vals = np.ascontiguousarray(c.noise[c.step][c.L.slice(key)])
        self.copy_(torch.from_numpy(vals).reshape(self.shape))
        return self

    def addnoise(self, w, mea, std_dev):
        c = _tls.ctx
        if c is not None:
            c.pending_keys = list(w.keys())
        return orig_addnoise(self, w, mea, std_dev)

    def langevin(self, x, w, optimizer_mode=1):
        out = orig_langevin(self, x, w, optimizer_mode)
        c = _tls.ctx
        if c is not None:
            c.grads.append(_flat_grad(self, c.L))
        return out

    def lik(cae, data, w, tau_sq, temp):
        res = orig_lik(cae, data, w, tau_sq, temp)
        c = _tls.ctx
        if c is not None and hasattr(c, "lik_calls"):
            c.lik_calls.append((data.shape[0], float(tau_sq), float(temp), float(res[0]), float(res[2])))
        return res

    def savetxt(fname, arr, *a, **k):
        c = _tls.ctx
        if c is not None and hasattr(c, "saved"):
            import os
            c.saved[os.path.basename(str(fname))] = np.array(arr, dtype=np.float64)
            return None
        return orig_savetxt(fname, arr, *a, **k)

    return [
        mock.patch.object(np.random, "uniform", uniform),
        mock.patch.object(np.random, "normal", normal),
        mock.patch.object(np.random, "randn", randn),
        mock.patch.object(ref.random, "uniform", py_uniform),
        mock.patch.object(torch.Tensor, "normal_", tnormal),
        mock.patch.object(ref.Model, "addnoiseandcopy", addnoise),
        mock.patch.object(ref.Model, "langevin_gradient", langevin),
        mock.patch.object(ref.ptReplica, "likelihood_func", staticmethod(lik)),
        mock.patch.object(np, "savetxt", savetxt),
    ]


class _MainCtx:
    def __init__(self, swap_u):
        self.swap_u = list(swap_u)


def run_reference_pt(dims4, train, test, temperatures, samples_per_chain, swap_interval, rnd,
                     swap_u=None, do_swaps=True, lrate=0.01, step_size=0.005, torch_seed=0, experiment_type=2):
    """Run `len(temperatures)` reference replicas for `samples_per_chain` iterations each.

    rnd: list (one per replica) of dict(w0_randn, lx, noise, eta_noise, u).
    swap_u: the uniforms `swap_procedure` consumes, in sweep order (round-major, pair-minor).
    Returns dict with per-replica captures and swap log.
    """
    ref = ref_import.load_reference()
    ref_import.set_shape(ref, dims4, lrate=lrate, step_size=step_size)
    ref.experiment_type = int(experiment_type)     # module global read inside ptReplica.run (MAIN:72, 454)
    L = Layout(dims4)
    n = len(temperatures)
    train = np.asarray(train, dtype=np.float64)
    test = np.asarray(test, dtype=np.float64)
    torch.manual_seed(torch_seed)

    class _Q:  # multiprocessing.Queue stand-in
        def __init__(self):
            self.q = queue.Queue()

        def put(self, x):
            self.q.put(x)

        def get(self):
            return self.q.get()

    queues = [_Q() for _ in range(n)]
    wait_chain = [threading.Event() for _ in range(n)]
    events = [threading.Event() for _ in range(n)]
    ctxs, reps = [], []
    path = ref._scratch
    for j in range(n):
        c = ReplicaCtx(L, **rnd[j])
        rep = ref.ptReplica(True, lrate, None, 0.0, 0.0, samples_per_chain, train, test, 0.5,
                            float(temperatures[j]), swap_interval, path, queues[j], wait_chain[j], events[j],
                            10, step_size)
        c.theta_init = np.asarray(rep.cae.getparameters(rep.cae.state_dict()), dtype=np.float64).copy()
        ctxs.append(c)
        reps.append(rep)

    # the main-process fp32 model used by swap_procedure (MAIN:830)
    pt = ref.ParallelTempering.__new__(ref.ParallelTempering)
    pt.cae = ref.Model()
    pt.traindata = train
    pt.num_param = L.w_size
    pt.num_swap = 0
    pt.total_swap_proposals = 0
    main_ctx = _MainCtx(swap_u or [])
    swap_log = []
    errors = []

    def worker(j):
        _tls.ctx = ctxs[j]
        try:
            reps[j].run()
        except Exception as e:  # the post-loop sklearn tail may raise after all traces are captured
            ctxs[j].tail_error = repr(e)
            if not any(k.startswith("mse_train_chain_") for k in ctxs[j].saved):
                errors.append((j, e))
        finally:
            wait_chain[j].set()
            _tls.ctx = None

    patches = _install_patches(ref)
    for p in patches:
        p.start()
    try:
        threads = [threading.Thread(target=worker, args=(j,), daemon=True) for j in range(n)]
        for j in range(n):
            wait_chain[j].clear()
            events[j].clear()
        for t in threads:
            t.start()
        n_rounds = samples_per_chain // swap_interval
        _tls.ctx = main_ctx
        for r in range(n_rounds):                      # MAIN:1035-1072
            for j in range(n):
                wait_chain[j].wait(timeout=600)
            if errors:
                break
            time.sleep(0.02)   # let every replica reach event.wait() (the reference has the same latent race, MAIN:598-600)
            if do_swaps:
                for index in range(n - 1):
                    p1, p2, swapped = pt.swap_procedure(queues[index], queues[index + 1])
                    swap_log.append((r, index, bool(swapped)))
                    queues[index].put(p1)
                    queues[index + 1].put(p2)
            for j in range(n):
                wait_chain[j].clear()
                events[j].set()
        _tls.ctx = None
        for t in threads:
            t.join(timeout=600)
    finally:
        for p in patches:
            p.stop()
        _tls.ctx = None
        ref.experiment_type = 2
    if errors:
        raise errors[0][1]
    return dict(ctxs=ctxs, swap_log=swap_log, num_swap=pt.num_swap, total_swap_proposals=pt.total_swap_proposals,
                layout=L)
