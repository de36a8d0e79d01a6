"""World-size-2 gloo test of the ladder-sharded exchange (host logic of bayesianautoencoders_b200.distributed):
both ranks must derive the same permutation from the all-gathered scalars, configurations must end where the
sequential sweep says, and the sweep itself must equal the oracle's `swap_procedure` chain of decisions."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bae_oracle as bo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeTransport:
    """numpy stand-in for the device: per local chain a weight vector tagged with its origin."""

    def __init__(self, rank, E, L, R, w_size, table):
        self.rank, self.E, self.L, self.R = rank, E, L, R
        self.w = {}
        self.sc = {}
        for e in range(E):
            for l in range(L):
                rung = rank * L + l
                c = e * L + l
                self.w[c] = torch.full((w_size,), float(e * 100 + rung))
                self.sc[c] = list(table[e, rung])
        self.finished = 0

    def collective_device(self):
        return torch.device("cpu")

    def scalars(self, c):
        return tuple(self.sc[c])

    def pack(self, c):
        return self.w[c]

    def unpack(self, c, tensor, eta, last_sse):
        self.w[c] = tensor.clone()
        self.sc[c][0] = eta
        self.sc[c][3] = last_sse

    def finish(self):
        self.finished += 1


def _make_table(E, R, seed=5):
    rng = np.random.default_rng(seed)
    tab = np.zeros((E, R, 4))
    tab[:, :, 0] = rng.normal(-1.0, 0.3, (E, R))            # eta
    tab[:, :, 3] = rng.uniform(50, 80, (E, R))               # SSE of w
    T = 2.0 ** (np.arange(R) / (R - 1))
    tab[:, :, 2] = T
    n = 400
    for e in range(E):
        for r in range(R):
            tab[e, r, 1] = (-(n / 2) * np.log(2 * np.pi * np.exp(tab[e, r, 0])) - 0.5 * np.exp(tab[e, r, 0]) * tab[e, r, 3]) / T[r]
    return tab, n


def _worker(rank, world, port, E, R, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayesianautoencoders_b200.distributed import ShardedLadder
    tab, n = _make_table(E, R)
    L = R // world
    t = FakeTransport(rank, E, L, R, 16, tab)
    sl = ShardedLadder(dist, rank, world, E, R, n, seed=0xBAE5EED, transport=t)
    perms = [sl.exchange() for _ in range(3)]
    tags = {c: float(t.w[c][0]) for c in t.w}
    q.put((rank, perms, tags, {c: t.sc[c] for c in t.sc}, sl.num_swap, sl.total_swap_proposals, t.finished))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_exchange_world2():
    E, R, world = 2, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, E, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)])
    for p in procs:
        p.join(timeout=30)
    (r0, perms0, tags0, sc0, ns0, nt0, f0), (r1, perms1, tags1, sc1, ns1, nt1, f1) = res
    assert perms0 == perms1 and ns0 == ns1 and nt0 == nt1 == 3 * E * (R - 1) and f0 == f1 == 3
    # replay on one process: positions hold configuration tags
    from bayesianautoencoders_b200.distributed import sweep, swap_uniform
    tab, n = _make_table(E, R)
    where = {(e, r): float(e * 100 + r) for e in range(E) for r in range(R)}
    for rnd in range(3):
        for e in range(E):
            u = [swap_uniform(0xBAE5EED, e, rnd, p) for p in range(R - 1)]
            src, _ = sweep(tab[e], n, u)
            assert src == perms0[rnd][e]
            assert sorted(src) == list(range(R))              # a permutation
            old = {r: where[(e, r)] for r in range(R)}
            old_sc = tab[e].copy()
            for p, s in enumerate(src):
                where[(e, p)] = old[s]
                tab[e, p, 0] = old_sc[s, 0]                   # eta and SSE travel with the configuration
                tab[e, p, 3] = old_sc[s, 3]
    L = R // world
    for rank, tags in ((0, tags0), (1, tags1)):
        for e in range(E):
            for l in range(L):
                assert tags[e * L + l] == where[(e, rank * L + l)]
    assert ns0 > 0    # the synthetic table produces at least one accepted exchange


def test_host_sweep_equals_oracle_swap_procedure_chain():
    """`distributed.sweep` against the oracle's sequential application of `swap_procedure` (MAIN:962-1011)
    on a tiny model where the fp32 forwards of the oracle are cheap."""
    from bayesianautoencoders_b200.distributed import sweep
    dims4 = (4, 3, 3, 2)
    L = bo.Layout(dims4)
    rng = np.random.default_rng(9)
    X = rng.random((40, 4), dtype=np.float32)
    R = 5
    T = bo.temperature_ladder(R, 2)
    params, table = [], np.zeros((R, 4))
    for r in range(R):
        w = (rng.standard_normal(L.w_size) * 0.5).astype(np.float32)
        w[L.nbt_index] = 2
        eta = rng.normal(-1, 0.2)
        out, _ = bo.forward(L, w.astype(np.float64), X.astype(np.float64), update_buffers=False)
        sse = float(np.sum((out - X) ** 2))
        lik = bo.loglik_from_sse(sse, X.size, np.exp(eta), T[r]) + rng.normal(0, 0.3)    # stored value may be stale
        params.append(np.concatenate([w.astype(np.float64), [eta, lik, T[r], 4]]))
        table[r] = (eta, lik, T[r], sse)
    u = list(rng.random(R - 1))
    src, nsw = sweep(table, X.size, u)
    # oracle: pass the messages through the sequential sweep; identify configurations by their first weight
    ident = [p[0] for p in params]
    msgs = [p.copy() for p in params]
    acc = 0
    for i in range(R - 1):
        p1, p2, swapped, _, _ = bo.swap_procedure(dims4, X, msgs[i], msgs[i + 1], u[i])
        msgs[i], msgs[i + 1] = p1, p2
        acc += swapped
    assert acc == nsw
    assert [ident.index(m[0]) for m in msgs] == src
