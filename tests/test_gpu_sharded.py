"""Ladder sharded over GPUs (`bae_shard_config` + `bae_swap_exchange`, NCCL): a 16-rung ladder cut over 2 GPUs must give
traces IDENTICAL to the same ladders on one GPU (same seed -> same Philox draws, same sweep, same data movement).
Needs >= 2 visible GPUs (`gpurun --gpus 2`); the log of the last run is kept under profiles/."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
        from bayesianautoencoders_b200.distributed import NcclComm
        from oracle import bae_oracle as bo
        from tests.parity_util import synth_data
        dims4, ntr, nte, gp = {"madelon": ((500, 450, 400, 300), 2000, 1800, 2), "mid": ((24, 20, 16, 12), 96, 40, 0)}[shape]
        train, test = synth_data(dims4, ntr, nte)
        E, L, S, si, seed = 2, 8, 10, 5, 77
        R = L * world
        T = bo.temperature_ladder(R, 2)
        Lay = bo.Layout(dims4)
        rng = np.random.default_rng(5)
        th0 = [[default_theta_init(dims4, rng) for _ in range(R)] for _ in range(E)]
        w0 = [[rng.standard_normal(Lay.w_size) for _ in range(R)] for _ in range(E)]
        # this rank's block of every ladder
        s = DeviceSampler(dims4, train, test, E * L, L, S, np.tile(T[rank * L:(rank + 1) * L], E), swap_interval=si, seed=seed,
                          device=rank, chain_id_base=0, gemm_path=gp)
        s.shard(rank, world)
        for e in range(E):
            for l in range(L):
                s.init_chain(e * L + l, th0[e][rank * L + l], w0[e][rank * L + l])
        comm = NcclComm(dist, rank, world, rank)
        for _ in range(S // si):
            s.step(si)
            s.swap_exchange(comm)
        s.synchronize()
        tr = [s.traces(c, 0, S) for c in range(E * L)]
        fin = [s.get_state(c) for c in range(E * L)]
        counts, sent = s.swap_counts(), s.vectors_sent()
        s.close()
        comm.close()
        # the same ladders whole on this GPU
        f = DeviceSampler(dims4, train, test, E * R, R, S, np.tile(T, E), swap_interval=si, seed=seed, device=rank, chain_id_base=0,
                          gemm_path=gp)
        for e in range(E):
            for r in range(R):
                f.init_chain(e * R + r, th0[e][r], w0[e][r])
        f.run(S)
        f.synchronize()
        ok_bits, worst = True, 0.0
        for e in range(E):
            for l in range(L):
                a, b = tr[e * L + l], f.traces(e * R + rank * L + l, 0, S)
                ok_bits &= a.tobytes() == b.tobytes()
                worst = max(worst, float(np.max(np.abs(a["likelihood_proposal"] - b["likelihood_proposal"]) / np.abs(b["likelihood_proposal"]))))
                assert np.array_equal(a["accepted"], b["accepted"]) and np.array_equal(a["langevin"], b["langevin"])
                fb = f.get_state(e * R + rank * L + l)
                fa = fin[e * L + l]
                assert fa["eta"] == fb["eta"] and fa["w_is_alias"] == fb["w_is_alias"]
                if ok_bits:
                    assert np.array_equal(fa["w"], fb["w"]) and np.array_equal(fa["m"], fb["m"])
        fc = f.swap_counts()
        f.close()
        q.put((rank, ok_bits, worst, counts, fc, sent, s.gemm_path() if False else gp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", ["madelon", "mid"])
def test_sharded_ladder_is_trace_identical_to_one_gpu(shape):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sent_total = 0
    for rank, ok_bits, worst, counts, fc, sent, gp in sorted(res):
        assert counts == fc, (counts, fc)                     # every rank counts the whole sweep
        if shape == "madelon":
            assert ok_bits, f"rank {rank}: traces of the sharded run differ from the one-GPU run (worst rel {worst:.2e})"
        else:
            # FP32 tiles: the K split of the weight-gradient GEMMs depends on the number of resident chains (summation
            # order), so the two runs agree to rounding, not to the bit; decisions, eta and swap counts are identical
            assert worst < 1e-5
        sent_total += sent
        print(f"[sharded {shape}] rank {rank}: bit-identical={ok_bits} worst rel {worst:.1e} swaps {counts} vectors sent {sent}")
    assert sent_total > 0 or shape != "madelon"
