#!/usr/bin/env python3
"""Ladder-sharded parallel tempering across GPUs (one process per GPU, NCCL): example / test driver.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_pt.py --shape madelon --ensembles 2 --rungs 8 --rounds 3

Every rank owns `rungs / world` consecutive temperatures of every ensemble; only the <= 32 B/chain scalar
all-gather and the weight vectors of configurations that cross a GPU boundary travel over NVLink.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianautoencoders_b200 import DeviceSampler, default_theta_init  # noqa: E402
from bayesianautoencoders_b200.distributed import DeviceTransport, ShardedLadder  # noqa: E402
from bench import SHAPES, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="madelon")
ap.add_argument("--ensembles", type=int, default=2)
ap.add_argument("--rungs", type=int, default=8)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--swap-interval", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dims4, ntr, nte = SHAPES[a.shape]
train, test = synth(dims4, ntr, nte)
L = a.rungs // world
T_all = 2.0 ** (np.arange(a.rungs) / (a.rungs - 1))
T_loc = np.tile(T_all[rank * L:(rank + 1) * L], a.ensembles)
# n_rungs = 1 on the handle: exchanges are driven from the host across ranks, not by the local swap kernel
s = DeviceSampler(dims4, train, test, a.ensembles * L, 1, a.rounds * a.swap_interval, T_loc, swap_interval=a.swap_interval,
                  device=local, chain_id_base=rank * a.ensembles * L)
rng = np.random.default_rng(1000 + rank)
for j in range(a.ensembles * L):
    s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(s.w_size))
sl = ShardedLadder(dist, rank, world, a.ensembles, a.rungs, train.size, 0xBAE5EED, DeviceTransport(s, local))
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
t_step = t_ex = 0.0
for r in range(a.rounds):
    t1 = time.perf_counter()
    s.step(a.swap_interval)
    s.synchronize()
    t2 = time.perf_counter()
    perms = sl.exchange()
    t3 = time.perf_counter()
    t_step += t2 - t1
    t_ex += t3 - t2
dist.barrier(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
tr = s.traces(0, 0, a.rounds * a.swap_interval)
ok = bool(np.all(np.isfinite(tr["sum_value"])))
sent = torch.tensor([sl.vectors_sent], device="cuda")
dist.all_reduce(sent)
if rank == 0:
    print(json.dumps({"world": world, "ensembles": a.ensembles, "rungs": a.rungs, "rounds": a.rounds,
                      "chain_steps_per_s": a.ensembles * a.rungs * a.rounds * a.swap_interval / wall,
                      "step_s": t_step, "exchange_s": t_ex, "num_swap": sl.num_swap,
                      "total_swap_proposals": sl.total_swap_proposals, "weight_vectors_over_nvlink": int(sent.item()),
                      "last_perm": perms, "finite": ok}))
s.close()
dist.destroy_process_group()
