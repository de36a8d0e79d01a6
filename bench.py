#!/usr/bin/env python3
"""Benchmark of the PT-Langevin hot path (BASELINE.json: chain-steps/sec at Madelon shape).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

One bench "step" = one parallel-tempering round: `swap_interval` (5) iterations of MAIN:432-592 for
every resident chain, then one exchange sweep (MAIN:1059-1065). Throughput = chain iterations / s
over all GPUs. Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {
    # name: dims4, n_train, n_test                                   SURVEY.md section 8
    "madelon": ((500, 450, 400, 300), 2000, 1800),
    "coil": ((85, 70, 60, 50), 4366, 1456),
    "swiss": ((3, 10, 5, 2), 3750, 1250),
    # BASELINE.json configs[4]: synthetic 2048-feature, 100k-row autoencoder; hidden widths = the Madelon ratios in multiples of
    # 256; BASELINE names no test set: a token 256-row test matrix keeps the step program unchanged
    "stress2048": ((2048, 1792, 1536, 1280), 100000, 256),
}
METRIC = "pt_langevin_chain_steps_per_sec"
UNIT = "chain-steps/s"
SWAP_INTERVAL = 5
N_RUNGS = 8
SAMPLES = 6000   # 48000 samples / 8 replicas (MAIN:846, 1333)


def flops_per_step(dims4, n_train, n_test):
    """Algorithmic FLOPs of one chain iteration: Langevin (canonical, de-duplicated) and random walk."""
    d0, d1, d2, d3 = dims4
    S = d0 * d1 + d1 * d2 + d2 * d3 + d3 * d2 + d2 * d1 + d1 * d0
    f_l = 2 * n_train * (6 * S - 2 * d0 * d1) + n_test * 2 * S
    f_r = (n_train + n_test) * 2 * S
    return f_l, f_r


def synth(dims4, n_train, n_test):
    train = np.random.default_rng(1234).random((n_train, dims4[0]), dtype=np.float32)
    test = np.random.default_rng(1235).random((n_test, dims4[0]), dtype=np.float32)
    return train, test


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), bf16=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def tf32_peak():
    """kind::tf32 dense peak measured on a B200 of this pool (tools/measure_tf32_peak.py -> profiles/r2_tf32_peak.json):
    cuBLAS TF32 8192^3 sustained / burst, and the raw tcgen05.mma.cta_group::2.kind::tf32 issue rate of all SM pairs."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_tf32_peak.json")) as f:
            d = json.load(f)
        issue = d.get("tcgen05_issue_rate", [{}])[0]
        return dict(sustained=d["cublas_tf32_8192"]["sustained_tflops"], burst=d["cublas_tf32_8192"]["burst_tflops"],
                    sustained_sm_mhz=d["cublas_tf32_8192"]["clocks_under_load"].get("sm_mhz"),
                    issue_rate=issue.get("tflops"), issue_rate_sm_mhz=issue.get("sm_mhz"))
    except Exception:
        return None


TRAFFIC_CAPTURE = "r2_v8_traffic.json"   # written by tools/make_profiles.py from the ncu --set full export of the current build


def ncu_traffic(shape, cpg, gemm_path):
    """DRAM bytes (read + write) of one launch of the step program from the committed `ncu --set full` capture
    (profiles/<TRAFFIC_CAPTURE>), scaled to this run's chains x iterations per launch; None if no capture fits."""
    p = os.path.join(ROOT, "profiles", TRAFFIC_CAPTURE)
    try:
        with open(p) as f:
            d = json.load(f)
        if d["shape"] != shape or d["gemm_path"] != gemm_path:
            return None
        per_chain_iter = (d["dram_bytes_read"] + d["dram_bytes_write"]) / (d["chains"] * d["iterations"])
        return {"bytes_per_launch": per_chain_iter * cpg * SWAP_INTERVAL, "bytes_per_chain_iteration": per_chain_iter,
                "source": d["source"]}
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    shape, n_iter, threads, seed = args
    from threadpoolctl import threadpool_limits
    from oracle import bae_oracle as bo
    from bayesianautoencoders_b200.sampler import default_theta_init
    dims4, n_train, n_test = SHAPES[shape]
    train, test = synth(dims4, n_train, n_test)
    L = bo.Layout(dims4)
    rng = np.random.default_rng(seed)
    with threadpool_limits(limits=threads):
        rep = bo.Replica(dims4, train, test, 1.0 + 0.1 * seed, SAMPLES)
        rep.initialise(default_theta_init(dims4, rng).astype(np.float64), rng.standard_normal(L.w_size))
        draws = [bo.StepRandomness(rng.random(), (rng.standard_normal(L.w_size) * 0.005).astype(np.float32),
                                   rng.standard_normal() * 0.2, rng.random()) for _ in range(n_iter)]
        t0 = time.perf_counter()
        for d in draws:
            rep.step(d)
        return time.perf_counter() - t0


def cpu_chain_steps_per_sec(shape, n_proc, iters_per_proc):
    """`n_proc` replica processes (the reference runs one OS process per replica, MAIN:1028-1031), each with
    cores//n_proc BLAS threads (timing (ii)); float64 as the reference runs."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    threads = max(1, cores // n_proc)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(n_proc) as pool:
        times = pool.map(_cpu_worker, [(shape, iters_per_proc, threads, s) for s in range(n_proc)])
    wall = time.perf_counter() - t0
    loop = max(times)
    return n_proc * iters_per_proc / loop, dict(cores=min(cores, threads * n_proc), procs=n_proc, threads_per_proc=threads,
                                                loop_s=loop, wall_s=wall)


def reference_real_chain_steps_per_sec(shape, iters_per_proc):
    """The UNMODIFIED reference (`/root/reference`, build container only): its own 8-process `ParallelTempering.run_chains()`
    (MAIN:1013-1090) on the synthetic matrices, `cores // 8` torch threads per replica process (timing
    (ii)). The sampling loop ends at the first `np.savetxt` of a replica (MAIN:632); the classification / plot tail is not timed.
    Returns None when the reference tree is not mounted (the GPU boxes): the caller falls back to the oracle port."""
    from oracle import ref_import
    if not ref_import.available():
        return None
    import glob
    import shutil
    import torch
    ref = ref_import.load_reference()
    dims4, n_train, n_test = SHAPES[shape]
    ref_import.set_shape(ref, dims4)
    train, test = synth(dims4, n_train, n_test)
    train, test = train.astype(np.float64), test.astype(np.float64)
    ref.data_load = lambda data='train': (train if data == 'train' else test)
    cores = os.cpu_count() or 1
    threads = max(1, cores // N_RUNGS)
    torch.set_num_threads(threads)
    stamp_dir = tempfile.mkdtemp(prefix="bae_ref_stamp_")
    orig_savetxt = np.savetxt

    def stamped_savetxt(*a, **k):   # inherited by the forked replica processes
        fn = os.path.join(stamp_dir, f"{os.getpid()}.t")
        if not os.path.exists(fn):
            with open(fn, "w") as f:
                f.write(repr(time.time()))
        return orig_savetxt(*a, **k)

    path = os.path.join(ref._scratch, "bench_res")
    os.makedirs(path, exist_ok=True)
    old_cwd = os.getcwd()
    os.chdir(ref._scratch)
    np.savetxt = stamped_savetxt
    try:
        pt = ref.ParallelTempering(True, N_RUNGS, 2, N_RUNGS * iters_per_proc, SWAP_INTERVAL, path, 10, 0.5, 0.005)
        pt.initialize_chains(0.5)
        t0 = time.time()
        try:
            pt.run_chains()
        except Exception:
            pass    # show_results() reads plots / files the stubs do not produce; the loop has been timed by then
    finally:
        np.savetxt = orig_savetxt
        os.chdir(old_cwd)
    stamps = [float(open(f).read()) for f in glob.glob(os.path.join(stamp_dir, "*.t"))]
    shutil.rmtree(stamp_dir, ignore_errors=True)
    if len(stamps) < N_RUNGS:
        return None
    loop = max(stamps) - t0
    return N_RUNGS * iters_per_proc / loop, dict(cores=min(cores, threads * N_RUNGS), procs=N_RUNGS, threads_per_proc=threads,
                                                  loop_s=loop, wall_s=time.time() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shape = args.shape
    n_proc = N_RUNGS
    iters = max(1, args.ref_iters)
    vals = []
    info = {}
    kind = "port"
    real = None
    try:
        real = reference_real_chain_steps_per_sec(shape, iters)
    except Exception as e:   # the reference cannot run here: fall back to the port and say so
        print(f"[bench] unmodified reference not runnable here ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    t_total = 0.0
    if real is not None:
        kind = "reference"
        for k in range(args.steps):
            v, info = real if k == 0 else reference_real_chain_steps_per_sec(shape, iters)
            vals.append(v)
            t_total += info["loop_s"]
    else:
        for _ in range(args.warmup if args.warmup < 2 else 1):
            cpu_chain_steps_per_sec(shape, n_proc, 1)
        for _ in range(args.steps):
            v, info = cpu_chain_steps_per_sec(shape, n_proc, iters)
            vals.append(v)
            t_total += info["loop_s"]
    value = float(np.mean(vals))
    dims4, n_train, n_test = SHAPES[shape]
    sample = (f"{n_proc} replica processes x {iters} iterations per bench step, " +
              ("the UNMODIFIED reference's ParallelTempering.run_chains() (torch float64)" if kind == "reference" else
               "float64 numpy oracle port (OpenBLAS)") + f", {info['threads_per_proc']} threads each")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(shape, args.chains_per_gpu), "shape": shape, "dims": list(dims4),
                   "n_train": n_train, "n_test": n_test},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_name(shape, cpg, world=1, sharded=False):
    if sharded:
        return (f"{shape} shape, {cpg * world} chains = {cpg // N_RUNGS} ensembles x {N_RUNGS * world}-rung geometric ladder (Tmax 2) "
                f"sharded over {world} GPUs ({cpg} chains per GPU), {SWAP_INTERVAL} iterations + 1 exchange sweep per bench step")
    return (f"{shape} shape, {cpg} chains per GPU = {cpg // N_RUNGS} ensembles x {N_RUNGS}-rung geometric ladder (Tmax 2), "
            f"{SWAP_INTERVAL} iterations + 1 exchange sweep per bench step")


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
class PtJob:
    """`n_ens` ensembles of `rungs_total` rungs on `world` GPUs: whole ladders per GPU (world == 1 or ladder == 'local'), or
    every ladder cut into contiguous blocks of rungs_total / world rungs per GPU (the exchange crosses NVLink via NCCL)."""

    def __init__(self, shape, n_ens, rungs_total, world, rank, local, sharded, comm, gemm_path, samples=SAMPLES, seed=0xBAE5EED):
        from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
        from oracle import bae_oracle as bo   # ladder helper + layout size only
        self.shape, self.world, self.rank, self.sharded, self.comm = shape, world, rank, sharded, comm
        dims4, n_train, n_test = SHAPES[shape]
        self.dims4, self.n_train, self.n_test = dims4, n_train, n_test
        self.train, self.test = synth(dims4, n_train, n_test)
        L = bo.Layout(dims4)
        if sharded:
            self.L = rungs_total // world
            t_all = bo.temperature_ladder(rungs_total, 2)
            temps = np.tile(t_all[rank * self.L:(rank + 1) * self.L], n_ens)
            self.s = DeviceSampler(dims4, self.train, self.test, n_ens * self.L, self.L, samples, temps, swap_interval=SWAP_INTERVAL,
                                   seed=seed, device=local, chain_id_base=0, gemm_path=gemm_path)
            self.s.shard(rank, world)
        else:
            self.L = rungs_total
            temps = np.tile(bo.temperature_ladder(rungs_total, 2), n_ens)
            self.s = DeviceSampler(dims4, self.train, self.test, n_ens * rungs_total, rungs_total, samples, temps,
                                   swap_interval=SWAP_INTERVAL, seed=seed, device=local, chain_id_base=rank * n_ens * rungs_total,
                                   gemm_path=gemm_path)
        self.chains = n_ens * self.L
        rng = np.random.default_rng(100 + rank)
        for j in range(self.chains):
            self.s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(L.w_size))
        self.s.synchronize()
        self.iters_done = 0

    def pt_round(self):
        """One bench step: 5 iterations of every resident chain + one exchange sweep; stream-ordered, returns at once."""
        if self.sharded:
            self.s.step(SWAP_INTERVAL)
            self.s.swap_exchange(self.comm)     # all-gather of 40 B/chain + device sweep + boundary-crossing weight vectors
        else:
            self.s.run(SWAP_INTERVAL)           # exchange phases are part of the step program
        self.iters_done += SWAP_INTERVAL

    def close(self):
        self.s.close()


def timed_rounds(job, steps, warmup, barrier, dist, world):
    """W untimed + K timed rounds, CUDA events on the library's stream, max over ranks. Returns ms for the K rounds."""
    import torch
    stream = torch.cuda.ExternalStream(job.s.lib.bae_stream(job.s._h))
    for _ in range(warmup):
        job.pt_round()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        job.pt_round()
    ev1.record(stream)
    job.s.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def roofline_block(job, kernel_ms, n_launch_rounds, sharded):
    """Algorithmic FLOPs of the launches just timed (Langevin / random-walk iterations counted from the traces) over the
    measured launch duration, against the measured peaks."""
    f_l, f_r = flops_per_step(job.dims4, job.n_train, job.n_test)
    n_l = n_r = 0
    first = job.iters_done - SWAP_INTERVAL * n_launch_rounds
    for j in range(job.chains):
        tr = job.s.traces(j, first, SWAP_INTERVAL * n_launch_rounds)
        n_l += int(tr["langevin"].sum())
        n_r += int((1 - tr["langevin"]).sum())
    flops_per_launch = (n_l * f_l + n_r * f_r) / n_launch_rounds
    achieved_tf = flops_per_launch / (kernel_ms / 1000.0) / 1e12
    peaks = measured_peaks()
    gemm_path = job.s.gemm_path()
    if gemm_path == 2:
        f16 = job.s.mma_kind() == 2
        kind = ("tcgen05 kind::f16, 3-way split of scaled FP32 operands (3xFP16: hi*hi, lo*hi, hi*lo), FP32 accumulate in TMEM" if f16
                else "tcgen05 kind::tf32, 3-way split (3xTF32), FP32 accumulate in TMEM")
        tf = None if f16 else tf32_peak()
        out = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
               "frac": achieved_tf / peaks["bf16_sustained"],
               "traffic": ncu_traffic(job.shape, job.chains, gemm_path),
               "kernel": ("bae_phase_kernel_tc (tcgen05 GEMM phases) + bae_aux_kernel (BatchNorm / Adam): the step program of 5 "
                          "iterations of all resident chains" + ("" if sharded else " + exchange sweep") +
                          ", one CUDA-graph launch of its per-phase kernels"),
               "peak_source": peaks["source"] + " bf16 dense sustained (MEASURED_PEAKS.json)", "mma_kind": kind}
        if f16:
            # kind::f16 issues at the bf16 rate: the measured bf16 peak IS the peak of the MMA kind issued; every algorithmic FLOP
            # costs three of its FLOPs, so the attainable algorithmic fraction is 1/3
            out.update({"frac_of_3xf16_ceiling": 3.0 * achieved_tf / peaks["bf16_sustained"], "mma_tflops_issued": 3.0 * achieved_tf})
        if tf:
            # the MMA kind actually issued, measured on this pool (profiles/r2_tf32_peak.json): cuBLAS TF32 sustained; every
            # algorithmic FLOP costs three kind::tf32 MMA FLOPs (hi*hi, hi*lo, lo*hi), so the attainable algorithmic rate is 1/3
            out.update({"peak_tf32_measured": tf["sustained"], "peak_tf32_burst": tf["burst"],
                        "peak_tf32_sm_mhz": tf["sustained_sm_mhz"], "frac_of_tf32": achieved_tf / tf["sustained"],
                        "frac_of_3xtf32_ceiling": 3.0 * achieved_tf / tf["sustained"],
                        "tcgen05_issue_rate_tflops": tf["issue_rate"], "tcgen05_issue_rate_sm_mhz": tf["issue_rate_sm_mhz"],
                        "mma_tflops_issued": 3.0 * achieved_tf})
    else:
        kind = "fp32 CUDA-core tiles (FFMA)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                mhz = float(json.load(f).get("sm_max_mhz", 1965.0))
        except Exception:
            mhz = 1965.0
        peak = 148 * 128 * 2 * mhz * 1e6 / 1e12
        narrow = job.s.step_kernel() == 3
        out = {"bound": "fp32", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak, "traffic": None,
               "kernel": (("bae_narrow_kernel (one thread-block cluster per chain, rows in registers, DSMEM reductions; one launch = 5 "
                           "iterations of all resident chains) + the exchange sweep kernel") if narrow else
                          ("bae_program_kernel (persistent FP32 step kernel, one cooperative launch = 5 iterations of all resident "
                           "chains" + ("" if sharded else " + exchange sweep") + ")")),
               "peak_source": f"derived: 148 SMs x 128 FP32 lanes x 2 FLOP x {mhz:.0f} MHz (max SM clock, MEASURED_PEAKS.json); "
                              "narrow layers are latency / SFU bound, see DESIGN.md", "mma_kind": kind}
    out.update({"kernel_ms_per_launch": kernel_ms, "algorithmic_gflop_per_launch": flops_per_launch / 1e9})
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from bayesianautoencoders_b200.distributed import NcclComm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = NcclComm(dist, rank, world, local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shape = args.shape
    cpg = args.chains_per_gpu
    sharded = world > 1 and args.ladder == "sharded"
    n_ens = cpg // N_RUNGS
    # BASELINE.json configs[3]: every ensemble's ladder has 8 x world rungs (64 at 8 GPUs), cut into blocks of 8 rungs per GPU
    job = PtJob(shape, n_ens, N_RUNGS * (world if sharded else 1), world, rank, local, sharded, comm, args.gemm_path)
    s = job.s
    dims4, n_train, n_test = job.dims4, job.n_train, job.n_test
    if sharded:
        for _ in range(2):     # opens the NCCL point-to-point connections of the pairs that exchange (lazy in NCCL)
            job.pt_round()
    clocks = ClockSampler(local)
    l0 = None
    for _ in range(args.warmup):
        job.pt_round()
    barrier()
    if rank == 0:
        clocks.start()
    l0 = s.launch_count()
    ms = timed_rounds(job, args.steps, 0, barrier, dist, world)
    launches = s.launch_count() - l0
    clk = clocks.stop() if rank == 0 else {}
    value = world * cpg * args.steps * SWAP_INTERVAL / (ms / 1000.0)

    # per-launch duration of the step program (graph of per-phase kernels / persistent kernel) and of the exchange, CUDA events
    stream = torch.cuda.ExternalStream(s.lib.bae_stream(s._h))
    per_launch, exch_ms = [], []
    for _ in range(min(3, args.steps)):
        if sharded:
            s.step(SWAP_INTERVAL)
            step_ms = s.last_ms()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(stream)
            s.swap_exchange(comm)
            eb.record(stream)
            s.synchronize()
            exch_ms.append(ea.elapsed_time(eb))
            job.iters_done += SWAP_INTERVAL
        else:
            job.pt_round()
            step_ms = s.last_ms()
        per_launch.append(step_ms)
    kernel_ms = float(np.mean(per_launch))
    roof = roofline_block(job, kernel_ms, len(per_launch), sharded)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    pin_train = torch.from_numpy(job.train).pin_memory()
    pin_test = torch.from_numpy(job.test).pin_memory()
    h2d = pin_train.numel() * 4 + pin_test.numel() * 4
    d2h = 0
    e2e_steps = max(1, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.upload_data(pin_train.numpy(), pin_test.numpy())
        job.pt_round()
        tr = s.traces_all(job.iters_done - SWAP_INTERVAL, SWAP_INTERVAL)   # device -> host read of the step's results, every chain
        d2h = tr.nbytes
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * cpg * e2e_steps * SWAP_INTERVAL / e2e_s
    assert job.iters_done <= SAMPLES
    vectors_sent = s.vectors_sent() if sharded else 0
    job.close()

    # ---- the other BASELINE.json shapes, 8 replicas in all (the reference's own configurations), on the same N GPUs: one
    # 8-rung ladder, cut over the GPUs when N > 1 (strong scaling of a latency-bound job: reported, not the scaling claim)
    extra = []
    if not args.no_extra:
        for xshape in ("swiss", "coil", "madelon", "stress2048"):
            if xshape == "stress2048" and args.no_stress:
                continue
            xjob = PtJob(xshape, 1, N_RUNGS, world, rank, local, world > 1, comm, 0, samples=4000 if xshape != "stress2048" else 64)
            xsteps = {"swiss": 40, "coil": 40, "madelon": 10, "stress2048": 2}[xshape]
            if world > 1:
                for _ in range(2):
                    xjob.pt_round()
            xms = timed_rounds(xjob, xsteps, 3 if xshape != "stress2048" else 1, barrier, dist, world)
            xk = []
            nk = 3 if xshape != "stress2048" else 1
            if world == 1:
                for _ in range(nk):
                    xjob.pt_round()
                    xk.append(xjob.s.last_ms())
            xroof = roofline_block(xjob, float(np.mean(xk)), nk, False) if xk else None
            extra.append({"workload": f"{xshape} shape, 8 replicas = one 8-rung ladder" + (f" cut over {world} GPUs" if world > 1 else "") +
                                      f" (BASELINE.json configs[{ {'swiss': 0, 'coil': 1, 'madelon': 2, 'stress2048': 4}[xshape] }]" +
                                      (": 2048-1792-1536-1280, 100 000 rows, 8 of the 1024 replicas" if xshape == "stress2048" else "") + ")",
                          "metric": METRIC, "value": N_RUNGS * xsteps * SWAP_INTERVAL / (xms / 1000.0), "unit": UNIT, "n_gpus": world,
                          "ms_per_step": xms / xsteps, "steps": xsteps, "scaling": "strong",
                          "gemm_path": xjob.s.gemm_path(), "roofline": xroof})
            xjob.close()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, info = cpu_chain_steps_per_sec(shape, N_RUNGS, args.ref_iters)
            cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port",
                   "sample": f"{N_RUNGS} replica processes x {args.ref_iters} iterations, float64 numpy oracle, "
                             f"{info['threads_per_proc']} BLAS threads each ({info['loop_s']:.1f} s)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(shape, cpg, world, sharded), "shape": shape, "dims": list(dims4), "n_train": n_train,
                       "n_test": n_test, "chains_per_gpu": cpg, "n_rungs": N_RUNGS * (world if sharded else 1), "swap_interval": SWAP_INTERVAL,
                       "parallelism": (f"{world} GPUs x {cpg} chains: {n_ens} ensembles, each ladder of {N_RUNGS * world} rungs cut into "
                                       f"{N_RUNGS}-rung blocks per GPU; per round one 40 B/chain ncclAllGather, the sweep as a device kernel on "
                                       f"every rank, and ncclSend/ncclRecv of the weight vectors that cross a GPU boundary straight between the "
                                       f"theta slabs ({vectors_sent} vectors sent by rank 0)"
                                       if sharded else
                                       f"{world} GPU(s) x {cpg} chains, whole ensembles per GPU (no data-path collective)"),
                       "l2_policy": "working set per step (chain state + activations) exceeds the 126 MB L2 many times over; no flush needed",
                       "gemm_path": roof["mma_kind"]},
            "clocks": {"sm_mhz": clk.get("sm_mhz"), "sm_max_mhz": clk.get("sm_max_mhz"), "reasons": clk.get("reasons", []),
                       "samples": clk.get("samples", 0)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "exchange_ms_per_round": (float(np.median(exch_ms)) if exch_ms else None),
            "roofline": roof,
            "cpu_baseline": cpu,
            "extra_configs": extra,
        }
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="madelon", choices=list(SHAPES))
    ap.add_argument("--chains-per-gpu", type=int, default=64)
    ap.add_argument("--gemm-path", type=int, default=0)
    ap.add_argument("--ref-iters", type=int, default=2, help="iterations per replica process in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the Swiss / Coil / Madelon 8-replica lines (extra_configs)")
    ap.add_argument("--no-stress", action="store_true", help="skip the 2048-feature / 100k-row line of extra_configs (about 25 s)")
    ap.add_argument("--ladder", default="sharded", choices=["sharded", "local"],
                    help="N > 1: 'sharded' cuts every ladder across the GPUs (NCCL exchange), 'local' keeps whole ensembles per GPU")
    args = ap.parse_args()
    if args.chains_per_gpu % N_RUNGS:
        raise SystemExit("--chains-per-gpu must be a multiple of 8")
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
