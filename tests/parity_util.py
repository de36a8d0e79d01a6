"""Shared helpers for the GPU parity tests: run the CUDA sampler and the CPU oracle on identical
weights, proposals and uniforms (the device's Philox draws are read back and fed to the oracle)."""
import numpy as np

from oracle import bae_oracle as bo


def synth_data(dims4, n_train, n_test, seed=1234):
    """SURVEY 8d synthetic inputs: U[0,1) float32 (mimics MinMaxScaler output, MAIN:108,126,133)."""
    train = np.random.default_rng(seed).random((n_train, dims4[0]), dtype=np.float32)
    test = np.random.default_rng(seed + 1).random((n_test, dims4[0]), dtype=np.float32)
    return train, test


def oracle_rnd_from_device(sampler, chain, n_steps, first_step=0):
    lx, noise, en, u = [], [], [], []
    for i in range(first_step, first_step + n_steps):
        d = sampler.debug_draws(chain, i)
        lx.append(d["lx"]); noise.append(d["noise"]); en.append(d["eta_noise"]); u.append(d["u"])
    return dict(lx=np.array(lx), noise=np.array(noise), eta_noise=np.array(en), u=np.array(u))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_device_and_oracle(dims4, train, test, temps, S, si, theta_inits, w0s, n_rungs, do_swaps, seed=21, hyper=None,
                          oracle_hyper=None, **sampler_kw):
    """Run the CUDA sampler (through the C-ABI) and the CPU oracle on identical weights, proposals and
    uniforms: the device's own Philox draws are read back (`bae_debug_draws`) and fed to `bo.run_pt`.
    Returns dict(dev, records, swap_log, counts, reps, final, gemm_path)."""
    from bayesianautoencoders_b200 import DeviceSampler
    n = len(temps)
    s = DeviceSampler(dims4, train, test, n, n_rungs, S, np.asarray(temps, float), seed=seed, swap_interval=si,
                      hyper=hyper, **sampler_kw)
    for j in range(n):
        s.init_chain(j, theta_inits[j], w0s[j])
    init_states = [s.get_state(j, vectors=False) for j in range(n)]
    rnd = []
    for j in range(n):
        d = oracle_rnd_from_device(s, j, S)
        d["w0_randn"] = w0s[j]
        rnd.append(d)
    # the oracle consumes swap uniforms round-major, pair-minor (one ensemble in these tests)
    swap_u = [s.debug_swap_uniform(0, r, p) for r in range(S // si) for p in range(n - 1)]
    s.run(S)
    s.synchronize()
    dev = [s.traces(j, 0, S) for j in range(n)]
    counts = s.swap_counts()
    final = [s.get_state(j) for j in range(n)]
    path = s.gemm_path()
    s.close()
    reps, records, swap_log = bo.run_pt(dims4, train, test, temps, S, si, theta_inits, rnd, swap_u=swap_u,
                                        do_swaps=do_swaps, hyper=oracle_hyper)
    return dict(dev=dev, records=records, swap_log=swap_log, counts=counts, reps=reps, final=final,
                init_states=init_states, gemm_path=path)


def trace_errors(dev, records):
    """Per-quantity worst errors of the device traces against the oracle's records, per iteration index.
    Relative errors are taken against the oracle value; `sum_value` against the scale of its terms."""
    S = len(records[0])
    out = dict(lik_prop=np.zeros(S), lik=np.zeros(S), prior=np.zeros(S), diff_prop=np.zeros(S), sum_value=np.zeros(S),
               mse_train=np.zeros(S), mse_test=np.zeros(S), log_u=np.zeros(S))
    decisions = []   # (chain, step, margin / scale, device accepted, oracle accepted)
    branch_ok = True
    for j in range(len(dev)):
        for i, rec in enumerate(records[j]):
            d = dev[j][i]
            branch_ok &= bool(d["langevin"]) == rec.langevin
            out["lik_prop"][i] = max(out["lik_prop"][i], abs(d["likelihood_proposal"] - rec.likelihood_proposal) / abs(rec.likelihood_proposal))
            out["lik"][i] = max(out["lik"][i], abs(d["likelihood"] - rec.likelihood) / abs(rec.likelihood))
            out["prior"][i] = max(out["prior"][i], abs(d["prior_proposal"] - rec.prior_prop) / abs(rec.prior_prop))
            scale = abs(rec.likelihood_proposal) + abs(rec.likelihood) + abs(rec.prior_prop) + abs(rec.diff_prop) + 1.0
            dscale = abs(rec.first) + abs(rec.second) + 1.0
            out["diff_prop"][i] = max(out["diff_prop"][i], abs(d["diff_prop"] - rec.diff_prop) / dscale)
            out["sum_value"][i] = max(out["sum_value"][i], abs(d["sum_value"] - rec.sum_value) / scale)
            out["mse_train"][i] = max(out["mse_train"][i], abs(d["mse_train_proposal"] - rec.mse_train) / rec.mse_train)
            out["mse_test"][i] = max(out["mse_test"][i], abs(d["mse_test_proposal"] - rec.mse_test) / rec.mse_test)
            out["log_u"][i] = max(out["log_u"][i], abs(d["log_u"] - rec.log_u) / max(1.0, abs(rec.log_u)))
            decisions.append((j, i, abs(rec.sum_value - rec.log_u) / scale, bool(d["accepted"]), bool(rec.accepted)))
    return out, decisions, branch_ok


def dump_report(name, report):
    """Measured parity errors go to gpurun_out/ (merged back from the GPU box) so that tolerances in the tests
    can be stated from measurements."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, f"parity_{name}.json"), "w") as f:
            json.dump(report, f, indent=1, default=lambda x: np.asarray(x).tolist())
    except OSError:
        pass


def _sync_oracle_to_device(rep, st):
    """Teacher forcing: the oracle replica takes the device chain's state (fp32 values widened to float64), so the
    next iteration starts from IDENTICAL weights, Adam moments and scalars on both sides. The oracle keeps its own
    `w_prop` (the chain's last proposal is not part of the device state, only its SSE) and its own `w_f32` flag."""
    rep.theta[:] = st["theta"].astype(np.float64)
    rep.m[:] = st["m"].astype(np.float64)
    rep.v[:] = st["v"].astype(np.float64)
    rep.t = int(st["adam_t"])
    rep.eta, rep.likelihood, rep.prior_current = st["eta"], st["likelihood"], st["prior_current"]
    rep.adapttemp = st["adapttemp"]
    rep.i, rep.init_count = int(st["step"]), int(st["init_count"])
    rep.num_accepted, rep.langevin_count = int(st["num_accepted"]), int(st["langevin_count"])
    rep.last_sse_train = st["last_sse_train"]
    rep.w_alias = bool(st["w_is_alias"])
    rep.w = None if rep.w_alias else st["w"].astype(np.float64)
    if rep.w_alias:
        rep.w_f32 = False
        # an aliased `w` IS the live model, which holds the chain's own last proposal: the i == pt re-evaluation
        # (MAIN:440-445) loads `w_proposal` back into it, so its trainable entries are synchronised as well
        msk = rep.L.trainable_mask
        rep.w_prop[msk] = rep.theta[msk]


def run_teacher_forced(dims4, train, test, temps, S, si, theta_inits, w0s, n_rungs, seed=21, hyper=None,
                       oracle_hyper=None, n_run=None, **sampler_kw):
    """Per-iteration parity with re-synchronisation: before EVERY iteration (and before every exchange sweep) the
    oracle takes the device's state, then both sides execute the same iteration on the same draws. What is compared
    is therefore the arithmetic of one iteration of MAIN:432-592 / one sweep of MAIN:962-1011 at a time - trajectory
    divergence (the two sides drifting apart over many iterations) cannot hide in, or leak into, the numbers.
    Returns a report dict with per-iteration worst errors over the chains."""
    from bayesianautoencoders_b200 import DeviceSampler
    n = len(temps)
    s = DeviceSampler(dims4, train, test, n, n_rungs, S, np.asarray(temps, float), seed=seed, swap_interval=si,
                      hyper=hyper, **sampler_kw)
    reps = []
    for j in range(n):
        s.init_chain(j, theta_inits[j], w0s[j])
        r = bo.Replica(dims4, train, test, temps[j], S, oracle_hyper)
        r.initialise(theta_inits[j], w0s[j])
        reps.append(r)
    keys = ("lik_prop", "lik_prop_terms", "lik", "prior", "diff_prop", "sum_value", "mse_train", "mse_test", "theta", "m", "v",
            "grad")
    err = {k: np.zeros(S if n_run is None else n_run) for k in keys}
    decisions, swaps, exchange_mismatch = [], [], []
    n_rw = 0
    train32 = np.asarray(train, dtype=np.float32)
    mask = reps[0].L.trainable_mask
    n_run = S if n_run is None else n_run     # iterations executed (the chain length S fixes the switch at i == S/2)
    for i in range(n_run):
        pre = [s.get_state(j) for j in range(n)]
        for j in range(n):
            _sync_oracle_to_device(reps[j], pre[j])
        draws = [s.debug_draws(j, i) for j in range(n)]
        # gradient of the MSE loss at the chain's current weights through the parity probe (same GEMM / BatchNorm phases
        # as the step program; does not disturb the chain): what the first drift call of this iteration differentiates
        g_dev = [s.eval(j, 0, w=(pre[j]["theta"] if pre[j]["w_is_alias"] else pre[j]["w"]), grads=True)["grads"]
                 if draws[j]["lx"] < float(np.float32(s.hyper.l_prob)) and s.hyper.use_langevin else None for j in range(n)]
        s.step(1)
        s.synchronize()
        for j in range(n):
            d = draws[j]
            rec = reps[j].step(bo.StepRandomness(d["lx"], d["noise"], d["eta_noise"], d["u"]))
            dv = s.traces(j, i, 1)[0]
            assert bool(dv["langevin"]) == rec.langevin
            n_rw += 0 if rec.langevin else 1
            scale = abs(rec.likelihood_proposal) + abs(rec.likelihood) + abs(rec.prior_prop) + abs(rec.diff_prop) + 1.0
            dscale = abs(rec.first) + abs(rec.second) + 1.0
            tau = float(np.exp(rec.eta_pro))
            nel = train.shape[0] * train.shape[1]
            terms = (abs((nel / 2.0) * np.log(2 * np.pi * tau)) + 0.5 * tau * rec.sse_train) / reps[j].adapttemp
            e = dict(lik_prop=abs(dv["likelihood_proposal"] - rec.likelihood_proposal) / abs(rec.likelihood_proposal),
                     lik_prop_terms=abs(dv["likelihood_proposal"] - rec.likelihood_proposal) / terms,
                     lik=abs(dv["likelihood"] - rec.likelihood) / abs(rec.likelihood),
                     prior=abs(dv["prior_proposal"] - rec.prior_prop) / abs(rec.prior_prop),
                     diff_prop=abs(dv["diff_prop"] - rec.diff_prop) / dscale,
                     sum_value=abs(dv["sum_value"] - rec.sum_value) / scale,
                     mse_train=abs(dv["mse_train_proposal"] - rec.mse_train) / rec.mse_train,
                     mse_test=abs(dv["mse_test_proposal"] - rec.mse_test) / rec.mse_test)
            post = s.get_state(j)
            # one-iteration errors of the live model and of the Adam moments (they carry both gradients of the iteration)
            e["theta"] = rel(post["theta"][mask], reps[j].theta[mask])
            if rec.langevin:
                e["m"] = rel(post["m"], reps[j].m)
                e["v"] = rel(post["v"], reps[j].v)
                e["grad"] = rel(g_dev[j], reps[j].grad1)
            for k, val in e.items():
                err[k][i] = max(err[k][i], float(val))
            decisions.append((j, i, float(abs(rec.sum_value - rec.log_u) / scale), bool(dv["accepted"]), bool(rec.accepted)))
        if (i + 1) % si == 0:
            pre = [s.get_state(j) for j in range(n)]
            for j in range(n):
                _sync_oracle_to_device(reps[j], pre[j])
            rnd = (i + 1) // si - 1
            s.swap()
            s.synchronize()
            src = s.last_sweep()
            for e0 in range(n // n_rungs):
                # The oracle evaluates every pair of the sequential sweep (two fp32 forwards per pair in the main process,
                # MAIN:980-983); the queue then advances with the DEVICE's decision for the pair, so each decision is
                # compared from identical queue contents. Near the threshold the reference's own fp32 re-evaluation of a
                # ~1e7-sized log-likelihood decides by rounding noise; the margin is recorded with the decision.
                b = e0 * n_rungs
                params = [reps[b + p].post() for p in range(n_rungs)]
                for p in range(n_rungs - 1):
                    u = s.debug_swap_uniform(e0, rnd, p)
                    dev_sw = bool(src[b + p] == b + p + 1)
                    p1, p2, sw, l12, l21 = bo.swap_procedure(dims4, train32, params[p], params[p + 1], u)
                    lh1, lh2 = params[p][-3], params[p + 1][-3]
                    ex = (l12 - lh1) + (l21 - lh2)
                    mag = abs(l12) + abs(lh1) + abs(l21) + abs(lh2)
                    if dev_sw != bool(sw):   # follow the device: re-run the pair with a uniform that forces its decision
                        p1, p2, _, _, _ = bo.swap_procedure(dims4, train32, params[p], params[p + 1], -1.0 if dev_sw else 2.0)
                    params[p], params[p + 1] = p1, p2
                    swaps.append((rnd, e0, p, dev_sw, bool(sw), float(abs(ex - np.log(u)) / mag), float(ex), float(u)))
                for p in range(n_rungs):
                    reps[b + p].take(params[p])
            post = [s.get_state(j) for j in range(n)]
            for j in range(n):
                if abs(post[j]["eta"] - reps[j].eta) > 1e-12 or post[j]["w_is_alias"] != 0 or rel(post[j]["w"], reps[j].w) != 0.0:
                    exchange_mismatch.append((rnd, j, float(post[j]["eta"]), float(reps[j].eta), float(rel(post[j]["w"], reps[j].w))))
    counts = s.swap_counts()
    path = s.gemm_path()
    dev_traces = [s.traces(j, 0, n_run) for j in range(n)]
    s.close()
    return dict(errors={k: v.tolist() for k, v in err.items()}, decisions=decisions, swaps=swaps, swap_counts=counts,
                exchange_mismatch=exchange_mismatch, random_walk_steps=n_rw, gemm_path=path), dev_traces
