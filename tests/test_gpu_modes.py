"""GPU tests of the widened rows of SURVEY.md 8f: alternative drift modes (experiment_type = 1, MALA), the linear
ladder, and the batched posterior-predictive / encoder-feature pass."""
import numpy as np
import pytest

from oracle import bae_oracle as bo
from tests.golden_util import GoldenCase
from tests.parity_util import rel, run_device_and_oracle, synth_data, trace_errors

pytestmark = pytest.mark.gpu


def _assert_traces(r, lik_rtol=2e-5, sum_rtol=5e-5, margin=5e-4):
    errs, decisions, branch_ok = trace_errors(r["dev"], r["records"])
    assert branch_ok
    assert max(errs["lik_prop"].max(), errs["lik"].max()) <= lik_rtol, (errs["lik_prop"], errs["lik"])
    assert errs["sum_value"].max() <= sum_rtol, errs["sum_value"]
    assert errs["diff_prop"].max() <= sum_rtol, errs["diff_prop"]
    n = 0
    for j, i, mg, a, b in decisions:
        if mg > margin:
            assert a == b, (j, i, mg)
            n += 1
    assert n >= 0.8 * len(decisions)
    assert r["counts"][1] == len(r["swap_log"]) and r["counts"][0] == sum(1 for s in r["swap_log"] if s[2])


def test_sgd_after_pt_matches_oracle():
    """experiment_type = 1 (MAIN:72, 454-470): Adam while i <= pt, then a stateless SGD(lr = 0.1) step. The oracle's
    mode is pinned to the reference by tests/golden/mid_sgd2.npz; inputs of that case, draws from the device."""
    from bayesianautoencoders_b200 import Hyper
    g = GoldenCase("mid_sgd2")
    assert g.experiment_type == 1
    r = run_device_and_oracle(g.dims4, g.train, g.test, list(g.temps), g.S, g.swap_interval,
                              [g.theta_init(j) for j in range(g.n)], [g.rnd(j)["w0_randn"] for j in range(g.n)],
                              n_rungs=g.n, do_swaps=True, hyper=Hyper(drift_mode=1), oracle_hyper=bo.Hyper(experiment_type=1))
    _assert_traces(r)
    n_sgd = sum(1 for rec in r["records"][0][g.S // 2 + 1:] if rec.langevin)
    assert n_sgd >= 2
    for j in range(g.n):
        # the Adam state stops moving once the chain has switched to SGD
        assert r["final"][j]["adam_t"] == r["reps"][j].t
        assert rel(r["final"][j]["m"], r["reps"][j].m) < 1e-4


def test_mala_drift_matches_oracle():
    """Plain-gradient drift on the tempered log-posterior (not in the reference: the oracle restates the
    same formula in float64)."""
    from bayesianautoencoders_b200 import Hyper
    g = GoldenCase("mid_pt3")
    r = run_device_and_oracle(g.dims4, g.train, g.test, list(g.temps), g.S, g.swap_interval,
                              [g.theta_init(j) for j in range(g.n)], [g.rnd(j)["w0_randn"] for j in range(g.n)],
                              n_rungs=g.n, do_swaps=True, hyper=Hyper(drift_mode=2), oracle_hyper=bo.Hyper(mala=True))
    _assert_traces(r)
    for j in range(g.n):
        assert r["final"][j]["adam_t"] == 0 and not np.any(r["final"][j]["m"])


def test_linear_ladder_matches_oracle():
    """`geometric = False` (MAIN:930-934): T = 1, 1 + maxtemp / n, ... through the device sweep."""
    g = GoldenCase("swiss_pt4")
    temps = list(bo.temperature_ladder(g.n, 2, geometric=False))
    assert temps == [1.0, 1.5, 2.0, 2.5]
    r = run_device_and_oracle(g.dims4, g.train, g.test, temps, g.S, g.swap_interval,
                              [g.theta_init(j) for j in range(g.n)], [g.rnd(j)["w0_randn"] for j in range(g.n)],
                              n_rungs=g.n, do_swaps=True)
    _assert_traces(r)


@pytest.mark.parametrize("shape", ["mid", "madelon_width"])
def test_posterior_predictive_batch_matches_oracle(shape):
    """`cae.encode(X)` / `cae.forward(X)` (MAIN:714, 752, 787) for K posterior samples in one chain-batched pass."""
    from bayesianautoencoders_b200 import PosteriorPredictive
    dims4, rows, K, gp = {"mid": ((24, 20, 16, 12), 150, 5, 0), "madelon_width": ((500, 450, 400, 300), 600, 3, 2)}[shape]
    X, _ = synth_data(dims4, rows, 4)
    L = bo.Layout(dims4)
    rng = np.random.default_rng(8)
    W = (rng.standard_normal((K, L.w_size)) * 0.3).astype(np.float32)
    pp = PosteriorPredictive(dims4, X, K, gemm_path=gp)
    pp.load_all(W)
    feats, recon, mse = pp.encode(), pp.reconstruct(), pp.mse()
    assert feats.shape == (K, rows, dims4[3]) and recon.shape == (K, rows, dims4[0])
    for k in range(K):
        out, c = bo.forward(L, W[k].astype(np.float64), X.astype(np.float64), update_buffers=False)
        z = c["zhat"] / c["inv"] + c["mu"]                         # encode(x): pre-BatchNorm bottleneck
        assert rel(feats[k], z) < 1e-5
        assert rel(recon[k], out) < 1e-5
        assert abs(mse[k] - np.mean((out - X) ** 2)) < 1e-5 * np.mean((out - X) ** 2)
    # the handle's chain state is untouched by the pass: same weights -> same answer twice
    assert np.array_equal(pp.encode(), feats)
    pp.close()


@pytest.mark.parametrize("shape", ["mid_fp32", "wide_tc"])
def test_chain_groups_are_bit_identical(shape):
    """Chain-group scheduling of the activation buffers (`act_group`; BASELINE.json configs[4] needs it: 7.7 GB of
    activations per chain): the step program run group after group gives the traces of the all-resident run, bit for bit."""
    from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
    dims4, ntr, nte, n, gp, grp = {"mid_fp32": ((24, 20, 16, 12), 96, 40, 3, 0, 2), "wide_tc": ((160, 144, 136, 128), 384, 256, 4, 2, 3)}[shape]
    train, test = synth_data(dims4, ntr, nte)
    L = bo.Layout(dims4)
    temps = bo.temperature_ladder(n, 2)
    runs = []
    for act_group in (0, grp, 1):
        s = DeviceSampler(dims4, train, test, n, n, 10, temps, swap_interval=5, seed=9, gemm_path=gp, act_group=act_group)
        rng = np.random.default_rng(2)
        for j in range(n):
            s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(L.w_size))
        s.run(10)
        s.synchronize()
        fa = s.forward_all(1)
        runs.append(([s.traces(j, 0, 10).tobytes() for j in range(n)], [s.get_state(j)["theta"].tobytes() for j in range(n)],
                     fa["features"].tobytes(), fa["recon"].tobytes(), s.swap_counts()))
        s.close()
    assert runs[0] == runs[1] == runs[2]


def test_narrow_kernel_is_used_and_agrees_with_the_generic_step_kernel(monkeypatch):
    """Swiss Roll widths run the cluster kernel of csrc/bae_narrow.cu (step_kernel 3). Same chains through the generic
    persistent FP32 tile kernel (BAE_NO_NARROW=1): the two programs agree to FP32 rounding on every iteration, take the same
    decisions and the same swaps; cluster sizes 8 (8 chains) and 1 (> 64 chains) give the same traces up to rounding."""
    from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
    dims4, ntr, nte = (3, 10, 5, 2), 3750, 1250
    train, test = synth_data(dims4, ntr, nte)
    L = bo.Layout(dims4)
    S = 15

    def run(n_ens, env):
        if env:
            monkeypatch.setenv("BAE_NO_NARROW", "1")
        else:
            monkeypatch.delenv("BAE_NO_NARROW", raising=False)
        n = 8 * n_ens
        s = DeviceSampler(dims4, train, test, n, 8, S, np.tile(bo.temperature_ladder(8, 2), n_ens), swap_interval=5, seed=3)
        kind = s.step_kernel()
        rng = np.random.default_rng(7)
        for j in range(n):
            s.init_chain(j, default_theta_init(dims4, rng), rng.standard_normal(L.w_size))
        s.run(S)
        s.synchronize()
        tr = [s.traces(j, 0, S) for j in range(8)]
        out = (kind, tr, s.swap_counts(), [s.get_state(j)["theta"] for j in range(8)])
        s.close()
        return out

    k3, tr3, sw3, th3 = run(1, False)        # cluster of 8 CTAs per chain
    k1, tr1, sw1, th1 = run(1, True)         # generic kernel
    k3b, tr3b, _, _ = run(9, False)          # 72 chains: one CTA per chain; the first ensemble has the same seeds and ids
    assert (k3, k1, k3b) == (3, 1, 3)
    for j in range(8):
        for a, b in ((tr3[j], tr1[j]), (tr3[j], tr3b[j])):
            assert np.array_equal(a["accepted"], b["accepted"]) and np.array_equal(a["langevin"], b["langevin"])
            np.testing.assert_allclose(a["mse_train_proposal"], b["mse_train_proposal"], rtol=2e-5)
            # the log-likelihood is the difference of two ~1.7e4-sized terms at this size: absolute tolerance
            np.testing.assert_allclose(a["likelihood_proposal"], b["likelihood_proposal"], rtol=2e-5, atol=0.05)
            scale = np.abs(a["likelihood_proposal"]) + np.abs(a["likelihood"]) + np.abs(a["prior_proposal"]) + np.abs(a["diff_prop"]) + 1.0
            assert np.all(np.abs(a["sum_value"] - b["sum_value"]) <= 2e-4 * scale), np.max(np.abs(a["sum_value"] - b["sum_value"]) / scale)
        assert rel(th3[j], th1[j]) < 1e-4
    assert sw3 == sw1
