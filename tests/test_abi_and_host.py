"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares, struct layouts match,
the library refuses to run without a GPU (no CPU fallback), and the host-side mirror logic (parameter
plumbing, ladder, result files, R-hat) matches the reference's definitions."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import bae_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    from bayesianautoencoders_b200 import _native
    return _native.load()


def test_every_header_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "bae_b200.h")).read()
    names = sorted(set(re.findall(r"\b(bae_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bae_b200.h but not exported"
    from bayesianautoencoders_b200 import _native
    assert set(_native.EXPORTS) <= set(names)


def test_struct_layouts_match_header():
    from bayesianautoencoders_b200 import _native as N
    assert C.sizeof(N.Trace) == 152            # 11 doubles + 13 floats + 3 ints
    assert C.sizeof(N.ChainScalars) == 96      # 8 doubles + 6 ints + 1 double
    assert C.sizeof(N.Config) == 104
    assert N.Config.seed.offset % 8 == 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bayesianautoencoders_b200 import BaeError, DeviceSampler
    with pytest.raises(BaeError, match="no CUDA device|no CPU fallback"):
        DeviceSampler((3, 10, 5, 2), np.zeros((10, 3), np.float32), np.zeros((10, 3), np.float32), 1, 1, 10, [1.0])


def test_invalid_config_is_rejected_before_touching_the_device(lib):
    from bayesianautoencoders_b200 import _native as N
    cfg = N.Config()
    h = C.c_void_p()
    assert lib.bae_create(C.byref(cfg), 0, C.byref(h)) == -1      # dims are zero
    assert b"dims" in lib.bae_last_error()
    for i, d in enumerate((3, 10, 5, 2)):
        cfg.dims[i] = d
    cfg.n_train, cfg.n_test, cfg.n_chains, cfg.n_rungs, cfg.samples, cfg.swap_interval = 10, 10, 6, 4, 5, 5
    assert lib.bae_create(C.byref(cfg), 0, C.byref(h)) == -1      # 6 chains is not a multiple of 4 rungs
    assert lib.bae_create(None, 0, C.byref(h)) == -1
    assert lib.bae_destroy(None) == 0


def test_model_parameter_plumbing_matches_reference_order():
    """getparameters = sorted-key order incl. BatchNorm buffers (MAIN:228-241); dictfromlist rounds to fp32
    (MAIN:243-251); load_state_dict truncates num_batches_tracked (int64 buffer)."""
    from bayesianautoencoders_b200 import Model, SORTED_KEYS
    dims4 = (6, 5, 4, 3)
    m = Model(dims4)
    L = bo.Layout(dims4)
    assert m.w_size == L.w_size
    assert list(SORTED_KEYS) == list(bo.SORTED_KEYS)
    vec = np.arange(L.w_size, dtype=np.float64) * 0.37 + 0.123456789123
    d = m.dictfromlist(vec)
    assert list(d.keys()) == bo.SORTED_KEYS
    for k in bo.SORTED_KEYS:
        assert d[k].dtype == np.float32 and d[k].shape == (L.shapes[k] if L.shapes[k] else ())
        np.testing.assert_array_equal(d[k].reshape(-1), vec[L.slice(k)].astype(np.float32))
    flat = m.getparameters(d)
    np.testing.assert_array_equal(flat, vec.astype(np.float32).astype(np.float64))
    m.loadparameters(d)
    assert m.getparameters()[L.nbt_index] == np.trunc(np.float32(vec[L.nbt_index]))
    sd = m.state_dict()
    assert list(sd.keys())[0] == "encode.0.weight" and "decode.0.num_batches_tracked" in sd
    w2 = m.addnoiseandcopy(d, 0, 0.005, rng=np.random.default_rng(1))
    assert list(w2.keys()) == list(d.keys())
    assert abs(float(np.std(m.getparameters(w2) - flat)) - 0.005) < 1e-3


def test_ladder_matches_reference_formula():
    from bayesianautoencoders_b200 import ParallelTempering
    pt = ParallelTempering(True, 8, 2, 48000, 5, None, 10, 0.5, 0.005, dims4=(3, 10, 5, 2),
                           traindata=np.zeros((4, 3)), testdata=np.zeros((4, 3)))
    pt.assign_temperatures()
    np.testing.assert_allclose(pt.temperatures, 2.0 ** (np.arange(8) / 7.0), rtol=1e-12)     # MAIN:914, 924-926
    np.testing.assert_allclose(pt.temperatures, bo.temperature_ladder(8, 2), rtol=1e-15)
    assert pt.NumSamples == 6000                                                               # MAIN:846
    for bad in ((0, 8, 2), (2, 8, 1), (2, 0, 2)):
        with pytest.raises(ValueError):
            pt.default_beta_ladder(*bad)
    with pytest.raises(ValueError):
        pt.default_beta_ladder(2, None, None)
    # every branch of MAIN:867-920 behaves like the reference, its dead ones included (values probed from the reference)
    np.testing.assert_allclose(pt.default_beta_ladder(3, 4, 4)[:3], [1.0, 0.6299605249474366, 0.3968502629920499], rtol=1e-14)
    np.testing.assert_allclose(pt.default_beta_ladder(5, 8, 3)[-1], 1.0 / 3.0, rtol=1e-14)
    with pytest.raises(TypeError):
        pt.default_beta_ladder(2, None, 3)          # `ntemps ** ...` with ntemps None: the reference raises here as well
    with pytest.raises(ZeroDivisionError):
        pt.default_beta_ladder(2, 1, 2)
    if os.path.exists("/root/reference/Bayesian/bayes_autoenc_langevincheck.py"):
        from oracle import ref_import
        ref = ref_import.load_reference()
        rp = ref.ParallelTempering.__new__(ref.ParallelTempering)
        for args in [(2, 8, 2), (2, 64, 2), (1, 8, 3), (5, 8, 3), (3, 4, 4), (40, 8, 2)]:
            np.testing.assert_array_equal(rp.default_beta_ladder(*args), pt.default_beta_ladder(*args))
    # linear ladder (geometric = False, MAIN:930-934)
    pt2 = ParallelTempering(True, 4, 2, 400, 5, None, 10, 0.5, 0.005, dims4=(3, 10, 5, 2),
                            traindata=np.zeros((4, 3)), testdata=np.zeros((4, 3)))
    pt2.geometric = False
    pt2.assign_temperatures()
    np.testing.assert_allclose(pt2.temperatures, [1.0, 1.5, 2.0, 2.5])
    np.testing.assert_allclose(pt2.temperatures, bo.temperature_ladder(4, 2, geometric=False))


def test_replica_files_names_and_formats(tmp_path):
    """MAIN:631-696: 22 text files per temperature, names `<stat>_<temperature>.txt`, printf formats."""
    from bayesianautoencoders_b200 import write_replica_files
    from bayesianautoencoders_b200.sampler import TRACE_DTYPE
    n = 7
    tr = np.zeros(n, dtype=TRACE_DTYPE)
    rng = np.random.default_rng(0)
    for k in ("sum_value", "likelihood", "likelihood_proposal", "log_u"):
        tr[k] = rng.standard_normal(n) * 100
    tr["mse_train"] = rng.random(n)
    tr["mse_test"] = rng.random(n)
    tr["weights"] = rng.standard_normal((n, 13))
    files = write_replica_files(str(tmp_path), 1.1040895136738123, tr, 37.9)
    names = sorted(os.path.basename(f) for f in files)
    t = "1.1040895136738123"
    expect = ["sum_value_", "likelihood_value_", "likelihood_value_proposal", "mse_test_chain_", "mse_train_chain_",
              "acc_test_chain_", "acc_train_chain_", "u_value_", "accept_percentage"] + \
             [f"weight[{i}]_" for i in (0, 100, 10000, 5000, 2000, 3000, 4000, 6000, 7000, 8000, 9000, 11000)]
    assert sorted(set(names)) == sorted(e + t + ".txt" for e in expect)
    got = np.loadtxt(tmp_path / f"mse_train_chain_{t}.txt")
    np.testing.assert_allclose(got, tr["mse_train"], atol=5e-6)                 # %1.5f
    line = open(tmp_path / f"sum_value_{t}.txt").readline().strip()
    assert re.fullmatch(r"-?\d+\.\d{2}", line)                                  # %1.2f
    assert re.fullmatch(r"-?\d+\.\d{4}", open(tmp_path / f"likelihood_value_{t}.txt").readline().strip())
    np.testing.assert_allclose(np.loadtxt(tmp_path / f"acc_train_chain_{t}.txt"), np.sqrt(tr["mse_train"]), atol=5e-3)
    assert open(tmp_path / f"accept_percentage{t}.txt").read() == "37"         # '%d' % accept_ratio
    # weight[5000] is written twice by the reference (arrays 3 and 7, same values): column 7 wins
    np.testing.assert_allclose(np.loadtxt(tmp_path / f"weight[5000]_{t}.txt"), tr["weights"][:, 7], atol=5e-3)


def test_gelman_rubin_matches_direct_formulas():
    """GR:690-735 restated independently on a small example."""
    from bayesianautoencoders_b200 import gelman_rubin
    rng = np.random.default_rng(3)
    data = rng.standard_normal((5, 400, 3)) + np.array([0.0, 1.0, -2.0])
    r = gelman_rubin(data)
    m, n = 5, 400
    means = data.mean(axis=1)
    B_on_n = means.var(axis=0)
    W = data.var(axis=1).mean(axis=0)
    Vhat = (n / (n - 1)) * W + B_on_n + B_on_n / m
    np.testing.assert_allclose(r["simple"], Vhat / W, rtol=1e-12)
    assert np.all(np.abs(r["simple"] - 1) < 0.05) and np.all(np.abs(r["advanced"] - 1) < 0.08)
    shifted = data + np.arange(5)[:, None, None] * 3.0        # chains that disagree -> R-hat >> 1
    assert np.all(gelman_rubin(shifted)["simple"] > 3)


def test_host_sweep_matches_python_restatement():
    """`bae_host_sweep` (C, used by the ladder-sharded exchange on GPU boxes) == `distributed.sweep` (Python, used by the
    gloo test) on random tables: same permutation and swap count (MAIN:962-1011, 1059-1065)."""
    from bayesianautoencoders_b200.distributed import sweep, sweep_native, swap_uniforms
    rng = np.random.default_rng(11)
    for R in (2, 8, 64):
        for trial in range(5):
            tab = np.zeros((R, 4))
            tab[:, 0] = rng.normal(-2.0, 0.3, R)                       # eta
            tab[:, 2] = 2.0 ** (np.arange(R) / max(1, R - 1))          # adapttemp
            tab[:, 3] = rng.uniform(8.0e4, 8.5e4, R)                   # SSE of w
            n = 1.0e6
            tab[:, 1] = [(-(n * 0.5) * np.log(2 * np.pi * np.exp(e)) - 0.5 * np.exp(e) * s) / t for e, t, s in tab[:, [0, 2, 3]]]
            u = swap_uniforms(123 + trial, [trial], 0, R - 1)[0]
            res = sweep_native(tab, n, u)
            assert res is not None
            ref = sweep(tab, n, [float(x) for x in u])
            assert res[0] == list(ref[0]) and res[1] == ref[1]


def test_sweep_non_finite_likelihoods_follow_the_reference_min():
    """MAIN:988 `min(1, exp(x))`: a NaN exponent gives probability 1 (Python's min keeps its first argument),
    +inf overflows to 1, -inf gives 0. The C host sweep, the Python sweep and the device decision (fmin) agree."""
    from bayesianautoencoders_b200.distributed import sweep, sweep_native
    n = 400.0
    base = np.array([[-1.0, -300.0, 1.0, 60.0], [-1.1, -280.0, 1.5, 70.0], [-0.9, -260.0, 2.0, 65.0]])
    for bad in (np.nan, np.inf, -np.inf):
        tab = base.copy()
        tab[1, 1] = bad                       # stored likelihood of the middle rung
        u = np.array([0.999, 0.999], np.float32)
        ref = sweep(tab, n, [float(x) for x in u])
        res = sweep_native(tab, n, u)
        assert res is not None and res[0] == list(ref[0]) and res[1] == ref[1], (bad, res, ref)
        with np.errstate(all="ignore"):
            lh12 = (-(n / 2) * np.log(2 * np.pi * np.exp(tab[0, 0])) - 0.5 * np.exp(tab[0, 0]) * tab[0, 3]) / tab[1, 2]
            lh21 = (-(n / 2) * np.log(2 * np.pi * np.exp(tab[1, 0])) - 0.5 * np.exp(tab[1, 0]) * tab[1, 3]) / tab[0, 2]
            want_first = 0.999 < min(1, np.exp((lh12 - tab[0, 1]) + (lh21 - tab[1, 1])))
        assert (ref[0][0] == 1) == bool(want_first), bad


def test_gelman_rubin_pinned_to_reference_script():
    """`gelman_rubin_reference` against the output of the UNMODIFIED last block of `Gelman Reuben.py` (GR:676-735),
    frozen by oracle/make_gr_golden.py: 2-decimal traces, axes swapped by the script's transposes."""
    from bayesianautoencoders_b200 import gelman_rubin, gelman_rubin_reference
    z = np.load(os.path.join(ROOT, "tests", "golden", "gr_rhat.npz"))
    w = np.stack([z["w_" + str(n)] for n in z["names"]], axis=2)          # [chains, samples, 12] in GR:690 order
    assert tuple(z["data_shape"]) == (w.shape[1], w.shape[0], 12)          # the script's "Nchains" is the sample count
    r = gelman_rubin_reference(w)
    np.testing.assert_allclose(r["simple"], z["simple"], rtol=1e-12)
    np.testing.assert_allclose(r["advanced"], z["advanced"], rtol=1e-10)
    # the textbook orientation gives a different number on the same traces
    assert np.abs(gelman_rubin(w)["simple"] - z["simple"]).max() > 0.05
    # live reference, when mounted (build container): regenerate and compare
    if os.path.exists("/root/reference/Bayesian/Gelman Reuben.py"):
        from oracle.make_gr_golden import run_reference_block
        live = run_reference_block({str(n): z["w_" + str(n)] for n in z["names"]})
        np.testing.assert_array_equal(live["simple"], z["simple"])
        np.testing.assert_array_equal(live["advanced"], z["advanced"])


def test_plan_exchange_matches_permutation_semantics():
    """`bae_plan_exchange` (host side of bae_swap_exchange): for random sweep results, every configuration that changes
    GPU appears exactly once as a send on its owner and once as a receive on its destination's owner, and the k-th
    message from rank a to rank b is the same (ensemble, configuration) on both sides (NCCL point-to-point matching)."""
    from bayesianautoencoders_b200.distributed import plan_exchange, sweep
    rng = np.random.default_rng(3)
    for world, L, E in ((2, 8, 3), (4, 2, 5), (8, 8, 8), (8, 1, 2)):
        R = world * L
        src = np.zeros((E, R), np.int32)
        for e in range(E):
            # a permutation the sequential sweep can produce: carried configuration + downward shifts
            cfg_prev = 0
            for p in range(R - 1):
                if rng.random() < 0.6:
                    src[e, p] = p + 1
                else:
                    src[e, p] = cfg_prev
                    cfg_prev = p + 1
            src[e, R - 1] = cfg_prev
            assert sorted(src[e]) == list(range(R))
        plans = [plan_exchange(src, world, r) for r in range(world)]
        crossing = [(e, int(src[e, p]), p) for e in range(E) for p in range(R) if src[e, p] // L != p // L]
        sends = [(r, peer, e, rung) for r in range(world) for kind, peer, e, rung in plans[r] if kind == "send"]
        recvs = [(r, peer, e, rung) for r in range(world) for kind, peer, e, rung in plans[r] if kind == "recv"]
        assert len(sends) == len(recvs) == len(crossing)
        for a in range(world):
            for b in range(world):
                if a == b:
                    continue
                s_ab = [(e, rung) for r, peer, e, rung in sends if r == a and peer == b]          # (ensemble, source rung)
                r_ab = [(e, int(src[e, rung])) for r, peer, e, rung in recvs if r == b and peer == a]  # destination -> its source
                assert s_ab == r_ab, (world, a, b)
