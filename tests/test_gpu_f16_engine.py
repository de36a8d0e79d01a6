"""The 3xFP16 split of the tcgen05 engine (kind::f16 MMAs on scaled operands, bae_tc.cuh): operand-layout probe, engine
selection, and that the tracked magnitude bounds keep the FP32-level accuracy when data, weights and deltas sit far from
unit scale (FP16 has the mantissa of TF32 but only 5 exponent bits: without the scales these cases overflow or lose bits)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import bae_oracle as bo
from tests.parity_util import rel, synth_data

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WIDE = (500, 450, 400, 300)


def _sampler(train, test, **kw):
    from bayesianautoencoders_b200 import DeviceSampler
    return DeviceSampler(WIDE, train, test, 1, 1, 10, np.ones(1), seed=7, gemm_path=2, **kw)


def test_engine_selection(monkeypatch):
    train, test = synth_data(WIDE, 256, 256)
    s = _sampler(train, test)
    assert s.gemm_path() == 2 and s.mma_kind() == 2          # production: per-phase step program, 3xFP16 split
    s.close()
    monkeypatch.setenv("BAE_TC_TF32", "1")
    s = _sampler(train, test)
    assert s.mma_kind() == 1                                  # 3xTF32 split on request
    s.close()
    monkeypatch.setenv("BAE_TC_TF32", "0")
    monkeypatch.setenv("BAE_PHASE_LAUNCH", "0")
    s = _sampler(train, test)
    assert s.mma_kind() == 1                                  # the persistent kernel keeps the 3xTF32 split
    s.close()


@pytest.mark.parametrize("data_scale,w_scale", [(1.0, 1.0), (700.0, 0.3), (1e-3, 0.3), (3e4, 0.1)])
def test_forward_backward_far_from_unit_scale(data_scale, w_scale):
    """Forward, SSE and the whole gradient against the float64 oracle with the data matrix scaled by several orders of
    magnitude in both directions: the first layer's weights move the other way (pre-activations stay O(1), no sigmoid
    saturates), the output deltas (out - x) 2 / (N D) and everything behind them move with the data."""
    L = bo.Layout(WIDE)
    train, test = synth_data(WIDE, 640, 384)
    train = (train * data_scale).astype(np.float32)
    test = (test * data_scale).astype(np.float32)
    s = _sampler(train, test)
    assert s.mma_kind() == 2
    rng = np.random.default_rng(11)
    w = rng.standard_normal(L.w_size) * w_scale
    w[L.slice("encode.0.weight")] /= data_scale
    w = w.astype(np.float32)
    w[L.nbt_index] = 3.0
    r = s.eval(0, 0, w=w, grads=True, out=True)
    th = w.astype(np.float64)
    out, cache = bo.forward(L, th, train.astype(np.float64), update_buffers=False)
    g = bo.backward(L, th, cache)
    sse = float(np.sum((out - train) ** 2))
    assert abs(r["sse"] - sse) <= 1e-5 * sse
    # the reconstruction comes back as delta / scale + x (FP32): next to data of size `data_scale` its own rounding is
    # 6e-8 * data_scale in absolute terms, whatever computed it
    assert np.abs(r["out"] - out).max() <= 1e-5 * max(1.0, float(np.abs(out).max())) + 2e-7 * data_scale
    gd = r["grads"].astype(np.float64)
    assert np.all(np.isfinite(gd))
    assert np.linalg.norm(gd - g) <= 2e-5 * np.linalg.norm(g), np.linalg.norm(gd - g) / np.linalg.norm(g)
    rt = s.eval(0, 1, w=w)
    out_t, _ = bo.forward(L, th, test.astype(np.float64), update_buffers=False)
    sse_t = float(np.sum((out_t - test) ** 2))
    assert abs(rt["sse"] - sse_t) <= 1e-5 * sse_t
    s.close()


def test_operand_layout_probe_runs():
    """tests/cuda/tc_probe_f16.cu: the no-swizzle FP16 tiles, the descriptors and the three-MMA recombination on one 256 x 128 x 64
    product, all four operand-major combinations, against a float64 product of the unsplit FP32 inputs."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = os.path.join(ROOT, "tests", "cuda", "tc_probe_f16")
    src = exe + ".cu"
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], check=True,
                       cwd=os.path.dirname(src), timeout=300)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "PROBE_F16 OK" in p.stdout, p.stdout + p.stderr


def _run_images(monkeypatch, preconv, dims, n_train, n_test, chains, rungs, steps, si, act_group=0, himg="1"):
    """One free-running PT run; returns the raw trace bytes and the final (theta, m, v) of every chain."""
    from bayesianautoencoders_b200 import DeviceSampler, default_theta_init
    monkeypatch.setenv("BAE_TC_PRECONV", preconv)
    monkeypatch.setenv("BAE_TC_HIMG", himg)
    train, test = synth_data(dims, n_train, n_test)
    T = np.tile(2.0 ** (np.arange(rungs) / max(1, rungs - 1)), chains // rungs)
    kw = dict(act_group=act_group) if act_group else {}
    s = DeviceSampler(dims, train, test, chains, rungs, steps, T, swap_interval=si, seed=5, gemm_path=2, **kw)
    assert s.mma_kind() == 2
    rng = np.random.default_rng(3)
    for j in range(chains):
        s.init_chain(j, default_theta_init(dims, rng), rng.standard_normal(s.w_size))
    s.run(steps - 2)
    s.run(2)          # a second call shape: the forced weight-bound / image pass in front of a program
    s.synchronize()
    out = []
    for j in range(chains):
        st = s.get_state(j)
        out.append((s.traces(j, 0, steps).tobytes(), st["theta"].copy(), st["m"].copy(), st["v"].copy()))
    ev = s.eval(0, 0, grads=True)          # the parity probe runs the image path on the scratch chain
    s.close()
    return out, ev


@pytest.mark.parametrize("case", ["madelon", "groups", "d3_mult16"])
def test_operand_images_bit_identical(monkeypatch, case):
    """Pre-split operand images (weights once per proposal, data once per upload, activations persisted by the converter warps:
    bae_tc.cuh "operand images") hold exactly the FP16 hi / lo units the converter warps produce in shared memory, so every
    trace record, the final weights and the Adam state must be BIT-identical to the converter-only engine (BAE_TC_PRECONV=0).
    Cases: Madelon widths over an exchange round; chain groups smaller than the ensemble (activation images are per group
    slot); a bottleneck width that is a multiple of 16 (the BatchNorm output then keeps the converter path, the sigmoid
    outputs' ones column is static)."""
    if case == "madelon":
        cfg = dict(dims=WIDE, n_train=2000, n_test=1800, chains=4, rungs=4, steps=8, si=3)
    elif case == "groups":
        cfg = dict(dims=WIDE, n_train=520, n_test=300, chains=6, rungs=3, steps=7, si=2, act_group=4)
    else:
        cfg = dict(dims=(272, 256, 240, 128), n_train=700, n_test=400, chains=2, rungs=2, steps=6, si=2)
    ref, ev_ref = _run_images(monkeypatch, "0", **cfg)
    for himg in ("0", "1"):          # weights / data images only, then with the activation images as well (the default)
        got, ev = _run_images(monkeypatch, "1", himg=himg, **cfg)
        for j, (a, b) in enumerate(zip(ref, got)):
            assert a[0] == b[0], f"trace records of chain {j} differ (BAE_TC_HIMG={himg})"
            for k in (1, 2, 3):
                assert a[k].tobytes() == b[k].tobytes(), f"state vector {k} of chain {j} differs (BAE_TC_HIMG={himg})"
        assert ev["sse"] == ev_ref["sse"] and ev["grads"].tobytes() == ev_ref["grads"].tobytes()
