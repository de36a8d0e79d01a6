"""On-disk result contract of a replica (MAIN:631-696) and the Gelman-Rubin statistic the reference's
offline script computes from those files (`Bayesian/Gelman Reuben.py`, "GR", lines 676-735)."""
from __future__ import annotations

import os

import numpy as np

# (file stem, trace field or weights column, printf format)   MAIN:631-692
_FILES = [
    ("sum_value_", "sum_value", "%1.2f"),
    ("likelihood_value_", "likelihood", "%1.4f"),
    ("likelihood_value_proposal", "likelihood_proposal", "%1.4f"),
    ("weight[0]_", 0, "%1.2f"), ("weight[100]_", 1, "%1.2f"), ("weight[10000]_", 2, "%1.2f"),
    ("weight[5000]_", 3, "%1.2f"), ("weight[2000]_", 4, "%1.2f"), ("weight[3000]_", 5, "%1.2f"),
    ("weight[4000]_", 6, "%1.2f"), ("weight[5000]_", 7, "%1.2f"),   # the reference writes weight[5000] twice (MAIN:649,661)
    ("weight[6000]_", 8, "%1.2f"), ("weight[7000]_", 9, "%1.2f"), ("weight[8000]_", 10, "%1.2f"),
    ("weight[9000]_", 11, "%1.2f"), ("weight[11000]_", 12, "%1.2f"),
    ("mse_test_chain_", "mse_test", "%1.5f"), ("mse_train_chain_", "mse_train", "%1.5f"),
    ("acc_test_chain_", "rmse_test", "%1.2f"), ("acc_train_chain_", "rmse_train", "%1.2f"),
    ("u_value_", "log_u", "%1.4f"),
]


def replica_columns(traces):
    """Per-iteration arrays in the reference's naming (acc_* is the RMSE, MAIN:544-545)."""
    cols = {k: np.asarray(traces[k], dtype=np.float64) for k in
            ("sum_value", "likelihood", "likelihood_proposal", "mse_test", "mse_train", "log_u")}
    cols["rmse_train"] = np.sqrt(cols["mse_train"])
    cols["rmse_test"] = np.sqrt(cols["mse_test"])
    return cols


def write_replica_files(path, temperature, traces, accept_ratio):
    """Write the 22 per-temperature text files of MAIN:631-696 with the same names and formats.
    `traces` is the structured array of `DeviceSampler.traces`."""
    os.makedirs(path, exist_ok=True)
    cols = replica_columns(traces)
    w = np.asarray(traces["weights"], dtype=np.float64)
    t = str(temperature)
    written = []
    for stem, key, fmt in _FILES:
        arr = cols[key] if isinstance(key, str) else w[:, key]
        name = os.path.join(path, stem + t + ".txt")
        np.savetxt(name, arr, fmt=fmt)
        written.append(name)
    name = os.path.join(path, "accept_percentage" + t + ".txt")
    with open(name, "w") as f:
        f.write("%d" % accept_ratio)                                                      # MAIN:694-696
    written.append(name)
    return written


def gelman_rubin(data, advanced=True):
    """R-hat of GR:690-735. `data` has shape [Nchains, Nsamples, Npars] (chains first).

    Returns dict(simple=..., advanced=...): the "simple version, as in Obsidian" (GR:699-702) and, when
    `advanced`, the degrees-of-freedom corrected value (GR:710-731). Note: GR itself stacks its arrays
    after a transpose, so its `Nchains`/`Nsamples` names refer to swapped axes; pass
    `data.transpose(1, 0, 2)` to reproduce that quirk.
    """
    data = np.asarray(data, dtype=np.float64)
    Nchains, Nsamples, Npars = data.shape
    B_on_n = data.mean(axis=1).var(axis=0)        # variance of in-chain means
    W = data.var(axis=1).mean(axis=0)             # mean of in-chain variances
    sig2 = (Nsamples / (Nsamples - 1)) * W + B_on_n
    Vhat = sig2 + B_on_n / Nchains
    with np.errstate(divide="ignore", invalid="ignore"):
        Rhat = Vhat / W
    out = {"simple": Rhat.copy()}
    if advanced:
        m, n = float(Nchains), float(Nsamples)
        si2 = data.var(axis=1)
        xi_bar = data.mean(axis=1)
        xi2_bar = data.mean(axis=1) ** 2
        var_si2 = data.var(axis=1).var(axis=0)
        allmean = data.mean(axis=1).mean(axis=0)
        cov_term1 = np.array([np.cov(si2[:, i], xi2_bar[:, i])[0, 1] for i in range(Npars)])
        cov_term2 = np.array([-2 * allmean[i] * (np.cov(si2[:, i], xi_bar[:, i])[0, 1]) for i in range(Npars)])
        var_Vhat = (((n - 1) / n) ** 2 * 1.0 / m * var_si2
                    + ((m + 1) / m) ** 2 * 2.0 / (m - 1) * B_on_n ** 2
                    + 2.0 * (m + 1) * (n - 1) / (m * n ** 2) * n / m * (cov_term1 + cov_term2))
        with np.errstate(divide="ignore", invalid="ignore"):
            df = 2 * Vhat ** 2 / var_Vhat
            out["advanced"] = Rhat * df / (df - 2)
    return out


# stacking order of the 12 traced weights in GR:690, as columns of `bae_trace.weights` (MAIN:564-592 order:
# 0, 100, 10000, 5000, 2000, 3000, 4000, 5000, 6000, 7000, 8000, 9000, 11000)
GR_WEIGHT_COLUMNS = [0, 1, 4, 5, 6, 3, 8, 9, 10, 11, 2, 12]


def quantise_like_files(x, fmt="%1.2f"):
    """Values as the reference's offline script reads them back from the `%1.2f` text files (MAIN:641-677)."""
    x = np.asarray(x, dtype=np.float64)
    return np.asarray(np.char.mod(fmt, x.reshape(-1)), dtype=np.float64).reshape(x.shape)


def gelman_rubin_reference(weights, quantise=True):
    """R-hat exactly as `Gelman Reuben.py` computes it (GR:676-735). `weights`: [chains, samples, 12] post-burn-in
    traces in GR:690 order (`GR_WEIGHT_COLUMNS`). Two properties of the script are reproduced: it reads the
    2-decimal text files, and it transposes every [chains, samples] array before stacking, so its
    `Nchains, Nsamples, Npars = data.shape` binds Nchains = samples and Nsamples = chains (in-"chain" statistics are
    taken ACROSS chains at a fixed iteration). Pinned by tests/golden/gr_rhat.npz (oracle/make_gr_golden.py)."""
    w = np.asarray(weights, dtype=np.float64)
    if quantise:
        w = quantise_like_files(w)
    return gelman_rubin(w.transpose(1, 0, 2))
