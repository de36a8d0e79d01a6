"""CPU oracle for the PT-Langevin hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

numpy restatement of the reference's parallel-tempering Langevin-MCMC sampler
over autoencoder weights, `/root/reference/Bayesian/bayes_autoenc_langevincheck.py`
("MAIN"). Every function cites the MAIN lines it follows. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker or the timed CPU baseline;
the product (`bayesianautoencoders_b200`) never does.

Parity status: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the pin is the reference itself: `oracle/make_golden.py`
imports MAIN unmodified in the build container, drives `Model`, `ptReplica.run`
and `ParallelTempering.swap_procedure` with injected randomness, and freezes the
results into `tests/golden/*.npz`; `tests/test_oracle_vs_golden.py` checks this
restatement against those files (and against the live reference when
/root/reference is mounted).

Third-party arithmetic: Linear/Sigmoid/BatchNorm1d/MSELoss/autograd/Adam come
from PyTorch (reference pins torch==1.5.1+cpu, requirements.txt:25; the goldens
were produced under torch 2.11.0). Their published algorithms are restated
below in closed form.

The reference semantics restated here include its quirks, because decisions
are compared step by step:
  * the sampled vector is the whole state_dict in sorted-key order, BatchNorm
    buffers included (MAIN:228-241);
  * BatchNorm1d runs in training mode in every forward (the model is never put
    in eval mode), so outputs use batch statistics and each forward mutates
    running_mean / running_var / num_batches_tracked of the live model;
  * the drift is one step of a persistent-state Adam on the MSE loss (MAIN:178,208-226);
  * on reject `w = old_w` aliases the live model tensors (MAIN:449,555), so the
    chain continues from the proposal; only bookkeeping scalars are gated;
  * tau_sq multiplies the SSE in the likelihood (MAIN:325-327).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------
# Parameter layout: sorted(state_dict.keys()) order                                MAIN:228-241
# ----------------------------------------------------------------------------------------------

SORTED_KEYS = [
    "decode.0.bias", "decode.0.num_batches_tracked", "decode.0.running_mean", "decode.0.running_var",
    "decode.0.weight", "decode.1.bias", "decode.1.weight", "decode.3.bias", "decode.3.weight",
    "decode.5.bias", "decode.5.weight", "encode.0.bias", "encode.0.weight", "encode.2.bias",
    "encode.2.weight", "encode.4.bias", "encode.4.weight",
]
BUFFER_KEYS = ("decode.0.num_batches_tracked", "decode.0.running_mean", "decode.0.running_var")


class Layout:
    """Offsets of the 17 state_dict tensors inside the flat vector of `getparameters` (MAIN:228-241).

    dims4 = (in_shape, in_one, in_two, enc_shape) = (d0, d1, d2, d3); layer widths are
    d0->d1->d2->d3->d2->d1->d0 (MAIN:151-176). Weights are torch `out x in`, row-major.
    """

    def __init__(self, dims4):
        d0, d1, d2, d3 = [int(x) for x in dims4]
        self.dims4 = (d0, d1, d2, d3)
        self.shapes = {
            "decode.0.bias": (d3,), "decode.0.num_batches_tracked": (), "decode.0.running_mean": (d3,),
            "decode.0.running_var": (d3,), "decode.0.weight": (d3,),
            "decode.1.bias": (d2,), "decode.1.weight": (d2, d3),
            "decode.3.bias": (d1,), "decode.3.weight": (d1, d2),
            "decode.5.bias": (d0,), "decode.5.weight": (d0, d1),
            "encode.0.bias": (d1,), "encode.0.weight": (d1, d0),
            "encode.2.bias": (d2,), "encode.2.weight": (d2, d1),
            "encode.4.bias": (d3,), "encode.4.weight": (d3, d2),
        }
        self.offsets: Dict[str, int] = {}
        off = 0
        for k in SORTED_KEYS:
            self.offsets[k] = off
            off += int(np.prod(self.shapes[k], dtype=np.int64)) if self.shapes[k] else 1
        self.w_size = off
        self.trainable_mask = np.ones(off, dtype=bool)
        for k in BUFFER_KEYS:
            self.trainable_mask[self.slice(k)] = False
        self.n_trainable = int(self.trainable_mask.sum())
        self.nbt_index = self.offsets["decode.0.num_batches_tracked"]

    def numel(self, k):
        return int(np.prod(self.shapes[k], dtype=np.int64)) if self.shapes[k] else 1

    def slice(self, k):
        return slice(self.offsets[k], self.offsets[k] + self.numel(k))

    def view(self, vec, k):
        return vec[self.slice(k)].reshape(self.shapes[k] if self.shapes[k] else (1,))

    def sum_S(self):
        d0, d1, d2, d3 = self.dims4
        return d0 * d1 + d1 * d2 + d2 * d3 + d3 * d2 + d2 * d1 + d1 * d0


# ----------------------------------------------------------------------------------------------
# Forward / backward / Adam                                                 MAIN:151-186, 208-226
# ----------------------------------------------------------------------------------------------

BN_EPS = 1e-5       # torch.nn.BatchNorm1d default eps
BN_MOMENTUM = 0.1   # torch.nn.BatchNorm1d default momentum
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8  # torch.optim.Adam defaults (MAIN:178)


def _sigmoid(x):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


def forward(L: Layout, theta: np.ndarray, X: np.ndarray, update_buffers: bool = True):
    """`Model.forward` (MAIN:183-186) with BatchNorm1d in TRAIN mode (MAIN:165; never `.eval()`).

    Returns (out, cache). When `update_buffers`, mutates theta's running_mean, running_var and
    num_batches_tracked exactly like torch's train-mode BatchNorm (momentum 0.1, unbiased variance
    for the running estimate).
    """
    v = lambda k: L.view(theta, k)
    W1, b1 = v("encode.0.weight"), v("encode.0.bias")
    W2, b2 = v("encode.2.weight"), v("encode.2.bias")
    W3, b3 = v("encode.4.weight"), v("encode.4.bias")
    gam, bet = v("decode.0.weight"), v("decode.0.bias")
    W4, b4 = v("decode.1.weight"), v("decode.1.bias")
    W5, b5 = v("decode.3.weight"), v("decode.3.bias")
    W6, b6 = v("decode.5.weight"), v("decode.5.bias")
    N = X.shape[0]
    h1 = _sigmoid(X @ W1.T + b1)
    h2 = _sigmoid(h1 @ W2.T + b2)
    z = h2 @ W3.T + b3
    mu = z.mean(axis=0)
    var = z.var(axis=0)  # biased, used for normalisation
    inv = 1.0 / np.sqrt(var + BN_EPS)
    zhat = (z - mu) * inv
    y = zhat * gam + bet
    h4 = _sigmoid(y @ W4.T + b4)
    h5 = _sigmoid(h4 @ W5.T + b5)
    out = h5 @ W6.T + b6
    if update_buffers:
        rm, rv = v("decode.0.running_mean"), v("decode.0.running_var")
        rm[:] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mu
        rv[:] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (N / (N - 1.0))
        theta[L.nbt_index] += 1
    cache = dict(X=X, h1=h1, h2=h2, zhat=zhat, inv=inv, y=y, h4=h4, h5=h5, out=out, mu=mu, var=var)
    return out, cache


def backward(L: Layout, theta: np.ndarray, cache) -> np.ndarray:
    """Gradient of `MSELoss(out, X)` (mean over all N*d0 elements, MAIN:150,223) w.r.t. every
    trainable entry, laid out like theta (buffer slots stay 0). Closed form of torch autograd for
    the network of MAIN:151-176; `encode.4.bias` is analytically zero (BatchNorm removes it) and
    is returned as exact 0 (the reference sees |g| ~ 1e-15 rounding noise there).
    """
    v = lambda k: L.view(theta, k)
    g = np.zeros_like(theta)
    gv = lambda k: L.view(g, k)
    X, h1, h2, zhat, inv, y, h4, h5, out = (cache[k] for k in
                                            ("X", "h1", "h2", "zhat", "inv", "y", "h4", "h5", "out"))
    N, d0 = X.shape
    dout = 2.0 * (out - X) / (N * d0)
    gv("decode.5.weight")[:] = dout.T @ h5
    gv("decode.5.bias")[:] = dout.sum(0)
    d5 = (dout @ v("decode.5.weight")) * h5 * (1 - h5)
    gv("decode.3.weight")[:] = d5.T @ h4
    gv("decode.3.bias")[:] = d5.sum(0)
    d4 = (d5 @ v("decode.3.weight")) * h4 * (1 - h4)
    gv("decode.1.weight")[:] = d4.T @ y
    gv("decode.1.bias")[:] = d4.sum(0)
    dy = d4 @ v("decode.1.weight")
    gv("decode.0.weight")[:] = (dy * zhat).sum(0)
    gv("decode.0.bias")[:] = dy.sum(0)
    dzhat = dy * v("decode.0.weight")
    dz = inv * (dzhat - dzhat.mean(0) - zhat * (dzhat * zhat).mean(0))
    gv("encode.4.weight")[:] = dz.T @ h2
    gv("encode.4.bias")[:] = 0.0
    d2 = (dz @ v("encode.4.weight")) * h2 * (1 - h2)
    gv("encode.2.weight")[:] = d2.T @ h1
    gv("encode.2.bias")[:] = d2.sum(0)
    d1 = (d2 @ v("encode.2.weight")) * h1 * (1 - h1)
    gv("encode.0.weight")[:] = d1.T @ X
    gv("encode.0.bias")[:] = d1.sum(0)
    return g


def adam_step(L: Layout, theta, g, m, v, t: int, lr: float) -> int:
    """One `torch.optim.Adam.step()` (defaults, single-tensor path) on the trainable entries;
    state (m, v, t) persists across calls (optimizer built once, MAIN:178). Returns the new t."""
    t = t + 1
    msk = L.trainable_mask
    m[msk] = m[msk] + (g[msk] - m[msk]) * (1 - ADAM_B1)          # exp_avg.lerp_(grad, 1-beta1)
    v[msk] = v[msk] * ADAM_B2 + (1 - ADAM_B2) * g[msk] * g[msk]
    bc1 = 1 - ADAM_B1 ** t
    bc2 = 1 - ADAM_B2 ** t
    step_size = lr / bc1
    denom = np.sqrt(v[msk]) / math.sqrt(bc2) + ADAM_EPS
    theta[msk] = theta[msk] - step_size * (m[msk] / denom)
    return t


# ----------------------------------------------------------------------------------------------
# Likelihood / prior                                                              MAIN:304-336
# ----------------------------------------------------------------------------------------------

def loglik_from_sse(sse: float, n: int, tau_sq: float, temp: float) -> float:
    """`likelihood_func` tail (MAIN:317-329): loss = -(n/2) log(2 pi tau) - 0.5 * tau * SSE; /temp.
    Note tau_sq MULTIPLIES the SSE, as written in the reference."""
    px = -(n / 2.0) * np.log(2 * math.pi * tau_sq)
    return float((px - 0.5 * tau_sq * sse) / temp)


def prior_likelihood(sigma_squared, w_list, tausq, nu_1, nu_2) -> float:
    """`ptReplica.prior_likelihood` (MAIN:331-336). `w_list` is the FULL flat vector incl. BN buffers."""
    part1 = -1 * (len(w_list) / 2) * np.log(sigma_squared)
    part2 = 1 / (2 * sigma_squared) * float(np.sum(np.square(np.asarray(w_list, dtype=np.float64))))
    return float(part1 - part2 - (1 + nu_1) * np.log(tausq) - (nu_2 / tausq))


def default_beta_ladder(ndim, ntemps, Tmax):
    """`ParallelTempering.default_beta_ladder` (MAIN:867-920) for finite Tmax and given ntemps:
    betas = logspace(0, -log10(Tmax), ntemps)  (MAIN:914)."""
    if type(ndim) != int or ndim < 1:
        raise ValueError('Invalid number of dimensions specified.')
    if ntemps is None and Tmax is None:
        raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
    if Tmax is not None and Tmax <= 1:
        raise ValueError('``Tmax`` must be greater than 1.')
    if ntemps is not None and (type(ntemps) != int or ntemps < 1):
        raise ValueError('Invalid number of temperatures specified.')
    return np.logspace(0, -np.log10(Tmax), ntemps)


def temperature_ladder(num_chains, maxtemp, geometric=True):
    """`assign_temperatures` (MAIN:922-935)."""
    if geometric:
        betas = default_beta_ladder(2, ntemps=num_chains, Tmax=maxtemp)
        return np.array([np.inf if b == 0 else 1.0 / b for b in betas])
    rate = maxtemp / num_chains
    return np.array([1 + i * rate for i in range(num_chains)], dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# One replica: the state machine of ptReplica.run()                               MAIN:348-603
# ----------------------------------------------------------------------------------------------

@dataclass
class StepRandomness:
    """The draws one iteration of MAIN:432-603 consumes, in consumption order."""
    lx: float            # np.random.uniform(0,1)           MAIN:447  (branch select)
    noise: np.ndarray    # fp32 N(0, step_w), flat sorted-key order, one per state_dict entry  MAIN:260
    eta_noise: float     # np.random.normal(0, step_eta)    MAIN:500
    u: float             # random.uniform(0,1)              MAIN:526  (MH test; log taken inside)


@dataclass
class Hyper:
    lr: float = 0.01          # MAIN:80,89,101
    step_w: float = 0.005     # MAIN:81,88,102
    step_eta: float = 0.2     # MAIN:387
    l_prob: float = 0.9       # MAIN:286
    sigma_squared: float = 25.0  # MAIN:396
    nu_1: float = 0.0         # MAIN:397
    nu_2: float = 3.0         # MAIN:398
    use_langevin: bool = True
    experiment_type: int = 2  # MAIN:72: 2 = Adam only; 1 = Adam while i <= pt, then `SGD(lr=0.1)` (MAIN:209-210, 454-470)
    mala: bool = False        # NOT in the reference: plain-gradient drift on the tempered log-posterior


@dataclass
class StepRecord:
    langevin: bool
    likelihood_proposal: float
    likelihood: float          # value BEFORE the accept update (what MAIN:521 records)
    prior_prop: float
    diff_prop: float
    sum_value: float
    log_u: float
    accepted: bool
    mse_train: float           # of the proposal
    mse_test: float
    sse_train: float
    sse_test: float
    eta_pro: float
    first: float = 0.0
    second: float = 0.0


class Replica:
    """Faithful restatement of one `ptReplica` (MAIN:265-603) over flat float64 vectors.

    State (all flat vectors are in sorted-key order, length w_size):
      theta   live model (`self.cae`) parameters AND BatchNorm buffers
      w       the "current" dict; `w_alias=True` means `w` IS the live model (MAIN:351, 449/555)
      w_prop  the last proposal dict (needed by the i == pt re-evaluation, MAIN:440-445)
      m, v, t persistent Adam state (MAIN:178)
      eta, likelihood, prior_current, adapttemp, temperature
    """

    def __init__(self, dims4, train, test, temperature, samples, hyper: Hyper = None,
                 pt_samples=0.5, dtype=np.float64):
        self.L = Layout(dims4)
        self.h = hyper or Hyper()
        self.dtype = dtype
        self.train = np.asarray(train, dtype=dtype)
        self.test = np.asarray(test, dtype=dtype)
        self.temperature = float(temperature)
        self.adapttemp = float(temperature)
        self.samples = int(samples)
        self.pt = int(pt_samples * samples)       # MAIN:429
        n = self.L.w_size
        self.theta = np.zeros(n, dtype)
        self.w: Optional[np.ndarray] = None
        self.w_alias = True                       # MAIN:351  w = cae.state_dict()
        self.w_f32 = False                        # True while `w` is a dict of torch.FloatTensor (after MAIN:602)
        self.w_prop = np.zeros(n, dtype)
        self.m = np.zeros(n, dtype)
        self.v = np.zeros(n, dtype)
        self.t = 0
        self.eta = 0.0
        self.likelihood = 0.0
        self.prior_current = 0.0
        self.i = 0
        self.init_count = 0
        self.num_accepted = 0
        self.langevin_count = 0
        self.mse_train_prev = 0.0   # mse_train[i-1]; numpy zeros + index -1 wrap at i == 0 gives 0 (MAIN:557)
        self.mse_test_prev = 0.0
        self.last_sse_train = 0.0   # SSE of the weights now held by `w` (== last proposal)
        self.grad_hook = None

    # -- dict <-> model plumbing -------------------------------------------------------------
    def _load(self, vec):
        """`load_state_dict` (MAIN:253-254): copy into the live model; the int64
        num_batches_tracked slot truncates toward zero."""
        self.theta[:] = vec
        self.theta[self.L.nbt_index] = np.trunc(vec[self.L.nbt_index])

    def current_w(self):
        return self.theta if self.w_alias else self.w

    def dictfromlist(self, param):
        """MAIN:243-251: through torch.FloatTensor -> every entry rounded to fp32."""
        return np.asarray(param, dtype=np.float32).astype(self.dtype)

    # -- operators --------------------------------------------------------------------------
    def evaluate_proposal(self, data, w=None):
        """MAIN:188-206."""
        if w is not None and w is not self.theta:
            self._load(w)
        out, _ = forward(self.L, self.theta, data)
        return out

    def langevin_gradient(self, x, w, optimizer_mode=1):
        """MAIN:208-226: load, forward (train-mode BN), MSE, backward, optimizer step; returns a deep copy of the
        state_dict (BN buffers as updated by the forward). optimizer_mode=1: the persistent Adam (MAIN:178);
        optimizer_mode=2: a fresh `torch.optim.SGD(lr=0.1)` (MAIN:209-210; no momentum, the Adam state is left alone)."""
        if w is not self.theta:
            self._load(w)
        _, cache = forward(self.L, self.theta, x)
        g = backward(self.L, self.theta, cache)
        if self.grad_hook is not None:      # test-only: lets the pin test inject the reference's ~1e-14
            g = self.grad_hook(g)           # rounding noise on the analytically-zero encode.4.bias gradient
        self.last_grad = g
        self.last_cache = cache
        msk = self.L.trainable_mask
        if self.h.mala:
            # extension, not in the reference: theta += (eps^2 / 2) grad[ loglik(theta; tau) / T + log prior(theta) ] with
            # grad loglik = -tau/2 * grad SSE = -tau/2 * n * grad MSE and grad log prior = -theta / sigma^2
            h2 = 0.5 * self.h.step_w ** 2
            n = x.shape[0] * x.shape[1]
            a = h2 * 0.5 * math.exp(self.eta) * n / self.adapttemp
            b = h2 / self.h.sigma_squared
            self.theta[msk] = self.theta[msk] - a * g[msk] - b * self.theta[msk]
        elif optimizer_mode == 2:
            self.theta[msk] = self.theta[msk] - 0.1 * g[msk]
        else:
            self.t = adam_step(self.L, self.theta, g, self.m, self.v, self.t, self.h.lr)
        return self.theta.copy()

    def addnoiseandcopy(self, w, noise, w_is_f32=False):
        """MAIN:256-262: every entry + fp32 noise; result loaded into the model; the returned dict keeps
        the un-truncated num_batches_tracked (int64 + float32 -> float32 in torch). When the incoming
        dict holds FloatTensors (it came from `dictfromlist`, MAIN:602) torch adds in fp32."""
        if w_is_f32:
            prop = (np.asarray(w, dtype=np.float32) + np.asarray(noise, dtype=np.float32)).astype(self.dtype)
        else:
            prop = np.asarray(w, dtype=self.dtype) + np.asarray(noise, dtype=np.float32).astype(self.dtype)
        k = self.L.nbt_index
        prop[k] = np.float32(np.float32(w[k]) + np.float32(noise[k]))
        self._load(prop)
        return prop

    def likelihood_func(self, data, w, tau_sq, temp):
        """MAIN:304-329 -> (loglik/temp, mse, sse)."""
        fx = self.evaluate_proposal(data, w)
        n = data.shape[0] * data.shape[1]
        d = fx - data
        sse = float(np.sum(np.square(d)))
        mse = float(np.mean(np.square(d)))
        return loglik_from_sse(sse, n, tau_sq, temp), mse, sse

    # -- prologue ---------------------------------------------------------------------------
    def initialise(self, theta_init, w0_randn):
        """`run()` prologue (MAIN:349-401). `theta_init`: the torch-default-initialised state_dict as a
        flat vector (rm=0, rv=1, nbt=0); `w0_randn`: the float64 `np.random.randn(w_size)` draw."""
        h = self.h
        self.theta[:] = np.asarray(theta_init, dtype=self.dtype)
        pred_train = self.evaluate_proposal(self.train, None)           # MAIN:381
        self.eta = float(np.log(np.var(pred_train - self.train)))       # MAIN:383
        tau_pro = np.exp(self.eta)
        w0_randn = np.asarray(w0_randn, dtype=np.float64)
        self.w_prop = self.dictfromlist(w0_randn)                       # MAIN:391
        self.prior_current = prior_likelihood(h.sigma_squared, w0_randn, tau_pro, h.nu_1, h.nu_2)  # MAIN:399
        self.likelihood, mse, sse = self.likelihood_func(self.train, self.w_prop, tau_pro, self.adapttemp)  # MAIN:401
        self.last_sse_train = sse
        self.w_alias = True

    # -- one iteration ------------------------------------------------------------------------
    def step(self, rnd: StepRandomness) -> StepRecord:
        h, L = self.h, self.L
        i = self.i
        if i < self.pt:
            self.adapttemp = self.temperature                                            # MAIN:437-438
        if i == self.pt and self.init_count == 0:                                        # MAIN:440-445
            self.adapttemp = 1.0
            self.likelihood, _, sse = self.likelihood_func(self.train, self.w_prop, 1.0, self.adapttemp)
            self.likelihood_func(self.test, self.w_prop, 1.0, self.adapttemp)
            self.init_count = 1
        first = second = 0.0
        prop_f32 = False
        langevin = bool(h.use_langevin and rnd.lx < h.l_prob)                            # MAIN:452
        if langevin:
            mode = 2 if (i > self.pt and h.experiment_type == 1) else 1                  # MAIN:454
            w_cur = self.current_w()
            w_gd = self.langevin_gradient(self.train, w_cur.copy(), mode)                # MAIN:473 / 455
            self.grad1 = self.last_grad
            w_proposal = self.addnoiseandcopy(w_gd, rnd.noise)                           # MAIN:474
            w_prop_gd = self.langevin_gradient(self.train, w_proposal.copy(), mode)      # MAIN:475 / 457
            self.grad2 = self.last_grad
            wc_delta = self.current_w() - w_prop_gd                                      # MAIN:476 (alias -> 0)
            wp_delta = w_proposal - w_gd                                                 # MAIN:477
            sigma_sq = h.step_w * h.step_w
            first = -0.5 * float(np.sum(wc_delta * wc_delta)) / sigma_sq                 # MAIN:481
            second = -0.5 * float(np.sum(wp_delta * wp_delta)) / sigma_sq                # MAIN:482
            diff_prop = first - second
            self.langevin_count += 1
        else:
            diff_prop = 0.0
            prop_f32 = (not self.w_alias) and self.w_f32
            w_proposal = self.addnoiseandcopy(self.current_w().copy(), rnd.noise, prop_f32)  # MAIN:492
        eta_pro = self.eta + rnd.eta_noise                                               # MAIN:500
        tau_pro = float(np.exp(eta_pro))
        lik_prop, msetrain, sse_train = self.likelihood_func(self.train, w_proposal.copy(), tau_pro, self.adapttemp)
        _, msetest, sse_test = self.likelihood_func(self.test, w_proposal.copy(), tau_pro, self.adapttemp)
        prior_prop = prior_likelihood(h.sigma_squared, w_proposal, tau_pro, h.nu_1, h.nu_2)  # MAIN:511
        diff_likelihood = lik_prop - self.likelihood
        diff_prior = prior_prop - self.prior_current
        sum_value = diff_likelihood + diff_prior + diff_prop                             # MAIN:524
        log_u = float(np.log(rnd.u))                                                     # MAIN:526
        rec = StepRecord(langevin, lik_prop, self.likelihood, prior_prop, diff_prop, sum_value, log_u,
                         False, msetrain, msetest, sse_train, sse_test, eta_pro, first, second)
        if log_u < sum_value:                                                            # MAIN:532-546
            rec.accepted = True
            self.num_accepted += 1
            self.likelihood = lik_prop
            self.prior_current = prior_prop
            self.w = w_proposal.copy()
            self.w_alias = False
            self.w_f32 = prop_f32
            self.eta = eta_pro
            self.mse_train_prev, self.mse_test_prev = msetrain, msetest
        else:                                                                            # MAIN:555
            self.w = None
            self.w_alias = True
            self.w_f32 = False
        self.w_prop = w_proposal
        self.last_sse_train = sse_train
        self.i += 1
        return rec

    def traced_weights(self):
        """MAIN:564-592: 13 entries of the LIVE model's flat vector."""
        idx = [0, 100, 10000, 5000, 2000, 3000, 4000, 5000, 6000, 7000, 8000, 9000, 11000]
        ll = self.theta
        out = np.zeros(13)
        for j, k in enumerate(idx):
            # MAIN guards with len(ll) >= idx (so idx == len would raise there; never hit at the 3 shapes)
            if k == 0 or len(ll) > k:
                out[j] = ll[k]
        return out

    # -- exchange -----------------------------------------------------------------------------
    def post(self):
        """MAIN:595-596: [vec(w), eta, likelihood, adapttemp, i]."""
        return np.concatenate([np.asarray(self.current_w(), dtype=np.float64),
                               [self.eta, self.likelihood, self.adapttemp, self.i - 1]])

    def take(self, result):
        """MAIN:601-603: w = dictfromlist(result[:w_size]) (fp32!), eta = result[w_size]."""
        n = self.L.w_size
        self.w = self.dictfromlist(result[:n])
        self.w_alias = False
        self.w_f32 = True
        self.eta = float(result[n])


def swap_procedure(dims4, train_f32, param1, param2, u: float):
    """`ParallelTempering.swap_procedure` (MAIN:962-1011). The main-process model is float32
    (MAIN:830) and the data is cast with `.astype(np.float32)` (MAIN:980-983), so the two extra
    forwards run in fp32; sums follow numpy's fp32 reductions. Returns (param1, param2, swapped,
    lhood12, lhood21)."""
    L = Layout(dims4)
    n_par = L.w_size
    X = np.asarray(train_f32, dtype=np.float32)
    n = X.shape[0] * X.shape[1]

    def lhood(param, T_other):
        th = np.asarray(param[:n_par], dtype=np.float32).copy()
        th[L.nbt_index] = np.trunc(th[L.nbt_index])
        out, _ = forward(L, th, X, update_buffers=False)
        tau_sq = np.exp(param[n_par])
        px = -(n / 2) * np.log(2 * math.pi * tau_sq)
        loss = px - 0.5 * tau_sq * np.sum(np.square(out - X))
        return float(loss / T_other)

    eta1, lhood1, T1 = param1[n_par], param1[n_par + 1], param1[n_par + 2]
    eta2, lhood2, T2 = param2[n_par], param2[n_par + 1], param2[n_par + 2]
    lhood12 = lhood(param1, T2)
    lhood21 = lhood(param2, T1)
    with np.errstate(over="ignore"):
        swap_proposal = min(1, np.exp((lhood12 - lhood1) + (lhood21 - lhood2)))   # MAIN:988 (inf, not OverflowError)
    if u < swap_proposal:
        p1, p2 = param2.copy(), param1.copy()
        p1[n_par + 1] = lhood21
        p1[n_par + 2] = T2
        p2[n_par + 1] = lhood12
        p2[n_par + 2] = T1
        return p1, p2, True, lhood12, lhood21
    return param1, param2, False, lhood12, lhood21


def swap_decision(L1_raw_other_T, lhood1, L2_raw_other_T, lhood2, u):
    """Scalar form of MAIN:988-995 given lhood12, lhood21."""
    with np.errstate(over="ignore"):
        p = min(1, np.exp((L1_raw_other_T - lhood1) + (L2_raw_other_T - lhood2)))
    return bool(u < p)


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 (host restatement of the device generator, for bit-exact stream checks)
# ----------------------------------------------------------------------------------------------

PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32(counter, key, rounds=10):
    """Philox4x32-10 (Salmon et al., SC'11). counter: uint32[...,4], key: uint32[...,2] -> uint32[...,4]."""
    c = np.array(counter, dtype=np.uint64).copy()
    k = np.array(key, dtype=np.uint64).copy()
    c = np.broadcast_to(c, np.broadcast_shapes(c.shape[:-1], k.shape[:-1]) + (4,)).copy()
    k = np.broadcast_to(k, c.shape[:-1] + (2,)).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = np.uint64(PHILOX_M0) * c[..., 0]
        p1 = np.uint64(PHILOX_M1) * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = (hi1 ^ c[..., 1] ^ k[..., 0]) & mask
        n1 = lo1
        n2 = (hi0 ^ c[..., 3] ^ k[..., 1]) & mask
        n3 = lo0
        c = np.stack([n0, n1, n2, n3], axis=-1)
        k[..., 0] = (k[..., 0] + np.uint64(PHILOX_W0)) & mask
        k[..., 1] = (k[..., 1] + np.uint64(PHILOX_W1)) & mask
    return c.astype(np.uint32)


def u32_to_unit_open(x):
    """(x + 0.5) * 2^-32 in fp32-safe form used on the device: ((x >> 8) + 0.5) * 2^-24  in (0,1)."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def box_muller(u1, u2):
    r = np.sqrt(-2.0 * np.log(u1.astype(np.float64)))
    a = 2.0 * math.pi * u2.astype(np.float64)
    return (r * np.cos(a)), (r * np.sin(a))


# ----------------------------------------------------------------------------------------------
# The PT driver loop                                                            MAIN:1013-1072
# ----------------------------------------------------------------------------------------------

def run_pt(dims4, train, test, temperatures, samples_per_chain, swap_interval, theta_inits, rnd,
           swap_u=None, do_swaps=True, hyper: Hyper = None, on_step=None, grad_hooks=None):
    """`ParallelTempering.run_chains` (MAIN:1013-1072) over `Replica` objects, with the SEQUENTIAL
    adjacent-pair sweep 0-1, 1-2, ... of MAIN:1059-1065. `rnd[j]` holds replica j's injected draws
    (dict of w0_randn, lx, noise, eta_noise, u); `swap_u` the swap uniforms in sweep order.
    Returns (replicas, records[j][i], swap_log)."""
    n = len(temperatures)
    reps = []
    for j in range(n):
        r = Replica(dims4, train, test, temperatures[j], samples_per_chain, hyper)
        r.initialise(theta_inits[j], rnd[j]["w0_randn"])
        if grad_hooks is not None:
            r.grad_hook = grad_hooks[j]
        reps.append(r)
    records = [[] for _ in range(n)]
    traces = [[] for _ in range(n)]
    swap_log = []
    swap_u = list(swap_u or [])
    train32 = np.asarray(train, dtype=np.float32)
    for i in range(samples_per_chain):
        for j, r in enumerate(reps):
            d = rnd[j]
            rec = r.step(StepRandomness(d["lx"][i], d["noise"][i], d["eta_noise"][i], d["u"][i]))
            records[j].append(rec)
            traces[j].append(r.traced_weights())
            if on_step is not None:
                on_step(j, i, r, rec)
        if (i + 1) % swap_interval == 0 and (i + 1) // swap_interval <= samples_per_chain // swap_interval:
            params = [r.post() for r in reps]
            if do_swaps:
                for index in range(n - 1):
                    p1, p2, swapped, _, _ = swap_procedure(dims4, train32, params[index], params[index + 1],
                                                           swap_u.pop(0))
                    params[index], params[index + 1] = p1, p2
                    swap_log.append(((i + 1) // swap_interval - 1, index, swapped))
            for r, p in zip(reps, params):
                r.take(p)
    return reps, records, swap_log
