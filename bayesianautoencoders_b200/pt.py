"""Host-side mirror of the reference's interface for the PT-Langevin path.

Same class names, method names, argument meaning and return shapes as
`Bayesian/bayes_autoenc_langevincheck.py` ("MAIN"): `Model` (MAIN:145-262), `ptReplica` (MAIN:265-823),
`ParallelTempering` (MAIN:827-1329). The arithmetic runs in the CUDA library behind
`include/bae_b200.h`; there is no torch model and no CPU fallback here.

Differences a caller must know (the reference reads these from module globals / the network):
  * shapes are passed explicitly (`dims4 = (in_shape, in_one, in_two, enc_shape)`), MAIN:75-102;
  * train / test matrices are passed explicitly (`data_load` downloads Madelon, MAIN:105-139);
  * state dicts are ordered dicts of numpy arrays keyed like the torch state_dict.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import numpy as np

from .results import GR_WEIGHT_COLUMNS, gelman_rubin, gelman_rubin_reference, write_replica_files  # noqa: F401
from .sampler import (SORTED_KEYS, DeviceSampler, Hyper, default_theta_init, param_offsets, param_shapes)

# torch's state_dict() iteration order for Model (MAIN:151-176)
STATE_DICT_KEYS = [
    "encode.0.weight", "encode.0.bias", "encode.2.weight", "encode.2.bias", "encode.4.weight", "encode.4.bias",
    "decode.0.weight", "decode.0.bias", "decode.0.running_mean", "decode.0.running_var", "decode.0.num_batches_tracked",
    "decode.1.weight", "decode.1.bias", "decode.3.weight", "decode.3.bias", "decode.5.weight", "decode.5.bias",
]
BUFFER_KEYS = ("decode.0.running_mean", "decode.0.running_var", "decode.0.num_batches_tracked")
ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8


class Model:
    """Mirror of `Model` (MAIN:145-262). Forward / backward run on the device through the parity-probe
    entry point `bae_eval`; the parameter plumbing (`getparameters`, `dictfromlist`, ...) is host logic."""

    def __init__(self, dims4, lrate=0.01, seed=0, device=0):
        self.dims4 = tuple(int(x) for x in dims4)
        self.lrate = lrate
        self.device = device
        self._offs, self.w_size = param_offsets(self.dims4)
        self._shapes = param_shapes(self.dims4)
        self._theta = default_theta_init(self.dims4, np.random.default_rng(seed)).astype(np.float64)
        self._m = np.zeros(self.w_size)
        self._v = np.zeros(self.w_size)
        self._t = 0                       # persistent Adam state (optimizer built once, MAIN:178)
        self._train_mask = np.ones(self.w_size, bool)
        for k in BUFFER_KEYS:
            o, n, _ = self._offs[k]
            self._train_mask[o:o + n] = False
        self._samplers = {}

    # -- parameter plumbing ---------------------------------------------------------------------
    def state_dict(self):
        return self._vec_to_dict(self._theta, STATE_DICT_KEYS, np.float64)

    def _vec_to_dict(self, vec, order, dtype):
        d = OrderedDict()
        for k in order:
            o, n, shp = self._offs[k]
            d[k] = np.array(vec[o:o + n], dtype=dtype).reshape(shp)
        return d

    def getparameters(self, w=None):
        """MAIN:228-241: flat float64 vector in sorted-key order."""
        dic = self.state_dict() if w is None else w
        return np.concatenate([np.asarray(dic[k], dtype=np.float64).reshape(-1) for k in sorted(dic.keys())])

    def dictfromlist(self, param):
        """MAIN:243-251: through torch.FloatTensor, i.e. every entry rounded to fp32; sorted-key order."""
        return self._vec_to_dict(np.asarray(param, dtype=np.float32), SORTED_KEYS, np.float32)

    def loadparameters(self, param):
        """MAIN:253-254 (`load_state_dict`): num_batches_tracked truncates toward zero (int64 buffer)."""
        for k, v in param.items():
            o, n, _ = self._offs[k]
            vals = np.asarray(v, dtype=np.float64).reshape(-1)
            if k == "decode.0.num_batches_tracked":
                vals = np.trunc(vals)
            self._theta[o:o + n] = vals

    def addnoiseandcopy(self, w, mea, std_dev, rng=None):
        """MAIN:256-262: every entry + fp32 N(mea, std_dev), loaded into the model, returned un-truncated."""
        rng = rng or np.random.default_rng()
        dic = OrderedDict()
        for name in w.keys():
            noise = rng.normal(mea, std_dev, np.shape(w[name])).astype(np.float32)
            dic[name] = np.asarray(w[name]) + noise
        self.loadparameters(dic)
        return dic

    # -- device forward / backward --------------------------------------------------------------
    def _sampler_for(self, data):
        data = np.ascontiguousarray(data, dtype=np.float32)
        key = (data.shape, hash(data.tobytes()[:4096]), float(data.sum()))
        s = self._samplers.get(key)
        if s is None:
            s = DeviceSampler(self.dims4, data, data, 1, 1, 1, [1.0], device=self.device)
            self._samplers[key] = s
        return s

    def _bn_side_effects(self, s, n_rows):
        # train-mode BatchNorm (MAIN:165) mutates the buffers of the live model on every forward (they never influence
        # outputs): running = 0.9 running + 0.1 batch statistic (unbiased variance), num_batches_tracked += 1. The state_dict
        # `langevin_gradient` returns carries them (MAIN:226).
        mu, var = s.eval_bn_stats(0)
        o, n, _ = self._offs["decode.0.running_mean"]
        self._theta[o:o + n] = 0.9 * self._theta[o:o + n] + 0.1 * mu.astype(np.float64)
        o, n, _ = self._offs["decode.0.running_var"]
        self._theta[o:o + n] = 0.9 * self._theta[o:o + n] + 0.1 * var.astype(np.float64) * (n_rows / (n_rows - 1.0))
        o, n, _ = self._offs["decode.0.num_batches_tracked"]
        self._theta[o] += 1

    def evaluate_proposal(self, data, w=None):
        """MAIN:188-206 -> numpy fx [rows, in_shape]."""
        if w is not None:
            self.loadparameters(w)
        s = self._sampler_for(data)
        r = s.eval(0, 0, w=self._theta.astype(np.float32), out=True)
        self._bn_side_effects(s, data.shape[0])
        return r["out"]

    def langevin_gradient(self, x, w, optimizer_mode=1):
        """MAIN:208-226: load, forward, MSELoss, backward (device), one persistent-state Adam step (host
        elementwise), returns a copy of the state_dict. optimizer_mode=2 is the stateless SGD(lr=0.1)."""
        self.loadparameters(w)
        s = self._sampler_for(x)
        g = s.eval(0, 0, w=self._theta.astype(np.float32), grads=True)["grads"].astype(np.float64)
        self._bn_side_effects(s, x.shape[0])
        self.last_grad = g
        msk = self._train_mask
        if optimizer_mode == 2:
            self._theta[msk] -= 0.1 * g[msk]
        else:
            self._t += 1
            self._m[msk] += (g[msk] - self._m[msk]) * (1 - ADAM_B1)
            self._v[msk] = self._v[msk] * ADAM_B2 + (1 - ADAM_B2) * g[msk] ** 2
            step = self.lrate / (1 - ADAM_B1 ** self._t)
            denom = np.sqrt(self._v[msk]) / math.sqrt(1 - ADAM_B2 ** self._t) + ADAM_EPS
            self._theta[msk] -= step * self._m[msk] / denom
        return self.state_dict()

    def close(self):
        for s in self._samplers.values():
            s.close()
        self._samplers = {}


class ptReplica:
    """Mirror of `ptReplica` (MAIN:265-823): one temperature's chain, resident on the device.

    `run()` executes the chain's `samples` iterations in blocks of `swap_interval` steps (one step-program
    launch per block) and performs the reference's rendezvous through the supplied queue / events
    after every block (MAIN:594-603). `multiprocessing.Process` is not involved: the chain state lives in HBM.
    """

    def __init__(self, use_langevin_gradients, learn_rate, w, minlim_param, maxlim_param, samples, traindata, testdata,
                 burn_in, temperature, swap_interval, path, parameter_queue, main_process, event, batch_size,
                 step_size, *, dims4, seed=0xBAE5EED, device=0, chain_id=0, theta_init=None, w0=None):
        self.samples = int(samples)
        self.temperature = float(temperature)
        self.adapttemp = float(temperature)
        self.swap_interval = int(swap_interval)
        self.path = path
        self.parameter_queue, self.signal_main, self.event = parameter_queue, main_process, event
        self.use_langevin_gradients = use_langevin_gradients
        self.step_size = step_size
        self.burn_in = burn_in
        self.traindata = np.ascontiguousarray(traindata, dtype=np.float32)
        self.testdata = np.ascontiguousarray(testdata, dtype=np.float32)
        self.dims4 = tuple(int(x) for x in dims4)
        self.l_prob = 0.9   # MAIN:286
        hyper = Hyper(lr=learn_rate, step_w=step_size, use_langevin=bool(use_langevin_gradients))
        self.sampler = DeviceSampler(self.dims4, self.traindata, self.testdata, 1, 1, self.samples, [self.temperature],
                                     swap_interval=self.swap_interval, seed=seed, device=device,
                                     chain_id_base=chain_id, hyper=hyper)
        rng = np.random.default_rng([seed & 0xFFFFFFFF, chain_id])
        self.w_size = self.sampler.w_size
        self._theta_init = default_theta_init(self.dims4, rng) if theta_init is None else theta_init
        self._w0 = rng.standard_normal(self.w_size) if w0 is None else w0

    @staticmethod
    def likelihood_func(cae, data, w, tau_sq, temp):
        """MAIN:304-329 -> [loglik / temp, fx, mse]; tau_sq multiplies the SSE as in the reference."""
        fx = cae.evaluate_proposal(data, w)
        n = data.shape[0] * data.shape[1]
        d = fx.astype(np.float64) - np.asarray(data, dtype=np.float64)
        mse = float(np.mean(np.square(d)))
        px = -(n / 2) * np.log(2 * math.pi * tau_sq)
        loss = px - 0.5 * tau_sq * float(np.sum(np.square(d)))
        return [loss / temp, fx, mse]

    def prior_likelihood(self, sigma_squared, w_list, tausq, nu_1, nu_2):
        """MAIN:331-336."""
        part1 = -1 * (len(w_list) / 2) * np.log(sigma_squared)
        part2 = 1 / (2 * sigma_squared) * float(np.sum(np.square(np.asarray(w_list, dtype=np.float64))))
        return part1 - part2 - (1 + nu_1) * np.log(tausq) - (nu_2 / tausq)

    def _post(self):
        st = self.sampler.get_state(0)
        vec = st["theta"] if st["w_is_alias"] else st["w"]
        return np.concatenate([vec.astype(np.float64), [st["eta"], st["likelihood"], st["adapttemp"], st["step"] - 1]]), st

    def run(self):
        s = self.sampler
        s.init_chain(0, self._theta_init, self._w0)
        done = 0
        while done < self.samples:
            n = min(self.swap_interval, self.samples - done)
            s.step(n)
            done += n
            if done % self.swap_interval == 0 and self.parameter_queue is not None:      # MAIN:594-603
                param, st = self._post()
                self.parameter_queue.put(param)
                self.signal_main.set()
                self.event.clear()
                self.event.wait()
                result = self.parameter_queue.get()
                w = np.asarray(result[:self.w_size], dtype=np.float32)          # dictfromlist: fp32
                theta = st["theta"].copy()
                mask = np.ones(self.w_size, bool)
                offs, _ = param_offsets(self.dims4)
                for k in BUFFER_KEYS:
                    o, nn, _ = offs[k]
                    mask[o:o + nn] = False
                theta[mask] = w[mask]     # trainables move into the live model (reloaded from `w` next step anyway)
                sc = {k: st[k] for k in st if k not in ("theta", "w", "m", "v")}
                sc["eta"] = float(result[self.w_size])
                sc["w_is_alias"] = 0
                s.set_state(0, theta=theta, w=w, scalars=sc)
        s.synchronize()
        if self.signal_main is not None:
            self.signal_main.set()                                                       # MAIN:619
        tr = s.traces(0, 0, self.samples)
        st = s.get_state(0, vectors=False)
        accept_ratio = st["num_accepted"] / (self.samples * 1.0) * 100
        if self.path is not None:
            write_replica_files(self.path, self.temperature, tr, accept_ratio)
        self.traces = tr
        w = tr["weights"]
        return (np.sqrt(tr["mse_train"]), np.sqrt(tr["mse_test"]), tr["mse_train"], tr["mse_test"], tr["sum_value"],
                w[:, 0], w[:, 1], w[:, 2], w[:, 3])                                     # MAIN:823


class ParallelTempering:
    """Mirror of `ParallelTempering` (MAIN:827-1329). All `num_chains` replicas live on one GPU handle;
    `run_chains()` is a sequence of step-program launches (5 iterations + exchange sweep each)."""

    def __init__(self, use_langevin_gradients, num_chains, maxtemp, NumSample, swap_interval, path, batch_size, bi,
                 step_size, *, dims4, traindata, testdata, lrate=0.01, burnin=0.5, seed=0xBAE5EED, device=0,
                 n_ensembles=1, chain_id_base=0):
        self.dims4 = tuple(int(x) for x in dims4)
        self.traindata = np.ascontiguousarray(traindata, dtype=np.float32)
        self.testdata = np.ascontiguousarray(testdata, dtype=np.float32)
        self.swap_interval = swap_interval
        self.path = path
        self.maxtemp = maxtemp
        self.num_swap = 0
        self.total_swap_proposals = 0
        self.num_chains = num_chains
        self.n_ensembles = n_ensembles
        self.chains = []
        self.temperatures = []
        self.NumSamples = int(NumSample / self.num_chains)                                # MAIN:846
        self.sub_sample_size = max(1, int(0.05 * self.NumSamples))
        self.geometric = True
        self.learn_rate = lrate
        self.use_langevin_gradients = use_langevin_gradients
        self.batch_size = batch_size
        self.masternumsample = NumSample
        self.burni = bi
        self.burnin = burnin
        self.step_size = step_size
        self.seed, self.device, self.chain_id_base = seed, device, chain_id_base
        _, self.num_param = param_offsets(self.dims4)
        self.sampler = None

    def default_beta_ladder(self, ndim, ntemps, Tmax):
        """MAIN:867-920, every branch. Inverse temperatures `logspace(0, -log10(Tmax), ntemps)`; when `Tmax` is
        None it is derived from `ntemps` and a per-dimension temperature step, when `ntemps` is None it is derived from
        `Tmax`; `Tmax = inf` replaces the coldest inverse temperature by 0. As in the reference the step table is built
        with `range(maxtemp)`, so `Tmax` must be an int (or None / inf raise TypeError there too) - the sampler itself
        always calls it with (2, num_chains, maxtemp) (MAIN:924)."""
        if type(ndim) != int or ndim < 1:
            raise ValueError('Invalid number of dimensions specified.')
        if ntemps is None and Tmax is None:
            raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
        if Tmax is not None and Tmax <= 1:
            raise ValueError('``Tmax`` must be greater than 1.')
        if ntemps is not None and (type(ntemps) != int or ntemps < 1):
            raise ValueError('Invalid number of temperatures specified.')
        # step table: Tmax * ntemps^(-k / (ntemps - 1)), k = 0 .. Tmax   (MAIN:878-885)
        steps = [Tmax]
        for _ in range(Tmax):
            steps.append(steps[-1] * (ntemps ** (-1 / (ntemps - 1))))
        steps = np.array(steps)
        if ndim > steps.shape[0]:
            tstep = 1.0 + 2.0 * np.sqrt(np.log(4.0)) / np.sqrt(ndim)      # large-dimension approximation (MAIN:887-890)
        else:
            tstep = steps[ndim - 1]
        append_inf = Tmax == np.inf
        if append_inf:
            Tmax = None
            ntemps = ntemps - 1
        if ntemps is not None:
            if Tmax is None:
                Tmax = tstep ** (ntemps - 1)
        else:
            if Tmax is None:
                raise ValueError('Must specify at least one of ``ntemps\'\' and finite ``Tmax``.')
            ntemps = int(np.log(Tmax) / np.log(tstep) + 2)
        betas = np.logspace(0, -np.log10(Tmax), ntemps)
        if append_inf:
            betas = np.concatenate((betas, [0]))
        return betas

    def assign_temperatures(self):
        """MAIN:922-935."""
        if self.geometric:
            betas = self.default_beta_ladder(2, ntemps=self.num_chains, Tmax=self.maxtemp)
            for i in range(self.num_chains):
                self.temperatures.append(np.inf if betas[i] == 0 else 1.0 / betas[i])
        else:
            rate = self.maxtemp / self.num_chains
            temp = 1
            for _ in range(self.num_chains):
                self.temperatures.append(temp)
                temp += rate

    def initialize_chains(self, burn_in, theta_inits=None, w0s=None):
        """MAIN:937-954: builds the ladder and the replicas (device-resident)."""
        self.burn_in = burn_in
        self.assign_temperatures()
        n = self.num_chains * self.n_ensembles
        hyper = Hyper(lr=self.learn_rate, step_w=self.step_size, use_langevin=bool(self.use_langevin_gradients))
        self.sampler = DeviceSampler(self.dims4, self.traindata, self.testdata, n, self.num_chains, self.NumSamples,
                                     np.tile(np.asarray(self.temperatures, dtype=np.float64), self.n_ensembles),
                                     swap_interval=self.swap_interval, seed=self.seed, device=self.device,
                                     chain_id_base=self.chain_id_base, hyper=hyper)
        rng = np.random.default_rng([self.seed & 0xFFFFFFFF, self.chain_id_base])
        for j in range(n):
            th = default_theta_init(self.dims4, rng) if theta_inits is None else theta_inits[j]
            w0 = rng.standard_normal(self.num_param) if w0s is None else w0s[j]
            self.sampler.init_chain(j, th, w0)
        self.chains = list(range(n))

    def swap_procedure(self, parameter_queue_1, parameter_queue_2, u=None):
        """MAIN:962-1011 on two posted parameter vectors (host scalar logic; the two likelihood
        evaluations go through the device parity probe). Returns (param1, param2, swapped)."""
        param1, param2 = parameter_queue_1.get(), parameter_queue_2.get()
        n = self.num_param
        s = self.sampler
        ntr = self.traindata.size

        def lhood(param, T_other):
            sse = s.eval(0, 0, w=np.asarray(param[:n], dtype=np.float32))["sse"]
            tau = np.exp(param[n])
            return (-(ntr / 2) * np.log(2 * math.pi * tau) - 0.5 * tau * sse) / T_other

        lhood1, T1 = param1[n + 1], param1[n + 2]
        lhood2, T2 = param2[n + 1], param2[n + 2]
        lhood12, lhood21 = lhood(param1, T2), lhood(param2, T1)
        with np.errstate(over="ignore"):
            swap_proposal = min(1, np.exp((lhood12 - lhood1) + (lhood21 - lhood2)))
        u = np.random.uniform(0, 1) if u is None else u
        self.total_swap_proposals += 1
        if u < swap_proposal:
            self.num_swap += 1
            param1, param2 = param2.copy(), param1.copy()
            param1[n + 1], param1[n + 2] = lhood21, T2
            param2[n + 1], param2[n + 2] = lhood12, T1
            return param1, param2, True
        return param1, param2, False

    def run_chains(self):
        """MAIN:1013-1090: all replicas advance `swap_interval` iterations per launch, then one sequential
        exchange sweep per ensemble on the device. Returns (mse_train, mse_test, acc_train, acc_test,
        accept_percentage_all_chains, swap_perc)."""
        s = self.sampler
        rounds = int(self.NumSamples / self.swap_interval)
        s.run(rounds * self.swap_interval)
        rest = self.NumSamples - rounds * self.swap_interval
        if rest:
            s.step(rest)
        s.synchronize()
        self.num_swap, self.total_swap_proposals = s.swap_counts()
        self._traces = [s.traces(j, 0, self.NumSamples) for j in self.chains]
        self._accept = []
        for j in self.chains:
            st = s.get_state(j, vectors=False)
            acc = st["num_accepted"] / (self.NumSamples * 1.0) * 100
            self._accept.append(acc)
            if self.path is not None:
                temp = self.temperatures[j % self.num_chains]
                sub = self.path if self.n_ensembles == 1 else os.path.join(self.path, f"ensemble_{j // self.num_chains}")
                os.makedirs(sub, exist_ok=True)
                write_replica_files(sub, temp, self._traces[j], acc)
        mse_train, mse_test, acc_train, acc_test, apal = self.show_results()
        total = self.total_swap_proposals if self.total_swap_proposals else 1
        swap_perc = self.num_swap * 100 / total
        return mse_train, mse_test, acc_train, acc_test, apal, swap_perc

    def show_results(self):
        """MAIN:1092-1329 without the plots: burn-in slicing of the per-chain traces."""
        b = int(self.NumSamples * self.burnin)
        mse_train = np.stack([t["mse_train"] for t in self._traces])[:, b:]
        mse_test = np.stack([t["mse_test"] for t in self._traces])[:, b:]
        acc_train = np.sqrt(np.stack([t["mse_train"] for t in self._traces]))[:, b:]
        acc_test = np.sqrt(np.stack([t["mse_test"] for t in self._traces]))[:, b:]
        return mse_train, mse_test, acc_train, acc_test, np.array([int(a) for a in self._accept], dtype=np.float64)

    def rhat(self, burnin=None, reference_compatible=False):
        """Gelman-Rubin R-hat over the 12 traced weights across the chains of this object (GR:690-735): the textbook
        statistic over [chains, samples], or - `reference_compatible` - what the reference's offline script prints
        (2-decimal traces, swapped axes; `results.gelman_rubin_reference`)."""
        b = int(self.NumSamples * (self.burnin if burnin is None else burnin))
        w = np.stack([t["weights"][b:, GR_WEIGHT_COLUMNS] for t in self._traces])
        if reference_compatible:
            return gelman_rubin_reference(w)
        return gelman_rubin(w.astype(np.float64))

    def close(self):
        if self.sampler is not None:
            self.sampler.close()
            self.sampler = None
