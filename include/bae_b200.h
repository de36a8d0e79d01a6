/*
 * bae_b200.h -- C-ABI of the B200-native parallel-tempering Langevin sampler over autoencoder weights.
 *
 * The reference (sydney-machine-learning/BayesianAutoencoders) has no FFI; its boundary for this path
 * is the Python method surface of Bayesian/bayes_autoenc_langevincheck.py ("MAIN"):
 *   Model.evaluate_proposal / langevin_gradient / addnoiseandcopy / getparameters   MAIN:188-262
 *   ptReplica.run / likelihood_func / prior_likelihood                              MAIN:304-336, 348-603
 *   ParallelTempering.swap_procedure / run_chains / default_beta_ladder             MAIN:867-1090
 * The entry points below are what a ctypes binding of that surface calls (INTEGRATION.md shows the
 * binding). Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative bae_status, with a per-thread message available from bae_last_error(); the caller owns
 * every host buffer, the library owns every device buffer; calls on one handle must be serialised by
 * the caller; one handle drives one GPU.
 *
 * Flat parameter vectors crossing this boundary are in the reference's `getparameters` order
 * (sorted state_dict keys, BatchNorm buffers included, MAIN:228-241), length bae_w_size().
 */
#ifndef BAE_B200_H
#define BAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bae_handle bae_handle;

enum bae_status {
  BAE_OK = 0,
  BAE_ERR_INVALID = -1,   /* bad argument */
  BAE_ERR_CUDA = -2,      /* CUDA runtime/driver error (message has the CUDA string) */
  BAE_ERR_NO_DEVICE = -3, /* no usable sm_100 device: there is NO CPU fallback */
  BAE_ERR_STATE = -4      /* call sequence error (e.g. step before data upload) */
};

/* Hyper-parameters and shapes. Defaults in comments are the reference's values. */
typedef struct bae_config {
  int32_t dims[4];        /* in_shape, in_one, in_two, enc_shape                 MAIN:76-79,84-87,97-100 */
  int32_t n_train;        /* rows of the training matrix                          MAIN:833 */
  int32_t n_test;         /* rows of the test matrix                              MAIN:834 */
  int32_t n_chains;       /* chains resident on this handle (multiple of n_rungs) */
  int32_t n_rungs;        /* ladder length = chains per ensemble (num_chains)     MAIN:64 */
  int32_t chain_id_base;  /* global id of local chain 0 (keys the Philox streams; multi-GPU sharding) */
  int32_t samples;        /* iterations per chain (NumSamples)                    MAIN:846 */
  int32_t pt_switch_step; /* int(pt_samples * samples): tempered -> canonical     MAIN:429,437-445 */
  int32_t swap_interval;  /* 5                                                    MAIN:69 */
  int32_t use_langevin;   /* 1                                                    MAIN:61 */
  float lr;               /* Adam lr 0.01                                         MAIN:80,89,101,178 */
  float step_w;           /* proposal sigma 0.005                                 MAIN:81,88,102 */
  float step_eta;         /* 0.2                                                  MAIN:387 */
  float l_prob;           /* 0.9                                                  MAIN:286 */
  float sigma_squared;    /* 25                                                   MAIN:396 */
  float nu_1;             /* 0                                                    MAIN:397 */
  float nu_2;             /* 3                                                    MAIN:398 */
  uint64_t seed;          /* Philox4x32-10 key */
  int32_t gemm_path;      /* 0 = auto, 1 = FP32 CUDA-core tiles only, 2 = tcgen05 3xTF32 where widths allow */
  int32_t drift_mode;     /* 0 = Adam drift in every Langevin iteration (experiment_type = 2, the reference's setting, MAIN:72)
                           * 1 = experiment_type = 1: Adam while i <= pt, then a stateless SGD(lr = 0.1) step (MAIN:209-210, 454-470)
                           * 2 = plain-gradient MALA on the tempered log-posterior (not in the reference):
                           *     drift = (step_w^2 / 2) * grad[ loglik(w; tau = exp(eta)) / adapttemp + log prior(w) ] */
  int32_t act_group;      /* chains whose activation / delta buffers are resident at a time (BASELINE.json configs[4]: 100k rows x
                           * 2048 features = 7.7 GB of activations per chain). 0 = automatic: all chains if they fit in 70% of the
                           * free device memory, else as many as fit. The step program then runs group after group. */
  int32_t reserved;
} bae_config;

/* Per-step record, one per (chain, iteration): what MAIN:520-592 stores into its trace arrays. */
typedef struct bae_trace {
  double likelihood_proposal; /* MAIN:520 */
  double likelihood;          /* MAIN:521 (value before the accept update) */
  double sum_value;           /* MAIN:524,530 */
  double log_u;               /* MAIN:526-527 */
  double prior_proposal;      /* MAIN:511 */
  double diff_prop;           /* MAIN:485,491 */
  double mse_train;           /* MAIN:542/557 (accepted value carried forward) */
  double mse_test;            /* MAIN:543/558 */
  double mse_train_proposal;  /* msetrain of this step's proposal, MAIN:505 */
  double mse_test_proposal;   /* MAIN:508 */
  double eta;                 /* eta after the step */
  float weights[13];          /* MAIN:564-592: live-model entries 0,100,10000,5000,2000,...,11000 */
  int32_t accepted;           /* MAIN:532 */
  int32_t langevin;           /* MAIN:452 */
  int32_t step;
} bae_trace;

/* Scalars of one chain (checkpoint / parity injection). */
typedef struct bae_chain_scalars {
  double eta, likelihood, prior_current, last_sse_train, mse_train, mse_test, temperature, adapttemp;
  int32_t adam_t, w_is_alias, init_count, num_accepted, langevin_count, step;
  double prop_sse_train; /* SSE of this chain's own last proposal (used by the i == pt re-evaluation, MAIN:442) */
} bae_chain_scalars;

const char* bae_last_error(void);
int bae_version(void);

/* Lifecycle. `device` is a CUDA ordinal. Fails with BAE_ERR_NO_DEVICE when no sm_100 GPU is present. */
int bae_create(const bae_config* cfg, int device, bae_handle** out);
int bae_destroy(bae_handle* h);
int bae_w_size(const bae_handle* h);       /* len(getparameters())            MAIN:352 */
int bae_n_trainable(const bae_handle* h);

/* Row-major float32 matrices, n_train x dims[0] and n_test x dims[0] (ParallelTempering.traindata/testdata, MAIN:833-834). */
int bae_upload_data(bae_handle* h, const float* train, const float* test);
/* Per-chain temperatures (assign_temperatures, MAIN:922-935), length n_chains. */
int bae_set_ladder(bae_handle* h, const double* temperatures);

/* run() prologue for one chain (MAIN:349-401): theta_init = torch-default-initialised state_dict (flat),
 * w0 = the float64 np.random.randn(w_size) start vector. Computes eta0, prior_current, likelihood. */
int bae_init_chain(bae_handle* h, int chain, const float* theta_init, const double* w0);

/* Full state of one chain. Any pointer may be NULL. theta/w/m/v: w_size floats, reference order.
 * `w` is meaningful only when scalars->w_is_alias == 0 (otherwise w IS theta, MAIN:449/555). */
int bae_get_state(bae_handle* h, int chain, float* theta, float* w, float* m, float* v, bae_chain_scalars* s);
int bae_set_state(bae_handle* h, int chain, const float* theta, const float* w, const float* m, const float* v,
                  const bae_chain_scalars* s);

/* n_steps iterations of MAIN:432-592 for every chain on the handle (forward/backward, Adam drift, noise, likelihood,
 * prior, reverse-proposal density, MH test): ONE device-side step program per call, no host work inside it - a replayed
 * CUDA graph of per-phase kernels on the tcgen05 path, one persistent cooperative kernel on the FP32 path. */
int bae_step(bae_handle* h, int n_steps);
/* One exchange round (MAIN:594-603 + 962-1011 + 1059-1065) for every ensemble resident on the handle. */
int bae_swap_local(bae_handle* h);
/* Steps with an exchange round after every swap_interval-th iteration, like run_chains (MAIN:1035-1072). */
int bae_run(bae_handle* h, int n_steps);
int bae_synchronize(bae_handle* h);

/* Parity probe (Model.evaluate_proposal + MSELoss.backward, MAIN:188-226): forward (and backward when
 * grads != NULL) of `w` (w_size floats, or NULL for the chain's live model) on the train (which=0) or test
 * (which=1) matrix. Does not disturb the chain state. out: n x dims[0] or NULL. */
int bae_eval(bae_handle* h, int chain, int which, const float* w, double* sse, double* mse, float* grads, float* out);

/* Batch statistics of the BatchNorm layer seen by the most recent bae_eval call on this handle (biased variance, as used
 * for normalisation): what a train-mode forward folds into running_mean / running_var (momentum 0.1, MAIN:165, 226).
 * mean, var: dims[3] floats each. */
int bae_eval_bn_stats(bae_handle* h, int which, float* mean, float* var);

/* Posterior-predictive pass, batched over every resident chain (MAIN:709-818 runs `cae.encode(X)` / `cae.forward(X)` on the
 * final weights of one chain): forward of the training (which = 0) or test (which = 1) matrix through each chain's live
 * model, BatchNorm in training mode on that batch as everywhere in the reference. Any output may be NULL.
 *   features: [n_chains][rows][dims[3]]  encoder output `encode(x)`, i.e. the pre-BatchNorm bottleneck      MAIN:714, 787
 *   recon:    [n_chains][rows][dims[0]]  reconstruction `forward(x)`                                       MAIN:752
 *   sse:      [n_chains]                 sum of squared reconstruction errors
 * Chain state (weights, BatchNorm buffers, Adam moments, scalars, traces) is not disturbed. */
int bae_forward_all(bae_handle* h, int which, float* features, float* recon, double* sse);

/* The draws iteration `step` of global chain id (chain_id_base + chain) consumes: noise = w_size fp32
 * normals * step_w in reference order (MAIN:260), scalars = {lx, eta_noise, u} (MAIN:447,500,526). */
int bae_debug_draws(bae_handle* h, int chain, int step, float* noise, double* scalars3);
/* Swap uniform of (ensemble, round, pair) (MAIN:994). */
int bae_debug_swap_uniform(bae_handle* h, int ensemble, int round, int pair, double* u);

/* Trace ring: records of iterations [first_step, first_step+n) of one chain. */
int bae_read_traces(bae_handle* h, int chain, int first_step, int n, bae_trace* out);
/* the same records of every resident chain, out[n_chains][n]: one copy, one synchronisation */
int bae_read_traces_all(bae_handle* h, int first_step, int n, bae_trace* out);
int bae_swap_counts(bae_handle* h, int64_t* num_swap, int64_t* total_swap_proposals);
/* Result of the most recent exchange sweep on this handle (MAIN:1059-1065): src[p] = local chain whose configuration
 * ended at ladder position p (length n_chains). Pair (p, p+1) of an ensemble swapped iff src[p] == p + 1 (`swapped`,
 * MAIN:995-1011). */
int bae_last_sweep(bae_handle* h, int32_t* src_out);

/* ---- Ladder sharded over GPUs, one process per GPU (BASELINE.json configs[3]) -------------------------
 * A handle created with n_chains = E * L, n_rungs = L and the SAME chain_id_base on every rank holds rungs
 * [rank * L, rank * L + L) of each of E ladders of R = L * world rungs. bae_shard_config declares that (it re-keys the
 * Philox streams to the ladder positions, so a sharded run draws exactly what a one-GPU run of the same ladders draws);
 * afterwards bae_run is refused: advance with bae_step(swap_interval), then call bae_swap_exchange on every rank.
 *
 * bae_swap_exchange = one exchange round (MAIN:594-603, 962-1011, 1059-1065), stream-ordered on the handle's stream:
 *   ncclAllGather of 40 B per chain (eta, stored likelihood, adapttemp, SSE of w, alias flag); the sequential sweep as a
 *   device kernel on every rank (identical permutation everywhere: same doubles, same Philox uniforms); ONE small
 *   device-to-host copy of the permutation (the only host synchronisation); grouped ncclSend / ncclRecv of the weight
 *   vectors whose destination is on another rank, straight from the theta slabs into the receiver's staging slabs (no
 *   packing); one gather and one scatter kernel for everything else. `nccl_comm` is an ncclComm_t created by the caller
 *   over exactly `world` ranks (never created here); NCCL is resolved from the already-loaded libnccl.so.2 at run time. */
int bae_shard_config(bae_handle* h, int rank, int world);
int bae_swap_exchange(bae_handle* h, void* nccl_comm);
/* HOST function: the transfers rank `rank` must post for the permutation `src` ([E][R]: rung whose configuration ends at
 * rung p): out_ops[4 * k] = (kind 0 send / 1 recv, peer rank, ensemble, rung = source rung of a send / destination rung of a
 * receive), in the order both sides issue them. Returns the number of operations (<= max_ops) or a negative bae_status. */
int bae_plan_exchange(const int32_t* src, int n_ensembles, int n_rungs_total, int world, int rank, int32_t* out_ops, int max_ops);
int64_t bae_exchange_vectors_sent(const bae_handle* h);   /* weight vectors this rank has sent over NCCL so far */

/* (older, host-driven variant of the same exchange, kept for callers that move the vectors themselves)
 * Multi-GPU ladder sharding (one process per GPU): export / import the exchange view of a chain
 * ([vec(w), eta, likelihood, adapttemp, last_sse], MAIN:595-596) as DEVICE buffers usable with NCCL. */
int bae_exchange_table(bae_handle* h, double* out_nchains_x4);   /* eta, likelihood, adapttemp, SSE of w for every chain (MAIN:594-597) */
int bae_exchange_pack(bae_handle* h, int chain, float* dev_w /*w_size*/, double* host_scalars4);
int bae_exchange_unpack(bae_handle* h, int chain, const float* dev_w, const double* host_scalars2 /*eta,last_sse*/);
/* Marks `w` as a detached copy for every chain, as MAIN:602 does after each exchange round. */
int bae_exchange_local(bae_handle* h, const int* src_chain /*[n_chains]*/);   /* every same-GPU move of a round + detach (MAIN:602), one gather/scatter pass */
int bae_host_sweep(const double* table_nrungs_x4, int n_rungs, double n_elems, const float* uniforms, int* src_out, int* n_swapped);   /* HOST function: the sequential swap sweep of one ensemble (MAIN:962-1011, 1059-1065) */
int bae_exchange_finish(bae_handle* h);

/* Debug probe: raw copy of a per-chain device array ("H1","H2","Z","Y","H4","H5","DOUT","D5","D4","DY","D2","D1",
 * "grad","theta","theta_hi","theta_lo") in the padded device layout; chain may be n_chains (probe scratch slot). */
int bae_debug_buffer(bae_handle* h, const char* name, int chain, float* out, int64_t n_floats, int* ld, int lo_part);

int bae_debug_trace(bae_handle* h, long long* out4096, int clear);   /* CTA-0 role timeline, BAE_TC_TRACE=1 */
int bae_debug_phase_times(bae_handle* h, double* ms64, int* n64);      /* per-phase milliseconds, BAE_PHASE_TIMES=1 */
int bae_debug_profiler(int on);                                        /* cudaProfilerStart (1) / Stop (0) */

/* Introspection for the benchmark: kernels launched so far by this handle, last step-kernel time (ms). */
int64_t bae_launch_count(const bae_handle* h);
int bae_last_step_ms(bae_handle* h, float* ms);
int bae_gemm_path(const bae_handle* h);  /* 1 = FP32 CUDA cores, 2 = tcgen05 tensor cores */
int bae_mma_kind(const bae_handle* h);   /* 0 = FFMA (gemm path 1), 1 = 3xTF32 split (kind::tf32), 2 = 3xFP16 split (kind::f16, scaled operands) */
int bae_step_kernel(const bae_handle* h); /* which step program runs bae_step / bae_run: 1 = persistent FP32 tile kernel, 2 = tcgen05
                                          * phase program, 3 = narrow-width cluster kernel (Swiss Roll widths, csrc/bae_narrow.cu) */
void* bae_stream(bae_handle* h);         /* cudaStream_t the handle launches on */

#ifdef __cplusplus
}
#endif
#endif /* BAE_B200_H */
