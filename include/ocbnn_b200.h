/*
 * ocbnn_b200.h — C-ABI of the B200-native OC-BNN posterior-inference engine.
 *
 * The reference (dtak/ocbnn-public) is pure Python/PyTorch and has no FFI; the
 * "plugin interface" this library sits behind is the inference-mixin contract of
 * bnn/inference.py:5-10 plus the model methods those mixins call
 * (bnn/base.py:89-133).  Each entry point below names the reference code it
 * replaces.  The host package (ocbnn-public_b200/ocbnn_b200) binds these with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types.
 *   - every call returns 0 on success, otherwise an OCBNN_E* code (or a
 *     cudaError_t offset by OCBNN_ECUDA); ocbnn_last_error() gives the text.
 *   - "device" pointers are CUDA device pointers on the current device,
 *     contiguous, fp32 unless stated.  "host" pointers are ordinary memory.
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as
 *     void*; NULL = legacy default stream).  Calls do not synchronise unless
 *     their name ends in _host.
 *   - the caller owns every array; the library owns only ocbnn_problem.
 *   - one host thread per problem handle.
 *
 * Packed weight layout (bnn/base.py:72-77): for each layer l in order,
 * W_l (s_{l-1} x s_l, row-major) then b_l (s_l).
 */
#ifndef OCBNN_B200_H
#define OCBNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OCBNN_API __attribute__((visibility("default")))
#else
#define OCBNN_API
#endif

#define OCBNN_ABI_VERSION 3 /* 3: ocbnn_predict, ocbnn_violation_flags, with_data = 2 (likelihood only) in ocbnn_logpost_grad */
#define OCBNN_MAX_LAYERS 8

/* error codes */
#define OCBNN_OK 0
#define OCBNN_EINVAL 1       /* bad argument */
#define OCBNN_EUNSUPPORTED 2 /* shape/feature not supported by this build */
#define OCBNN_ENOMEM 3
#define OCBNN_ENODEVICE 4    /* no CUDA device / wrong architecture */
#define OCBNN_ECUDA 1000     /* + cudaError_t */

/* activation (bnn/base.py:79-87) */
#define OCBNN_ACT_IDENTITY 0
#define OCBNN_ACT_RBF 1
#define OCBNN_ACT_RELU 2

/* per-point head kinds (likelihoods and output-constrained prior terms) */
#define OCBNN_HEAD_LIK_GAUSS 0     /* RegressorMixin.log_likelihood        base.py:252-265 */
#define OCBNN_HEAD_NEG_EXP 1       /* negative_exponential_cocp            base.py:360-388 */
#define OCBNN_HEAD_POS_GAUSS 2     /* positive_gaussian_cocp               base.py:344-358 */
#define OCBNN_HEAD_DIRICHLET 3     /* positive_dirichlet_cocp              base.py:498-511 */
#define OCBNN_HEAD_LIK_SOFTMAX 4   /* ClassifierMixin.log_likelihood       base.py:432-444 */
#define OCBNN_HEAD_LIK_BERNOULLI 5 /* BinaryClassifierMixin.log_likelihood base.py:524-542 */

typedef struct ocbnn_problem ocbnn_problem; /* opaque */

/*
 * Problem description ("compiled" model + prior + data).  All pointers are HOST
 * pointers; ocbnn_problem_create copies everything to the device.
 *
 * log posterior(w) = prior_const
 *                  + sum_i -(w_i - mean_i)^2 / (2 std_i^2)                     (isotropic or AOCP prior)
 *                  + sum_c cweight_c * head_c(f(cx_c; w))                      (COCP constraint points)
 *                  + [with_data] (N/|B|) * sum_{n in B} lik(f(x_n; w), y_n) + lik_const
 * B = {offset, offset+stride, ...} reproduces torch.arange(t % nbatches, N, nbatches).
 */
typedef struct {
    int32_t abi_version;                     /* OCBNN_ABI_VERSION */
    int32_t n_layers;                        /* weight layers K (>=1) */
    int32_t layer_sizes[OCBNN_MAX_LAYERS + 1]; /* s_0..s_K */
    int32_t activation;                      /* OCBNN_ACT_* */
    int32_t lik_head;                        /* OCBNN_HEAD_LIK_* */
    float lik_scale;                         /* gauss: 1/(2 var) ; unused otherwise */
    double lik_const;                        /* data-independent constant of the full-data log-likelihood */

    /* prior over weights: scalar, or per-weight vectors (AOCP, base.py:111-112) when the pointers are non-NULL */
    float prior_mean, prior_std;
    const float *prior_mean_vec, *prior_std_vec; /* D each or NULL */
    double prior_const;                      /* all data-independent constants of log_prior */

    /* training data */
    int64_t n_train;
    const float *x_train;                    /* n_train x s_0 */
    const float *y_train;                    /* n_train x y_width: targets, or class labels stored as float */
    int32_t y_width;

    /* constraint points of the output-constrained priors (already sampled by add_*_constraint) */
    int32_t n_cpoints;
    const float *cx;                         /* n_cpoints x s_0 */
    const int32_t *ckind;                    /* n_cpoints, OCBNN_HEAD_* */
    const float *cweight;                    /* n_cpoints */
    int32_t cparam_stride;
    const float *cparam;                     /* n_cpoints x cparam_stride:
                                                NEG_EXP  : ngaps, a_0, b_0, a_1, b_1, ... (penalise a < f < b; +-inf allowed)
                                                POS_GAUSS: K, y_0..y_{K-1}
                                                DIRICHLET: conc_k - 1, k < s_K */
    float neg_tau0, neg_tau1;                /* cocp_expo_tau */
    float pos_inv_var;                       /* 1 / cocp_gaussian_sigma_c (the reference passes sigma_c as a VARIANCE) */
} ocbnn_desc;

OCBNN_API const char *ocbnn_last_error(void);
OCBNN_API int ocbnn_abi_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py: gpu_launches) */
OCBNN_API int64_t ocbnn_launch_count(void);

OCBNN_API int ocbnn_problem_create(const ocbnn_desc *desc, ocbnn_problem **out);
OCBNN_API int ocbnn_problem_destroy(ocbnn_problem *p);
OCBNN_API int ocbnn_problem_nweights(const ocbnn_problem *p, int64_t *out_d);

/* K1 — log posterior / log prior value and gradient for M packed weight vectors.
 * Replaces forward + log_prior + log_likelihood + autograd .backward() (base.py:89-133,252-265,344-388,432-444,498-511,528-542).
 * with_data: 0 log prior, 1 log posterior, 2 log likelihood alone (log_likelihood, base.py:252-265,432-444,524-542 - evaluated on its
 * own rather than as posterior - prior, which cancels digits when the prior term is large).
 * batch_stride <= 1 means full batch.  out_lp (M) and out_grad (M x D) may each be NULL.
 * lanes: threads cooperating on one weight vector (0 = choose). */
OCBNN_API int ocbnn_logpost_grad(ocbnn_problem *p, const float *w, int64_t m, int32_t batch_offset, int32_t batch_stride,
                       int32_t with_data, float *out_lp, float *out_grad, int32_t lanes, void *stream);

/* forward pass f(x; w) for M weight vectors on P points: out (M x P x s_K).  Replaces forward/predict (base.py:89-104,267-277). */
OCBNN_API int ocbnn_forward(ocbnn_problem *p, const float *w, int64_t m, const float *x, int64_t npoints, float *out, void *stream);

/* predict(): forward + the output head, in the library.  Replaces RegressorMixin.predict (base.py:267-277),
 * ClassifierMixin.predict (base.py:446-461: softmax over the s_K logits, argmax) and BinaryClassifierMixin.predict
 * (base.py:544-559: sigmoid of the single logit, p >= 0.5).
 *   head = OCBNN_PREDICT_RAW   : out (M x P x s_K) network outputs
 *          OCBNN_PREDICT_PROBS : out (M x P x s_K) softmax probabilities, or (M x P) p(Y=1) for a Bernoulli problem
 *          OCBNN_PREDICT_CLASS : out holds the raw outputs (scratch), out_class (M x P, int32) the decisions
 * out_class may also be given with OCBNN_PREDICT_PROBS.  out_mean (P x s_K, or NULL): mean over the M samples of what `out`
 * holds — the posterior-mean predictive every evaluation of the reference starts from (bnn/utils.py:217-232). */
#define OCBNN_PREDICT_RAW 0
#define OCBNN_PREDICT_PROBS 1
#define OCBNN_PREDICT_CLASS 2
OCBNN_API int ocbnn_predict(ocbnn_problem *p, const float *w, int64_t m, const float *x, int64_t npoints, int32_t head,
                  float *out, int32_t *out_class, float *out_mean, void *stream);

/* Constraint-violation flags of posterior samples (the rejection count of run_toys.py:229-250, Figure 3c):
 * out_flags[m] = 1 when lo < f(x_p; w_m)[0] < hi for any of the P domain points.  scratch: M x P x s_K floats. */
OCBNN_API int ocbnn_violation_flags(ocbnn_problem *p, const float *w, int64_t m, const float *x, int64_t npoints, float lo, float hi,
                          float *scratch, uint8_t *out_flags, void *stream);

/* K2 — n_it HMC iterations (L leapfrog steps + Metropolis-Hastings each) for M independent chains.
 * Replaces HMCMixin.single (inference.py:39-78).  q (M x D) in/out; p0 (n_it x M x D) momenta ~N(0,1) and
 * u (n_it x M, float64) uniforms are injected by the host.  restore_on_reject = 0 reproduces the reference
 * (rejected proposals are kept, inference.py:46,78), 1 is textbook MH.
 * out_accept (n_it x M, uint8), out_h (n_it x M x 4: U0,K0,U1,K1), out_flags (M, bit0 = NaN seen) may be NULL.
 * out_samples (n_it/thin x M x D) receives q after every thin-th iteration when non-NULL. */
OCBNN_API int ocbnn_hmc_iterate(ocbnn_problem *p, float *q, const float *p0, const double *u, int64_t m, int32_t n_it,
                      double eps, int32_t leapfrog_l, int32_t with_data, int32_t restore_on_reject,
                      uint8_t *out_accept, float *out_h, int32_t *out_flags, float *out_samples, int32_t thin,
                      int32_t lanes, void *stream);

/* K4 — T SGLD steps for M independent chains.  Replaces SGLDMixin.single + the update (inference.py:304-340).
 * noise (T x M x D) is already scaled ~N(0, eps_t); half_eps (T, host) = (float)(eps_t/2); batch_offsets (T, host,
 * may be NULL when batch_stride <= 1).  out_samples (T/thin x M x D) receives w after every thin-th step when non-NULL. */
OCBNN_API int ocbnn_sgld_iterate(ocbnn_problem *p, float *w, const float *noise, const float *half_eps, const int32_t *batch_offsets,
                       int32_t batch_stride, int64_t m, int32_t t_steps, int32_t with_data, float *out_samples,
                       int32_t thin, int32_t lanes, void *stream);

/* Communicator for the two samplers with an exchange step (SVGD, BBB): one per process (= per GPU), NCCL underneath (the library
 * binds libnccl.so.2 at run time - the copy already loaded in the process if there is one).  Rank 0 obtains an id, the host
 * program hands it to the other ranks by whatever means it has (torch.distributed / MPI / a file), every rank calls
 * ocbnn_comm_create (collective).  ocbnn_comm_wrap adopts a communicator the caller already owns (nccl_comm = ncclComm_t).
 * NULL stands for "one rank" wherever a communicator is taken. */
typedef struct ocbnn_comm ocbnn_comm;
#define OCBNN_COMM_ID_BYTES 128
OCBNN_API int ocbnn_comm_unique_id(uint8_t id[OCBNN_COMM_ID_BYTES]);
OCBNN_API int ocbnn_comm_create(const uint8_t id[OCBNN_COMM_ID_BYTES], int32_t rank, int32_t world, ocbnn_comm **out);
OCBNN_API int ocbnn_comm_wrap(void *nccl_comm, int32_t rank, int32_t world, ocbnn_comm **out);
OCBNN_API int ocbnn_comm_destroy(ocbnn_comm *comm);
/* rows [*row_begin, *row_begin + *n_rows) of n units that rank `rank` of `world` owns: contiguous shares, the first n % world
 * ranks hold one more.  The layout every sharded entry point below assumes. */
OCBNN_API int ocbnn_comm_shard(int64_t n, int32_t rank, int32_t world, int64_t *row_begin, int64_t *n_rows);

/* K3 — SVGD (inference.py:208-292).  Two implementations behind one interface, selected by use_tensor:
 *   0  exact fp32 SIMT kernels (distances recomputed from x on every pass);
 *   1  tcgen05 path: x is centred and split into tf32 hi/lo parts.  D <= 64: neither the distances nor the kernel matrix are
 *      ever written to memory - the median select classifies 128 x 128 Gram tiles recomputed on the tensor cores (pairs near
 *      the bracket are re-evaluated with exact differences) and phi is a flash-attention shaped kernel, S = Xc Xc^T -> TMEM ->
 *      K = exp(..) -> TMEM -> O += K [G | Xc].  D > 64: distances of this rank's rows by a 3xTF32 Gram GEMM, K materialised as
 *      tf32 hi/lo, O by a TMA-fed tcgen05 GEMM;
 *  -1  choose (tensor path from 1024 particles up).
 * workspace >= ocbnn_svgd_workspace_bytes(n, d, n_rows, use_tensor) device bytes, 256-byte aligned.
 *
 * ocbnn_svgd_step — ONE SVGD iteration (SVGDMixin.single + the Adagrad step of SVGDMixin.infer, inference.py:225-259,272-283)
 * for particles sharded over the ranks of `comm`, everything enqueued on `stream` (and on an internal side stream forked from /
 * joined to it by events), no host synchronisation, capturable in a CUDA graph:
 *   gradients of this rank's rows (K1)  ->  all-gather G  ->  exact median bandwidth over ALL pairs (K3a; counters and radix
 *   histograms all-reduced)  ->  phi of this rank's rows (K3b)  ->  Adagrad on this rank's rows (K3c)  ->  all-gather X.
 * x_all (n x D, in/out): ALL particles, identical on every rank at entry and at exit; this rank owns the rows ocbnn_comm_shard
 * gives it (row_begin / n_rows are passed for checking).  g_all (n x D, out): gradients of all particles.  acc, phi
 * (n_rows x D): Adagrad accumulator (in/out) and update direction (out) of the owned rows.  h: device scalar (out).
 * batch_offset / batch_stride / with_data as in ocbnn_logpost_grad. */
OCBNN_API int64_t ocbnn_svgd_workspace_bytes(int64_t n, int64_t d, int64_t n_rows, int32_t use_tensor);
OCBNN_API int ocbnn_svgd_step(ocbnn_problem *p, float *x_all, float *g_all, float *acc, float *phi, float *h, int64_t n,
                              int64_t row_begin, int64_t n_rows, float lr, int32_t batch_offset, int32_t batch_stride,
                              int32_t with_data, int32_t use_tensor, void *workspace, int64_t workspace_bytes, ocbnn_comm *comm,
                              void *stream);

/* Building blocks of the step on one GPU (tests, profiling).
 * K3a — bandwidth h = med^2 / ln n: exact LOWER median of ||x_i - x_j + 1e-6|| over i<j, zeros dropped (inference.py:208-217).
 * SIMT arm: 3-pass radix select (11+11+10 bits of the fp32 pattern).  Tensor path: sampled bracket, one counting/compaction pass
 * over the pairs, select among the ~3 % candidates; in-kernel fallback to the radix select if the bracket misses.  On return
 * the first 4 x u64 after the 3 x 2048 u32 at the start of the workspace hold (bracket hit, candidates, values below the
 * bracket, exact zeros).  out_h: device scalar. */
OCBNN_API int ocbnn_svgd_bandwidth(const float *x, int64_t n, int64_t d, float *out_h, int32_t use_tensor, void *workspace,
                                   int64_t workspace_bytes, void *stream);
/* median-select algorithm of the tensor path: 1 bracket select (default), 0 the plain radix select, 2 bracket select with a
 * sabotaged bracket (exercises the fallback; tests).  mode < 0 queries.  Returns the previous mode. */
OCBNN_API int ocbnn_svgd_select_mode(int32_t mode);

/* K3b — SVGD update direction phi_i = (1/n) sum_j [K_ji g_j + grad_{x_j} K_ji] (inference.py:219-259) for rows
 * [row_begin, row_begin + n_rows) of the particle matrix x (n x D) with gradients g (n x D).
 * use_tensor = 2: tensor path, and the workspace was already prepared for this x and these rows by ocbnn_svgd_bandwidth;
 * 1 prepares it itself. */
OCBNN_API int ocbnn_svgd_phi(const float *x, const float *g, int64_t n, int64_t d, int64_t row_begin, int64_t n_rows,
                             const float *h, float *out_phi, int32_t use_tensor, void *workspace, int64_t workspace_bytes, void *stream);

/* tcgen05/TMA building block of the K3b tensor path, exported for testing: C (m x n) = A (m x k) * B (n x k)^T with every
 * fp32 operand split into tf32 hi + lo parts and three tensor-core MMAs per K step (fp32-class accuracy, ~2^-21 relative).
 * Row-major device arrays; workspace >= ocbnn_gemm_tf32x3_workspace_bytes(m, n, k) device bytes. */
OCBNN_API int64_t ocbnn_gemm_tf32x3_workspace_bytes(int64_t m, int64_t n, int64_t k);
OCBNN_API int ocbnn_gemm_tf32x3(const float *a, const float *b, float *c, int64_t m, int64_t n, int64_t k, void *workspace,
                                int64_t workspace_bytes, void *stream);

/* K3c — Adagrad (torch.optim.Adagrad defaults, inference.py:173,272): acc += g^2 ; x -= lr * g / (sqrt(acc)+1e-10),
 * with g = grad_sign * dir (SVGD passes dir = phi, grad_sign = -1). */
OCBNN_API int ocbnn_adagrad_step(float *x, float *acc, const float *dir, int64_t n_elem, float lr, float grad_sign, void *stream);

/* K5 — BBB: w_j = mu + eps_j * softplus(rho) (inference.py:132-140) and the closed-form ELBO gradient
 * reduction over the k Monte-Carlo samples (inference.py:142-153 + autograd). */
OCBNN_API int ocbnn_bbb_sample(const float *q_params, const float *eps, int64_t k, int64_t d, float *out_w, void *stream);
/* k local samples of k_total; writes out_grad_q (D x 2) = this rank's partial sums / k_total (the entropy term
 * sigmoid(rho)/sigma is added only when add_entropy != 0, i.e. on one rank) and out_elbo (device scalar) likewise;
 * colstat (device, D floats) is scratch.  Partial results of several ranks add up (all-reduce sum). */
OCBNN_API int ocbnn_bbb_reduce(const float *q_params, const float *eps, const float *lp, const float *grad, int64_t k, int64_t d,
                     int64_t k_total, int32_t add_entropy, float *out_grad_q, float *out_elbo, float *colstat, void *stream);

/* ocbnn_bbb_step — ONE Bayes-by-Backprop epoch (BBBMixin.compute_elbo + backward + the Adagrad step, inference.py:142-187) with
 * the Monte-Carlo samples sharded over the ranks of `comm`: w_j = mu + eps_j softplus(rho) for this rank's k_local draws
 * eps (k_local x D), K1 on them, the closed-form (D x 2) ELBO gradient, ONE all-reduce of gradient + ELBO, Adagrad on q_params
 * (D x 2, in/out, identical on every rank) with accumulator acc (D x 2, in/out).  out_elbo: device scalar.
 * workspace >= ocbnn_bbb_workspace_bytes(k_local, d) device bytes.  Enqueued on `stream`, no host synchronisation. */
OCBNN_API int64_t ocbnn_bbb_workspace_bytes(int64_t k_local, int64_t d);
OCBNN_API int ocbnn_bbb_step(ocbnn_problem *p, float *q_params, float *acc, const float *eps, int64_t k_local, int64_t k_total,
                             float lr, int32_t batch_offset, int32_t batch_stride, int32_t with_data, float *out_elbo,
                             void *workspace, int64_t workspace_bytes, ocbnn_comm *comm, void *stream);

/* K6 — Gaussian AOCP objective and its gradient w.r.t. params (D x 2 = mean, rho; std = softplus(rho)).
 * Replaces gaussian_aocp_objective + autograd (base.py:390-419 regression, :624-672 binary) for C constraint points cx (C x s_0):
 * mode 0: binary deterministic, targets (C) = desired class, loss |class - p| ; mode 1: binary probabilistic, targets (C) =
 * desired p(Y=1), symmetric Bernoulli KL with both sides clamped to [0.001, 0.999] ; mode 2: regression, targets
 * (C x 2*max_iv) = permitted (lb, ub) pairs, NaN padded, loss = 1 - permitted mass of N(f(x;mean), sigma_noise + sigma^2).
 * Writes (or, with accumulate != 0, adds) sum_c loss_c * inv_denom to out_loss (device scalar) and the gradient * inv_denom to
 * out_grad (D x 2).  workspace >= ocbnn_aocp_workspace_bytes(C, D) device bytes.  Uses only the net shape of `p`. */
OCBNN_API int64_t ocbnn_aocp_workspace_bytes(int64_t c, int64_t d);
OCBNN_API int ocbnn_aocp_objective(ocbnn_problem *p, const float *params, const float *cx, const float *targets, int64_t c,
                                   int32_t mode, int32_t max_iv, float sigma_noise, float inv_denom, int32_t accumulate,
                                   float *out_loss, float *out_grad, void *workspace, int64_t workspace_bytes, void *stream);

/* Diagnostics for the roofline report: measured instruction peaks of this GPU.  mode 0: scalar fp32 FMA, 1: packed
 * f32x2 FMA (both in FLOP/s, fma = 2), 2: MUFU.EX2 (op/s).  Synchronous. */
OCBNN_API int ocbnn_probe_peak(int32_t mode, double *out_rate);

/* Host-buffer convenience used for end-to-end timing: H2D of q/p0/u, K2, D2H of q/accept inside the call (synchronous). */
OCBNN_API int ocbnn_hmc_iterate_host(ocbnn_problem *p, float *q_host, const float *p0_host, const double *u_host, int64_t m,
                           int32_t n_it, double eps, int32_t leapfrog_l, int32_t with_data, int32_t restore_on_reject,
                           uint8_t *accept_host, int32_t lanes);

#ifdef __cplusplus
}
#endif
#endif /* OCBNN_B200_H */
