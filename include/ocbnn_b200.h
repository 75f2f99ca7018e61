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

#define OCBNN_ABI_VERSION 1
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
 * batch_stride <= 1 means full batch.  out_lp (M) and out_grad (M x D) may each be NULL.
 * lanes: threads cooperating on one weight vector (0 = choose). */
OCBNN_API int ocbnn_logpost_grad(ocbnn_problem *p, const float *w, int64_t m, int32_t batch_offset, int32_t batch_stride,
                       int32_t with_data, float *out_lp, float *out_grad, int32_t lanes, void *stream);

/* forward pass f(x; w) for M weight vectors on P points: out (M x P x s_K).  Replaces forward/predict (base.py:89-104,267-277). */
OCBNN_API int ocbnn_forward(ocbnn_problem *p, const float *w, int64_t m, const float *x, int64_t npoints, float *out, void *stream);

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

/* K3 — SVGD (inference.py:208-292).  Two implementations behind one interface, selected by use_tensor:
 *   0  exact fp32 SIMT kernels (distances recomputed from x on every pass);
 *   1  tcgen05 path: x is centred and split into tf32 hi/lo parts, squared distances of this rank's rows are formed once
 *      (exact differences for D <= 64, 3xTF32 Gram GEMM on the tensor cores above that) and kept in the workspace, the
 *      kernel matrix K = exp(-d2/h) and O = K [G | Xc] run as a TMA-fed tcgen05 GEMM in 3xTF32 precision;
 *  -1  choose (tensor path from 1024 particles up).
 * workspace >= ocbnn_svgd_workspace_bytes(n, d, n_rows, use_tensor) device bytes, 256-byte aligned; it starts with the
 * select histogram (3 x 2048 u32) and the select state (4 x u64), which a multi-GPU caller all-reduces between passes.
 *
 * K3a — bandwidth h = med^2 / ln n: exact LOWER median of ||x_i - x_j + 1e-6|| over i<j, zeros dropped, by a 3-pass radix
 * select (11+11+10 bits of the fp32 pattern).  out_h: device scalar.  A SIMT pass histograms rows i = row_start + t*row_step,
 * t < n_rows, against all j > i; a "prepared" pass reads the distances ocbnn_svgd_prepare left in the workspace for rows
 * [row_begin, row_begin + n_rows). */
OCBNN_API int64_t ocbnn_svgd_workspace_bytes(int64_t n, int64_t d, int64_t n_rows, int32_t use_tensor);
OCBNN_API int ocbnn_svgd_prepare(const float *x, int64_t n, int64_t d, int64_t row_begin, int64_t n_rows, int32_t use_tensor,
                                 void *workspace, int64_t workspace_bytes, void *stream);
OCBNN_API int ocbnn_svgd_hist_pass(const float *x, int64_t n, int64_t d, int64_t row_start, int64_t row_step, int64_t n_rows,
                                   int32_t pass, uint32_t *hist, uint64_t *select_state, void *stream);
OCBNN_API int ocbnn_svgd_hist_pass_prepared(int64_t n, int64_t d, int64_t row_begin, int64_t n_rows, int32_t pass, uint32_t *hist,
                                            uint64_t *select_state, void *workspace, void *stream);
OCBNN_API int ocbnn_svgd_select_update(int32_t pass, uint32_t *hist, uint64_t *select_state, int64_t n, float *out_h, void *stream);
/* single-GPU convenience: prepare (tensor path) + the select.  With all rows on one GPU the tensor path finds the same exact
 * median in three launches (sampled bracket, one counting/compaction pass over the distances, one-CTA select among the ~2.5 %
 * candidates; in-kernel fallback to the radix select if the bracket misses).  On return the first 4 x u64 of the select state
 * hold (bracket hit, candidates, values below the bracket, exact zeros). */
OCBNN_API int ocbnn_svgd_bandwidth(const float *x, int64_t n, int64_t d, float *out_h, int32_t use_tensor, void *workspace,
                                   int64_t workspace_bytes, void *stream);
/* Bracket select across ranks (tensor path; every rank holds all particles and the distances of rows [row_begin, row_begin + n_rows),
 * left in `workspace` by ocbnn_svgd_prepare).  Stage 1 samples the bracket from X (identical on every rank), counts this rank's pairs
 * and returns the device reduce buffer (reduce_words u64: counters + linear histogram) which the caller all-reduces as int64 sums;
 * stage 2 gathers this rank's candidates of the hit bin; stage 3 is a 3-pass radix select over those short lists: _sel_hist fills
 * hist[pass] (2048 u32, all-reduced by the caller), _sel_update advances the select state and, after pass 2, writes h.  If the bracket
 * missed, stage 3 runs over the distance rows instead (same result).  Replaces inference.py:208-217 for sharded particles. */
OCBNN_API int ocbnn_svgd_sel_count(int64_t n, int64_t d, int64_t row_begin, int64_t n_rows, void *workspace, uint64_t **reduce_buf,
                                   int64_t *reduce_words, void *stream);
OCBNN_API int ocbnn_svgd_sel_gather(int64_t n, int64_t d, int64_t row_begin, int64_t n_rows, void *workspace, void *stream);
OCBNN_API int ocbnn_svgd_sel_hist(int64_t n, int64_t d, int64_t row_begin, int64_t n_rows, int32_t pass, uint32_t *hist, uint64_t *select_state,
                                  void *workspace, void *stream);
OCBNN_API int ocbnn_svgd_sel_update(int64_t n, int64_t d, int64_t row_begin, int64_t n_rows, int32_t pass, const uint32_t *hist,
                                    uint64_t *select_state, void *workspace, float *out_h, void *stream);
/* median-select algorithm of ocbnn_svgd_bandwidth's tensor path: 1 bracket select (default), 0 the 3-pass radix select,
 * 2 bracket select with a sabotaged bracket (exercises the fallback; tests).  mode < 0 queries.  Returns the previous mode. */
OCBNN_API int ocbnn_svgd_select_mode(int32_t mode);

/* K3b — SVGD update direction phi_i = (1/n) sum_j [K_ji g_j + grad_{x_j} K_ji] (inference.py:219-259) for rows
 * [row_begin, row_begin + n_rows) of the (gathered) particle matrix x (n x D) with gradients g (n x D).
 * use_tensor = 2: tensor path, and the workspace was already prepared for this x and these rows (by ocbnn_svgd_prepare
 * or ocbnn_svgd_bandwidth); 1 prepares it itself. */
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
