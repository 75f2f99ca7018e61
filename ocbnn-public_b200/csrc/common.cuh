// common.cuh — shared definitions of the OC-BNN B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/ocbnn_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "ocbnn_b200 is written for sm_100a (B200) only"
#endif

namespace ocbnn {

// ---- error plumbing -------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define OCBNN_CUDA(call)                                                        \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return ::ocbnn::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define OCBNN_LAUNCH_CHECK()                                                    \
    do {                                                                        \
        ::ocbnn::count_launch();                                                \
        cudaError_t _e = cudaGetLastError();                                    \
        if (_e != cudaSuccess) return ::ocbnn::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define OCBNN_REQUIRE(cond, code, ...)                                          \
    do {                                                                        \
        if (!(cond)) { ::ocbnn::set_error(__VA_ARGS__); return (code); }        \
    } while (0)

// ---- point table ----------------------------------------------------------------------------------
// One record per point (training point or constraint point), RS floats, RS odd so that lanes reading
// consecutive records hit distinct shared-memory banks:
//   [0, S0)      x
//   [S0]         head kind (int bits)
//   [S0+1]       weight  (likelihood rows: 1, multiplied by N/|B| when staged; constraint rows: cweight)
//   [S0+2, RS)   head parameters
// Rows [0, N) are the training set, rows [N, N+C) the constraint points.

struct EvalArgs {
    const float* pts;      // (N + C) x RS
    const float2* pri;     // D x (mean, 1/var)
    long long n_train;
    int n_cpoints;
    int n_pos, n_neg, n_dir;   // constraint rows are grouped: positive-Gaussian | negative-exponential | Dirichlet
    int neg_n1, neg_n2;        // term structure shared by every negative row (one-sided, two-sided terms), -1 if the rows differ
    int lik_head;          // OCBNN_HEAD_LIK_* of the training rows
    int rs;                // record stride
    int s0;
    int batch_offset;      // B = {offset, offset+stride, ...}
    int batch_stride;      // <=1: full batch
    int with_data;
    double const_prior;    // data-independent constants (added once)
    double const_lik;      // added when with_data
    float lik_scale;       // gauss: 1/(2 var)
    float tau0, tau1;      // negative exponential COCP
    float pos_inv_var;     // positive gaussian COCP
};

__host__ __device__ inline int batch_count(long long n, int offset, int stride) {
    if (stride <= 1) return (int)n;
    if (offset >= n) return 0;
    return (int)((n - offset + stride - 1) / stride);
}

struct Problem {
    ocbnn_desc desc;       // host copy (pointers cleared)
    int d;                 // number of weights
    int rs;                // record stride
    int np;                // params per record
    float* pts = nullptr;
    float2* pri = nullptr;
    float* x_train = nullptr;  // contiguous copy (n_train x s0) for the forward/predict path
    EvalArgs base;         // template args (batch fields filled per call)
    int sm_count = 148;
    int max_smem_optin = 0;
};

constexpr int kMaxSmemTable = 160 * 1024;   // bytes of shared memory a kernel may use for staged point records

// launch helpers implemented per translation unit
int small_net_supported(const Problem& p);
int launch_logpost_small(const Problem& p, const EvalArgs& a, const float* w, long long m, float* out_lp, float* out_grad,
                         int lanes, cudaStream_t st);
int launch_hmc_small(const Problem& p, const EvalArgs& a, float* q, const float* p0, const double* u, long long m, int n_it,
                     float eps, float eps_half, int L, int restore, uint8_t* out_acc, float* out_h, int* out_flags,
                     float* out_samples, int thin, int lanes, cudaStream_t st);
int launch_sgld_small(const Problem& p, const EvalArgs& a, float* w, const float* noise, const float* half_eps_dev,
                      const int* offsets_dev, long long m, int t_steps, float* out_samples, int thin, int lanes,
                      cudaStream_t st);
int launch_forward_small(const Problem& p, const float* w, long long m, const float* x, long long npts, float* out,
                         cudaStream_t st);

int launch_logpost_generic(const Problem& p, const EvalArgs& a, const float* w, long long m, float* out_lp, float* out_grad,
                           cudaStream_t st);
int launch_forward_generic(const Problem& p, const float* w, long long m, const float* x, long long npts, float* out,
                           cudaStream_t st);
// gen_tc.cu: the tcgen05 K1 of the application nets.  Returns 1 when it took the launch (*rc = its result), 0 when the net does
// not fit it and the caller should use generic_net.cu's kernels.
int try_launch_logpost_gen_tc(const Problem& p, const EvalArgs& a, const float* w, long long m, float* out_lp, float* out_grad,
                              cudaStream_t st, int* rc);
int try_launch_forward_gen_tc(const Problem& p, const float* w, long long m, const float* x, long long npts, float* out, cudaStream_t st, int* rc);

int choose_lanes(const Problem& p, long long m, int points);
int choose_lanes_hmc(const Problem& p, long long m, int points);

// abi.cu: evaluation arguments of a (minibatch, with_data) pair and K1 on whichever kernel family serves the net
EvalArgs eval_args(const Problem& p, int offset, int stride, int with_data);
int logpost_any(const Problem& p, const EvalArgs& a, const float* w, long long m, float* lp, float* grad, int lanes, cudaStream_t st);
const Problem* problem_of(const ocbnn_problem* h);

}  // namespace ocbnn
