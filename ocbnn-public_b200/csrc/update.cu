// update.cu — K5: Bayes-by-Backprop reparameterised sampling and the closed-form ELBO-gradient reduction.
//   w_j = mu + eps_j * softplus(rho)                                                    (inference.py:132-140)
//   dELBO/dmu = mean_j g_j ; dELBO/drho = sigmoid(rho) (mean_j g_j*eps_j + 1/sigma)      (autograd of :142-153)
//   ELBO = mean_j [ log post(w_j) - sum_i log N(w_ji; mu_i, sigma_i) ]
#include "comm.h"
#include "svgd_tc.h"

namespace ocbnn {

__device__ __forceinline__ float softplus_f(float r) { return fmaxf(r, 0.f) + log1pf(expf(-fabsf(r))); }   // logsumexp(rho, 0)
__device__ __forceinline__ float sigmoid_f(float r) {
    float e = expf(-fabsf(r));
    float s = 1.f / (1.f + e);
    return r >= 0.f ? s : e * s;
}

__global__ void __launch_bounds__(256)
k_bbb_sample(const float* __restrict__ qp, const float* __restrict__ eps, long long k, long long d, float* __restrict__ w) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = k * d, stride = (long long)gridDim.x * blockDim.x;
    for (; idx < tot; idx += stride) {
        long long i = idx % d;
        float mu = __ldg(qp + 2 * i), sd = softplus_f(__ldg(qp + 2 * i + 1));
        w[idx] = __fadd_rn(mu, __fmul_rn(eps[idx], sd));
    }
}

// one block per column i: deterministic tree reduction over the k samples
__global__ void __launch_bounds__(256)
k_bbb_reduce_cols(const float* __restrict__ qp, const float* __restrict__ eps, const float* __restrict__ grad, long long k, long long d,
                  float inv_ktot, int add_entropy, float* __restrict__ out_gq, float* __restrict__ colstat) {
    const long long i = blockIdx.x;
    float sg = 0.f, sge = 0.f, se2 = 0.f;
    for (long long j = threadIdx.x; j < k; j += blockDim.x) {
        float gv = __ldg(grad + j * d + i), ev = __ldg(eps + j * d + i);
        sg += gv;
        sge = fmaf(gv, ev, sge);
        se2 = fmaf(ev, ev, se2);
    }
    __shared__ float red[3][8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, off);
        sge += __shfl_xor_sync(0xffffffffu, sge, off);
        se2 += __shfl_xor_sync(0xffffffffu, se2, off);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sg; red[1][threadIdx.x >> 5] = sge; red[2][threadIdx.x >> 5] = se2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        sg = sge = se2 = 0.f;
        for (int w = 0; w < 8; ++w) { sg += red[0][w]; sge += red[1][w]; se2 += red[2][w]; }
        const float rho = qp[2 * i + 1];
        const float sd = softplus_f(rho), sig = sigmoid_f(rho);
        out_gq[2 * i] = sg * inv_ktot;
        out_gq[2 * i + 1] = sig * (sge * inv_ktot + (add_entropy ? 1.f / sd : 0.f));
        // sum_j log q(w_ji) = -se2/2 - k (log sd + 1/2 log 2pi)
        colstat[i] = -0.5f * se2 - (float)k * (logf(sd) + 0.91893853320467274178f);
    }
}

__global__ void __launch_bounds__(256)
k_bbb_elbo(const float* __restrict__ lp, long long k, const float* __restrict__ colstat, long long d, float inv_ktot, float* __restrict__ out) {
    double s = 0.0;
    for (long long j = threadIdx.x; j < k; j += blockDim.x) s += (double)lp[j];
    for (long long i = threadIdx.x; i < d; i += blockDim.x) s -= (double)colstat[i];
    __shared__ double red[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        *out = (float)(s * (double)inv_ktot);
    }
}

}  // namespace ocbnn

using namespace ocbnn;

extern "C" OCBNN_API int ocbnn_bbb_sample(const float* q_params, const float* eps, int64_t k, int64_t d, float* out_w, void* stream) {
    OCBNN_REQUIRE(q_params && eps && out_w && k >= 1 && d >= 1, OCBNN_EINVAL, "ocbnn_bbb_sample: bad argument");
    long long blocks = (k * d + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_bbb_sample<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(q_params, eps, k, d, out_w);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_bbb_reduce(const float* q_params, const float* eps, const float* lp, const float* grad, int64_t k, int64_t d,
                                int64_t k_total, int32_t add_entropy, float* out_grad_q, float* out_elbo, float* colstat, void* stream) {
    OCBNN_REQUIRE(q_params && eps && lp && grad && out_grad_q && out_elbo && colstat && k >= 1 && d >= 1 && k_total >= k, OCBNN_EINVAL,
                  "ocbnn_bbb_reduce: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const float inv = 1.f / (float)k_total;
    k_bbb_reduce_cols<<<(unsigned)d, 256, 0, st>>>(q_params, eps, grad, k, d, inv, add_entropy, out_grad_q, colstat);
    OCBNN_LAUNCH_CHECK();
    k_bbb_elbo<<<1, 256, 0, st>>>(lp, k, colstat, d, inv, out_elbo);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

// ---- one BBB epoch in one call ---------------------------------------------------------------------------------------------
// workspace: W[k x D] | G[k x D] | lp[k] | colstat[D] | red[2 D + 1] (gradient and ELBO, summed over the ranks in ONE all-reduce)
static long long up256(long long floats) { return (floats * 4 + 255) / 256 * 256; }
extern "C" OCBNN_API int64_t ocbnn_bbb_workspace_bytes(int64_t k_local, int64_t d) {
    return 2 * up256(k_local * d) + up256(k_local) + up256(d) + up256(2 * d + 1) + 256;
}

__global__ void k_copy_scalar(const float* __restrict__ src, float* __restrict__ dst) { *dst = *src; }

extern "C" OCBNN_API int ocbnn_bbb_step(ocbnn_problem* ph, float* q_params, float* acc, const float* eps, int64_t k_local, int64_t k_total, float lr,
                                        int32_t batch_offset, int32_t batch_stride, int32_t with_data, float* out_elbo, void* workspace,
                                        int64_t workspace_bytes, ocbnn_comm* comm_h, void* stream) {
    OCBNN_REQUIRE(ph && q_params && acc && eps && out_elbo && k_local >= 1 && k_total >= k_local, OCBNN_EINVAL, "ocbnn_bbb_step: bad argument");
    OCBNN_REQUIRE(batch_stride <= 1 || (batch_offset >= 0 && batch_offset < batch_stride), OCBNN_EINVAL, "batch_offset must be in [0, stride)");
    const Problem& p = *problem_of(ph);
    const long long d = p.d;
    OCBNN_REQUIRE(workspace && workspace_bytes >= ocbnn_bbb_workspace_bytes(k_local, d), OCBNN_EINVAL, "ocbnn_bbb_step: workspace too small");
    Comm* comm = reinterpret_cast<Comm*>(comm_h);
    cudaStream_t st = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* w = reinterpret_cast<float*>(base);        base += up256(k_local * d);
    float* g = reinterpret_cast<float*>(base);        base += up256(k_local * d);
    float* lp = reinterpret_cast<float*>(base);       base += up256(k_local);
    float* colstat = reinterpret_cast<float*>(base);  base += up256(d);
    float* red = reinterpret_cast<float*>(base);
    int rc;
    if ((rc = ocbnn_bbb_sample(q_params, eps, k_local, d, w, stream))) return rc;
    const EvalArgs a = eval_args(p, batch_offset, batch_stride, with_data);
    if ((rc = logpost_any(p, a, w, k_local, lp, g, 0, st))) return rc;
    // the entropy term of the gradient is added once: on rank 0
    if ((rc = ocbnn_bbb_reduce(q_params, eps, lp, g, k_local, d, k_total, comm_rank(comm) == 0 ? 1 : 0, red, red + 2 * d, colstat, stream))) return rc;
    if ((rc = comm_all_sum(comm, red, 2 * d + 1, kCommF32, st))) return rc;
    if ((rc = adagrad_launch(q_params, acc, red, 2 * d, lr, -1.f, st))) return rc;              // loss = -ELBO
    k_copy_scalar<<<1, 1, 0, st>>>(red + 2 * d, out_elbo);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
