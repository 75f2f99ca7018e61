// update.cu — K5: Bayes-by-Backprop reparameterised sampling and the closed-form ELBO-gradient reduction.
//   w_j = mu + eps_j * softplus(rho)                                                    (inference.py:132-140)
//   dELBO/dmu = mean_j g_j ; dELBO/drho = sigmoid(rho) (mean_j g_j*eps_j + 1/sigma)      (autograd of :142-153)
//   ELBO = mean_j [ log post(w_j) - sum_i log N(w_ji; mu_i, sigma_i) ]
#include "common.cuh"

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
