// aocp.cu — K6: objective and gradient of the Gaussian amortised output-constrained prior (AOCP).
// Replaces gaussian_aocp_objective + autograd (bnn/base.py:390-419 regression, :624-672 binary), which loops over the
// constraint points in Python and builds a dense D x D matrix s^2*I per point (500 MB at D=11201).  Here:
//   1. k_point_jacobian   one warp per constraint point: f(x; mu) and b = df/dmu (D values) at the current mean
//   2. k_aocp_point       one warp per point: sigma^2 = sum_j b_j^2 s_j^2, a = b.mu, the point's loss and the two
//                          coefficients of its gradient (w.r.t. the mean path and w.r.t. sigma^2)
//   3. k_aocp_reduce      one thread per weight j: grad mu_j = sum_c cm_c b_cj ; grad rho_j = sum_c cs_c b_cj^2 2 s_j sigmoid(rho_j)
// b is treated as a constant (the reference evaluates it on a DETACHED copy of the mean).
#include "common.cuh"

namespace ocbnn {

struct JacLayout {
    int nl;
    int s[OCBNN_MAX_LAYERS + 1];
    int coff[OCBNN_MAX_LAYERS + 1];   // compact offset of W_l
    int aoff[OCBNN_MAX_LAYERS + 1];   // offset of a_l in the per-warp activation buffer
    int atot, maxw, act, d;
};

__device__ __forceinline__ float jac_act(int act, float z) {
    if (act == OCBNN_ACT_RBF) return expf(-z * z);
    if (act == OCBNN_ACT_RELU) return fmaxf(z, 0.f);
    return z;
}
__device__ __forceinline__ float jac_dact(int act, float z, float a) {
    if (act == OCBNN_ACT_RBF) return a * (-2.f * z);
    if (act == OCBNN_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return 1.f;
}

constexpr int kJacWarps = 4;

__global__ void __launch_bounds__(kJacWarps * 32)
k_point_jacobian(JacLayout L, const float* __restrict__ params, const float* __restrict__ cx, long long C, float* __restrict__ B,
                 float* __restrict__ fout) {
    extern __shared__ float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = 2 * L.atot + 2 * L.maxw;
    float* a = sm + warp * per;          // activations a_0..a_K
    float* z = a + L.atot;               // pre-activations (same offsets)
    float* dz = z + L.atot;              // ping-pong
    const long long c = (long long)blockIdx.x * kJacWarps + warp;
    if (c >= C) return;
    for (int d = lane; d < L.s[0]; d += 32) a[d] = cx[c * L.s[0] + d];
    __syncwarp();
    for (int l = 1; l <= L.nl; ++l) {
        const int si = L.s[l - 1], so = L.s[l];
        const float* ain = a + L.aoff[l - 1];
        for (int o = lane; o < so; o += 32) {
            float acc = params[2 * (L.coff[l] + si * so + o)];
            for (int d = 0; d < si; ++d) acc = fmaf(params[2 * (L.coff[l] + d * so + o)], ain[d], acc);
            z[L.aoff[l] + o] = acc;
            a[L.aoff[l] + o] = l < L.nl ? jac_act(L.act, acc) : acc;
        }
        __syncwarp();
    }
    if (lane == 0) fout[c] = z[L.aoff[L.nl]];
    float* dcur = dz;
    float* dnext = dz + L.maxw;
    for (int o = lane; o < L.s[L.nl]; o += 32) dcur[o] = o == 0 ? 1.f : 0.f;   // d f_0 / d z_K
    __syncwarp();
    float* Bc = B + c * L.d;
    for (int l = L.nl; l >= 1; --l) {
        const int si = L.s[l - 1], so = L.s[l];
        const float* ain = a + L.aoff[l - 1];
        for (int i = lane; i < si * so; i += 32) Bc[L.coff[l] + i] = ain[i / so] * dcur[i % so];
        for (int o = lane; o < so; o += 32) Bc[L.coff[l] + si * so + o] = dcur[o];
        if (l > 1) {
            for (int d = lane; d < si; d += 32) {
                float acc = 0.f;
                for (int o = 0; o < so; ++o) acc = fmaf(params[2 * (L.coff[l] + d * so + o)], dcur[o], acc);
                dnext[d] = acc * jac_dact(L.act, z[L.aoff[l - 1] + d], ain[d]);
            }
            __syncwarp();
            float* t = dcur; dcur = dnext; dnext = t;
        }
    }
}

__device__ __forceinline__ float softplus_a(float r) { return fmaxf(r, 0.f) + log1pf(expf(-fabsf(r))); }
__device__ __forceinline__ float sigmoid_a(float r) {
    float e = expf(-fabsf(r));
    float s = 1.f / (1.f + e);
    return r >= 0.f ? s : e * s;
}

// stats[c] = (loss_c, cm_c, cs_c)
__global__ void __launch_bounds__(128)
k_aocp_point(const float* __restrict__ params, const float* __restrict__ B, const float* __restrict__ fvals, const float* __restrict__ targets,
             long long C, int d, int mode, int max_iv, float sigma_noise, float* __restrict__ stats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long c = (long long)blockIdx.x * 4 + warp;
    if (c >= C) return;
    double s2 = 0.0, a = 0.0;
    for (int j = lane; j < d; j += 32) {
        const float b = B[c * d + j];
        const float sj = softplus_a(params[2 * j + 1]);
        s2 += (double)(b * b) * (double)(sj * sj);
        a += (double)b * (double)params[2 * j];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        a += __shfl_xor_sync(0xffffffffu, a, off);
    }
    if (lane != 0) return;
    const float sig2 = (float)s2, am = (float)a;
    float loss = 0.f, cm = 0.f, cs = 0.f;
    if (mode == 0 || mode == 1) {
        const float base = 1.f + 3.14159265358979323846f * sig2 / 8.f;
        const float kappa = rsqrtf(base);                                   // (1 + pi sigma^2 / 8)^(-1/2)   base.py:641
        const float p = sigmoid_a(kappa * am);                              // sigmoid(kappa * (b . mu))      base.py:642
        float dp;
        if (mode == 0) {                                                    // |class - p|                     base.py:645
            const float cls = targets[c];
            loss = fabsf(cls - p);
            dp = cls > p ? -1.f : (cls < p ? 1.f : 0.f);
        } else {                                                            // symmetric Bernoulli KL          base.py:663-670
            const bool inside = p > 0.001f && p < 0.999f;
            const float ph = fminf(fmaxf(p, 0.001f), 0.999f);
            const float p1 = fminf(fmaxf(targets[c], 0.001f), 0.999f);
            const float lph = logf(ph), lqh = logf(1.f - ph), lp1 = logf(p1), lq1 = logf(1.f - p1);
            loss = p1 * (lp1 - lph) + (1.f - p1) * (lq1 - lqh) + ph * (lph - lp1) + (1.f - ph) * (lqh - lq1);
            dp = inside ? (-p1 / ph + (1.f - p1) / (1.f - ph) + (lph - lp1) - (lqh - lq1)) : 0.f;
        }
        const float dzv = dp * p * (1.f - p);
        cm = dzv * kappa;
        cs = dzv * am * (-0.5f * kappa / base * 3.14159265358979323846f / 8.f);
    } else {                                                                // regression, base.py:405-417
        const float m = fvals[c];
        const float sd = sqrtf(sigma_noise + sig2);                         // sigma_noise enters unsquared (:405)
        float permitted = 0.f, dm = 0.f, dsd = 0.f;
        for (int k = 0; k < max_iv; ++k) {
            const float lb = targets[c * 2 * max_iv + 2 * k], ub = targets[c * 2 * max_iv + 2 * k + 1];
            if (isnan(lb) || isnan(ub)) continue;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float bound = e == 0 ? ub : lb, sign = e == 0 ? 1.f : -1.f;
                if (isinf(bound)) {
                    permitted += sign * (bound > 0.f ? 1.f : 0.f);
                } else {
                    const float t = (bound - m) / sd;
                    const float pdf = expf(-0.5f * t * t) * 0.3989422804014327f;
                    permitted += sign * normcdff(t);
                    dm += sign * pdf / sd;
                    dsd += sign * pdf * t / sd;
                }
            }
        }
        loss = 1.f - permitted;
        cm = dm;
        cs = dsd / (2.f * sd);
    }
    stats[3 * c + 0] = loss;
    stats[3 * c + 1] = cm;
    stats[3 * c + 2] = cs;
}

__global__ void __launch_bounds__(256)
k_aocp_reduce(const float* __restrict__ params, const float* __restrict__ B, const float* __restrict__ stats, long long C, int d, float inv_denom,
              int accumulate, float* __restrict__ out_loss, float* __restrict__ out_grad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < d) {
        float gm = 0.f, gs = 0.f;
        for (long long c = 0; c < C; ++c) {
            const float b = B[c * d + j];
            gm = fmaf(stats[3 * c + 1], b, gm);
            gs = fmaf(stats[3 * c + 2], b * b, gs);
        }
        const float rho = params[2 * j + 1];
        const float g0 = gm * inv_denom, g1 = gs * 2.f * softplus_a(rho) * sigmoid_a(rho) * inv_denom;
        out_grad[2 * j] = accumulate ? out_grad[2 * j] + g0 : g0;
        out_grad[2 * j + 1] = accumulate ? out_grad[2 * j + 1] + g1 : g1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (long long c = 0; c < C; ++c) s += (double)stats[3 * c];
        const float v = (float)(s * (double)inv_denom);
        *out_loss = accumulate ? *out_loss + v : v;
    }
}

}  // namespace ocbnn

using namespace ocbnn;

extern "C" OCBNN_API int64_t ocbnn_aocp_workspace_bytes(int64_t c, int64_t d) { return (c * d + 4 * c) * (int64_t)sizeof(float); }

extern "C" OCBNN_API int ocbnn_aocp_objective(ocbnn_problem* h, const float* params, const float* cx, const float* targets, int64_t c, int32_t mode,
                                              int32_t max_iv, float sigma_noise, float inv_denom, int32_t accumulate, float* out_loss,
                                              float* out_grad, void* workspace, int64_t workspace_bytes, void* stream) {
    OCBNN_REQUIRE(h && params && cx && targets && out_loss && out_grad && c >= 1 && mode >= 0 && mode <= 2, OCBNN_EINVAL,
                  "ocbnn_aocp_objective: bad argument");
    const Problem& p = *reinterpret_cast<Problem*>(h);
    OCBNN_REQUIRE(p.desc.layer_sizes[p.desc.n_layers] == 1, OCBNN_EUNSUPPORTED, "the AOCP objective is defined for scalar-output nets");
    OCBNN_REQUIRE(workspace && workspace_bytes >= ocbnn_aocp_workspace_bytes(c, p.d), OCBNN_EINVAL, "ocbnn_aocp_objective: workspace too small");
    OCBNN_REQUIRE(mode != 2 || max_iv >= 1, OCBNN_EINVAL, "ocbnn_aocp_objective: max_iv must be >= 1 in regression mode");
    cudaStream_t st = (cudaStream_t)stream;
    JacLayout L{};
    L.nl = p.desc.n_layers;
    int coff = 0, aoff = 0, maxw = 1;
    for (int l = 0; l <= L.nl; ++l) {
        L.s[l] = p.desc.layer_sizes[l];
        L.aoff[l] = aoff;
        aoff += L.s[l];
        if (L.s[l] > maxw) maxw = L.s[l];
        if (l >= 1) {
            L.coff[l] = coff;
            coff += (L.s[l - 1] + 1) * L.s[l];
        }
    }
    L.atot = aoff;
    L.maxw = maxw;
    L.act = p.desc.activation;
    L.d = p.d;
    float* B = reinterpret_cast<float*>(workspace);
    float* fvals = B + c * p.d;
    float* stats = fvals + c;
    const size_t sm = sizeof(float) * kJacWarps * (2 * L.atot + 2 * L.maxw);
    OCBNN_REQUIRE(sm <= 200 * 1024, OCBNN_EUNSUPPORTED, "network too wide for the Jacobian kernel");
    OCBNN_CUDA(cudaFuncSetAttribute(k_point_jacobian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_point_jacobian<<<(unsigned)((c + kJacWarps - 1) / kJacWarps), kJacWarps * 32, sm, st>>>(L, params, cx, c, B, fvals);
    OCBNN_LAUNCH_CHECK();
    k_aocp_point<<<(unsigned)((c + 3) / 4), 128, 0, st>>>(params, B, fvals, targets, c, p.d, mode, max_iv, sigma_noise, stats);
    OCBNN_LAUNCH_CHECK();
    k_aocp_reduce<<<(unsigned)((p.d + 255) / 256), 256, 0, st>>>(params, B, stats, c, p.d, inv_denom, accumulate, out_loss, out_grad);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
