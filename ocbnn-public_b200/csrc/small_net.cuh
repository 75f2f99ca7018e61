// small_net.cuh — register-resident evaluator for one-hidden-layer nets (the repro toy shapes [1,10,1],
// [2,10,3], [2,10,1]): every chain keeps its packed weight vector, momentum and gradient in registers;
// training points and constraint points are staged once in shared memory and streamed through.
//
// G lanes cooperate on one chain: lane r evaluates points r, r+G, ... and the partial gradients are combined
// with a warp-shuffle butterfly, so that 4096 chains (G=16..32) and 10^6 chains (G=1) both fill the GPU.
//
// Replaces forward/_unpack_layers/_nonlinearity (bnn/base.py:72-104), the heads (heads.cuh) and the autograd
// backward pass that follows every log_posterior() in bnn/inference.py.
#pragma once
#include "heads.cuh"

namespace ocbnn {

template <int S0_, int H_, int SK_, int ACT_>
struct Net1 {
    static constexpr int S0 = S0_, H = H_, SK = SK_, ACT = ACT_;
    static constexpr int OW1 = 0, OB1 = S0 * H, OW2 = OB1 + H, OB2 = OW2 + H * SK, D = OB2 + SK;
};

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
    if (ACT == OCBNN_ACT_RBF) return expf(-z * z);      // base.py:84
    if (ACT == OCBNN_ACT_RELU) return fmaxf(z, 0.f);    // base.py:86
    return z;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float z, float a) {
    if (ACT == OCBNN_ACT_RBF) return a * (-2.f * z);
    if (ACT == OCBNN_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return 1.f;
}

// value + gradient contribution of ONE point record
template <class N>
__device__ __forceinline__ void point_eval(const float (&w)[N::D], float (&g)[N::D], double& lp, const float* __restrict__ rec,
                                           const HeadConsts& hc) {
    float x[N::S0];
#pragma unroll
    for (int d = 0; d < N::S0; ++d) x[d] = rec[d];
    const int kind = __float_as_int(rec[N::S0]);
    const float wgt = rec[N::S0 + 1];

    float z[N::H], a[N::H];
#pragma unroll
    for (int h = 0; h < N::H; ++h) {
        float s = w[N::OB1 + h];
#pragma unroll
        for (int d = 0; d < N::S0; ++d) s = fmaf(w[N::OW1 + d * N::H + h], x[d], s);
        z[h] = s;
        a[h] = act_fwd<N::ACT>(s);
    }
    float f[N::SK];
#pragma unroll
    for (int k = 0; k < N::SK; ++k) {
        float s = w[N::OB2 + k];
#pragma unroll
        for (int h = 0; h < N::H; ++h) s = fmaf(w[N::OW2 + h * N::SK + k], a[h], s);
        f[k] = s;
    }
    float val, df[N::SK];
    eval_head<N::SK>(kind, f, rec + N::S0 + 2, hc, val, df);
    lp += (double)(wgt * val);
#pragma unroll
    for (int k = 0; k < N::SK; ++k) {
        df[k] *= wgt;
        g[N::OB2 + k] += df[k];
    }
#pragma unroll
    for (int h = 0; h < N::H; ++h) {
        float da = 0.f;
#pragma unroll
        for (int k = 0; k < N::SK; ++k) {
            da = fmaf(df[k], w[N::OW2 + h * N::SK + k], da);
            g[N::OW2 + h * N::SK + k] = fmaf(df[k], a[h], g[N::OW2 + h * N::SK + k]);
        }
        float dz = da * act_bwd<N::ACT>(z[h], a[h]);
        g[N::OB1 + h] += dz;
#pragma unroll
        for (int d = 0; d < N::S0; ++d) g[N::OW1 + d * N::H + h] = fmaf(dz, x[d], g[N::OW1 + d * N::H + h]);
    }
}

// forward only (predict path)
template <class N>
__device__ __forceinline__ void point_forward(const float (&w)[N::D], const float* __restrict__ x, float (&f)[N::SK]) {
    float a[N::H];
#pragma unroll
    for (int h = 0; h < N::H; ++h) {
        float s = w[N::OB1 + h];
#pragma unroll
        for (int d = 0; d < N::S0; ++d) s = fmaf(w[N::OW1 + d * N::H + h], x[d], s);
        a[h] = act_fwd<N::ACT>(s);
    }
#pragma unroll
    for (int k = 0; k < N::SK; ++k) {
        float s = w[N::OB2 + k];
#pragma unroll
        for (int h = 0; h < N::H; ++h) s = fmaf(w[N::OW2 + h * N::SK + k], a[h], s);
        f[k] = s;
    }
}

// ---- staging of point records into shared memory -----------------------------------------------------
// points [begin, begin+count) of the evaluation set: first nb = |B| training rows (strided), then the constraint rows.
__device__ __forceinline__ void stage_points(float* __restrict__ tab, const EvalArgs& a, int nb, float mult, int begin, int count) {
    const int rs = a.rs;
    const int stride = a.batch_stride > 1 ? a.batch_stride : 1;
    const int offset = a.batch_stride > 1 ? a.batch_offset : 0;
    const int wcol = a.s0 + 1;
    for (int idx = threadIdx.x; idx < count * rs; idx += blockDim.x) {
        int j = idx / rs, k = idx - j * rs;
        int pj = begin + j;
        long long row = pj < nb ? (long long)offset + (long long)pj * stride : a.n_train + (pj - nb);
        float v = __ldg(a.pts + row * rs + k);
        if (k == wcol && pj < nb) v *= mult;
        tab[idx] = v;
    }
}

struct EvalCtx {
    float* tab;            // shared: staged records (tile_cap x rs)
    const float2* pri;     // shared: D x (mean, 1/var)
    int nb;                // training points in this evaluation
    int npts;              // nb + constraint points
    int tile_cap;          // records that fit in `tab`
    float mult;            // N/|B|
    double lp_const;
    HeadConsts hc;
};

__device__ __forceinline__ EvalCtx make_ctx(const EvalArgs& a, float* tab, const float2* pri, int tile_cap) {
    EvalCtx c;
    c.tab = tab;
    c.pri = pri;
    c.nb = a.with_data ? batch_count(a.n_train, a.batch_offset, a.batch_stride) : 0;
    c.npts = c.nb + a.n_cpoints;
    c.tile_cap = tile_cap;
    c.mult = (a.with_data && a.batch_stride > 1) ? (float)((double)a.n_train / (double)c.nb) : 1.f;
    c.lp_const = a.const_prior + (a.with_data ? a.const_lik : 0.0);
    c.hc.lik_scale = a.lik_scale;
    c.hc.tau0x2 = 2.f * a.tau0;
    c.hc.tau1x2 = 2.f * a.tau1;
    c.hc.pos_inv_var = a.pos_inv_var;
    return c;
}

// log posterior value and gradient at w.  Block-uniform control flow (contains __syncthreads when the point set
// does not fit one tile or `restage` is set).  On return every lane of the group holds the full g and lp.
template <class N, int G>
__device__ __forceinline__ void evaluate(const EvalArgs& a, const EvalCtx& c, bool restage, const float (&w)[N::D], float (&g)[N::D],
                                         float& lp_out, int lane) {
#pragma unroll
    for (int i = 0; i < N::D; ++i) g[i] = 0.f;
    double lp = 0.0;
    const bool multi = c.npts > c.tile_cap;
    for (int begin = 0; begin < c.npts; begin += c.tile_cap) {
        int count = min(c.tile_cap, c.npts - begin);
        if (multi || restage) {
            __syncthreads();
            stage_points(c.tab, a, c.nb, c.mult, begin, count);
            __syncthreads();
        }
        for (int j = lane; j < count; j += G) point_eval<N>(w, g, lp, c.tab + j * a.rs, c.hc);
    }
    if (G > 1) {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
            for (int i = 0; i < N::D; ++i) g[i] += __shfl_xor_sync(0xffffffffu, g[i], off);
            lp += __shfl_xor_sync(0xffffffffu, lp, off);
        }
    }
    // Gaussian prior over the weights (isotropic, or the AOCP diagonal Gaussian): base.py:106-133
    double pr = 0.0;
#pragma unroll
    for (int i = 0; i < N::D; ++i) {
        float2 mv = c.pri[i];
        float diff = w[i] - mv.x;
        float t = diff * mv.y;
        g[i] -= t;
        pr += (double)(diff * t);
    }
    lp += c.lp_const - 0.5 * pr;
    lp_out = (float)lp;
}

}  // namespace ocbnn
