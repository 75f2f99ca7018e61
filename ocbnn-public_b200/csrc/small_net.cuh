// small_net.cuh — register-resident evaluator for one-hidden-layer nets (the repro toy shapes [1,10,1],
// [2,10,3], [2,10,1]): every chain keeps its packed weight vector, momentum and gradient in registers;
// training points and constraint points are staged once in shared memory and streamed through.
//
// B200 specifics:
//   * all per-weight state lives in float2 register pairs (two adjacent hidden units per pair) and the MLP
//     forward/backward, the leapfrog update and the kinetic energy use the packed fp32x2 instructions of
//     sm_100 (FFMA2/FMUL2/FADD2): half the issue slots of scalar code on the fma pipe;
//   * exp(-z^2) is one FMUL2 + two MUFU.EX2 (ex2.approx, 2^-22 relative error), reciprocals are MUFU.RCP;
//   * points are grouped by head kind (likelihood | positive | negative | Dirichlet rows), so every inner
//     loop has a compile-time head and no per-point dispatch;
//   * G lanes cooperate on one chain: lane r evaluates points r, r+G, ... and the partial gradients are
//     combined with a warp-shuffle butterfly, so that 4096 chains (G=8..32) and 10^6 chains (G=1) both fill
//     the 148 SMs.
//
// Replaces forward/_unpack_layers/_nonlinearity (bnn/base.py:72-104), the heads (heads.cuh) and the autograd
// backward pass that follows every log_posterior() in bnn/inference.py.
#pragma once
#include "heads.cuh"

namespace ocbnn {

// Slot layout (float2 slots): W1 pairs over h for every input d | b1 pairs | W2 pairs over h for every output k | b2 pairs.
template <int S0_, int H_, int SK_, int ACT_>
struct Net1 {
    static constexpr int S0 = S0_, H = H_, SK = SK_, ACT = ACT_;
    static constexpr int HP = (H + 1) / 2, SKP = (SK + 1) / 2;
    static constexpr int SW1 = 0, SB1 = S0 * HP, SW2 = SB1 + HP, SB2 = SW2 + SK * HP, NS = SB2 + SKP;
    static constexpr int D = S0 * H + H + H * SK + SK;
    // index into the packed (reference) layout of element e (0/1) of slot s; -1 for padding
    __host__ __device__ static constexpr int cidx(int s, int e) {
        if (s < SB1) {
            int d = s / HP, h = 2 * (s % HP) + e;
            return h < H ? d * H + h : -1;
        }
        if (s < SW2) {
            int h = 2 * (s - SB1) + e;
            return h < H ? S0 * H + h : -1;
        }
        if (s < SB2) {
            int k = (s - SW2) / HP, h = 2 * ((s - SW2) % HP) + e;
            return h < H ? S0 * H + H + h * SK + k : -1;
        }
        int k = 2 * (s - SB2) + e;
        return k < SK ? S0 * H + H + H * SK + k : -1;
    }
};

template <class N>
__device__ __forceinline__ void load_state(float2 (&v)[N::NS], const float* __restrict__ src) {
#pragma unroll
    for (int s = 0; s < N::NS; ++s) {
        v[s].x = N::cidx(s, 0) >= 0 ? __ldg(src + N::cidx(s, 0)) : 0.f;
        v[s].y = N::cidx(s, 1) >= 0 ? __ldg(src + N::cidx(s, 1)) : 0.f;
    }
}
template <class N>
__device__ __forceinline__ void store_state(const float2 (&v)[N::NS], float* __restrict__ dst) {
#pragma unroll
    for (int s = 0; s < N::NS; ++s) {
        if (N::cidx(s, 0) >= 0) dst[N::cidx(s, 0)] = v[s].x;
        if (N::cidx(s, 1) >= 0) dst[N::cidx(s, 1)] = v[s].y;
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

// activation on a pair; returns a and m with da/dz = kActScale<ACT> * m (the constant is folded into the upstream gradient)
template <int ACT>
__device__ __forceinline__ void act_pair(float2 z, float2& a, float2& m) {
    if (ACT == OCBNN_ACT_RBF) {                       // exp(-z^2), base.py:84 ; d/dz = -2 z a
        float2 t = __fmul2_rn(z, z);
        float2 u = __fmul2_rn(t, f2(-1.4426950408889634f));
        a = make_float2(ex2_approx(u.x), ex2_approx(u.y));
        m = __fmul2_rn(z, a);
    } else if (ACT == OCBNN_ACT_RELU) {               // max(z,0), base.py:86
        a = make_float2(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f));
        m = make_float2(z.x > 0.f ? 1.f : 0.f, z.y > 0.f ? 1.f : 0.f);
    } else {
        a = z;
        m = f2(1.f);
    }
}
template <int ACT>
__device__ __forceinline__ constexpr float act_scale() { return ACT == OCBNN_ACT_RBF ? -2.f : 1.f; }

// Weight views: the evaluator reads weights through operator[] so that they can live in registers (RegView) or in a
// per-thread column of shared memory (SmemView: slot s of thread t at base[s * stride], conflict-free float2 loads), which
// frees 2 x NS registers per thread for occupancy.
template <int NS>
struct RegView {
    const float2 (&a)[NS];
    __device__ __forceinline__ float2 operator[](int s) const { return a[s]; }
};
struct SmemView {
    const float2* base;
    int stride;
    // volatile asm: the loads must stay inside the point loop (a plain load is loop invariant and would be hoisted back
    // into 2 x NS registers, which is exactly what this view is meant to avoid)
    __device__ __forceinline__ float2 operator[](int s) const {
        float2 v;
        const unsigned addr = (unsigned)__cvta_generic_to_shared(base + s * stride);
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
        return v;
    }
};

// ---- fast heads for the register kernels (same formulas as heads.cuh, MUFU approximations) ----------------------
struct FastConsts {
    float2 tau_l2e;      // 2*tau_i*log2(e)
    float2 tau2;         // 2*tau_i
    float lik_scale, pos_inv_var;
};

// Phi(z) = 1/((1+e^{2 tau0 z})(1+e^{2 tau1 z})) and dPhi/dz (see heads.cuh); exponent clamped at 2^57.7 ~ e^40
__device__ __forceinline__ void phi_fast(float z, const FastConsts& c, float& P, float& dP) {
    float2 u = __fmul2_rn(f2(z), c.tau_l2e);
    float2 e = make_float2(ex2_approx(fminf(u.x, 57.7f)), ex2_approx(fminf(u.y, 57.7f)));
    float2 ab = __fadd2_rn(e, f2(1.f));
    float r = rcp_approx(ab.x * ab.y);
    float2 t = __fmul2_rn(e, c.tau2);
    float s = fmaf(t.x, ab.y, t.y * ab.x);
    P = r;
    dP = -(r * r) * s;
}

// two Phi evaluations packed in f32x2 lanes: 4 EX2 + 2 RCP, every other op a packed instruction
__device__ __forceinline__ void phi_fast2(float2 z, const FastConsts& c, float2& P, float2& dP) {
    const float2 u0 = __fmul2_rn(z, f2(c.tau_l2e.x)), u1 = __fmul2_rn(z, f2(c.tau_l2e.y));
    const float2 e0 = make_float2(ex2_approx(fminf(u0.x, 57.7f)), ex2_approx(fminf(u0.y, 57.7f)));
    const float2 e1 = make_float2(ex2_approx(fminf(u1.x, 57.7f)), ex2_approx(fminf(u1.y, 57.7f)));
    const float2 a = __fadd2_rn(e0, f2(1.f)), b = __fadd2_rn(e1, f2(1.f));
    const float2 ab = __fmul2_rn(a, b);
    const float2 r = make_float2(rcp_approx(ab.x), rcp_approx(ab.y));
    const float2 s = __ffma2_rn(__fmul2_rn(e0, f2(c.tau2.x)), b, __fmul2_rn(__fmul2_rn(e1, f2(c.tau2.y)), a));
    P = r;
    dP = __fmul2_rn(__fmul2_rn(r, make_float2(-r.x, -r.y)), s);
}

// Heads are evaluated for U points at once (run_many) so that their long serial MUFU chains interleave.
template <int KIND, int SK>
struct FastHead;

template <int KIND, int SK, int U>
__device__ __forceinline__ void head_many_generic(const float (&f)[U][SK], const float* const (&prm)[U], const FastConsts& c, float (&val)[U],
                                                  float (&df)[U][SK]) {
#pragma unroll
    for (int u = 0; u < U; ++u) FastHead<KIND, SK>::run(f[u], prm[u], c, val[u], df[u]);
}

template <int SK>
struct FastHead<OCBNN_HEAD_LIK_GAUSS, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts& c, float& val, float (&df)[SK]) {
        val = 0.f;
#pragma unroll
        for (int k = 0; k < SK; ++k) {
            float r = f[k] - prm[k];
            val = fmaf(-r * r, c.lik_scale, val);
            df[k] = -2.f * c.lik_scale * r;
        }
    }
};
// internal parameter layout of a negative-COCP row (built by ocbnn_problem_create from the ABI's gap list):
//   n1, n2 (int bits), n1 x (c, s) one-sided terms Phi(s (f - c)), n2 x (a, b) two-sided terms Phi(a - f) Phi(f - b)
template <int SK>
struct FastHead<OCBNN_HEAD_NEG_EXP, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts& c, float& val, float (&df)[SK]) {
        const int n1 = __float_as_int(prm[0]), n2 = __float_as_int(prm[1]);
        val = 0.f;
        float d = 0.f;
        const float* q = prm + 2;
        for (int k = 0; k < n1; ++k, q += 2) {
            float P, D;
            phi_fast(q[1] * (f[0] - q[0]), c, P, D);
            val += P;
            d = fmaf(q[1], D, d);
        }
        for (int k = 0; k < n2; ++k, q += 2) {
            float P1, D1, P2, D2;
            phi_fast(q[0] - f[0], c, P1, D1);
            phi_fast(f[0] - q[1], c, P2, D2);
            val = fmaf(P1, P2, val);
            d += P1 * D2 - D1 * P2;
        }
#pragma unroll
        for (int k = 0; k < SK; ++k) df[k] = 0.f;
        df[0] = d;
    }
    // two points in lockstep: term k of both points is one packed Phi evaluation; a point with fewer terms is masked
    static __device__ __forceinline__ void run_pair(float fa, float fb, const float* __restrict__ pa, const float* __restrict__ pb,
                                                    const FastConsts& c, float2& val, float2& d) {
        const int n1a = __float_as_int(pa[0]), n2a = __float_as_int(pa[1]);
        const int n1b = __float_as_int(pb[0]), n2b = __float_as_int(pb[1]);
        const int n1 = max(n1a, n1b), n2 = max(n2a, n2b);
        const float2 ff = make_float2(fa, fb);
        val = f2(0.f);
        d = f2(0.f);
        for (int k = 0; k < n1; ++k) {
            const bool va = k < n1a, vb = k < n1b;
            const float2 cc = make_float2(va ? pa[2 + 2 * k] : 0.f, vb ? pb[2 + 2 * k] : 0.f);
            const float2 ss = make_float2(va ? pa[3 + 2 * k] : 0.f, vb ? pb[3 + 2 * k] : 0.f);
            float2 P, D;
            phi_fast2(__fmul2_rn(ss, __fadd2_rn(ff, make_float2(-cc.x, -cc.y))), c, P, D);
            val = __ffma2_rn(P, make_float2(va ? 1.f : 0.f, vb ? 1.f : 0.f), val);
            d = __ffma2_rn(ss, D, d);
        }
        const float* qa = pa + 2 + 2 * n1a;
        const float* qb = pb + 2 + 2 * n1b;
        for (int k = 0; k < n2; ++k) {
            const bool va = k < n2a, vb = k < n2b;
            const float2 lo = make_float2(va ? qa[2 * k] : 0.f, vb ? qb[2 * k] : 0.f);
            const float2 hi = make_float2(va ? qa[2 * k + 1] : 0.f, vb ? qb[2 * k + 1] : 0.f);
            const float2 msk = make_float2(va ? 1.f : 0.f, vb ? 1.f : 0.f);
            float2 P1, D1, P2, D2;
            phi_fast2(__fadd2_rn(lo, make_float2(-ff.x, -ff.y)), c, P1, D1);
            phi_fast2(__fadd2_rn(ff, make_float2(-hi.x, -hi.y)), c, P2, D2);
            val = __ffma2_rn(__fmul2_rn(P1, P2), msk, val);
            const float2 dd = __ffma2_rn(P1, D2, make_float2(-D1.x * P2.x, -D1.y * P2.y));
            d = __ffma2_rn(dd, msk, d);
        }
    }
};
template <int SK, int U>
__device__ __forceinline__ void head_many_neg(const float (&f)[U][SK], const float* const (&prm)[U], const FastConsts& c, float (&val)[U],
                                              float (&df)[U][SK]) {
#pragma unroll
    for (int u = 0; u + 1 < U; u += 2) {
        float2 v, d;
        FastHead<OCBNN_HEAD_NEG_EXP, SK>::run_pair(f[u][0], f[u + 1][0], prm[u], prm[u + 1], c, v, d);
        val[u] = v.x;
        val[u + 1] = v.y;
#pragma unroll
        for (int k = 0; k < SK; ++k) df[u][k] = df[u + 1][k] = 0.f;
        df[u][0] = d.x;
        df[u + 1][0] = d.y;
    }
    if (U & 1) FastHead<OCBNN_HEAD_NEG_EXP, SK>::run(f[U - 1], prm[U - 1], c, val[U - 1], df[U - 1]);
}
template <int SK>
struct FastHead<OCBNN_HEAD_POS_GAUSS, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts& c, float& val, float (&df)[SK]) {
        HeadConsts hc;
        hc.pos_inv_var = c.pos_inv_var;
#pragma unroll
        for (int k = 0; k < SK; ++k) df[k] = 0.f;
        head_pos_gauss(f[0], prm, hc, val, df[0]);
    }
};
template <int SK>
struct FastHead<OCBNN_HEAD_DIRICHLET, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts&, float& val, float (&df)[SK]) {
        head_dirichlet<SK>(f, prm, val, df);
    }
};
template <int SK>
struct FastHead<OCBNN_HEAD_LIK_SOFTMAX, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts&, float& val, float (&df)[SK]) {
        head_lik_softmax<SK>(f, prm, val, df);
    }
};
template <int SK>
struct FastHead<OCBNN_HEAD_LIK_BERNOULLI, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts&, float& val, float (&df)[SK]) {
#pragma unroll
        for (int k = 0; k < SK; ++k) df[k] = 0.f;
        head_lik_bernoulli(f[0], prm, val, df[0]);
    }
};

#ifndef OCBNN_POINT_UNROLL
#define OCBNN_POINT_UNROLL 2
#endif

// value + gradient contribution of U point records with a compile-time head.  The U points are independent until the
// gradient accumulation, which gives the scheduler U forward/head dependency chains to interleave (the head is a long
// serial chain of MUFU ops).  Records beyond the segment are passed with live[u] = 0 and contribute nothing.
template <class N, int KIND, int U, class WV>
__device__ __forceinline__ void point_eval(const WV& w, float2 (&g)[N::NS], double& lp, const float* const (&rec)[U],
                                           const float (&live)[U], const FastConsts& fc) {
    float2 xx[U][N::S0], a[U][N::HP], m[U][N::HP];
    float f[U][N::SK], df[U][N::SK], val[U];
    const float* prm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int d = 0; d < N::S0; ++d) xx[u][d] = f2(rec[u][d]);
        prm[u] = rec[u] + N::S0 + 2;
        float2 facc[N::SK];
#pragma unroll
        for (int k = 0; k < N::SK; ++k) facc[k] = f2(0.f);
#pragma unroll
        for (int hp = 0; hp < N::HP; ++hp) {
            float2 z = w[N::SB1 + hp];
#pragma unroll
            for (int d = 0; d < N::S0; ++d) z = __ffma2_rn(w[N::SW1 + d * N::HP + hp], xx[u][d], z);
            act_pair<N::ACT>(z, a[u][hp], m[u][hp]);
#pragma unroll
            for (int k = 0; k < N::SK; ++k) facc[k] = __ffma2_rn(w[N::SW2 + k * N::HP + hp], a[u][hp], facc[k]);
        }
#pragma unroll
        for (int k = 0; k < N::SK; ++k) {
            const float b2 = (k & 1) ? w[N::SB2 + k / 2].y : w[N::SB2 + k / 2].x;
            f[u][k] = (facc[k].x + facc[k].y) + b2;
        }
    }
    if (KIND == OCBNN_HEAD_NEG_EXP) head_many_neg<N::SK, U>(f, prm, fc, val, df);
    else head_many_generic<KIND, N::SK, U>(f, prm, fc, val, df);
    float vsum = 0.f;                      // the U values are added in fp32, then once into the float64 accumulator
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float wgt = rec[u][N::S0 + 1] * live[u];
        vsum = fmaf(wgt, val[u], vsum);
#pragma unroll
        for (int k = 0; k < N::SK; ++k) df[u][k] *= wgt;
    }
    lp += (double)vsum;
#pragma unroll
    for (int k = 0; k < N::SK; ++k) {
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) s += df[u][k];
        if (k & 1) g[N::SB2 + k / 2].y += s; else g[N::SB2 + k / 2].x += s;
    }
#pragma unroll
    for (int hp = 0; hp < N::HP; ++hp) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float2 da = f2(0.f);
#pragma unroll
            for (int k = 0; k < N::SK; ++k) {
                const float2 dk = f2(df[u][k]);
                da = __ffma2_rn(dk, w[N::SW2 + k * N::HP + hp], da);
                g[N::SW2 + k * N::HP + hp] = __ffma2_rn(dk, a[u][hp], g[N::SW2 + k * N::HP + hp]);
            }
            const float2 dz = __fmul2_rn(__fmul2_rn(da, f2(act_scale<N::ACT>())), m[u][hp]);
            g[N::SB1 + hp] = __fadd2_rn(g[N::SB1 + hp], dz);
#pragma unroll
            for (int d = 0; d < N::S0; ++d) g[N::SW1 + d * N::HP + hp] = __ffma2_rn(dz, xx[u][d], g[N::SW1 + d * N::HP + hp]);
        }
    }
}

// forward only (predict path), scalar
template <class N>
__device__ __forceinline__ void point_forward(const float2 (&w)[N::NS], const float* __restrict__ x, float (&f)[N::SK]) {
    float2 facc[N::SK];
#pragma unroll
    for (int k = 0; k < N::SK; ++k) facc[k] = f2(0.f);
#pragma unroll
    for (int hp = 0; hp < N::HP; ++hp) {
        float2 z = w[N::SB1 + hp];
#pragma unroll
        for (int d = 0; d < N::S0; ++d) z = __ffma2_rn(w[N::SW1 + d * N::HP + hp], f2(x[d]), z);
        float2 a;
        if (N::ACT == OCBNN_ACT_RBF) a = make_float2(expf(-z.x * z.x), expf(-z.y * z.y));      // accurate exp on the predict path
        else if (N::ACT == OCBNN_ACT_RELU) a = make_float2(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f));
        else a = z;
#pragma unroll
        for (int k = 0; k < N::SK; ++k) facc[k] = __ffma2_rn(w[N::SW2 + k * N::HP + hp], a, facc[k]);
    }
#pragma unroll
    for (int k = 0; k < N::SK; ++k) f[k] = (facc[k].x + facc[k].y) + ((k & 1) ? w[N::SB2 + k / 2].y : w[N::SB2 + k / 2].x);
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
    const float4* pri;     // shared: per slot (mean.x, mean.y, 1/var.x, 1/var.y)   [register kernels]
    int nb;                // training points in this evaluation
    int npts;              // nb + constraint points
    int seg_end[4];        // end (exclusive) of the likelihood | positive | negative | Dirichlet rows in evaluation order
    int tile_cap;          // records that fit in `tab`
    float mult;            // N/|B|
    double lp_const;
    HeadConsts hc;
    FastConsts fc;
};

__device__ __forceinline__ EvalCtx make_ctx(const EvalArgs& a, float* tab, const float4* pri, int tile_cap) {
    EvalCtx c;
    c.tab = tab;
    c.pri = pri;
    c.nb = a.with_data ? batch_count(a.n_train, a.batch_offset, a.batch_stride) : 0;
    c.npts = c.nb + a.n_cpoints;
    c.seg_end[0] = c.nb;
    c.seg_end[1] = c.seg_end[0] + a.n_pos;
    c.seg_end[2] = c.seg_end[1] + a.n_neg;
    c.seg_end[3] = c.seg_end[2] + a.n_dir;
    c.tile_cap = tile_cap;
    c.mult = (a.with_data && a.batch_stride > 1) ? (float)((double)a.n_train / (double)c.nb) : 1.f;
    c.lp_const = a.const_prior + (a.with_data ? a.const_lik : 0.0);
    c.hc.lik_scale = a.lik_scale;
    c.hc.tau0x2 = 2.f * a.tau0;
    c.hc.tau1x2 = 2.f * a.tau1;
    c.hc.pos_inv_var = a.pos_inv_var;
    c.fc.tau2 = make_float2(2.f * a.tau0, 2.f * a.tau1);
    c.fc.tau_l2e = make_float2(2.f * a.tau0 * 1.4426950408889634f, 2.f * a.tau1 * 1.4426950408889634f);
    c.fc.lik_scale = a.lik_scale;
    c.fc.pos_inv_var = a.pos_inv_var;
    return c;
}

// prior table in slot order
template <class N>
__device__ __forceinline__ void load_pri_slots(float4* pri_s, const EvalArgs& a) {
    for (int s = threadIdx.x; s < N::NS; s += blockDim.x) {
        int i0 = N::cidx(s, 0), i1 = N::cidx(s, 1);
        float2 m0 = i0 >= 0 ? a.pri[i0] : make_float2(0.f, 0.f);
        float2 m1 = i1 >= 0 ? a.pri[i1] : make_float2(0.f, 0.f);
        pri_s[s] = make_float4(m0.x, m1.x, m0.y, m1.y);
    }
}

template <class N, int G, int KIND, class WV>
__device__ __forceinline__ void eval_segment(const EvalCtx& c, int rs, int lo, int hi, const WV& w, float2 (&g)[N::NS],
                                             double& lp, int lane) {
    constexpr int U = OCBNN_POINT_UNROLL;
    for (int j = lo + lane; j < hi; j += U * G) {
        const float* rec[U];
        float live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = j + u * G < hi;
            rec[u] = c.tab + (ok ? j + u * G : j) * rs;
            live[u] = ok ? 1.f : 0.f;
        }
        point_eval<N, KIND, U, WV>(w, g, lp, rec, live, c.fc);
    }
}

// log posterior value and gradient at w.  Block-uniform control flow (contains __syncthreads when the point set
// does not fit one tile or `restage` is set).  On return every lane of the group holds the full g and lp.
template <class N, int G, class WV>
__device__ __forceinline__ void evaluate(const EvalArgs& a, const EvalCtx& c, bool restage, const WV& w, float2 (&g)[N::NS],
                                         float& lp_out, int lane) {
#pragma unroll
    for (int s = 0; s < N::NS; ++s) g[s] = f2(0.f);
    double lp = 0.0;
    const bool multi = c.npts > c.tile_cap;
    for (int begin = 0; begin < c.npts; begin += c.tile_cap) {
        const int count = min(c.tile_cap, c.npts - begin);
        if (multi || restage) {
            __syncthreads();
            stage_points(c.tab, a, c.nb, c.mult, begin, count);
            __syncthreads();
        }
        int seg_lo = 0;
#pragma unroll
        for (int sg = 0; sg < 4; ++sg) {
            const int lo = max(seg_lo, begin) - begin, hi = min(c.seg_end[sg], begin + count) - begin;
            seg_lo = c.seg_end[sg];
            if (lo >= hi) continue;
            if (sg == 0) {
                if (a.lik_head == OCBNN_HEAD_LIK_GAUSS) eval_segment<N, G, OCBNN_HEAD_LIK_GAUSS, WV>(c, a.rs, lo, hi, w, g, lp, lane);
                else if (a.lik_head == OCBNN_HEAD_LIK_SOFTMAX) eval_segment<N, G, OCBNN_HEAD_LIK_SOFTMAX, WV>(c, a.rs, lo, hi, w, g, lp, lane);
                else eval_segment<N, G, OCBNN_HEAD_LIK_BERNOULLI, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            } else if (sg == 1) {
                eval_segment<N, G, OCBNN_HEAD_POS_GAUSS, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            } else if (sg == 2) {
                eval_segment<N, G, OCBNN_HEAD_NEG_EXP, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            } else {
                eval_segment<N, G, OCBNN_HEAD_DIRICHLET, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            }
        }
    }
    if (G > 1) {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
            for (int s = 0; s < N::NS; ++s) {
                g[s].x += __shfl_xor_sync(0xffffffffu, g[s].x, off);
                g[s].y += __shfl_xor_sync(0xffffffffu, g[s].y, off);
            }
            lp += __shfl_xor_sync(0xffffffffu, lp, off);
        }
    }
    // Gaussian prior over the weights (isotropic, or the AOCP diagonal Gaussian): base.py:106-133.
    // Padding slots have 1/var = 0 and are masked so that they never move.
    double pr = 0.0;
#pragma unroll
    for (int s = 0; s < N::NS; ++s) {
        const float4 mv = c.pri[s];
        const float2 diff = __fadd2_rn(w[s], make_float2(-mv.x, -mv.y));
        const float2 t = __fmul2_rn(diff, make_float2(mv.z, mv.w));
        g[s] = __fadd2_rn(g[s], make_float2(-t.x, -t.y));
        const float2 e = __fmul2_rn(diff, t);
        pr += (double)e.x + (double)e.y;
        if (N::cidx(s, 0) < 0) g[s].x = 0.f;
        if (N::cidx(s, 1) < 0) g[s].y = 0.f;
    }
    lp += c.lp_const - 0.5 * pr;
    lp_out = (float)lp;
}

}  // namespace ocbnn
