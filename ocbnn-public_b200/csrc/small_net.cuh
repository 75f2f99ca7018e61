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

// ---- activation, evaluated in SCALED pre-activation units --------------------------------------------------------
// The evaluator works on z' = c z with c = sqrt(log2 e) folded into W1 and b1 once per evaluation, so that the rbf
// activation exp(-z^2) = 2^(-z'^2) is one FMUL2 + two MUFU.EX2 (the negation is an operand modifier) and its derivative
// d/dz = -2 z a = -(2/c) z' a.  The constant -(2/c) c = ... is applied once to the accumulated W1/b1 gradients:
// dL/dW1 = act_grad_scale * sum_points dL/da * (z' a) * x   with act_grad_scale = -2 sqrt(ln 2).
template <int ACT>
__host__ __device__ constexpr float act_in_scale() { return ACT == OCBNN_ACT_RBF ? 1.2011224087864498f : 1.f; }
template <int ACT>
__host__ __device__ constexpr float act_grad_scale() { return ACT == OCBNN_ACT_RBF ? -1.6651092223153954f : 1.f; }
// the same two constants when the evaluator reads UNSCALED weights (wide nets keep them in shared memory, see SmemView)
template <int ACT, bool SCALED>
__host__ __device__ constexpr float grad_scale_for() { return SCALED ? act_grad_scale<ACT>() : (ACT == OCBNN_ACT_RBF ? -2.f : 1.f); }

// returns a = act(z) and m with da/dz' (up to act_grad_scale) = m
template <int ACT, bool SCALED = true>
__device__ __forceinline__ void act_pair(float2 zs, float2& a, float2& m) {
    if (ACT == OCBNN_ACT_RBF) {                       // exp(-z^2), base.py:84
        float2 t = __fmul2_rn(zs, zs);
        if (!SCALED) t = __fmul2_rn(t, f2(1.4426950408889634f));
        a = make_float2(ex2_approx(-t.x), ex2_approx(-t.y));
        m = __fmul2_rn(zs, a);
    } else if (ACT == OCBNN_ACT_RELU) {               // max(z,0), base.py:86
        a = make_float2(fmaxf(zs.x, 0.f), fmaxf(zs.y, 0.f));
        m = make_float2(zs.x > 0.f ? 1.f : 0.f, zs.y > 0.f ? 1.f : 0.f);
    } else {
        a = zs;
        m = f2(1.f);
    }
}

// Weight views: the evaluator reads the chain's packed weights through operator[] so that the master copy can live in
// registers (RegView) or in a per-thread column of shared memory (SmemView: slot s of thread t at base[s * stride],
// conflict-free float2 loads).  The evaluation copy (W1, b1 scaled) is always built in registers.
template <int NS>
struct RegView {
    const float2 (&a)[NS];
    __device__ __forceinline__ float2 operator[](int s) const { return a[s]; }
};
template <int NS>
struct RegW {                                   // evaluation weights in registers, W1 / b1 pre-scaled
    static constexpr bool kScaled = true;
    const float2 (&a)[NS];
    __device__ __forceinline__ float2 operator[](int s) const { return a[s]; }
};
struct SmemView {                               // as evaluation weights: unscaled, one LDS.64 per use
    static constexpr bool kScaled = false;
    const float2* base;
    int stride;
    // volatile asm: every use is its own load (a plain load at the top of an evaluation would be kept live in 2 x NS
    // registers until the prior term at the end, which is exactly what this view is meant to avoid)
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
    float zmax;          // Phi argument clamp: 40 / (2 max tau)  (Phi < e^-40 beyond; the reference's fp32 tanh form is exactly 0)
    float lik_scale, pos_inv_var;
};

// Phi(z) = 1/((1+e^{2 tau0 z})(1+e^{2 tau1 z})) and dPhi/dz (see heads.cuh)
__device__ __forceinline__ void phi_fast(float z, const FastConsts& c, float& P, float& dP) {
    const float2 u = __fmul2_rn(f2(fminf(z, c.zmax)), c.tau_l2e);
    const float2 e = make_float2(ex2_approx(u.x), ex2_approx(u.y));
    const float2 ab = __fadd2_rn(e, f2(1.f));
    const float r = rcp_approx(ab.x * ab.y);
    const float2 t = __fmul2_rn(e, c.tau2);
    const float s = fmaf(t.x, ab.y, t.y * ab.x);
    P = r;
    dP = -(r * r) * s;
}

// two Phi evaluations packed in f32x2 lanes: 4 EX2 + 2 RCP, every other op a packed instruction
__device__ __forceinline__ void phi_fast2(float2 z, const FastConsts& c, float2& P, float2& dP) {
    z = make_float2(fminf(z.x, c.zmax), fminf(z.y, c.zmax));
    const float2 u0 = __fmul2_rn(z, f2(c.tau_l2e.x)), u1 = __fmul2_rn(z, f2(c.tau_l2e.y));
    const float2 e0 = make_float2(ex2_approx(u0.x), ex2_approx(u0.y));
    const float2 e1 = make_float2(ex2_approx(u1.x), ex2_approx(u1.y));
    const float2 a = __fadd2_rn(e0, f2(1.f)), b = __fadd2_rn(e1, f2(1.f));
    const float2 ab = __fmul2_rn(a, b);
    const float2 r = make_float2(rcp_approx(ab.x), rcp_approx(ab.y));
    const float2 s = __ffma2_rn(__fmul2_rn(e0, f2(c.tau2.x)), b, __fmul2_rn(__fmul2_rn(e1, f2(c.tau2.y)), a));
    P = r;
    dP = __fmul2_rn(__fmul2_rn(r, make_float2(-r.x, -r.y)), s);
}

// phi_fast2 with ONE reciprocal for the pair: r = 1/(ab.x ab.y), 1/ab.x = r ab.y, 1/ab.y = r ab.x (two FMULs instead of
// a MUFU op; MUFU is the binding pipe of the toy-net kernels).  ab <= (1+e^40)^2 per lane, so the product can overflow only
// when both lanes are saturated; then r = 0 and both Phi (< e^-40, exactly 0 in the reference's fp32 form) come out as 0.
__device__ __forceinline__ void phi_fast2_1r(float2 z, const FastConsts& c, float2& P, float2& dP) {
    z = make_float2(fminf(z.x, c.zmax), fminf(z.y, c.zmax));
    const float2 u0 = __fmul2_rn(z, f2(c.tau_l2e.x)), u1 = __fmul2_rn(z, f2(c.tau_l2e.y));
    const float2 e0 = make_float2(ex2_approx(u0.x), ex2_approx(u0.y));
    const float2 e1 = make_float2(ex2_approx(u1.x), ex2_approx(u1.y));
    const float2 a = __fadd2_rn(e0, f2(1.f)), b = __fadd2_rn(e1, f2(1.f));
    const float2 ab = __fmul2_rn(a, b);
    const float rx = rcp_approx(ab.x * ab.y);
    const float2 r = make_float2(rx * ab.y, rx * ab.x);
    const float2 s = __ffma2_rn(__fmul2_rn(e0, f2(c.tau2.x)), b, __fmul2_rn(__fmul2_rn(e1, f2(c.tau2.y)), a));
    P = r;
    dP = __fmul2_rn(__fmul2_rn(r, make_float2(-r.x, -r.y)), s);
}

template <int KIND, int SK>
struct FastHead;

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
//   n1, n2 (int bits), n1 x (s c, s) one-sided terms Phi(s f - s c), n2 x (a, b) two-sided terms Phi(a - f) Phi(f - b)
template <int SK>
struct FastHead<OCBNN_HEAD_NEG_EXP, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts& c, float& val, float (&df)[SK]) {
        const int n1 = __float_as_int(prm[0]), n2 = __float_as_int(prm[1]);
        val = 0.f;
        float d = 0.f;
        const float* q = prm + 2;
        for (int k = 0; k < n1; ++k, q += 2) {
            float P, D;
            phi_fast(fmaf(q[1], f[0], -q[0]), c, P, D);
            val += P;
            d = fmaf(q[1], D, d);
        }
        for (int k = 0; k < n2; ++k, q += 2) {
            float2 P, D;
            phi_fast2(make_float2(q[0] - f[0], f[0] - q[1]), c, P, D);
            val = fmaf(P.x, P.y, val);
            d += P.x * D.y - D.x * P.y;
        }
#pragma unroll
        for (int k = 0; k < SK; ++k) df[k] = 0.f;
        df[0] = d;
    }
};
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
// log-softmax with MUFU approximations (ex2 / lg2 / rcp, ~2^-22 relative): ~25 instructions for SK = 3 instead of ~100 with
// expf / logf / IEEE reciprocal; the classifier kernels are fma-pipe/issue bound, and the heads were 40 % of their instructions
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int SK>
__device__ __forceinline__ void log_softmax_fast(const float (&f)[SK], float (&p)[SK], float (&logp)[SK]) {
    float mx = f[0];
#pragma unroll
    for (int k = 1; k < SK; ++k) mx = fmaxf(mx, f[k]);
    float t[SK], se = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        t[k] = (f[k] - mx) * 1.4426950408889634f;
        p[k] = ex2_approx(t[k]);
        se += p[k];
    }
    const float l2 = lg2_approx(se), rs = rcp_approx(se);
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        logp[k] = (t[k] - l2) * 0.6931471805599453f;
        p[k] *= rs;
    }
}
// positive_dirichlet_cocp (base.py:498-511), see heads.cuh::head_dirichlet
template <int SK>
struct FastHead<OCBNN_HEAD_DIRICHLET, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts&, float& val, float (&df)[SK]) {
        float p[SK], logp[SK];
        log_softmax_fast<SK>(f, p, logp);
        float tot = 0.f;
        val = 0.f;
#pragma unroll
        for (int k = 0; k < SK; ++k) {
            tot += prm[k];
            val = fmaf(prm[k], logp[k], val);
        }
#pragma unroll
        for (int k = 0; k < SK; ++k) df[k] = fmaf(-tot, p[k], prm[k]);
    }
};
// ClassifierMixin.log_likelihood (base.py:432-444), see heads.cuh::head_lik_softmax
template <int SK>
struct FastHead<OCBNN_HEAD_LIK_SOFTMAX, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts&, float& val, float (&df)[SK]) {
        float p[SK], logp[SK];
        log_softmax_fast<SK>(f, p, logp);
        const int y = (int)prm[0];
        val = 0.f;
#pragma unroll
        for (int k = 0; k < SK; ++k) {
            const bool hit = k == y;
            val = hit ? logp[k] : val;
            df[k] = (hit ? 1.f : 0.f) - p[k];
        }
    }
};
// BinaryClassifierMixin.log_likelihood (base.py:524-542) in the overflow-free form t z - softplus(z), see heads.cuh::head_lik_bernoulli
template <int SK>
struct FastHead<OCBNN_HEAD_LIK_BERNOULLI, SK> {
    static __device__ __forceinline__ void run(const float (&f)[SK], const float* __restrict__ prm, const FastConsts&, float& val, float (&df)[SK]) {
#pragma unroll
        for (int k = 0; k < SK; ++k) df[k] = 0.f;
        const float z = f[0], t = prm[0];
        const float e = ex2_approx(-fabsf(z) * 1.4426950408889634f);        // in (0,1]
        const float ope = 1.f + e;
        const float sp = fmaf(lg2_approx(ope), 0.6931471805599453f, fmaxf(z, 0.f));   // softplus(z)
        const float r = rcp_approx(ope);
        val = fmaf(t, z, -sp);
        df[0] = t - ((z >= 0.f) ? r : e * r);
    }
};

// Head policies: run() evaluates the head of U points at once so that their serial MUFU chains interleave.
template <int KIND>
struct HeadEach {                                  // any head kind, one point after the other
    template <int SK, int U>
    static __device__ __forceinline__ void run(const float (&f)[U][SK], const float* const (&prm)[U], const FastConsts& c, float (&val)[U],
                                               float (&df)[U][SK]) {
#pragma unroll
        for (int u = 0; u < U; ++u) FastHead<KIND, SK>::run(f[u], prm[u], c, val[u], df[u]);
    }
};
// Negative COCP whose rows all have the same term structure (N1 one-sided, N2 two-sided terms; found by
// ocbnn_problem_create): the N1 + 2 N2 Phi evaluations of the U points are packed pairwise into f32x2 lanes, no loop, no
// per-row counts.  F1a/example.yaml (permitted [2.5,3]) is <2,0>; F3b (forbidden gap (1,2.5)) is <0,1>.
template <int N1, int N2>
struct HeadNegFixed {
    template <int SK, int U>
    static __device__ __forceinline__ void run(const float (&f)[U][SK], const float* const (&prm)[U], const FastConsts& c, float (&val)[U],
                                               float (&df)[U][SK]) {
        constexpr int NP = N1 + 2 * N2, T = U * NP;
        float z[T], sg[U][N1 > 0 ? N1 : 1], P[T], D[T];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float* q = prm[u] + 2;
#pragma unroll
            for (int k = 0; k < N1; ++k) {
                sg[u][k] = q[2 * k + 1];
                z[u * NP + k] = fmaf(sg[u][k], f[u][0], -q[2 * k]);
            }
#pragma unroll
            for (int k = 0; k < N2; ++k) {
                z[u * NP + N1 + 2 * k] = q[2 * N1 + 2 * k] - f[u][0];
                z[u * NP + N1 + 2 * k + 1] = f[u][0] - q[2 * N1 + 2 * k + 1];
            }
        }
#pragma unroll
        for (int i = 0; i + 1 < T; i += 2) {
            float2 Pp, Dp;
            phi_fast2_1r(make_float2(z[i], z[i + 1]), c, Pp, Dp);
            P[i] = Pp.x, P[i + 1] = Pp.y, D[i] = Dp.x, D[i + 1] = Dp.y;
        }
        if (T & 1) phi_fast(z[T - 1], c, P[T - 1], D[T - 1]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v = 0.f, d = 0.f;
#pragma unroll
            for (int k = 0; k < N1; ++k) {
                v = k == 0 ? P[u * NP] : v + P[u * NP + k];
                d = k == 0 ? sg[u][0] * D[u * NP] : fmaf(sg[u][k], D[u * NP + k], d);
            }
#pragma unroll
            for (int k = 0; k < N2; ++k) {
                const int i = u * NP + N1 + 2 * k;
                const float pd = P[i] * D[i + 1] - D[i] * P[i + 1];
                if (N1 == 0 && k == 0) {
                    v = P[i] * P[i + 1];
                    d = pd;
                } else {
                    v = fmaf(P[i], P[i + 1], v);
                    d += pd;
                }
            }
            val[u] = v;
#pragma unroll
            for (int k = 0; k < SK; ++k) df[u][k] = 0.f;
            df[u][0] = d;
        }
    }
};

#ifndef OCBNN_POINT_UNROLL
#define OCBNN_POINT_UNROLL 2
#endif

// value + gradient contribution of U point records with a compile-time head.  `w` are the EVALUATION weights (W1, b1 in
// scaled units, see act_in_scale); g accumulates the W1/b1 gradients in the same scaled units (evaluate() rescales once).
// The U points are independent until the gradient accumulation, which gives the scheduler U forward/head dependency
// chains to interleave.  Per hidden pair and point: S0 + 3 packed fma-pipe instructions + 2 MUFU forward, S0 + 2 backward
// for single-output nets (dL/dz' = df w2 m: the factor w2 is constant over the points, so the W1/b1 accumulators hold
// sum df m [x] and evaluate() multiplies by w2 once), S0 + 2 SK + 2 otherwise.
// Single-input, single-output rbf nets ([1,H,1], every 1-D toy of the reference): dL/dz'_h = df w2_h z'_h a_h with
// z'_h = w1'_h x + b1'_h, so the sums over points that the W1/b1 gradients need are moments of a in x,
//   sum_p df_p z'_h a_h [x_p] = w1'_h A_{1[+1]} + b1'_h A_{0[+1]},   A_k = sum_p df_p x_p^k a_h(x_p),   A_0 = dL/dw2_h,
// and the per-point work is three FFMA2 per hidden pair (A0, A1, A2) with neither z' nor z' a kept: one packed multiply and
// two saved registers less per hidden pair and point.  evaluate() turns (A0, A1, A2) into the gradient once per evaluation.
template <class N>
__host__ __device__ constexpr bool use_moments() { return N::S0 == 1 && N::SK == 1 && N::ACT == OCBNN_ACT_RBF; }

template <class N, class HEAD, int U, class WV>
__device__ __forceinline__ void point_eval(const WV& w, float2 (&g)[N::NS], float& vacc, const float* const (&rec)[U],
                                           const FastConsts& fc) {
    constexpr bool MOM = use_moments<N>() && WV::kScaled;
    float2 a[U][N::HP], m[U][MOM ? 1 : N::HP];
    float x[U][N::S0], f[U][N::SK], df[U][N::SK], val[U];
    const float* prm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int d = 0; d < N::S0; ++d) x[u][d] = rec[u][d];
        prm[u] = rec[u] + N::S0 + 2;
        float2 facc[N::SK];
#pragma unroll
        for (int k = 0; k < N::SK; ++k) facc[k] = f2(0.f);
#pragma unroll
        for (int hp = 0; hp < N::HP; ++hp) {
            float2 z = w[N::SB1 + hp];
#pragma unroll
            for (int d = 0; d < N::S0; ++d) z = __ffma2_rn(w[N::SW1 + d * N::HP + hp], f2(x[u][d]), z);
            if (MOM) {
                const float2 t = __fmul2_rn(z, z);
                a[u][hp] = make_float2(ex2_approx(-t.x), ex2_approx(-t.y));
            } else {
                act_pair<N::ACT, WV::kScaled>(z, a[u][hp], m[u][MOM ? 0 : hp]);
            }
#pragma unroll
            for (int k = 0; k < N::SK; ++k)
                facc[k] = (hp == 0) ? __fmul2_rn(w[N::SW2 + k * N::HP], a[u][0]) : __ffma2_rn(w[N::SW2 + k * N::HP + hp], a[u][hp], facc[k]);
        }
#pragma unroll
        for (int k = 0; k < N::SK; ++k) {
            const float b2 = (k & 1) ? w[N::SB2 + k / 2].y : w[N::SB2 + k / 2].x;
            f[u][k] = (facc[k].x + facc[k].y) + b2;
        }
    }
    HEAD::template run<N::SK, U>(f, prm, fc, val, df);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float wgt = rec[u][N::S0 + 1];
        vacc = fmaf(wgt, val[u], vacc);
#pragma unroll
        for (int k = 0; k < N::SK; ++k) df[u][k] *= wgt;
    }
#pragma unroll
    for (int k = 0; k < N::SK; ++k) {
        float s = df[0][k];
#pragma unroll
        for (int u = 1; u < U; ++u) s += df[u][k];
        if (k & 1) g[N::SB2 + k / 2].y += s; else g[N::SB2 + k / 2].x += s;
    }
    if (MOM) {
        float2 d0[U], d1[U], d2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float dx = df[u][0] * x[u][0];
            d0[u] = f2(df[u][0]);
            d1[u] = f2(dx);
            d2[u] = f2(dx * x[u][0]);
        }
#pragma unroll
        for (int hp = 0; hp < N::HP; ++hp) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                g[N::SW2 + hp] = __ffma2_rn(d0[u], a[u][hp], g[N::SW2 + hp]);     // A0
                g[N::SB1 + hp] = __ffma2_rn(d1[u], a[u][hp], g[N::SB1 + hp]);     // A1
                g[N::SW1 + hp] = __ffma2_rn(d2[u], a[u][hp], g[N::SW1 + hp]);     // A2
            }
        }
        return;
    }
#pragma unroll
    for (int hp = 0; hp < N::HP; ++hp) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float2 mm = m[u][MOM ? 0 : hp];
            if (N::SK == 1) {
                const float2 dk = f2(df[u][0]);
                g[N::SW2 + hp] = __ffma2_rn(dk, a[u][hp], g[N::SW2 + hp]);
                g[N::SB1 + hp] = __ffma2_rn(dk, mm, g[N::SB1 + hp]);
#pragma unroll
                for (int d = 0; d < N::S0; ++d) g[N::SW1 + d * N::HP + hp] = __ffma2_rn(f2(df[u][0] * x[u][d]), mm, g[N::SW1 + d * N::HP + hp]);
            } else {
                float2 da = f2(0.f);
#pragma unroll
                for (int k = 0; k < N::SK; ++k) {
                    const float2 dk = f2(df[u][k]);
                    da = (k == 0) ? __fmul2_rn(dk, w[N::SW2 + hp]) : __ffma2_rn(dk, w[N::SW2 + k * N::HP + hp], da);
                    g[N::SW2 + k * N::HP + hp] = __ffma2_rn(dk, a[u][hp], g[N::SW2 + k * N::HP + hp]);
                }
                const float2 dz = __fmul2_rn(da, mm);
                g[N::SB1 + hp] = __fadd2_rn(g[N::SB1 + hp], dz);
#pragma unroll
                for (int d = 0; d < N::S0; ++d) g[N::SW1 + d * N::HP + hp] = __ffma2_rn(dz, f2(x[u][d]), g[N::SW1 + d * N::HP + hp]);
            }
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
    c.fc.zmax = 40.f / fmaxf(2.f * fmaxf(a.tau0, a.tau1), 1e-30f);
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

// points [lo, hi) of the staged tile with one compile-time head: lane r of the chain's G lanes takes points lo + r + k G,
// U at a time; the (at most one per lane) leftover point goes through the U = 1 instantiation.
// The point values are added in fp32 over up to 4 trips (4 U points) and then into the float64 sum: one F2F + DADD per 4
// trips instead of per trip (F2F issues on the XU pipe, which the MUFU ops of these kernels already load the most).
template <class N, int G, int U, class HEAD, class WV>
__device__ __forceinline__ void eval_segment(const EvalCtx& c, int rs, int lo, int hi, const WV& w, float2 (&g)[N::NS],
                                             double& lp, int lane) {
    int j = lo + lane;
    float vacc = 0.f;
    while (j + (U - 1) * G < hi) {
#pragma unroll 1
        for (int trip = 0; trip < 4 && j + (U - 1) * G < hi; ++trip, j += U * G) {
            const float* rec[U];
#pragma unroll
            for (int u = 0; u < U; ++u) rec[u] = c.tab + (j + u * G) * rs;
            point_eval<N, HEAD, U, WV>(w, g, vacc, rec, c.fc);
        }
        lp += (double)vacc;
        vacc = 0.f;
    }
    if (U > 1) {
        for (; j < hi; j += G) {
            const float* rec[1] = {c.tab + j * rs};
            point_eval<N, HEAD, 1, WV>(w, g, vacc, rec, c.fc);
        }
    }
    lp += (double)vacc;
}

// Point loop of one evaluation: raw accumulators g (see point_eval / use_moments) and the float64 value sum of the points
// lane `lane` of the chain's G lanes is responsible for.  Block-uniform control flow (contains __syncthreads when the point
// set does not fit one tile or `restage` is set).
template <class N, int G, int U, class WV>
__device__ __forceinline__ void eval_points(const EvalArgs& a, const EvalCtx& c, bool restage, const WV& w, float2 (&g)[N::NS],
                                            double& lp, int lane) {
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
                if (a.lik_head == OCBNN_HEAD_LIK_GAUSS) eval_segment<N, G, U, HeadEach<OCBNN_HEAD_LIK_GAUSS>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
                else if (a.lik_head == OCBNN_HEAD_LIK_SOFTMAX) eval_segment<N, G, U, HeadEach<OCBNN_HEAD_LIK_SOFTMAX>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
                else eval_segment<N, G, U, HeadEach<OCBNN_HEAD_LIK_BERNOULLI>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            } else if (sg == 1) {
                eval_segment<N, G, U, HeadEach<OCBNN_HEAD_POS_GAUSS>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            } else if (sg == 2) {
                if (a.neg_n1 == 2 && a.neg_n2 == 0) eval_segment<N, G, U, HeadNegFixed<2, 0>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
                else if (a.neg_n1 == 0 && a.neg_n2 == 1) eval_segment<N, G, U, HeadNegFixed<0, 1>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
                else if (a.neg_n1 == 1 && a.neg_n2 == 0) eval_segment<N, G, U, HeadNegFixed<1, 0>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
                else eval_segment<N, G, U, HeadEach<OCBNN_HEAD_NEG_EXP>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            } else {
                eval_segment<N, G, U, HeadEach<OCBNN_HEAD_DIRICHLET>, WV>(c, a.rs, lo, hi, w, g, lp, lane);
            }
        }
    }
}

// log posterior value and gradient at the weights behind the view `q`.  Block-uniform control flow (see eval_points).  On
// return every lane of the group holds the full g and lp.
template <class N, int G, int U, class QV, bool WSM = false>
__device__ __forceinline__ void evaluate(const EvalArgs& a, const EvalCtx& c, bool restage, const QV& q, float2 (&g)[N::NS],
                                         float& lp_out, int lane) {
    // WSM: the point loop reads the (unscaled) weights through the view q itself - one LDS.64 per use - instead of a scaled
    // register copy.  For nets whose 2 x NS weight + accumulator registers do not fit next to the point temporaries.
    constexpr bool SC = !WSM;
    double lp = 0.0;
#pragma unroll
    for (int s = 0; s < N::NS; ++s) g[s] = f2(0.f);
    float2 w[WSM ? 1 : N::NS];                        // evaluation copy: W1 and b1 in scaled units (act_in_scale)
    if constexpr (WSM) {
        eval_points<N, G, U>(a, c, restage, q, g, lp, lane);
    } else {
#pragma unroll
        for (int s = 0; s < N::NS; ++s) {
            const float2 v = q[s];
            w[WSM ? 0 : s] = (s < N::SW2 && act_in_scale<N::ACT>() != 1.f) ? __fmul2_rn(v, f2(act_in_scale<N::ACT>())) : v;
        }
        eval_points<N, G, U>(a, c, restage, RegW<WSM ? 1 : N::NS>{w}, g, lp, lane);
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
    if (use_moments<N>() && !WSM) {      // (A0, A1, A2) in the (W2, b1, W1) slots -> sum_p df z' a [x] (see point_eval); W2 slot already final
#pragma unroll
        for (int hp = 0; hp < N::HP; ++hp) {
            const float2 A0 = g[N::SW2 + hp], A1 = g[N::SB1 + hp], A2 = g[N::SW1 + hp];
            g[N::SB1 + hp] = __ffma2_rn(w[WSM ? 0 : N::SW1 + hp], A1, __fmul2_rn(w[WSM ? 0 : N::SB1 + hp], A0));
            g[N::SW1 + hp] = __ffma2_rn(w[WSM ? 0 : N::SW1 + hp], A2, __fmul2_rn(w[WSM ? 0 : N::SB1 + hp], A1));
        }
    }
    // Gaussian prior over the weights (isotropic, or the AOCP diagonal Gaussian): base.py:106-133, added to the rescaled
    // data gradient.  Padding slots have 1/var = 0 and are masked so that they never move.
    float pr = 0.f;              // fp32 like the reference's Normal.log_prob(...).sum(); added once to the float64 value
#pragma unroll
    for (int s = 0; s < N::NS; ++s) {
        const float4 mv = c.pri[s];
        const float2 diff = __fadd2_rn(q[s], make_float2(-mv.x, -mv.y));
        const float2 t = __fmul2_rn(diff, make_float2(mv.z, mv.w));
        if (s < N::SW2 && N::SK == 1) {               // accumulated without the w2 factor (point_eval)
            const float2 sc = __fmul2_rn(q[N::SW2 + (s % N::HP)], f2(grad_scale_for<N::ACT, SC>()));
            g[s] = __ffma2_rn(g[s], sc, make_float2(-t.x, -t.y));
        } else if (s < N::SW2 && grad_scale_for<N::ACT, SC>() != 1.f) {
            g[s] = __ffma2_rn(g[s], f2(grad_scale_for<N::ACT, SC>()), make_float2(-t.x, -t.y));
        } else {
            g[s] = __fadd2_rn(g[s], make_float2(-t.x, -t.y));
        }
        const float2 e = __fmul2_rn(diff, t);
        pr += e.x + e.y;
        if (N::cidx(s, 0) < 0) g[s].x = 0.f;
        if (N::cidx(s, 1) < 0) g[s].y = 0.f;
    }
    lp += c.lp_const - 0.5 * (double)pr;
    lp_out = (float)lp;
}

}  // namespace ocbnn
