// heads.cuh — per-point "heads": value and d/df of the likelihood / output-constrained prior term at one
// point, given the network output f (SK values).  All heads return the UNWEIGHTED term; the caller scales
// by the record weight.  Reference formulas: bnn/base.py (line numbers at each head).
#pragma once
#include "common.cuh"

namespace ocbnn {

struct HeadConsts {
    float lik_scale;     // 1/(2 var)
    float tau0x2, tau1x2;  // 2*tau
    float pos_inv_var;
};

// Phi(z) = 1/4 (tanh(-tau0 z)+1)(tanh(-tau1 z)+1)  (base.py:360-362), evaluated in the cancellation-free form
// Phi = 1 / ((1+e^{2 tau0 z})(1+e^{2 tau1 z})) with one reciprocal; exponents clamped at 40 (Phi < 2e-35 there, the
// reference's fp32 tanh form returns exactly 0).  Returns Phi and dPhi/dz.
__device__ __forceinline__ void phi_classifier(float z, const HeadConsts& c, float& P, float& dP) {
    float e0 = expf(fminf(c.tau0x2 * z, 40.f));
    float e1 = expf(fminf(c.tau1x2 * z, 40.f));
    float A = 1.f + e0, B = 1.f + e1;
    float r = __frcp_rn(A * B);
    P = r;
    dP = -(r * r) * (c.tau0x2 * e0 * B + c.tau1x2 * e1 * A);
}

// negative_exponential_cocp (base.py:364-388): sum over the forbidden gaps (a,b) of Phi(a-f) Phi(f-b).
// Internal row layout (built by ocbnn_problem_create from the ABI's gap list): n1, n2 (int bits), n1 x (s c, s) one-sided terms
// Phi(s (f - c)) [gaps with one infinite end], n2 x (a, b) two-sided terms.
__device__ __forceinline__ void head_neg_exp(float f, const float* __restrict__ prm, const HeadConsts& c, float& val, float& df) {
    const int n1 = __float_as_int(prm[0]), n2 = __float_as_int(prm[1]);
    val = 0.f;
    df = 0.f;
    const float* q = prm + 2;
    for (int k = 0; k < n1; ++k, q += 2) {
        float P, D;
        phi_classifier(fmaf(q[1], f, -q[0]), c, P, D);
        val += P;
        df = fmaf(q[1], D, df);
    }
    for (int k = 0; k < n2; ++k, q += 2) {
        float P1, D1, P2, D2;
        phi_classifier(q[0] - f, c, P1, D1);
        phi_classifier(f - q[1], c, P2, D2);
        val = fmaf(P1, P2, val);
        df += P1 * D2 - D1 * P2;
    }
}

// positive_gaussian_cocp (base.py:344-358): logsumexp_k -(f-y_k)^2/(2 sigma_c)  [sigma_c is a VARIANCE, :354];
// the constants -1/2 log(2 pi sigma_c) + log(1/K) live in prior_const.
__device__ __forceinline__ void head_pos_gauss(float f, const float* __restrict__ prm, const HeadConsts& c, float& val, float& df) {
    int K = (int)prm[0];
    float iv = c.pos_inv_var;
    if (K == 1) {
        float r = f - prm[1];
        val = -0.5f * r * r * iv;
        df = -r * iv;
        return;
    }
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
        float r = f - prm[1 + k];
        mx = fmaxf(mx, -0.5f * r * r * iv);
    }
    float se = 0.f, sd = 0.f;
    for (int k = 0; k < K; ++k) {
        float r = f - prm[1 + k];
        float e = expf(-0.5f * r * r * iv - mx);
        se += e;
        sd = fmaf(e, -r * iv, sd);
    }
    val = mx + logf(se);
    df = sd * __frcp_rn(se);
}

// RegressorMixin.log_likelihood (base.py:252-265): -(f-y)^2 * lik_scale per output (constants in lik_const).
template <int SK>
__device__ __forceinline__ void head_lik_gauss(const float (&f)[SK], const float* __restrict__ prm, const HeadConsts& c, float& val,
                                               float (&df)[SK]) {
    val = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        float r = f[k] - prm[k];
        val = fmaf(-r * r, c.lik_scale, val);
        df[k] = -2.f * c.lik_scale * r;
    }
}

template <int SK>
__device__ __forceinline__ void log_softmax(const float (&f)[SK], float (&p)[SK], float (&logp)[SK]) {
    float mx = f[0];
#pragma unroll
    for (int k = 1; k < SK; ++k) mx = fmaxf(mx, f[k]);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        p[k] = expf(f[k] - mx);
        se += p[k];
    }
    float ls = logf(se), rs = __frcp_rn(se);
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        logp[k] = f[k] - mx - ls;
        p[k] *= rs;
    }
}

// ClassifierMixin.log_likelihood (base.py:432-444): log_softmax(f)[y]
template <int SK>
__device__ __forceinline__ void head_lik_softmax(const float (&f)[SK], const float* __restrict__ prm, float& val, float (&df)[SK]) {
    float p[SK], logp[SK];
    log_softmax<SK>(f, p, logp);
    int y = (int)prm[0];
    val = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        bool hit = (k == y);
        val = hit ? logp[k] : val;
        df[k] = (hit ? 1.f : 0.f) - p[k];
    }
}

// positive_dirichlet_cocp (base.py:498-511): sum_k (conc_k - 1) log softmax(f)_k (lgamma terms in prior_const)
template <int SK>
__device__ __forceinline__ void head_dirichlet(const float (&f)[SK], const float* __restrict__ prm, float& val, float (&df)[SK]) {
    float p[SK], logp[SK];
    log_softmax<SK>(f, p, logp);
    float tot = 0.f;
    val = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k) {
        float cm1 = prm[k];
        tot += cm1;
        val = fmaf(cm1, logp[k], val);
    }
#pragma unroll
    for (int k = 0; k < SK; ++k) df[k] = prm[k] - tot * p[k];
}

// BinaryClassifierMixin.log_likelihood (base.py:524-542): t log s(z) + (1-t) log(1-s(z)).  Evaluated in the
// overflow-free form t*z - softplus(z); the reference's naive e^z/(e^z+1) returns NaN/-inf for |z| >~ 17..88.
__device__ __forceinline__ void head_lik_bernoulli(float z, const float* __restrict__ prm, float& val, float& df) {
    float t = prm[0];
    float e = expf(-fabsf(z));               // in (0,1]
    float sp = fmaxf(z, 0.f) + log1pf(e);    // softplus(z)
    float r = __frcp_rn(1.f + e);
    float s = (z >= 0.f) ? r : e * r;        // sigmoid(z)
    val = fmaf(t, z, -sp);
    df = t - s;
}

// dispatch on the record's head kind
template <int SK>
__device__ __forceinline__ void eval_head(int kind, const float (&f)[SK], const float* __restrict__ prm, const HeadConsts& c,
                                          float& val, float (&df)[SK]) {
#pragma unroll
    for (int k = 0; k < SK; ++k) df[k] = 0.f;
    val = 0.f;
    switch (kind) {
        case OCBNN_HEAD_LIK_GAUSS: head_lik_gauss<SK>(f, prm, c, val, df); break;
        case OCBNN_HEAD_NEG_EXP: head_neg_exp(f[0], prm, c, val, df[0]); break;
        case OCBNN_HEAD_POS_GAUSS: head_pos_gauss(f[0], prm, c, val, df[0]); break;
        case OCBNN_HEAD_DIRICHLET: head_dirichlet<SK>(f, prm, val, df); break;
        case OCBNN_HEAD_LIK_SOFTMAX: head_lik_softmax<SK>(f, prm, val, df); break;
        case OCBNN_HEAD_LIK_BERNOULLI: head_lik_bernoulli(f[0], prm, val, df[0]); break;
        default: break;
    }
}

}  // namespace ocbnn
