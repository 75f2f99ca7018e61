// small_kernels.cuh — K1 (log-posterior + gradient), K2 (fused HMC trajectory + Metropolis-Hastings) and
// K4 (fused SGLD steps) for the register-resident one-hidden-layer nets.  Included by one .cu per net shape.
#pragma once
#include <cstdlib>
#include "small_net.cuh"

namespace ocbnn {

constexpr int kSmallThreads = 128;
constexpr int kPointUnroll = OCBNN_POINT_UNROLL;
// measured on B200 (W2, 262144 chains, G chain-leapfrog-steps/s): (MB,U) = (4,2) 2.03, (3,4) 2.06, (4,4) 2.05, (5,2) 1.98, (5,1) 1.93,
// (3,6) 1.95.  (4,2) is the default: within 1.5 % of the best and it keeps short point segments (6 training points) in full groups.
constexpr int kDefaultHmcOwn = 1;                    // 16 / 32 lanes per chain: k_hmc_own with shuffle (1) or shared-memory (2) exchanges
constexpr int kDefaultHmcMb = 4, kDefaultHmcU = 2;   // points in flight per thread in K1 / K4

__host__ __device__ constexpr int kSmemFixedFloats(int ns) { return 4 * ns; }   // float4 prior table per slot

// ---- K1 ------------------------------------------------------------------------------------------------
template <class N, int G>
__global__ void __launch_bounds__(kSmallThreads)
k_logpost_small(EvalArgs a, const float* __restrict__ w, long long m, float* __restrict__ out_lp, float* __restrict__ out_grad,
                int tile_cap) {
    extern __shared__ __align__(16) float smem[];
    float4* pri_s = reinterpret_cast<float4*>(smem);
    float* tab = smem + kSmemFixedFloats(N::NS);
    EvalCtx c = make_ctx(a, tab, pri_s, tile_cap);
    load_pri_slots<N>(pri_s, a);
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    const bool active = gid < m;
    const long long chain = active ? gid : m - 1;
    float2 wv[N::NS], g[N::NS];
    load_state<N>(wv, w + chain * N::D);
    float lp;
    evaluate<N, G, kPointUnroll>(a, c, /*restage=*/true, RegView<N::NS>{wv}, g, lp, lane);
    if (active && lane == 0) {
        if (out_lp) out_lp[chain] = lp;
        if (out_grad) store_state<N>(g, out_grad + chain * N::D);
    }
}

// ---- K2 ------------------------------------------------------------------------------------------------
// HMC (bnn/inference.py:39-78):  p ~ injected;  U0,K0;  p -= eps/2 dU;  L x { q += eps p ; [l<L] p -= eps dU };
// p -= eps/2 dU;  U1,K1;  accept iff u < exp(U0 - U1 + K0 - K1)  (fp32 exp, float64 compare, strict <).
// dU = -grad log posterior, so "p -= c*dU" is "p += c*g" with identical rounding (fl(c*g) then one add, like the
// reference's `y += c * x` on fp32 tensors).  The value+gradient at q_L is reused as U0/grad of the next iteration (the
// reference recomputes both).
//
// One launch runs n_it iterations x L leapfrog steps of every chain.  The master copies of q and p of every thread live in
// its own column of shared memory (slot-major float2, conflict-free): they are touched once per leapfrog step, while the
// evaluator keeps its own (scaled) copy of the weights, the gradient accumulators and the per-point temporaries in
// registers.  The trajectory is a single loop around ONE inlined evaluate() (phase l = 0 is the evaluation at the start
// of an iteration, needed at launch and after a restore-on-reject), which keeps the kernel small enough for the
// instruction cache and the register allocator.
template <class N, int G, int MB, int U, bool WSM>
__global__ void __launch_bounds__(kSmallThreads, MB)
k_hmc_small(EvalArgs a, float* __restrict__ q_io, const float* __restrict__ p0, const double* __restrict__ u, long long m, int n_it,
            float eps, float eps_half, int L, int restore, uint8_t* __restrict__ out_acc, float* __restrict__ out_h,
            int* __restrict__ out_flags, float* __restrict__ out_samples, int thin, int tile_cap) {
    extern __shared__ __align__(16) float smem[];
    float4* pri_s = reinterpret_cast<float4*>(smem);
    float2* qs = reinterpret_cast<float2*>(smem + kSmemFixedFloats(N::NS)) + threadIdx.x;     // [NS][kSmallThreads]
    float2* ps = qs + N::NS * kSmallThreads;
    float* tab = smem + kSmemFixedFloats(N::NS) + 4 * N::NS * kSmallThreads;
    EvalCtx c = make_ctx(a, tab, pri_s, tile_cap);
    load_pri_slots<N>(pri_s, a);
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    const bool active = gid < m;
    const long long chain = active ? gid : m - 1;
    const bool writer = active && lane == 0;
    const bool multi = c.npts > c.tile_cap;
    const SmemView qv{qs, kSmallThreads};

    auto load_to = [&](float2* dst, const float* __restrict__ src) {
#pragma unroll
        for (int s = 0; s < N::NS; ++s)
            dst[s * kSmallThreads] = make_float2(N::cidx(s, 0) >= 0 ? __ldg(src + N::cidx(s, 0)) : 0.f, N::cidx(s, 1) >= 0 ? __ldg(src + N::cidx(s, 1)) : 0.f);
    };
    auto store_from = [&](const float2* srcv, float* __restrict__ dst) {
#pragma unroll
        for (int s = 0; s < N::NS; ++s) {
            const float2 v = srcv[s * kSmallThreads];
            if (N::cidx(s, 0) >= 0) dst[N::cidx(s, 0)] = v.x;
            if (N::cidx(s, 1) >= 0) dst[N::cidx(s, 1)] = v.y;
        }
    };

    float2 g[N::NS];
    float lp = 0.f, U0 = 0.f, K0 = 0.f;
    int nan_flag = 0, it = 0, l = 0;
    bool first = true;
    load_to(qs, q_io + chain * N::D);

    // start of iteration `it`: momenta in, U0/K0, backup of q for restore-on-reject, first half kick
    auto begin_iteration = [&]() {
        const float* __restrict__ src = p0 + ((long long)it * m + chain) * N::D;
        U0 = -lp;
        if (restore) {
            if (writer) store_from(qs, q_io + chain * N::D);
            __syncwarp();
        }
        double ks = 0.0;
#pragma unroll
        for (int s = 0; s < N::NS; ++s) {
            const float2 pv = make_float2(N::cidx(s, 0) >= 0 ? __ldg(src + N::cidx(s, 0)) : 0.f, N::cidx(s, 1) >= 0 ? __ldg(src + N::cidx(s, 1)) : 0.f);
            const float2 e = __fmul2_rn(pv, pv);
            ks += (double)e.x + (double)e.y;
            ps[s * kSmallThreads] = __fadd2_rn(pv, __fmul2_rn(f2(eps_half), g[s]));
        }
        K0 = 0.5f * (float)ks;
    };

    for (;;) {
        if (l >= 1) {
#pragma unroll
            for (int s = 0; s < N::NS; ++s) qs[s * kSmallThreads] = __fadd2_rn(qs[s * kSmallThreads], __fmul2_rn(f2(eps), ps[s * kSmallThreads]));
        }
        evaluate<N, G, U, SmemView, WSM>(a, c, first, qv, g, lp, lane);      // the first call stages the (fixed) point set
        first = false;
        if (l == 0) {
            begin_iteration();
            l = 1;
            continue;
        }
        {
            const float2 e2 = f2((l < L) ? eps : eps_half);
#pragma unroll
            for (int s = 0; s < N::NS; ++s) ps[s * kSmallThreads] = __fadd2_rn(ps[s * kSmallThreads], __fmul2_rn(e2, g[s]));
        }
        if (l < L) {
            ++l;
            continue;
        }
        // end of the trajectory: Metropolis-Hastings
        bool isn = false;
        double ks = 0.0;
#pragma unroll
        for (int s = 0; s < N::NS; ++s) {
            const float2 v = qs[s * kSmallThreads];
            isn |= isnan(v.x) | isnan(v.y);
            const float2 pv = ps[s * kSmallThreads];
            const float2 e = __fmul2_rn(pv, pv);
            ks += (double)e.x + (double)e.y;
        }
        nan_flag |= isn ? 1 : 0;
        const float U1 = -lp;
        const float K1 = 0.5f * (float)ks;
        const float dH = __fsub_rn(__fadd_rn(__fsub_rn(U0, U1), K0), K1);
        const double alpha = (double)expf(dH);
        const double ui = __ldg(u + (long long)it * m + chain);
        const bool acc = ui < alpha;
        if (writer) {
            if (out_acc) out_acc[(long long)it * m + chain] = acc ? 1 : 0;
            if (out_h) reinterpret_cast<float4*>(out_h)[(long long)it * m + chain] = make_float4(U0, K0, U1, K1);
        }
        const bool need_eval = restore && !acc;
        if (need_eval) load_to(qs, q_io + chain * N::D);
        if (out_samples && thin > 0 && ((it + 1) % thin == 0) && writer)
            store_from(qs, out_samples + ((long long)((it + 1) / thin - 1) * m + chain) * N::D);
        if (++it == n_it) break;
        // a restored chain needs lp and g at the restored state; with a multi-tile point set evaluate() contains
        // __syncthreads, so the decision must then be block-uniform
        const bool again = multi ? (__syncthreads_or(need_eval ? 1 : 0) != 0) : (__any_sync(0xffffffffu, need_eval) != 0);
        if (again) {
            l = 0;
            continue;
        }
        begin_iteration();
        l = 1;
    }
    if (writer) {
        store_from(qs, q_io + chain * N::D);
        if (out_flags) out_flags[chain] = nan_flag;
    }
}

// ---- K2, few chains: 16 / 32 lanes per chain with slot ownership ---------------------------------------------------------------------
// With a few thousand chains the one-lane and 8-lane kernels above leave < 2 warps per scheduler and run at the latency of one
// leapfrog step.  More lanes per chain hide that latency, but k_hmc_small replicates the whole leapfrog update, the prior and the
// gradient butterfly (log2 G x NS float2 shuffles) on every lane - at 16 / 32 lanes that costs more than it hides (measured:
// 0.72 / 0.47 G steps/s against 0.88 G at 8 lanes).  Here lane r of a chain's 16-lane group OWNS slot r (one float2 of q, p, g):
//   * the gradient partials of the lanes meet by a recursive-halving REDUCE-SCATTER (8 + 4 + 2 + 1 float2 exchanges instead of
//     4 x 16), after which each lane holds the total of its own slot;
//   * moments -> gradient, prior and the leapfrog update touch the own slot only (one extra exchange brings the neighbouring
//     moment the transform needs);
//   * the evaluation copy of all weights is rebuilt by an all-gather of the owned slots (16 float2 shuffles) each step;
//   * q and p live in registers (no shared-memory columns); 32 lanes = two 16-lane groups that split the points and first add
//     their partials.
// Single-input single-output rbf nets (the moment form of the gradient, use_moments<N>) with <= 16 slots: the [1,10,1] repro nets.
// XS = true: the reduce-scatter and the all-gather go through shared memory (one STS.64 per slot and lane, then each lane sums its own
// slot over the 16 lanes: 16 LDS.64 at an odd slot stride, conflict-free per half-warp) instead of shuffles + selects.
constexpr int kOwnXsFloat2PerWarp = 16 * 33 + 2 * 16;      // exchange buffer of one warp: [16 slots][33] partials + [2 groups][16] weights
template <class N, int G, int MB, int U, bool XS>
__global__ void __launch_bounds__(kSmallThreads, MB)
k_hmc_own(EvalArgs a, float* __restrict__ q_io, const float* __restrict__ p0, const double* __restrict__ u, long long m, int n_it,
          float eps, float eps_half, int L, int restore, uint8_t* __restrict__ out_acc, float* __restrict__ out_h,
          int* __restrict__ out_flags, float* __restrict__ out_samples, int thin, int tile_cap) {
    static_assert(use_moments<N>() && N::NS <= 16 && (G == 16 || G == 32), "slot-ownership HMC: moment-form nets with <= 16 slots, 16 or 32 lanes");
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) float smem[];
    float4* pri_s = reinterpret_cast<float4*>(smem);
    float2* xs = reinterpret_cast<float2*>(smem + kSmemFixedFloats(N::NS)) + (threadIdx.x >> 5) * kOwnXsFloat2PerWarp;   // this warp's exchange buffer
    float* tab = smem + kSmemFixedFloats(N::NS) + (XS ? 2 * kOwnXsFloat2PerWarp * (kSmallThreads / 32) : 0);
    EvalCtx c = make_ctx(a, tab, pri_s, tile_cap);
    load_pri_slots<N>(pri_s, a);
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    const int wl = threadIdx.x & 31;                             // lane of the warp (a warp holds two 16-lane groups when G = 16)
    const int own = lane & 15;                                   // the slot this lane owns
    const bool active = gid < m;
    const long long chain = active ? gid : m - 1;
    const bool owner = active && lane < 16 && own < N::NS;       // writes its slot; lane 0 also writes the per-chain scalars
    const bool multi = c.npts > c.tile_cap;
    const int i0 = own < N::NS ? N::cidx(own, 0) : -1, i1 = own < N::NS ? N::cidx(own, 1) : -1;
    const int hp = own % N::HP;

    auto load_own = [&](const float* __restrict__ src) { return make_float2(i0 >= 0 ? __ldg(src + i0) : 0.f, i1 >= 0 ? __ldg(src + i1) : 0.f); };
    auto store_own = [&](float2 v, float* __restrict__ dst) {
        if (i0 >= 0) dst[i0] = v.x;
        if (i1 >= 0) dst[i1] = v.y;
    };
    auto sum16 = [&](double v) {                                  // over the 16 lanes of the group (every lane gets the same bits)
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
        return v;
    };

    float2 q_own = load_own(q_io + chain * N::D), p_own = f2(0.f), g_own = f2(0.f), q0_own = q_own;
    float lp = 0.f, U0 = 0.f, K0 = 0.f;
    int nan_flag = 0, it = 0, l = 0;
    bool first = true;
    __syncthreads();                                              // pri_s

    // value and gradient at q: the gradient of the OWN slot in g_own, the value on every lane
    auto evaluate_own = [&]() {
        float2 w[N::NS], g[N::NS];
        float2* wbuf = xs + 16 * 33 + (wl & 16);                  // this group's 16 evaluation weights
        if (XS) {
            // all-gather through shared memory: the owner stores its slot in evaluation units, everyone reads the 16 slots
            __syncwarp();
            wbuf[own] = (own < N::SW2 && act_in_scale<N::ACT>() != 1.f) ? __fmul2_rn(q_own, f2(act_in_scale<N::ACT>())) : q_own;
            __syncwarp();
#pragma unroll
            for (int s = 0; s < N::NS; ++s) {
                w[s] = wbuf[s];
                g[s] = f2(0.f);
            }
        } else {
#pragma unroll
            for (int s = 0; s < N::NS; ++s) {                     // all-gather of the owned slots -> evaluation copy (W1, b1 scaled)
                const float2 v = make_float2(__shfl_sync(kFull, q_own.x, s, 16), __shfl_sync(kFull, q_own.y, s, 16));
                w[s] = (s < N::SW2 && act_in_scale<N::ACT>() != 1.f) ? __fmul2_rn(v, f2(act_in_scale<N::ACT>())) : v;
                g[s] = f2(0.f);
            }
        }
        double lpd = 0.0;
        eval_points<N, G, U>(a, c, first, RegW<N::NS>{w}, g, lpd, lane);
        float2 acc, w1s = f2(0.f), b1s = f2(0.f), w2s = f2(0.f);
        if (XS) {
            // reduce-scatter through shared memory: slot s of warp lane L at xs[s * 33 + L] (odd slot stride: the 16 lanes of a
            // half-warp read 16 different bank pairs); lane r then sums slot r over the lanes of its chain
            __syncwarp();
#pragma unroll
            for (int s = 0; s < N::NS; ++s) xs[s * 33 + wl] = g[s];
            __syncwarp();
            const float2* col = xs + (own < N::NS ? own : 0) * 33 + (G == 32 ? 0 : (wl & 16));
            float2 t0 = f2(0.f), t1 = f2(0.f);
#pragma unroll
            for (int j = 0; j < (G == 32 ? 32 : 16); j += 2) {
                t0 = __fadd2_rn(t0, col[j]);
                t1 = __fadd2_rn(t1, col[j + 1]);
            }
            acc = __fadd2_rn(t0, t1);
            w1s = wbuf[N::SW1 + hp];                              // this slot's hidden pair: evaluation weights as stored by their owners
            b1s = wbuf[N::SB1 + hp];
            w2s = wbuf[N::SW2 + hp];
        } else {
            float2 r[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) r[s] = s < N::NS ? g[s] : f2(0.f);
            if (G == 32) {
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    r[s].x += __shfl_xor_sync(kFull, r[s].x, 16);
                    r[s].y += __shfl_xor_sync(kFull, r[s].y, 16);
                }
            }
            // recursive halving: after the stage with offset o the lane keeps the slots whose index agrees with its own in bit o
            float2 r8[8], r4[4], r2[2];
            {
                const bool up = (lane & 8) != 0;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const float2 send = up ? r[s] : r[s + 8], keep = up ? r[s + 8] : r[s];
                    r8[s] = make_float2(keep.x + __shfl_xor_sync(kFull, send.x, 8), keep.y + __shfl_xor_sync(kFull, send.y, 8));
                }
            }
            {
                const bool up = (lane & 4) != 0;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const float2 send = up ? r8[s] : r8[s + 4], keep = up ? r8[s + 4] : r8[s];
                    r4[s] = make_float2(keep.x + __shfl_xor_sync(kFull, send.x, 4), keep.y + __shfl_xor_sync(kFull, send.y, 4));
                }
            }
            {
                const bool up = (lane & 2) != 0;
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const float2 send = up ? r4[s] : r4[s + 2], keep = up ? r4[s + 2] : r4[s];
                    r2[s] = make_float2(keep.x + __shfl_xor_sync(kFull, send.x, 2), keep.y + __shfl_xor_sync(kFull, send.y, 2));
                }
            }
            {
                const bool up = (lane & 1) != 0;
                const float2 send = up ? r2[0] : r2[1], keep = up ? r2[1] : r2[0];
                acc = make_float2(keep.x + __shfl_xor_sync(kFull, send.x, 1), keep.y + __shfl_xor_sync(kFull, send.y, 1));
            }
#pragma unroll
            for (int h = 0; h < N::HP; ++h)
                if (hp == h) {
                    w1s = w[N::SW1 + h];
                    b1s = w[N::SB1 + h];
                    w2s = w[N::SW2 + h];
                }
        }
        // moments -> gradient (see evaluate()): the W1 slot needs A1 of the b1 slot, the b1 slot A0 of the W2 slot - both HP lanes up
        const float2 other = make_float2(__shfl_sync(kFull, acc.x, (own + N::HP) & 15, 16), __shfl_sync(kFull, acc.y, (own + N::HP) & 15, 16));
        float2 gm = acc;
        if (own < N::SW2) gm = __ffma2_rn(w1s, acc, __fmul2_rn(b1s, other));
        // Gaussian prior on the own slot, rescaling of the data gradient (evaluate())
        float pr = 0.f;
        if (own < N::NS) {
            const float4 mv = c.pri[own];
            const float2 diff = __fadd2_rn(q_own, make_float2(-mv.x, -mv.y));
            const float2 t = __fmul2_rn(diff, make_float2(mv.z, mv.w));
            if (own < N::SW2) gm = __ffma2_rn(gm, __fmul2_rn(w2s, f2(grad_scale_for<N::ACT, true>())), make_float2(-t.x, -t.y));
            else gm = __fadd2_rn(gm, make_float2(-t.x, -t.y));
            const float2 e = __fmul2_rn(diff, t);
            pr = e.x + e.y;
        }
        if (i0 < 0) gm.x = 0.f;
        if (i1 < 0) gm.y = 0.f;
        g_own = gm;
        if (lane < 16) lpd -= 0.5 * (double)pr;
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) lpd += __shfl_xor_sync(kFull, lpd, off);
        lp = (float)(lpd + c.lp_const);
    };

    auto begin_iteration = [&]() {
        const float2 pv = load_own(p0 + ((long long)it * m + chain) * N::D);
        U0 = -lp;
        q0_own = q_own;                                           // for restore-on-reject
        const float2 e = __fmul2_rn(pv, pv);
        K0 = 0.5f * (float)sum16((double)e.x + (double)e.y);
        p_own = __fadd2_rn(pv, __fmul2_rn(f2(eps_half), g_own));
    };

    for (;;) {
        if (l >= 1) q_own = __fadd2_rn(q_own, __fmul2_rn(f2(eps), p_own));
        evaluate_own();
        first = false;
        if (l == 0) {
            begin_iteration();
            l = 1;
            continue;
        }
        p_own = __fadd2_rn(p_own, __fmul2_rn(f2((l < L) ? eps : eps_half), g_own));
        if (l < L) {
            ++l;
            continue;
        }
        // end of the trajectory: Metropolis-Hastings (same arithmetic as k_hmc_small)
        const bool isn = __any_sync(kFull, isnan(q_own.x) | isnan(q_own.y));       // (a 32-lane warp holds one or two chains: a NaN in
        nan_flag |= isn ? 1 : 0;                                                    //  either flags both - diagnostics only)
        const float2 e = __fmul2_rn(p_own, p_own);
        const float U1 = -lp;
        const float K1 = 0.5f * (float)sum16((double)e.x + (double)e.y);
        const float dH = __fsub_rn(__fadd_rn(__fsub_rn(U0, U1), K0), K1);
        const double alpha = (double)expf(dH);
        const double ui = __ldg(u + (long long)it * m + chain);
        const bool acc = ui < alpha;
        if (active && lane == 0) {
            if (out_acc) out_acc[(long long)it * m + chain] = acc ? 1 : 0;
            if (out_h) reinterpret_cast<float4*>(out_h)[(long long)it * m + chain] = make_float4(U0, K0, U1, K1);
        }
        const bool need_eval = restore && !acc;
        if (need_eval) q_own = q0_own;
        if (out_samples && thin > 0 && ((it + 1) % thin == 0) && owner) store_own(q_own, out_samples + ((long long)((it + 1) / thin - 1) * m + chain) * N::D);
        if (++it == n_it) break;
        const bool again = multi ? (__syncthreads_or(need_eval ? 1 : 0) != 0) : (__any_sync(kFull, need_eval) != 0);
        if (again) {
            l = 0;
            continue;
        }
        begin_iteration();
        l = 1;
    }
    if (owner) store_own(q_own, q_io + chain * N::D);
    if (active && lane == 0 && out_flags) out_flags[chain] = nan_flag;
}

// ---- K4 ------------------------------------------------------------------------------------------------
// SGLD (bnn/inference.py:304-340): upd = (grad log prior + grad log lik) * (eps_t/2) + noise_t ; w += upd.
template <class N, int G>
__global__ void __launch_bounds__(kSmallThreads)
k_sgld_small(EvalArgs a, float* __restrict__ w_io, const float* __restrict__ noise, const float* __restrict__ half_eps,
             const int* __restrict__ offsets, long long m, int t_steps, float* __restrict__ out_samples, int thin, int tile_cap) {
    extern __shared__ __align__(16) float smem[];
    float4* pri_s = reinterpret_cast<float4*>(smem);
    float* tab = smem + kSmemFixedFloats(N::NS);
    load_pri_slots<N>(pri_s, a);
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    const bool active = gid < m;
    const long long chain = active ? gid : m - 1;
    const bool writer = active && lane == 0;
    float2 w[N::NS], g[N::NS];
    load_state<N>(w, w_io + chain * N::D);
    const bool batched = a.batch_stride > 1;
    EvalCtx c = make_ctx(a, tab, pri_s, tile_cap);
    for (int t = 0; t < t_steps; ++t) {
        if (batched) {
            a.batch_offset = offsets[t];
            c = make_ctx(a, tab, pri_s, tile_cap);
        }
        float lp;
        evaluate<N, G, kPointUnroll>(a, c, /*restage=*/batched || t == 0, RegView<N::NS>{w}, g, lp, lane);
        const float2 he = f2(half_eps[t]);
        const float* __restrict__ nz = noise + ((long long)t * m + chain) * N::D;
#pragma unroll
        for (int s = 0; s < N::NS; ++s) {
            const float2 nv = make_float2(N::cidx(s, 0) >= 0 ? __ldg(nz + N::cidx(s, 0)) : 0.f, N::cidx(s, 1) >= 0 ? __ldg(nz + N::cidx(s, 1)) : 0.f);
            w[s] = __fadd2_rn(w[s], __fadd2_rn(__fmul2_rn(g[s], he), nv));
        }
        if (out_samples && thin > 0 && ((t + 1) % thin == 0) && writer)
            store_state<N>(w, out_samples + ((long long)((t + 1) / thin - 1) * m + chain) * N::D);
    }
    if (writer) store_state<N>(w, w_io + chain * N::D);
}

// ---- forward (predict) -----------------------------------------------------------------------------------
template <class N>
__global__ void __launch_bounds__(kSmallThreads)
k_forward_small(const float* __restrict__ w, long long m, const float* __restrict__ x, long long npts, float* __restrict__ out) {
    const long long chain = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (chain >= m) return;
    float2 wv[N::NS];
    load_state<N>(wv, w + chain * N::D);
    for (long long j = 0; j < npts; ++j) {
        float xv[N::S0], f[N::SK];
#pragma unroll
        for (int d = 0; d < N::S0; ++d) xv[d] = __ldg(x + j * N::S0 + d);
        point_forward<N>(wv, xv, f);
#pragma unroll
        for (int k = 0; k < N::SK; ++k) out[(chain * npts + j) * N::SK + k] = f[k];
    }
}

// ---- host-side launchers for one net shape ---------------------------------------------------------------
template <class N>
struct SmallLaunch {
    static size_t smem_bytes(const Problem& p, int npts, int& tile_cap, size_t reserved = 0) {
        size_t fixed = sizeof(float) * kSmemFixedFloats(N::NS) + reserved;
        int cap = (int)((kMaxSmemTable - fixed) / (sizeof(float) * p.rs));
        tile_cap = npts < cap ? (npts > 0 ? npts : 1) : cap;
        return fixed - reserved + sizeof(float) * (size_t)tile_cap * p.rs;
    }
    static int npts(const EvalArgs& a) {
        return (a.with_data ? batch_count(a.n_train, 0, a.batch_stride) : 0) + a.n_cpoints;   // largest batch: offset 0
    }

    template <int G>
    static int logpost(const Problem& p, const EvalArgs& a, const float* w, long long m, float* lp, float* grad, cudaStream_t st) {
        int tile_cap;
        size_t sm = smem_bytes(p, npts(a), tile_cap);
        auto k = k_logpost_small<N, G>;
        OCBNN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        long long blocks = (m * G + kSmallThreads - 1) / kSmallThreads;
        k<<<(unsigned)blocks, kSmallThreads, sm, st>>>(a, w, m, lp, grad, tile_cap);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    }
    template <int G, int MB, int U, bool WSM = false>
    static int hmc_mb(const Problem& p, const EvalArgs& a, float* q, const float* p0, const double* u, long long m, int n_it, float eps,
                      float eps_half, int L, int restore, uint8_t* acc, float* h, int* flags, float* samples, int thin, cudaStream_t st) {
        int tile_cap;
        const size_t extra = sizeof(float) * 4 * N::NS * kSmallThreads;        // q and p columns of every thread
        size_t sm = smem_bytes(p, npts(a), tile_cap, extra) + extra;
        auto k = k_hmc_small<N, G, MB, U, WSM>;
        OCBNN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        long long blocks = (m * G + kSmallThreads - 1) / kSmallThreads;
        k<<<(unsigned)blocks, kSmallThreads, sm, st>>>(a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin,
                                                       tile_cap);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    }
    template <int G, int MB, int U, bool XS = false>
    static int hmc_own(const Problem& p, const EvalArgs& a, float* q, const float* p0, const double* u, long long m, int n_it, float eps,
                       float eps_half, int L, int restore, uint8_t* acc, float* h, int* flags, float* samples, int thin, cudaStream_t st) {
        if constexpr (use_moments<N>() && N::NS <= 16 && (G == 16 || G == 32)) {
            int tile_cap;
            const size_t extra = XS ? sizeof(float2) * kOwnXsFloat2PerWarp * (kSmallThreads / 32) : 0;     // per-warp exchange buffers
            size_t sm = smem_bytes(p, npts(a), tile_cap, extra) + extra;
            auto k = k_hmc_own<N, G, MB, U, XS>;
            OCBNN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            long long blocks = (m * G + kSmallThreads - 1) / kSmallThreads;
            k<<<(unsigned)blocks, kSmallThreads, sm, st>>>(a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, tile_cap);
            OCBNN_LAUNCH_CHECK();
            return OCBNN_OK;
        } else {
            return OCBNN_EUNSUPPORTED;
        }
    }
    // One chain per thread (G = 1, the large-M case) has tuning variants selected by OCBNN_HMC_MB (resident CTAs per SM
    // the kernel is compiled for: register cap 65536 / (128 MB)) and OCBNN_HMC_U (points in flight per thread); measured
    // defaults in kDefaultHmc*.  Lane-split chains (G > 1, few chains) use one configuration.
    // One chain per thread (G = 1, the large-M case): variants selected by OCBNN_HMC_MB (resident CTAs per SM the kernel is
    // compiled for: register cap 65536 / (128 MB)) and OCBNN_HMC_U (points in flight per thread); measured defaults in
    // kDefaultHmc*.  Lane-split chains (G > 1, few chains) run the same kernel with the gradient combined by a shuffle butterfly.  (A variant
    // with per-lane slot ownership and a shared-memory reduction removed ~40 % of the per-lane instructions but was 25 % SLOWER at
    // 4096 chains: at <2 warps per scheduler the serial shared-memory round trips cost more than the replicated arithmetic.)
    template <int G>
    static int hmc(const Problem& p, const EvalArgs& a, float* q, const float* p0, const double* u, long long m, int n_it, float eps,
                   float eps_half, int L, int restore, uint8_t* acc, float* h, int* flags, float* samples, int thin, cudaStream_t st) {
        if (G == 1) {
            // wide nets (32 float2 slots for [2,10,3]: 64 registers of weights + 64 of accumulators) spill at 128 registers/thread:
            // they default to 2 CTAs/SM (255 registers)
            // (measured, W3 = F2a classifier + Dirichlet COCP: 0.21 G steps/s at (4,2) with 4 KB of spills, 0.76 G at (2,2)); [2,10,1] (21 slots)
            // gets 3 CTAs/SM (168 registers, no spills)
            static const int mb = [] { const char* e = getenv("OCBNN_HMC_MB"); return e ? atoi(e) : (N::NS > 24 ? 2 : N::NS > 18 ? 3 : kDefaultHmcMb); }();
            static const int uu = [] { const char* e = getenv("OCBNN_HMC_U"); return e ? atoi(e) : kDefaultHmcU; }();
#define OCBNN_HMC_GO(MBV, UV) return hmc_mb<1, MBV, UV>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, st)
            // wide nets: OCBNN_HMC_WSM=1 selects the variant that reads the weights from the thread's shared-memory column (WSM) at
            // 3 CTAs/SM.  Measured on W3 after the MUFU heads: register weights at 2 CTAs/SM 0.974 G steps/s, WSM at 3 CTAs/SM 0.890 G
            // (45 extra LDS per point), WSM at 4 CTAs/SM spills (0.53-0.63 G).  Default: register weights.
            static const bool wsm = N::NS > 24 && getenv("OCBNN_HMC_WSM") != nullptr;
            if (wsm) return hmc_mb<1, (N::NS > 24 ? 3 : 4), 2, (N::NS > 24)>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, st);
            if (mb <= 2) { if (uu <= 1) OCBNN_HMC_GO(2, 1); OCBNN_HMC_GO(2, 2); }
            if (mb == 3) { if (uu <= 1) OCBNN_HMC_GO(3, 1); if (uu <= 2) OCBNN_HMC_GO(3, 2); OCBNN_HMC_GO(3, 4); }
            if (uu >= 4) OCBNN_HMC_GO(3, 4);
            // long point segments: four points in flight at 3 CTAs/SM are ~1 % faster than two at 4 (W2: 2.06 vs 2.04 G steps/s); short
            // segments (W1: 6 points) lose 8 % to the partial groups
            static const bool auto_cfg = !getenv("OCBNN_HMC_MB") && !getenv("OCBNN_HMC_U");
            if (auto_cfg && N::NS <= 18 && a.n_cpoints >= 64) OCBNN_HMC_GO(3, 4);
            OCBNN_HMC_GO(4, 2);
#undef OCBNN_HMC_GO
        }
        if constexpr ((G == 16 || G == 32) && use_moments<N>() && N::NS <= 16) {
            // 16 / 32 lanes per chain: slot ownership (k_hmc_own).  OCBNN_HMC_OWN=0 selects the replicated k_hmc_small; OCBNN_HMC_GU the points in flight
            // OCBNN_HMC_OWN: 0 replicated kernel, 1 ownership with shuffle exchanges, 2 ownership with shared-memory exchanges
            static const int own_mode = [] { const char* e = getenv("OCBNN_HMC_OWN"); return e ? atoi(e) : kDefaultHmcOwn; }();
            static const int gu = [] { const char* e = getenv("OCBNN_HMC_GU"); return e ? atoi(e) : 0; }();
            if (own_mode == 2) return hmc_own<G, 4, 2, true>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, st);
            if (own_mode) {
                if (gu == 4) return hmc_own<G, 3, 4>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, st);
                return hmc_own<G, 4, 2>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, st);
            }
        }
        if constexpr (G > 1 && G <= 8 && N::NS <= 18) {
            // few chains, 2-8 lanes per chain: four points in flight per lane at 3 CTAs/SM, as on the one-lane path (4096 chains x 8 lanes on
            // W2: 0.915 vs 0.872 G steps/s; at 16 lanes the segments are too short: 0.62 vs 0.74, so G >= 16 keeps two).  OCBNN_HMC_GU=2 | 4 forces.
            static const int gu = [] { const char* e = getenv("OCBNN_HMC_GU"); return e ? atoi(e) : 0; }();
            if (gu == 4 || (gu == 0 && a.n_cpoints >= 64))
                return hmc_mb<G, 3, 4>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, st);
        }
        return hmc_mb<G, (N::NS > 24 ? 2 : N::NS > 18 ? 3 : 4), 2>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, st);
    }
    template <int G>
    static int sgld(const Problem& p, const EvalArgs& a, float* w, const float* noise, const float* he, const int* offs, long long m,
                    int t_steps, float* samples, int thin, cudaStream_t st) {
        int tile_cap;
        size_t sm = smem_bytes(p, npts(a), tile_cap);
        auto k = k_sgld_small<N, G>;
        OCBNN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        long long blocks = (m * G + kSmallThreads - 1) / kSmallThreads;
        k<<<(unsigned)blocks, kSmallThreads, sm, st>>>(a, w, noise, he, offs, m, t_steps, samples, thin, tile_cap);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    }
    static int forward(const float* w, long long m, const float* x, long long np, float* out, cudaStream_t st) {
        long long blocks = (m + kSmallThreads - 1) / kSmallThreads;
        k_forward_small<N><<<(unsigned)blocks, kSmallThreads, 0, st>>>(w, m, x, np, out);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    }
};

}  // namespace ocbnn

#define OCBNN_LANE_CASES(EXPR)                                                                   \
    switch (lanes) {                                                                             \
        case 1: { constexpr int G = 1; return EXPR; }                                            \
        case 2: { constexpr int G = 2; return EXPR; }                                            \
        case 4: { constexpr int G = 4; return EXPR; }                                            \
        case 8: { constexpr int G = 8; return EXPR; }                                            \
        case 16: { constexpr int G = 16; return EXPR; }                                          \
        case 32: { constexpr int G = 32; return EXPR; }                                          \
        default: set_error("lanes must be 1,2,4,8,16 or 32 (got %d)", lanes); return OCBNN_EINVAL; \
    }

// Defines the non-template entry points of one net shape (collected in the table of small_dispatch.cu).
#define OCBNN_DEFINE_SMALL_NET(NAME, S0, H, SK, ACT)                                                                          \
    namespace ocbnn {                                                                                                         \
    using NAME##_N = Net1<S0, H, SK, ACT>;                                                                                    \
    int NAME##_logpost(const Problem& p, const EvalArgs& a, const float* w, long long m, float* lp, float* grad, int lanes,    \
                       cudaStream_t st) {                                                                                     \
        OCBNN_LANE_CASES((SmallLaunch<NAME##_N>::template logpost<G>(p, a, w, m, lp, grad, st)))                             \
    }                                                                                                                         \
    int NAME##_hmc(const Problem& p, const EvalArgs& a, float* q, const float* p0, const double* u, long long m, int n_it,     \
                   float eps, float eps_half, int L, int restore, uint8_t* acc, float* h, int* flags, float* samples,          \
                   int thin, int lanes, cudaStream_t st) {                                                                    \
        OCBNN_LANE_CASES((SmallLaunch<NAME##_N>::template hmc<G>(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h,  \
                                                                  flags, samples, thin, st)))                                 \
    }                                                                                                                         \
    int NAME##_sgld(const Problem& p, const EvalArgs& a, float* w, const float* noise, const float* he, const int* offs,       \
                    long long m, int t_steps, float* samples, int thin, int lanes, cudaStream_t st) {                         \
        OCBNN_LANE_CASES((SmallLaunch<NAME##_N>::template sgld<G>(p, a, w, noise, he, offs, m, t_steps, samples, thin, st)))  \
    }                                                                                                                         \
    int NAME##_forward(const float* w, long long m, const float* x, long long np, float* out, cudaStream_t st) {              \
        return SmallLaunch<NAME##_N>::forward(w, m, x, np, out, st);                                                          \
    }                                                                                                                         \
    }
