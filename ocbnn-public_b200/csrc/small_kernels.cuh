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
