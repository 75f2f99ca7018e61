// gen_tc.cu — K1 (log posterior + gradient) for the application-sized MLPs on tcgen05 / TMEM.
//
// One CTA per weight vector, points in tiles of 128 (= the 128 TMEM lanes).  Every layer is three small GEMMs
//     GEMM1  Z_l   [p][n] = sum_k A_{l-1}[p][k] W_l[k][n]          M = 128 points, N = n, K = k     (forward)
//     GEMM2  dA_l-1[p][k] = sum_n dZ_l[p][n]   W_l[k][n]           M = 128 points, N = k, K = n     (input gradient)
//     GEMM3  dW_l  [k][n] = sum_p A_{l-1}[p][k] dZ_l[p][n]         M = k (64|128), N = n, K = points (weight gradient, accumulated in
//                                                                  TMEM over the point tiles of the weight vector)
// issued by one elected thread as tcgen05.mma kind::f16 with fp32 accumulators in TMEM.  The bias is a weight row: A carries a
// column of ones, so b_l is row s_{l-1} of W_l (exactly the reference's packed layout, base.py:72-77) and db_l falls out of GEMM3.
//
// fp32 accuracy from bf16 tensor cores: every fp32 operand x is split into three bf16 parts x = x0 + x1 + x2 (24 mantissa bits) and a
// product is the six MMAs x0y0 + x0y1 + x1y0 + x1y1 + x0y2 + x2y0 (dropped terms <= 2^-26 relative; bf16 x bf16 is exact in the fp32
// accumulator).  Six bf16 MMAs cost the tensor pipe what three tf32 MMAs do, and the parts take 6 bytes per element instead of 8.
//
// Shared-memory layout: ONE array per matrix serves both operand roles.  With 16-bit elements and no swizzle a core matrix is
// 8 rows x 16 bytes; stored "core-tiled" ([row/8][col/8] cores, row%8 x col%8 inside), an array [r][c] is a K-major operand
// (MN = r, K = c: LBO = column-core stride, SBO = row-group stride) AND an MN-major operand (MN = c, K = r: LBO = row-group stride,
// SBO = column-core stride) on the same bytes — measured on the B200, scripts/probe/umma_bf16_probe.cu /
// profiles/r02_umma_bf16_noswizzle_probe.txt (kind::tf32 accepts MN-major only in the 128B_BASE32B swizzle, which is why this is
// a bf16x3 and not a 3xTF32 kernel).  Activations / dZ are [point][feature] with point-groups contiguous, so a warp's 32 threads
// (one point each) store 512 contiguous bytes per 8-feature chunk: conflict-free 16-byte stores.
//
// Epilogues (512 threads: warp w owns TMEM lanes 32 (w%4).., column quarter w/4): tcgen05.ld the accumulator row, activation /
// activation derivative (Z_l stays in TMEM for the backward pass), head at the output layer, split into the three bf16 parts,
// store.  dZ_l overwrites A_l in place (A_l is dead once GEMM3_{l+1} has retired).
//
// Replaces forward (bnn/base.py:89-104) + log_likelihood / COCP terms + autograd .backward() for nets served by generic_net.cu.
#include <cuda_bf16.h>
#include <cstdlib>
#include "heads.cuh"
#include "small_net.cuh"
#include "tc_gemm.cuh"

namespace ocbnn {
namespace gtc {

using tc::smem_u32;

constexpr int kTPmax = 128;         // points per tile: 128 (= the TMEM lanes, MMA M = 128) or 64 (M = 64: 16 lanes of every subpartition)
#ifndef OCBNN_GTC_THREADS
#define OCBNN_GTC_THREADS 512
#endif
constexpr int kThreads = OCBNN_GTC_THREADS;   // 16 epilogue warps: 4 per TMEM subpartition, a quarter of the columns each (warp 0 also issues the MMAs)
constexpr int kParts = kThreads / 128;        // column shares per TMEM lane
constexpr int kBatch = kThreads >= 512 ? 2 : 4;   // 8-column chunks loaded from TMEM before one wait
constexpr int kRG = 128;            // bytes between 8-row groups of a core-tiled array (row groups contiguous)
constexpr int kCutTiles = 4;        // dW accumulation chains are drained after this many tiles (48 MMAs per tile and layer)
constexpr int kMaxRt = 10;

struct Layer {
    int kin, kp;        // s_{l-1} + 1 (bias row), rounded up to 16: K of GEMM1, rows of the W array, features of A_{l-1}
    int n, np;          // s_l, rounded up to 16
    int w_off, w_part, w_cg;   // W_l: smem byte offset of part 0, bytes per part, column-core stride
    int a_off, a_part;  // A_{l-1} (and, in place, dZ_{l-1}): smem byte offset of part 0, bytes per part
    int g_off, g_part;  // dZ_l: smem byte offset of part 0, bytes per part
    int g_overlap;      // 1: dZ_l has an array of its own while GEMM3_{l+1} still reads A_l, so that GEMM3 runs behind the epilogue
    int z_col;          // TMEM column of Z_l
    int dw_col, dw_m, dw_n, dw_t;   // dW_l accumulator: TMEM column, MMA M (64 | 128), MMA N, transposed ([n][k]) orientation
    int coff;           // offset of W_l (then b_l) in the packed weight vector
};
struct Layout {
    int nl, act, d, s0, sk;
    Layer L[OCBNN_MAX_LAYERS + 1];      // index 1..nl
    int tp;                             // points per tile
    int rec_off, rec_bytes, bar_off, zero_bytes, smem_bytes;    // rec: two buffers (the next tile's records are prefetched)
    int da_col, tmem_cols;
};

__host__ __device__ inline int up16(int v) { return (v + 15) & ~15; }

// ---- descriptors ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One K = 16 step of an fp32-accurate product: the six part pairs.  a / b: descriptors of part 0; parts are a_part / b_part bytes apart.
__device__ __forceinline__ void umma_x3(uint32_t d, uint64_t a, uint64_t b, uint32_t a_part16, uint32_t b_part16, uint32_t idesc, uint32_t accumulate) {
    const uint64_t a1 = a + a_part16, a2 = a + 2 * (uint64_t)a_part16, b1 = b + b_part16, b2 = b + 2 * (uint64_t)b_part16;
    umma_bf16(d, a1, b1, idesc, accumulate);      // small terms first
    umma_bf16(d, a, b2, idesc, 1);
    umma_bf16(d, a2, b, idesc, 1);
    umma_bf16(d, a, b1, idesc, 1);
    umma_bf16(d, a1, b, idesc, 1);
    umma_bf16(d, a, b, idesc, 1);
}

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- fp32 -> three bf16 parts -------------------------------------------------------------------------------------------
// two values at a time: p0 = bf16x2(x), r = x - p0 (exact), p1 = bf16x2(r), p2 = bf16x2(r - p1)
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(x1), "f"(x0));
    float r0 = x0 - __uint_as_float(p0 << 16), r1 = x1 - __uint_as_float(p0 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(r1), "f"(r0));
    r0 -= __uint_as_float(p1 << 16);
    r1 -= __uint_as_float(p1 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(r1), "f"(r0));
}
// eight consecutive features of one point -> one 16-byte store per part
__device__ __forceinline__ void store_chunk(uint8_t* part0, int part_bytes, int byte_off, const float (&v)[8]) {
    uint4 q0, q1, q2;
    split3_pair(v[0], v[1], q0.x, q1.x, q2.x);
    split3_pair(v[2], v[3], q0.y, q1.y, q2.y);
    split3_pair(v[4], v[5], q0.z, q1.z, q2.z);
    split3_pair(v[6], v[7], q0.w, q1.w, q2.w);
    *reinterpret_cast<uint4*>(part0 + byte_off) = q0;
    *reinterpret_cast<uint4*>(part0 + part_bytes + byte_off) = q1;
    *reinterpret_cast<uint4*>(part0 + 2 * part_bytes + byte_off) = q2;
}
template <int TP>
__device__ __forceinline__ int act_chunk_off(int p, int chunk) { return chunk * ((TP / 8) * kRG) + (p >> 3) * kRG + (p & 7) * 16; }

__device__ __forceinline__ float act_f(int act, float z) {
    if (act == OCBNN_ACT_RBF) return ex2_approx(z * z * -1.4426950408889634f);
    if (act == OCBNN_ACT_RELU) return fmaxf(z, 0.f);
    return z;
}
__device__ __forceinline__ float dact_f(int act, float z) {
    if (act == OCBNN_ACT_RBF) return ex2_approx(z * z * -1.4426950408889634f) * (-2.f * z);
    if (act == OCBNN_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return 1.f;
}

template <int SK>
__device__ __forceinline__ float head_regs(int kind, const float (&z)[16], const float* prm, const HeadConsts& hc, float wgt, float (&dz)[16]) {
    float f[SK], df[SK], val;
#pragma unroll
    for (int k = 0; k < SK; ++k) f[k] = z[k];
    eval_head<SK>(kind, f, prm, hc, val, df);
#pragma unroll
    for (int k = 0; k < SK; ++k) dz[k] = wgt * df[k];
    return wgt * val;
}
__device__ __forceinline__ float head_switch(int sk, int kind, const float (&z)[16], const float* prm, const HeadConsts& hc, float wgt, float (&dz)[16]) {
    switch (sk) {
        case 1: return head_regs<1>(kind, z, prm, hc, wgt, dz);
        case 2: return head_regs<2>(kind, z, prm, hc, wgt, dz);
        case 3: return head_regs<3>(kind, z, prm, hc, wgt, dz);
        case 4: return head_regs<4>(kind, z, prm, hc, wgt, dz);
        case 5: return head_regs<5>(kind, z, prm, hc, wgt, dz);
        case 6: return head_regs<6>(kind, z, prm, hc, wgt, dz);
        case 7: return head_regs<7>(kind, z, prm, hc, wgt, dz);
        case 8: return head_regs<8>(kind, z, prm, hc, wgt, dz);
        case 9: return head_regs<9>(kind, z, prm, hc, wgt, dz);
        default: return head_regs<10>(kind, z, prm, hc, wgt, dz);
    }
}

// kWLoad: weights (and prior entries) a thread has in flight while it loads a weight vector - 8 where a weight vector meets only a tile or
// two of points (recid minibatches: the load is a quarter of the chain's time), 4 otherwise
template <int kTP, int ACT, int kWLoad>
__global__ void __launch_bounds__(kThreads, 1)
k_logpost_gen_tc(EvalArgs a, const __grid_constant__ Layout Yp, const float* __restrict__ w, long long m, float* __restrict__ out_lp,
                 float* __restrict__ out_grad) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ Layout Y;                               // indexed by a run-time layer number: shared memory, not the parameter bank
    for (int i = threadIdx.x; i < (int)(sizeof(Layout) / 4); i += kThreads) reinterpret_cast<int*>(&Y)[i] = reinterpret_cast<const int*>(&Yp)[i];
    __syncthreads();
    float* rec0 = reinterpret_cast<float*>(sm + Y.rec_off);
    const uint32_t bar = smem_u32(sm + Y.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + Y.bar_off + 8);
    double* red = reinterpret_cast<double*>(sm + Y.bar_off + 16);
    const uint32_t sm_base = smem_u32(sm);

    const int warp = tc::warp_idx_sync(), lane = threadIdx.x & 31;
    const int q = warp & 3, part = warp >> 2;
    constexpr int kActCG = (kTP / 8) * kRG;            // bytes between 8-feature column cores of an activation array
    // this thread's point of the tile <-> its TMEM lane.  M = 64 accumulators keep 16 rows in lanes 0..15 of every subpartition:
    // the upper half-warp has no point (it still takes part in the warp-wide tcgen05.ld)
    const int pt = kTP == 128 ? q * 32 + lane : q * 16 + (lane & 15);
    const bool has_pt = kTP == 128 || lane < 16;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int nl = Y.nl, sk = Y.sk;
    const bool want_grad = out_grad != nullptr;

    // pad rows / columns of every operand array are zero for the whole kernel (only real entries are rewritten)
    for (int i = threadIdx.x; i < Y.zero_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(smem_u32(tmem_slot), (uint32_t)Y.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t phase = 0;                                 // parity of the next completion of `bar`

    EvalCtx c = make_ctx(a, rec0, nullptr, kTP);
    const int rec_floats = Y.rec_bytes >> 2;
    // records of the tile starting at `begin` -> buffer `buf` by cp.async (the N/|B| factor of likelihood rows is applied at use)
    auto prefetch_records = [&](int begin, int buf) {
        const int count = min(kTP, c.npts - begin);
        const int rs = a.rs, stride = a.batch_stride > 1 ? a.batch_stride : 1, offset = a.batch_stride > 1 ? a.batch_offset : 0;
        const uint32_t dst0 = smem_u32(rec0 + buf * rec_floats);
        if ((int)threadIdx.x < count) {                 // one thread per record: the row address is formed once
            const int j = threadIdx.x, pj = begin + j;
            const long long row = pj < c.nb ? (long long)offset + (long long)pj * stride : a.n_train + (pj - c.nb);
            const float* src = a.pts + row * rs;
            uint32_t dst = dst0 + (uint32_t)(j * rs * 4);
            for (int k = 0; k < rs; ++k, dst += 4, ++src) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (long long chain = blockIdx.x; chain < m; chain += gridDim.x) {
        const float* wc = w + chain * Y.d;
        float* gc = want_grad ? out_grad + chain * Y.d : nullptr;
        if (chain + gridDim.x < m) {          // the next weight vector of this CTA: into L2 behind this one's work
            const float* nx = w + (chain + gridDim.x) * Y.d;
            for (int i = threadIdx.x * 32; i < Y.d; i += kThreads * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + i));
        }
        // ---- weights -> three bf16 parts, core-tiled [k][n]; prior value and prior gradient ----------------------------------
        double lp = 0.0;
        {
            double pr = 0.0;
            for (int l = 1; l <= nl; ++l) {
                const Layer& L = Y.L[l];
                uint8_t* wp = sm + L.w_off;
                // thread -> (row k0 + j * rows_per_pass, column n): consecutive threads read consecutive n of one row
                int nn = 8;
                while (nn < L.n) nn <<= 1;
                const int sh = __ffs(nn) - 1, rpp = max(1, kThreads >> sh);
                const int n = threadIdx.x & (nn - 1), k0 = threadIdx.x >> sh;
                if (n < L.n && k0 < rpp) {
                    const int coloff = (n >> 3) * L.w_cg + (n & 7) * 2;
                    for (int kb = k0; kb < L.kin; kb += kWLoad * rpp) {
                        float v[kWLoad];
                        float2 mv[kWLoad];
#pragma unroll
                        for (int u = 0; u < kWLoad; ++u) {
                            const int k = kb + u * rpp;
                            if (k < L.kin) {
                                v[u] = __ldg(wc + L.coff + k * L.n + n);
                                mv[u] = a.pri[L.coff + k * L.n + n];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kWLoad; ++u) {
                            const int k = kb + u * rpp;
                            if (k >= L.kin) break;
                            const float diff = v[u] - mv[u].x, t = diff * mv[u].y;
                            pr += (double)(diff * t);
                            if (gc) gc[L.coff + k * L.n + n] = -t;
                            __nv_bfloat16 b0 = __float2bfloat16_rn(v[u]);
                            float r = v[u] - __bfloat162float(b0);
                            __nv_bfloat16 b1 = __float2bfloat16_rn(r);
                            r -= __bfloat162float(b1);
                            __nv_bfloat16 b2 = __float2bfloat16_rn(r);
                            const int off = coloff + (k >> 3) * kRG + (k & 7) * 16;
                            *reinterpret_cast<__nv_bfloat16*>(wp + off) = b0;
                            *reinterpret_cast<__nv_bfloat16*>(wp + L.w_part + off) = b1;
                            *reinterpret_cast<__nv_bfloat16*>(wp + 2 * L.w_part + off) = b2;
                        }
                    }
                }
            }
            lp = -0.5 * pr;
        }
        int chain_tiles = 0;          // tiles accumulated in the dW accumulators since the last drain
        bool pending = false;         // a commit (GEMM3 of layer 1) has not been waited for yet

        int buf = 0;
        if (c.npts > 0) prefetch_records(0, 0);
        for (int begin = 0; begin < c.npts; begin += kTP, buf ^= 1) {
            const int count = min(kTP, c.npts - begin);
            const float* rec = rec0 + buf * rec_floats;
            asm volatile("cp.async.wait_group 0;" ::: "memory");        // this tile's records (issued one tile ago) have landed
            if (pending) {            // the previous tile's GEMM3s still read A_0 / the dZ arrays
                tc::mbar_wait(bar, phase);
                phase ^= 1;
                pending = false;
            }
            __syncthreads();
            if (begin + kTP < c.npts) prefetch_records(begin + kTP, buf ^ 1);      // lands behind this tile's work
            // ---- A_0 = [x | 1 | 0..] ------------------------------------------------------------------------------------------
            {
                const Layer& L1 = Y.L[1];
                const int nch = L1.kp >> 3, per = (nch + kParts - 1) / kParts;
                const float* r = rec + pt * a.rs;
                const int a_off = L1.a_off, a_part = L1.a_part;
                for (int ch = part * per; ch < min(nch, (part + 1) * per); ++ch) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int f = ch * 8 + i;
                        v[i] = f < Y.s0 ? (pt < count ? r[f] : 0.f) : (f == Y.s0 ? 1.f : 0.f);
                    }
                    if (has_pt) store_chunk(sm + a_off, a_part, act_chunk_off<kTP>(pt, ch), v);
                }
            }
            fence_async_smem();
            __syncthreads();

            // ---- forward ------------------------------------------------------------------------------------------------------------
            for (int l = 1; l <= nl; ++l) {
                const Layer& L = Y.L[l];
                const uint32_t zc = l < nl ? (uint32_t)L.z_col : (uint32_t)Y.da_col;
                if (warp == 0) {
                    if (tc::elect_one_sync()) {
                        tc::tc_fence_after();
                        const uint32_t idesc = idesc_bf16(kTP, L.np, 0, 1);
                        const uint64_t da = desc_noswz(sm_base + L.a_off, kActCG, kRG);          // A_{l-1}: K-major
                        const uint64_t db = desc_noswz(sm_base + L.w_off, kRG, L.w_cg);          // W_l: MN-major (MN = n, K = k)
                        for (int ks = 0; ks < (L.kp >> 4); ++ks)
                            umma_x3(tmem + zc, da + (uint64_t)((ks * 2 * kActCG) >> 4), db + (uint64_t)((ks * 2 * kRG) >> 4), L.a_part >> 4,
                                    L.w_part >> 4, idesc, ks != 0);
                        tc::umma_commit(bar);
                    }
                    __syncwarp();
                }
                tc::mbar_wait(bar, phase);
                phase ^= 1;
                tc::tc_fence_after();
                if (l < nl) {
                    // A_l = [act(Z_l) | 1 | 0..]: features of layer l+1's input
                    const Layer& N = Y.L[l + 1];
                    const int nch = N.kp >> 3, per = (nch + kParts - 1) / kParts;
                    const int c_lo = part * per, c_hi = min(nch, (part + 1) * per);
                    const int n_real = L.n, np = L.np, a_off = N.a_off, a_part = N.a_part;
                    constexpr int act = ACT;                 // compile-time activation: the per-element dispatch folds away
                    for (int c0 = c_lo; c0 < c_hi; c0 += kBatch) {
                        uint32_t raw[kBatch][8];
#pragma unroll
                        for (int j = 0; j < kBatch; ++j)
                            if (c0 + j < c_hi && (c0 + j) * 8 < np) tmem_ld8_nowait(tmem + lane_base + zc + (uint32_t)((c0 + j) * 8), raw[j]);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < kBatch; ++j) {
                            if (c0 + j >= c_hi) break;
                            const int f0 = (c0 + j) * 8;
                            float v[8];
                            if (f0 + 8 <= n_real) {             // interior chunk: no masks
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = act_f(act, __uint_as_float(raw[j][i]));
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = f0 + i < n_real ? act_f(act, __uint_as_float(raw[j][i])) : (f0 + i == n_real ? 1.f : 0.f);
                            }
                            if (has_pt) store_chunk(sm + a_off, a_part, act_chunk_off<kTP>(pt, c0 + j), v);
                        }
                    }
                } else if (part == 0) {
                    // heads: value and dZ of the output layer (weighted); dead points get a zero gradient
                    uint32_t raw[2][8];
                    tmem_ld8_nowait(tmem + lane_base + zc, raw[0]);
                    tmem_ld8_nowait(tmem + lane_base + zc + 8, raw[1]);
                    tmem_ld_wait();
                    float z[16], dz[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        z[i] = __uint_as_float(raw[i >> 3][i & 7]);
                        dz[i] = 0.f;
                    }
                    if (has_pt && pt < count) {
                        const float* r = rec + pt * a.rs;
                        const float wgt = begin + pt < c.nb ? r[a.s0 + 1] * c.mult : r[a.s0 + 1];      // likelihood rows: N/|B|
                        lp += (double)head_switch(sk, __float_as_int(r[a.s0]), z, r + a.s0 + 2, c.hc, wgt, dz);
                    }
                    if (want_grad && has_pt) {
                        float v0[8], v1[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v0[i] = dz[i];
                            v1[i] = dz[8 + i];
                        }
                        store_chunk(sm + Y.L[nl].g_off, Y.L[nl].g_part, act_chunk_off<kTP>(pt, 0), v0);
                        store_chunk(sm + Y.L[nl].g_off, Y.L[nl].g_part, act_chunk_off<kTP>(pt, 1), v1);
                    }
                }
                tc::tc_fence_before();
                fence_async_smem();
                __syncthreads();
            }
            if (!want_grad) continue;

            // ---- backward ---------------------------------------------------------------------------------------------------------
            // Step l: GEMM2_l (dA_{l-1}, the dependent chain) is committed on its own; GEMM3_l (dW_l) is issued behind it and retires
            // under the epilogue that turns dA_{l-1} into dZ_{l-1} - unless dZ_{l-1} has to overwrite A_{l-1}, which GEMM3_l reads.
            for (int l = nl; l >= 1; --l) {
                const Layer& L = Y.L[l];
                const int g_off = L.g_off, g_part = L.g_part;                        // dZ_l
                const bool dw_first = l >= 2 && !Y.L[l - 1].g_overlap;                // dZ_{l-1} goes over A_{l-1}: GEMM3_l must retire first
                if (warp == 0) {
                    if (tc::elect_one_sync()) {
                        tc::tc_fence_after();
                        auto gemm3 = [&]() {   // dW_l (+)= A_{l-1}^T dZ_l : both MN-major (K = points)
                            const uint64_t dact = desc_noswz(sm_base + L.a_off, kRG, kActCG);
                            const uint64_t ddz = desc_noswz(sm_base + g_off, kRG, kActCG);
                            const uint32_t idesc = idesc_bf16(L.dw_m, L.dw_n, 1, 1);
                            for (int ks = 0; ks < kTP / 16; ++ks) {
                                const uint64_t adv = (uint64_t)((ks * 2 * kRG) >> 4);
                                if (!L.dw_t) umma_x3(tmem + (uint32_t)L.dw_col, dact + adv, ddz + adv, L.a_part >> 4, g_part >> 4, idesc, (chain_tiles | ks) != 0);
                                else umma_x3(tmem + (uint32_t)L.dw_col, ddz + adv, dact + adv, g_part >> 4, L.a_part >> 4, idesc, (chain_tiles | ks) != 0);
                            }
                        };
                        if (dw_first) gemm3();
                        if (l >= 2) {       // dA_{l-1} = dZ_l W_l^T : A = dZ_l K-major, B = W_l K-major (MN = k, K = n)
                            const int nprev = Y.L[l - 1].np;
                            const uint32_t idesc = idesc_bf16(kTP, nprev, 0, 0);
                            const uint64_t da = desc_noswz(sm_base + g_off, kActCG, kRG);
                            const uint64_t db = desc_noswz(sm_base + L.w_off, L.w_cg, kRG);
                            for (int ks = 0; ks < (L.np >> 4); ++ks)
                                umma_x3(tmem + (uint32_t)Y.da_col, da + (uint64_t)((ks * 2 * kActCG) >> 4), db + (uint64_t)((ks * 2 * L.w_cg) >> 4),
                                        g_part >> 4, L.w_part >> 4, idesc, ks != 0);
                            tc::umma_commit(bar);
                        }
                        if (!dw_first) gemm3();
                        if (l == 1) tc::umma_commit(bar);       // covers every GEMM3 of this tile
                    }
                    __syncwarp();
                }
                if (l == 1) {
                    pending = true;
                    break;
                }
                tc::mbar_wait(bar, phase);
                phase ^= 1;
                tc::tc_fence_after();
                // dZ_{l-1} = dA_{l-1} * act'(Z_{l-1})
                {
                    const Layer& P = Y.L[l - 1];
                    const int nch = P.np >> 3, per = (nch + kParts - 1) / kParts;
                    const int c_lo = part * per, c_hi = min(nch, (part + 1) * per);
                    const int n_real = P.n, pg_off = P.g_off, pg_part = P.g_part;
                    constexpr int act = ACT;
                    const uint32_t t_da = tmem + lane_base + (uint32_t)Y.da_col, t_z = tmem + lane_base + (uint32_t)P.z_col;
                    for (int c0 = c_lo; c0 < c_hi; c0 += kBatch) {
                        uint32_t rd[kBatch][8], rz[kBatch][8];
#pragma unroll
                        for (int j = 0; j < kBatch; ++j)
                            if (c0 + j < c_hi) {
                                tmem_ld8_nowait(t_da + (uint32_t)((c0 + j) * 8), rd[j]);
                                tmem_ld8_nowait(t_z + (uint32_t)((c0 + j) * 8), rz[j]);
                            }
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < kBatch; ++j) {
                            if (c0 + j >= c_hi) break;
                            const int f0 = (c0 + j) * 8;
                            float v[8];
                            if (f0 + 8 <= n_real) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(rd[j][i]) * dact_f(act, __uint_as_float(rz[j][i]));
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = f0 + i < n_real ? __uint_as_float(rd[j][i]) * dact_f(act, __uint_as_float(rz[j][i])) : 0.f;
                            }
                            if (has_pt) store_chunk(sm + pg_off, pg_part, act_chunk_off<kTP>(pt, c0 + j), v);
                        }
                    }
                }
                tc::tc_fence_before();
                fence_async_smem();
                __syncthreads();
            }

            // ---- drain the dW accumulation chains into the gradient ---------------------------------------------------------------
            ++chain_tiles;
            if (chain_tiles == kCutTiles || begin + kTP >= c.npts) {
                tc::mbar_wait(bar, phase);
                phase ^= 1;
                pending = false;
                tc::tc_fence_after();
                for (int l = 1; l <= nl; ++l) {
                    const Layer& L = Y.L[l];
                    const int row = L.dw_m == 128 ? q * 32 + lane : q * 16 + lane;           // M = 64: 16 rows per TMEM subpartition
                    const bool row_ok = L.dw_m == 128 || lane < 16;
                    const int nch = L.dw_n >> 3, per = (nch + kParts - 1) / kParts;
                    const int c_lo = part * per, c_hi = min(nch, (part + 1) * per);
                    for (int ch = c_lo; ch < c_hi; ++ch) {
                        uint32_t raw[8];
                        tmem_ld8_nowait(tmem + lane_base + (uint32_t)L.dw_col + (uint32_t)(ch * 8), raw);
                        tmem_ld_wait();
                        if (!row_ok) continue;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int col = ch * 8 + i;
                            const int k = L.dw_t ? col : row, n = L.dw_t ? row : col;
                            if (k < L.kin && n < L.n) atomicAdd(gc + L.coff + k * L.n + n, __uint_as_float(raw[i]));      // RED: no round trip
                        }
                    }
                }
                chain_tiles = 0;
                tc::tc_fence_before();
                __syncthreads();
            }
        }
        if (pending) {      // (only when the gradient is not wanted this cannot happen; kept for symmetry)
            tc::mbar_wait(bar, phase);
            phase ^= 1;
            pending = false;
        }
        // ---- value ---------------------------------------------------------------------------------------------------------------
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, off);
        if (lane == 0) red[warp] = lp;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = c.lp_const;
            for (int i = 0; i < kThreads / 32; ++i) s += red[i];
            if (out_lp) out_lp[chain] = (float)s;
        }
        __syncthreads();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)Y.tmem_cols);
}

// ---- forward only (predict / forward, base.py:89-104, 267-277): the forward half of the kernel above -------------------------------
// out[chain][point][k] = f(x_point; w_chain).  Same operand arrays, GEMM1 and epilogues; the points come straight from x (npts x s_0,
// row-major) and the rows of the output layer's accumulator go to global memory.  One CTA per weight vector at a time.
template <int kTP, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
k_forward_gen_tc(const __grid_constant__ Layout Yp, const float* __restrict__ w, long long m, const float* __restrict__ x, long long npts,
                 float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ Layout Y;
    for (int i = threadIdx.x; i < (int)(sizeof(Layout) / 4); i += kThreads) reinterpret_cast<int*>(&Y)[i] = reinterpret_cast<const int*>(&Yp)[i];
    __syncthreads();
    float* rec0 = reinterpret_cast<float*>(sm + Y.rec_off);
    const uint32_t bar = smem_u32(sm + Y.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + Y.bar_off + 8);
    const uint32_t sm_base = smem_u32(sm);
    const int warp = tc::warp_idx_sync(), lane = threadIdx.x & 31;
    const int q = warp & 3, part = warp >> 2;
    constexpr int kActCG = (kTP / 8) * kRG;
    const int pt = kTP == 128 ? q * 32 + lane : q * 16 + (lane & 15);
    const bool has_pt = kTP == 128 || lane < 16;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int nl = Y.nl, sk = Y.sk, s0 = Y.s0;

    for (int i = threadIdx.x; i < Y.zero_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(smem_u32(tmem_slot), (uint32_t)Y.tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t phase = 0;
    const int rec_floats = Y.rec_bytes >> 2;
    // rows [begin, begin + count) of x -> buffer `buf`: contiguous in memory, copied 4 bytes at a time (s_0 is not a multiple of 4 in general)
    auto prefetch_x = [&](long long begin, int buf) {
        const int count = (int)min((long long)kTP, npts - begin);
        const uint32_t dst0 = smem_u32(rec0 + buf * rec_floats);
        const float* src = x + begin * s0;
        for (int i = threadIdx.x; i < count * s0; i += kThreads) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + (uint32_t)(i * 4)), "l"(src + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (long long chain = blockIdx.x; chain < m; chain += gridDim.x) {
        const float* wc = w + chain * Y.d;
        __syncthreads();                                  // the previous weight vector's last GEMM has been waited for by every thread
        for (int l = 1; l <= nl; ++l) {
            const Layer& L = Y.L[l];
            uint8_t* wp = sm + L.w_off;
            int nn = 8;
            while (nn < L.n) nn <<= 1;
            const int sh = __ffs(nn) - 1, rpp = max(1, kThreads >> sh);
            const int n = threadIdx.x & (nn - 1), k0 = threadIdx.x >> sh;
            if (n < L.n && k0 < rpp) {
                const int coloff = (n >> 3) * L.w_cg + (n & 7) * 2;
                for (int k = k0; k < L.kin; k += rpp) {
                    const float v = __ldg(wc + L.coff + k * L.n + n);
                    __nv_bfloat16 b0 = __float2bfloat16_rn(v);
                    float r = v - __bfloat162float(b0);
                    __nv_bfloat16 b1 = __float2bfloat16_rn(r);
                    r -= __bfloat162float(b1);
                    __nv_bfloat16 b2 = __float2bfloat16_rn(r);
                    const int off = coloff + (k >> 3) * kRG + (k & 7) * 16;
                    *reinterpret_cast<__nv_bfloat16*>(wp + off) = b0;
                    *reinterpret_cast<__nv_bfloat16*>(wp + L.w_part + off) = b1;
                    *reinterpret_cast<__nv_bfloat16*>(wp + 2 * L.w_part + off) = b2;
                }
            }
        }
        int buf = 0;
        if (npts > 0) prefetch_x(0, 0);
        for (long long begin = 0; begin < npts; begin += kTP, buf ^= 1) {
            const int count = (int)min((long long)kTP, npts - begin);
            const float* rec = rec0 + buf * rec_floats;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            if (begin + kTP < npts) prefetch_x(begin + kTP, buf ^ 1);
            {
                const Layer& L1 = Y.L[1];
                const int nch = L1.kp >> 3, per = (nch + kParts - 1) / kParts;
                const float* r = rec + pt * s0;
                const int a_off = L1.a_off, a_part = L1.a_part;
                for (int ch = part * per; ch < min(nch, (part + 1) * per); ++ch) {
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int f = ch * 8 + i;
                        v[i] = f < s0 ? (pt < count ? r[f] : 0.f) : (f == s0 ? 1.f : 0.f);
                    }
                    if (has_pt) store_chunk(sm + a_off, a_part, act_chunk_off<kTP>(pt, ch), v);
                }
            }
            fence_async_smem();
            __syncthreads();
            for (int l = 1; l <= nl; ++l) {
                const Layer& L = Y.L[l];
                const uint32_t zc = l < nl ? (uint32_t)L.z_col : (uint32_t)Y.da_col;
                if (warp == 0) {
                    if (tc::elect_one_sync()) {
                        tc::tc_fence_after();
                        const uint32_t idesc = idesc_bf16(kTP, L.np, 0, 1);
                        const uint64_t da = desc_noswz(sm_base + L.a_off, kActCG, kRG);
                        const uint64_t db = desc_noswz(sm_base + L.w_off, kRG, L.w_cg);
                        for (int ks = 0; ks < (L.kp >> 4); ++ks)
                            umma_x3(tmem + zc, da + (uint64_t)((ks * 2 * kActCG) >> 4), db + (uint64_t)((ks * 2 * kRG) >> 4), L.a_part >> 4,
                                    L.w_part >> 4, idesc, ks != 0);
                        tc::umma_commit(bar);
                    }
                    __syncwarp();
                }
                tc::mbar_wait(bar, phase);
                phase ^= 1;
                tc::tc_fence_after();
                if (l < nl) {
                    const Layer& N = Y.L[l + 1];
                    const int nch = N.kp >> 3, per = (nch + kParts - 1) / kParts;
                    const int c_lo = part * per, c_hi = min(nch, (part + 1) * per);
                    const int n_real = L.n, np = L.np, a_off = N.a_off, a_part = N.a_part;
                    constexpr int act = ACT;
                    for (int c0 = c_lo; c0 < c_hi; c0 += kBatch) {
                        uint32_t raw[kBatch][8];
#pragma unroll
                        for (int j = 0; j < kBatch; ++j)
                            if (c0 + j < c_hi && (c0 + j) * 8 < np) tmem_ld8_nowait(tmem + lane_base + zc + (uint32_t)((c0 + j) * 8), raw[j]);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < kBatch; ++j) {
                            if (c0 + j >= c_hi) break;
                            const int f0 = (c0 + j) * 8;
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = f0 + i < n_real ? act_f(act, __uint_as_float(raw[j][i])) : (f0 + i == n_real ? 1.f : 0.f);
                            if (has_pt) store_chunk(sm + a_off, a_part, act_chunk_off<kTP>(pt, c0 + j), v);
                        }
                    }
                } else if (part == 0) {
                    uint32_t raw[2][8];
                    tmem_ld8_nowait(tmem + lane_base + zc, raw[0]);
                    tmem_ld8_nowait(tmem + lane_base + zc + 8, raw[1]);
                    tmem_ld_wait();
                    if (has_pt && pt < count) {
                        float* o = out + ((long long)chain * npts + begin + pt) * sk;
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            if (k < sk) o[k] = __uint_as_float(raw[k >> 3][k & 7]);
                    }
                }
                tc::tc_fence_before();
                fence_async_smem();
                __syncthreads();
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)Y.tmem_cols);
}


// ---- host side --------------------------------------------------------------------------------------------------------------------
static bool make_layout(const Problem& p, Layout& Y, bool overlap, int tp) {
    Y = Layout{};
    Y.tp = tp;
    Y.nl = p.desc.n_layers;
    Y.act = p.desc.activation;
    Y.s0 = p.desc.layer_sizes[0];
    Y.sk = p.desc.layer_sizes[Y.nl];
    if (Y.nl < 1 || Y.sk > kMaxRt) return false;
    int off = 0, coff = 0, col = 0;
    // activation arrays first: GEMM3 reads dw_m feature rows of A_{l-1}, possibly beyond its kp features (rows that are never
    // used) - whatever follows must still be shared memory of this CTA
    for (int l = 1; l <= Y.nl; ++l) {
        Layer& L = Y.L[l];
        L.kin = p.desc.layer_sizes[l - 1] + 1;
        L.kp = up16(L.kin);
        L.n = p.desc.layer_sizes[l];
        L.np = up16(L.n);
        if (L.kin > 128 || L.np > 256) return false;
        L.a_off = off;
        L.a_part = tp * L.kp * 2;
        off += 3 * L.a_part;
        L.coff = coff;
        coff += L.kin * L.n;
    }
    Y.d = coff;
    // dZ arrays.  Output layer: 16 feature columns of its own.  Hidden layers: dZ_l may be written while GEMM3_{l+1} still reads A_l,
    // so it cannot simply overwrite A_l: dZ_{nl-1} gets an array X of its own and dZ_l below it goes to the array of A_{l+1}, whose
    // last reader (GEMM3_{l+2}) retired a step earlier - if it is wide enough; otherwise in place over A_l after a full wait.
    {
        Layer& O = Y.L[Y.nl];
        O.g_off = off;
        O.g_part = tp * 16 * 2;
        O.g_overlap = 1;
        off += 3 * O.g_part;
    }
    for (int l = Y.nl - 1; l >= 1; --l) {
        Layer& L = Y.L[l];
        L.g_part = tp * L.np * 2;
        if (l == Y.nl - 1 && overlap) {
            L.g_off = off;
            L.g_overlap = 1;
            off += 3 * L.g_part;
        } else if (overlap && l + 2 <= Y.nl && Y.L[l + 2].kp >= L.np && Y.L[l + 1].g_overlap) {
            L.g_off = Y.L[l + 2].a_off;              // the array of A_{l+1}
            L.g_part = Y.L[l + 2].a_part;
            L.g_overlap = 1;
        } else {
            L.g_off = Y.L[l + 1].a_off;              // in place over A_l
            L.g_part = Y.L[l + 1].a_part;
            L.g_overlap = 0;
        }
    }
    for (int l = 1; l <= Y.nl; ++l) {
        Layer& L = Y.L[l];
        // W_l as the K-major B operand of GEMM2 is read for N = np_{l-1} rows: the array has max(kp, np_{l-1}) = kp rows
        L.w_off = off;
        L.w_part = L.kp * L.np * 2;
        L.w_cg = (L.kp / 8) * kRG;
        off += 3 * L.w_part;
    }
    Y.zero_bytes = off;
    Y.rec_off = off;
    Y.rec_bytes = ((tp * p.rs * 4) + 15) & ~15;
    off += 2 * Y.rec_bytes;
    Y.bar_off = off;
    off += 16 + 8 * (kThreads / 32) + 16;
    // GEMM3 reads up to 128 feature rows (32 KB) from the start of an activation / dZ part, possibly beyond its features (rows
    // that are never used): those bytes must still be shared memory of this CTA
    int need_end = 0;
    for (int l = 1; l <= Y.nl; ++l) need_end = max(need_end, max(Y.L[l].a_off + 2 * Y.L[l].a_part, Y.L[l].g_off + 2 * Y.L[l].g_part) + 32 * 1024);
    Y.smem_bytes = max(off, need_end);
    // TMEM columns: Z_1 .. Z_{nl-1}, one region for dA_l / Z_nl, the dW accumulators
    for (int l = 1; l < Y.nl; ++l) {
        Y.L[l].z_col = col;
        col += Y.L[l].np;
    }
    int da = 16;
    for (int l = 1; l < Y.nl; ++l) da = max(da, Y.L[l].np);
    Y.da_col = col;
    col += da;
    for (int l = 1; l <= Y.nl; ++l) {
        Layer& L = Y.L[l];
        const int m_n = L.kin <= 64 ? 64 : 128, m_t = L.n <= 64 ? 64 : 128;     // normal: rows = k ; transposed: rows = n
        const int cols_n = L.np, cols_t = L.kp;
        const bool can_t = L.n <= 128;
        L.dw_t = can_t && cols_t < cols_n;
        L.dw_m = L.dw_t ? m_t : m_n;
        L.dw_n = L.dw_t ? cols_t : cols_n;
        L.dw_col = col;
        col += L.dw_n;
    }
    if (col > 512) return false;
    Y.tmem_cols = col <= 32 ? 32 : col <= 64 ? 64 : col <= 128 ? 128 : col <= 256 ? 256 : 512;
    return true;
}

}  // namespace gtc

// 1 = launched, 0 = this net does not fit the tcgen05 kernel (the caller falls back), < 0 = error code negated
int try_launch_logpost_gen_tc(const Problem& p, const EvalArgs& a, const float* w, long long m, float* out_lp, float* out_grad, cudaStream_t st,
                              int* rc_out) {
    static const bool enabled = [] { const char* e = getenv("OCBNN_GEN_TC"); return e ? atoi(e) != 0 : true; }();
    *rc_out = OCBNN_OK;
    if (!enabled) return 0;
    gtc::Layout Y;
    static const bool want_overlap = [] { const char* e = getenv("OCBNN_GEN_TC_OVERLAP"); return e ? atoi(e) != 0 : true; }();
    // 128-point tiles (all TMEM lanes, full warps in the epilogues) when the operand arrays fit one SM, else 64-point tiles; with
    // an array of its own for dZ (GEMM3 retires under the epilogue) when that fits too, else dZ in place.  A point set of at most
    // 64 points (recid: 62 per minibatch + constraint points) never needs the larger tile.
    static const int force_tp = [] { const char* e = getenv("OCBNN_GEN_TC_TP"); return e ? atoi(e) : 0; }();
    const int npts = (a.with_data ? batch_count(a.n_train, a.batch_offset, a.batch_stride) : 0) + a.n_cpoints;
    bool ok = false;
    for (int tp : {128, 64}) {
        if (force_tp ? tp != force_tp : (tp == 128 && npts <= 64)) continue;
        for (int ov = want_overlap ? 1 : 0; ov >= 0 && !ok; --ov) ok = gtc::make_layout(p, Y, ov != 0, tp) && Y.smem_bytes + 2048 <= p.max_smem_optin;
        if (ok) break;
    }
    if (!ok) return 0;
    auto launch = [&]() -> int {
        void (*kern)(EvalArgs, const gtc::Layout, const float*, long long, float*, float*) = nullptr;
        const int act = p.desc.activation;
        const bool deep = npts <= 2 * Y.tp;              // few point tiles per weight vector: load it with more requests in flight
#define OCBNN_GTC_PICK(TP, WL)                                                                                                     \
    (act == OCBNN_ACT_RBF ? gtc::k_logpost_gen_tc<TP, OCBNN_ACT_RBF, WL>                                                            \
                          : act == OCBNN_ACT_RELU ? gtc::k_logpost_gen_tc<TP, OCBNN_ACT_RELU, WL> : gtc::k_logpost_gen_tc<TP, OCBNN_ACT_IDENTITY, WL>)
        if (Y.tp == 128) kern = deep ? OCBNN_GTC_PICK(128, 8) : OCBNN_GTC_PICK(128, 4);
        else kern = deep ? OCBNN_GTC_PICK(64, 8) : OCBNN_GTC_PICK(64, 4);
#undef OCBNN_GTC_PICK
        OCBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Y.smem_bytes));
        const long long grid = m < (long long)p.sm_count ? m : (long long)p.sm_count;
        kern<<<(unsigned)grid, gtc::kThreads, Y.smem_bytes, st>>>(a, Y, w, m, out_lp, out_grad);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    };
    *rc_out = launch();
    return 1;
}

// forward / predict of the application nets on the same tcgen05 machinery; 1 = launched, 0 = use generic_net.cu's SIMT forward
int try_launch_forward_gen_tc(const Problem& p, const float* w, long long m, const float* x, long long npts, float* out, cudaStream_t st, int* rc_out) {
    static const bool enabled = [] { const char* e = getenv("OCBNN_GEN_TC"); return e ? atoi(e) != 0 : true; }();
    *rc_out = OCBNN_OK;
    if (!enabled || npts < 32) return 0;                  // a handful of points: the SIMT kernel has no tile to fill
    gtc::Layout Y;
    bool ok = false;
    Problem q = p;
    q.rs = p.desc.layer_sizes[0];                         // the staged "records" are the rows of x
    for (int tp : {128, 64}) {
        if (tp == 128 && npts <= 64) continue;
        ok = gtc::make_layout(q, Y, false, tp) && Y.smem_bytes + 2048 <= p.max_smem_optin;
        if (ok) break;
    }
    if (!ok) return 0;
    auto launch = [&]() -> int {
        void (*kern)(const gtc::Layout, const float*, long long, const float*, long long, float*) = nullptr;
        const int act = p.desc.activation;
#define OCBNN_GTC_FWD(TP) (act == OCBNN_ACT_RBF ? gtc::k_forward_gen_tc<TP, OCBNN_ACT_RBF> : act == OCBNN_ACT_RELU ? gtc::k_forward_gen_tc<TP, OCBNN_ACT_RELU> : gtc::k_forward_gen_tc<TP, OCBNN_ACT_IDENTITY>)
        kern = Y.tp == 128 ? OCBNN_GTC_FWD(128) : OCBNN_GTC_FWD(64);
#undef OCBNN_GTC_FWD
        OCBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Y.smem_bytes));
        const long long grid = m < (long long)p.sm_count ? m : (long long)p.sm_count;
        kern<<<(unsigned)grid, gtc::kThreads, Y.smem_bytes, st>>>(Y, w, m, x, npts, out);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    };
    *rc_out = launch();
    return 1;
}

}  // namespace ocbnn
