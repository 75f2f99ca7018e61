// svgd_flash.cuh — SVGD for D <= 64 without the n x n distance matrix (included by svgd_tc.cu).
//
// Round 1 formed d2[n x n] once (exact differences), ran the median select over it and fed it to the phi kernel: 67 MB of HBM
// traffic for 1 MB of algorithmic input at n = 4096, 4.3 GB at n = 32768.  Here both consumers recompute their tiles on the
// tensor cores instead:
//
//   k_count_tc   (K3a) one CTA per 128 x 128 tile ON OR ABOVE the diagonal: S = Xc_i Xc_j^T on tcgen05 (3xTF32, operands by
//                TMA), w~ = r_i + r_j - 2 S (+ the eps terms) classifies every pair i < j against the sampled bracket
//                [lo, hi] of the median select.  Pairs whose w~ is within a rigorous margin of the bracket (or of zero) are
//                re-evaluated EXACTLY (fp32 direct differences, the same arithmetic as the round-1 distance kernels) before
//                they are counted / compacted, so the select stays exact while ~97 % of the pairs cost one FMA + two compares.
//   k_phi_flash  (K3b) flash-attention shaped: per 128-row tile and 64-column block  S -> TMEM (UMMA #1, A = the tile's Xc
//                rows RESIDENT IN TMEM, B = Xc_j by TMA),  K = exp2((2 S - r_i - r_j) c) evaluated by 8 SIMT warps straight out
//                of TMEM, split into tf32 hi / lo and written back to TMEM,  O += K [G | Xc]_j (UMMA #2, A = K from TMEM,
//                B = [G | Xc]^T by TMA).  Neither d2 nor K ever leaves the SM.  Two S/K buffers ping-pong (two row tiles per
//                CTA when D <= 32, else two column blocks of one tile) so the tensor pipe runs UMMA #1 of the next item and
//                UMMA #2 of the previous one while the SIMT warps exponentiate the current one.
//
// Accumulation chains are kept short on purpose: the tensor core adds into its fp32 accumulator with truncation, so a chain
// of m UMMAs carries a bias of ~m 2^-24 relative to the partial sums (measured: 1.4e-5 on a K = 11232 Gram).  A CTA therefore
// accumulates at most kFlashChainCols columns (64 K steps x 3 MMAs) into one TMEM accumulator; partial results of the column
// splits are summed in fp32 round-to-nearest by the finalize kernel.
#pragma once

namespace ocbnn {
namespace tc {

// ---- TMEM store, UMMA with the A operand in TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem: lane = row, column = k] * B[smem desc]^T, kind::tf32
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// exact squared distance of two centred particles: sequential fp32 fma over the coordinates, the arithmetic of k_d2_direct /
// k_d2_sym / k_sel_sample_x (bit-identical values whichever kernel forms them)
template <int DP>
__device__ __forceinline__ float exact_d2(const float* __restrict__ xc, long long i, long long j) {
    const float4* a4 = reinterpret_cast<const float4*>(xc + i * DP);
    const float4* b4 = reinterpret_cast<const float4*>(xc + j * DP);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < DP / 4; ++c) {
        const float4 a = __ldg(a4 + c), b = __ldg(b4 + c);
        float e;
        e = a.x - b.x; acc = fmaf(e, e, acc);
        e = a.y - b.y; acc = fmaf(e, e, acc);
        e = a.z - b.z; acc = fmaf(e, e, acc);
        e = a.w - b.w; acc = fmaf(e, e, acc);
    }
    return acc;
}
__device__ __forceinline__ float exact_d2_rt(const float* __restrict__ xc, int dp, long long i, long long j) {
    const float4* a4 = reinterpret_cast<const float4*>(xc + i * dp);
    const float4* b4 = reinterpret_cast<const float4*>(xc + j * dp);
    float acc = 0.f;
    for (int c = 0; c < dp / 4; ++c) {
        const float4 a = __ldg(a4 + c), b = __ldg(b4 + c);
        float e;
        e = a.x - b.x; acc = fmaf(e, e, acc);
        e = a.y - b.y; acc = fmaf(e, e, acc);
        e = a.z - b.z; acc = fmaf(e, e, acc);
        e = a.w - b.w; acc = fmaf(e, e, acc);
    }
    return acc;
}

// =========================================================================================================================
// K3a counting pass on recomputed tiles
// =========================================================================================================================
constexpr int kCntThreads = 320;          // warp 0: TMA, warp 1: TMEM + MMA issue, warps 2..9: classification
constexpr int kCntPairCap = 1024;         // flagged pairs per warp and 32-column chunk (32 rows x 32 columns: cannot overflow)
constexpr int kCntValCap = 512;           // exact candidates staged per warp before one global append
// |w~ - w_exact| bound used for the classification: 3xTF32 Gram (<= 2^-20 sum|a||b| incl. the truncating accumulate over <= 24
// MMAs), the rounding of r_i + r_j - 2 S and of the exact fp32 evaluation itself (<= DP 2^-24 d2) are all below
// 2^-19 (r_i + r_j); 2^-17 leaves a factor 4.
constexpr float kCntMargin = 7.62939453125e-06f;

template <int DP>
struct CountCfg {
    static constexpr int kKB = DP / 32;
    static constexpr int kTile = 128 * 128;                       // bytes of one operand K block: 128 rows x 32 floats
    static constexpr int kOperandBytes = 4 * kKB * kTile;          // A hi | A lo | B hi | B lo
    static constexpr int kListBytes = 8 * kCntPairCap * 2;         // per-warp flagged (row, column) lists, u16
    static constexpr int kValBytes = 8 * kCntValCap * 4;           // per-warp exact candidates
    static constexpr int kSmemBytes = kOperandBytes + kListBytes + kValBytes + 3 * 128 * 4 + 1024 + 256;
};

// linear index over the tiles (ti <= tj) of a T x T tile grid, row-major over the upper triangle
__device__ __forceinline__ void tri_tile(long long lin, int T, int* ti, int* tj) {
    int i = 0;
    long long rowlen = T;
    while (lin >= rowlen) { lin -= rowlen; --rowlen; ++i; }
    *ti = i;
    *tj = i + (int)lin;
}

template <int DP>
__global__ void __launch_bounds__(kCntThreads, (DP == 32 ? 2 : 1))
k_count_tc(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const float* __restrict__ xc,
           const float* __restrict__ r, const float* __restrict__ ssum, long long n, int d, int T, long long lin0, long long lin_step,
           SelHdr* __restrict__ hdr, float* __restrict__ cand, long long cand_cap, unsigned* __restrict__ ghist) {
    using Cfg = CountCfg<DP>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint16_t* lists = reinterpret_cast<uint16_t*>(gen + Cfg::kOperandBytes);
    float* vals = reinterpret_cast<float*>(gen + Cfg::kOperandBytes + Cfg::kListBytes);
    float* rj_s = reinterpret_cast<float*>(gen + Cfg::kOperandBytes + Cfg::kListBytes + Cfg::kValBytes);
    float* sj_s = rj_s + 128;
    float* si_s = sj_s + 128;
    const uint32_t bars = base + Cfg::kOperandBytes + Cfg::kListBytes + Cfg::kValBytes + 3 * 128 * 4;     // full, tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + Cfg::kOperandBytes + Cfg::kListBytes + Cfg::kValBytes + 3 * 128 * 4 + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ti, tj;
    tri_tile(lin0 + (long long)blockIdx.x * lin_step, T, &ti, &tj);
    const long long i0 = (long long)ti * 128, j0 = (long long)tj * 128;
    const uint32_t full_bar = bars, tmem_full_bar = bars + 8;

    if (warp == 0 && lane == 0) {
        mbar_init(full_bar, 1);
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(full_bar, Cfg::kOperandBytes);
#pragma unroll
            for (int kb = 0; kb < Cfg::kKB; ++kb) {
                tma_load_2d(base + (0 * Cfg::kKB + kb) * Cfg::kTile, &tm_hi, full_bar, kb * 32, (int)i0);
                tma_load_2d(base + (1 * Cfg::kKB + kb) * Cfg::kTile, &tm_lo, full_bar, kb * 32, (int)i0);
                tma_load_2d(base + (2 * Cfg::kKB + kb) * Cfg::kTile, &tm_hi, full_bar, kb * 32, (int)j0);
                tma_load_2d(base + (3 * Cfg::kKB + kb) * Cfg::kTile, &tm_lo, full_bar, kb * 32, (int)j0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(128, 128);
            mbar_wait(full_bar, 0);
            tc_fence_after();
#pragma unroll
            for (int kb = 0; kb < Cfg::kKB; ++kb) {
                const uint64_t a_hi = make_kmajor_sw128_desc(base + (0 * Cfg::kKB + kb) * Cfg::kTile);
                const uint64_t a_lo = make_kmajor_sw128_desc(base + (1 * Cfg::kKB + kb) * Cfg::kTile);
                const uint64_t b_hi = make_kmajor_sw128_desc(base + (2 * Cfg::kKB + kb) * Cfg::kTile);
                const uint64_t b_lo = make_kmajor_sw128_desc(base + (3 * Cfg::kKB + kb) * Cfg::kTile);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);
                    umma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0);
                    umma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
                    umma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, 1);
                }
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        // ---- classification: warp cw owns TMEM lane quarter q (rows 32 q .. 32 q + 31 of the tile) x column half ch ----
        const int cw = warp - 2, q = warp & 3, ch = cw >> 2;
        const int t256 = threadIdx.x - 64;
        if (t256 < 128) {
            const long long j = j0 + t256;
            rj_s[t256] = j < n ? r[j] : 0.f;
            sj_s[t256] = j < n ? ssum[j] : 0.f;
        } else {
            const long long i = i0 + (t256 - 128);
            si_s[t256 - 128] = i < n ? ssum[i] : 0.f;
        }
        const int row = 32 * q + lane;
        const long long gi = i0 + row;
        const float ri = gi < n ? r[gi] : 0.f;
        const float lo_b = hdr->lo, hi_b = hdr->hi;
        const float scale = hi_b > lo_b ? (float)kSelBins / (hi_b - lo_b) : 0.f;
        named_bar_sync(1, 256);
        const float si = si_s[row];
        uint16_t* my_list = lists + cw * kCntPairCap;
        float* my_vals = vals + cw * kCntValCap;
        const unsigned lt_mask = (1u << lane) - 1u;
        unsigned zeros = 0, below = 0, vcnt = 0;      // vcnt: warp-uniform
        auto flush_vals = [&]() {
            unsigned gb = 0;
            if (lane == 0) gb = atomicAdd(&hdr->ncand, vcnt);
            gb = __shfl_sync(0xffffffffu, gb, 0);
            if ((long long)gb + vcnt > cand_cap) { if (lane == 0) hdr->overflow = 1u; }
            else for (unsigned k = lane; k < vcnt; k += 32) cand[gb + k] = my_vals[k];
            __syncwarp();
            vcnt = 0;
        };
        const bool interior = ti != tj && i0 + 128 <= n && j0 + 128 <= n;     // no bounds / j > i tests needed
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = 0; chunk < 2; ++chunk) {
            const int col0 = 64 * ch + 32 * chunk;
            float acc[32];
            tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, acc);
            unsigned cnt = 0;                         // warp-uniform
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int col = col0 + e;
                const float rj = rj_s[col];
                const float d2 = fmaf(-2.f, acc[e], ri + rj);
                const float w = pair_w(d2, si, sj_s[col], d);
                const float mg = kCntMargin * (ri + rj);
                const bool valid = interior || (gi < n && j0 + col < n && j0 + col > gi);
                const bool flag = valid && (w <= mg || (w >= lo_b - mg && w <= hi_b + mg));
                below += (valid && !flag && w < lo_b) ? 1u : 0u;
                const unsigned m = __ballot_sync(0xffffffffu, flag);
                if (flag) my_list[cnt + __popc(m & lt_mask)] = (uint16_t)((row << 8) | col);
                cnt += __popc(m);
            }
            __syncwarp();
            // exact re-evaluation of the flagged pairs, one pair per lane
            for (unsigned b0 = 0; b0 < cnt; b0 += 32) {
                const unsigned idx = b0 + lane;
                bool is_c = false;
                float v = 0.f;
                if (idx < cnt) {
                    const unsigned pc = my_list[idx];
                    const int prow = pc >> 8, pcol = pc & 255;
                    v = pair_w(exact_d2<DP>(xc, i0 + prow, j0 + pcol), si_s[prow], sj_s[pcol], d);
                    zeros += v == 0.f ? 1u : 0u;
                    below += (v > 0.f && v < lo_b) ? 1u : 0u;
                    is_c = v > 0.f && v >= lo_b && v <= hi_b;
                }
                const unsigned m = __ballot_sync(0xffffffffu, is_c);
                if (m) {
                    if (is_c) {
                        my_vals[vcnt + __popc(m & lt_mask)] = v;
                        atomicAdd(&ghist[sel_lbin(v, lo_b, scale)], 1u);
                    }
                    vcnt += __popc(m);
                    __syncwarp();
                    if (vcnt > kCntValCap - 32) flush_vals();
                }
            }
            __syncwarp();
        }
        if (vcnt) flush_vals();
        unsigned long long z = zeros, b = below;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            z += __shfl_xor_sync(0xffffffffu, z, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        if (lane == 0) {
            if (z) atomicAdd(&hdr->zeros, z);
            if (b) atomicAdd(&hdr->below, b);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 128);
}

// =========================================================================================================================
// K3b flash kernel
// =========================================================================================================================
constexpr int kFlashThreads = 320;        // warp 0: TMA, warp 1: TMEM + MMA issue, warps 2..9: two SIMT groups of four warps
constexpr int kFlashBJ = 64;              // columns (particles j) per block
constexpr int kFlashChainBlocks = 8;      // column blocks accumulated into one TMEM accumulator (512 columns, see header)

template <int DP, int TILES>
struct FlashCfg {
    static constexpr int NV = 2 * DP;                              // [G | Xc] columns = N of UMMA #2
    static constexpr int kKB1 = DP / 32;                           // K blocks of UMMA #1
    static constexpr int kKB2 = kFlashBJ / 32;                     // K blocks of UMMA #2
    static constexpr int kXjBytes = kFlashBJ * 128;                // one K block of the Xc_j tile: 64 rows x 32 floats
    static constexpr int kVtBytes = NV * 128;                      // one K block of the [G | Xc]^T tile: NV rows x 32 floats
    static constexpr int kStageBytes = 2 * kKB1 * kXjBytes + 2 * kKB2 * kVtBytes;
    static constexpr int kStages = DP == 32 ? 4 : 2;
    static constexpr int kNumBars = 2 * kStages + 6;               // full[], empty[], s_full[2], k_full[2], xi_full, o_full
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
    static constexpr int kColS0 = 0;                               // buffer b: S / K_hi at 128 b, K_lo at 128 b + 64
    static constexpr int kColO = 256;                              // O of tile t at + t NV
    static constexpr int kColXi = 256 + TILES * NV;                // Xc_i hi of tile t at + 2 DP t, lo at + DP
    static_assert(kColXi + TILES * 2 * DP <= 512, "TMEM budget");
};

template <int DP, int TILES>
__global__ void __launch_bounds__(kFlashThreads, 1)
k_phi_flash(const __grid_constant__ CUtensorMap tm_xj_hi, const __grid_constant__ CUtensorMap tm_xj_lo,
            const __grid_constant__ CUtensorMap tm_vt_hi, const __grid_constant__ CUtensorMap tm_vt_lo,
            const float* __restrict__ xc_hi, const float* __restrict__ xc_lo, const float* __restrict__ r, const float* __restrict__ ncr,
            const float* __restrict__ hptr, long long n, long long row_begin, long long n_rows, float* __restrict__ o,
            float* __restrict__ rs_part, int nj_total, int jb_per_split) {
    using Cfg = FlashCfg<DP, TILES>;
    constexpr int NV = Cfg::NV;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + Cfg::kStages * Cfg::kStageBytes;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + Cfg::kStages * Cfg::kStageBytes + 8 * Cfg::kNumBars);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * 128 * TILES;       // first local row of this CTA
    const int jb0 = blockIdx.y * jb_per_split;
    const int nj = min(nj_total, jb0 + jb_per_split) - jb0;         // >= 1 (host)
    const int W = nj * TILES;                                       // work items: (column block, row tile)

    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::kStages + s); };
    auto s_full = [&](int b) { return bars + 8u * (2 * Cfg::kStages + b); };
    auto k_full = [&](int b) { return bars + 8u * (2 * Cfg::kStages + 2 + b); };
    const uint32_t xi_full = bars + 8u * (2 * Cfg::kStages + 4), o_full = bars + 8u * (2 * Cfg::kStages + 5);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(k_full(b), 4); }
        mbar_init(xi_full, 8);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer: per column block the Xc_j tile (hi, lo) and the [G | Xc]^T tile (hi, lo) ----
        if (lane == 0) {
            for (int j = 0; j < nj; ++j) {
                const int s = j % Cfg::kStages;
                mbar_wait(empty_bar(s), ((j / Cfg::kStages) & 1) ^ 1);
                const uint32_t sb = base + s * Cfg::kStageBytes;
                const int jx = (jb0 + j) * kFlashBJ;
                mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
#pragma unroll
                for (int kb = 0; kb < Cfg::kKB1; ++kb) {
                    tma_load_2d(sb + kb * Cfg::kXjBytes, &tm_xj_hi, full_bar(s), kb * 32, jx);
                    tma_load_2d(sb + (Cfg::kKB1 + kb) * Cfg::kXjBytes, &tm_xj_lo, full_bar(s), kb * 32, jx);
                }
                const uint32_t vb = sb + 2 * Cfg::kKB1 * Cfg::kXjBytes;
#pragma unroll
                for (int kb = 0; kb < Cfg::kKB2; ++kb) {
                    tma_load_2d(vb + kb * Cfg::kVtBytes, &tm_vt_hi, full_bar(s), jx + kb * 32, 0);
                    tma_load_2d(vb + (Cfg::kKB2 + kb) * Cfg::kVtBytes, &tm_vt_lo, full_bar(s), jx + kb * 32, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: S(w + 1) is queued before O += K(w) so that the tensor pipe never waits for the SIMT warps twice ----
        if (lane == 0) {
            constexpr uint32_t idesc1 = make_idesc_tf32(128, kFlashBJ), idesc2 = make_idesc_tf32(128, NV);
            auto issue_s = [&](int w) {
                const int j = w / TILES, t = w % TILES, b = w & 1, s = j % Cfg::kStages;
                if (t == 0) {
                    mbar_wait(full_bar(s), (j / Cfg::kStages) & 1);
                    tc_fence_after();
                }
                const uint32_t sb = base + s * Cfg::kStageBytes;
                const uint32_t d_s = tmem + Cfg::kColS0 + 128 * b;
                const uint32_t a_hi = tmem + Cfg::kColXi + 2 * DP * t, a_lo = a_hi + DP;
#pragma unroll
                for (int kb = 0; kb < Cfg::kKB1; ++kb) {
                    const uint64_t b_hi = make_kmajor_sw128_desc(sb + kb * Cfg::kXjBytes);
                    const uint64_t b_lo = make_kmajor_sw128_desc(sb + (Cfg::kKB1 + kb) * Cfg::kXjBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);
                        const uint32_t ka = (uint32_t)(kb * 32 + k * kUmmaK);
                        umma_tf32_ts(d_s, a_hi + ka, b_hi + adv, idesc1, (kb | k) != 0);
                        umma_tf32_ts(d_s, a_hi + ka, b_lo + adv, idesc1, 1);
                        umma_tf32_ts(d_s, a_lo + ka, b_hi + adv, idesc1, 1);
                    }
                }
                umma_commit(s_full(b));
            };
            auto issue_o = [&](int w) {
                const int j = w / TILES, t = w % TILES, b = w & 1, s = j % Cfg::kStages;
                mbar_wait(k_full(b), (w >> 1) & 1);
                tc_fence_after();
                const uint32_t vb = base + s * Cfg::kStageBytes + 2 * Cfg::kKB1 * Cfg::kXjBytes;
                const uint32_t d_o = tmem + Cfg::kColO + NV * t;
                const uint32_t a_hi = tmem + Cfg::kColS0 + 128 * b, a_lo = a_hi + kFlashBJ;
#pragma unroll
                for (int kb = 0; kb < Cfg::kKB2; ++kb) {
                    const uint64_t b_hi = make_kmajor_sw128_desc(vb + kb * Cfg::kVtBytes);
                    const uint64_t b_lo = make_kmajor_sw128_desc(vb + (Cfg::kKB2 + kb) * Cfg::kVtBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);
                        const uint32_t ka = (uint32_t)(kb * 32 + k * kUmmaK);
                        umma_tf32_ts(d_o, a_hi + ka, b_hi + adv, idesc2, (j | kb | k) != 0);
                        umma_tf32_ts(d_o, a_hi + ka, b_lo + adv, idesc2, 1);
                        umma_tf32_ts(d_o, a_lo + ka, b_hi + adv, idesc2, 1);
                    }
                }
                if (t == TILES - 1) umma_commit(empty_bar(s));       // the stage's tiles are no longer read
            };
            mbar_wait(xi_full, 0);
            tc_fence_after();
            issue_s(0);
            for (int w = 0; w < W; ++w) {
                if (w + 1 < W) issue_s(w + 1);
                issue_o(w);
            }
            umma_commit(o_full);
        }
    } else {
        // ---- SIMT warps: group g = buffer g; thread <-> row 32 q + lane of a 128-row tile ----
        const int g = (warp - 2) >> 2, q = warp & 3;
        const int row = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        const float c = 1.4426950408889634f / __ldg(hptr), c2 = 2.f * c;
        // Xc rows of the CTA's tile(s) -> TMEM (A operand of UMMA #1 for the whole kernel)
        {
            const int t = TILES == 2 ? g : 0;
            const long long lr = m0 + 128 * t + row;
            const bool ok = lr < n_rows;
            const long long gi = row_begin + lr;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                if (TILES == 1 && part != g) continue;             // one tile: group 0 stores hi, group 1 lo
                const float* src = (part == 0 ? xc_hi : xc_lo) + gi * DP;
#pragma unroll
                for (int c0 = 0; c0 < DP; c0 += 32) {
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 x4 = ok ? __ldg(reinterpret_cast<const float4*>(src + c0) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[4 * k] = x4.x; v[4 * k + 1] = x4.y; v[4 * k + 2] = x4.z; v[4 * k + 3] = x4.w;
                    }
                    tmem_st32(tmem + lane_addr + (uint32_t)(Cfg::kColXi + 2 * DP * t + DP * part + c0), v);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(xi_full);
        }
        const int my_t = TILES == 2 ? g : 0;
        const long long my_lr = m0 + 128 * my_t + row;
        const long long my_gi = row_begin + my_lr;
        const float nri = my_lr < n_rows ? -c * __ldg(r + my_gi) : 0.f;
        const long long tile_gi0 = row_begin + m0 + 128 * my_t;
        float rs = 0.f;
        for (int w = g; w < W; w += 2) {
            const int j = w / TILES;
            const long long jx = (long long)(jb0 + j) * kFlashBJ;
            mbar_wait(s_full(g), (w >> 1) & 1);
            tc_fence_after();
            const bool diag = jx < tile_gi0 + 128 && jx + kFlashBJ > tile_gi0;
            const uint32_t t_s = tmem + lane_addr + (uint32_t)(Cfg::kColS0 + 128 * g);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[32], lo[32];
                tmem_ld32(t_s + 32 * half, v);
                const float4* nc4 = reinterpret_cast<const float4*>(ncr + jx + 32 * half);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 nc = __ldg(nc4 + k);
                    const float ncv[4] = {nc.x, nc.y, nc.z, nc.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float arg = fminf(fmaf(v[4 * k + e], c2, nri + ncv[e]), 0.f);
                        if (diag && jx + 32 * half + 4 * k + e == my_gi) arg = 0.f;      // K_ii = 1 exactly
                        const float kv = ex2_fast(arg);
                        const float hi = __uint_as_float(__float_as_uint(kv) & 0xFFFFE000u);
                        rs += kv;
                        v[4 * k + e] = hi;
                        lo[4 * k + e] = kv - hi;
                    }
                }
                tmem_st32(t_s + 32 * half, v);
                tmem_st32(t_s + kFlashBJ + 32 * half, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(k_full(g));
        }
        // ---- epilogue: partial O and partial row sums of this column split ----
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int c_begin = TILES == 2 ? 0 : g * (NV / 2), c_end = TILES == 2 ? NV : (g + 1) * (NV / 2);
        float* dst = o + ((long long)blockIdx.y * n_rows + my_lr) * NV;
#pragma unroll 1
        for (int cc = c_begin; cc < c_end; cc += 32) {
            float acc[32];
            tmem_ld32(tmem + lane_addr + (uint32_t)(Cfg::kColO + NV * my_t + cc), acc);
            if (my_lr < n_rows) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + cc + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            }
        }
        if (my_lr < n_rows) {
            const long long slot = TILES == 2 ? (long long)blockIdx.y : 2ll * blockIdx.y + g;
            rs_part[slot * n_rows + my_lr] = rs;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace ocbnn
