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

// =========================================================================================================================
// K3a counting pass on recomputed tiles
// =========================================================================================================================
constexpr int kCntThreads = 320;          // warp 0: TMA, warp 1: TMEM + MMA issue, warps 2..9: classification
constexpr int kCntPairCap = 1024;         // flagged pairs per warp and 32-column chunk (32 rows x 32 columns: cannot overflow)
constexpr int kCntValCap = 256;           // exact candidates staged per warp before one global append
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
    static constexpr int kHistBytes = kSelBins * 4;                // per-CTA copy of the linear histogram over the bracket
    static constexpr int kSmemBytes = kOperandBytes + kListBytes + kValBytes + kHistBytes + 4 * 128 * 4 + 1024 + 256;
    static constexpr int kCtasPerSm = DP == 32 ? 2 : 1;
};

// linear index over the tiles (ti <= tj) of a T x T tile grid, row-major over the upper triangle
__device__ __forceinline__ void tri_tile(long long lin, int T, int* ti, int* tj) {
    // row i starts at i T - i (i - 1) / 2: invert with a float estimate, then fix up
    const double Td = (double)T + 0.5;
    int i = (int)(Td - sqrt(fmax(Td * Td - 2.0 * (double)lin, 0.0)));
    if (i < 0) i = 0;
    if (i > T - 1) i = T - 1;
    auto start = [&](int k) { return (long long)k * T - (long long)k * (k - 1) / 2; };
    while (i > 0 && start(i) > lin) --i;
    while (i < T - 1 && start(i + 1) <= lin) ++i;
    *ti = i;
    *tj = i + (int)(lin - start(i));
}

// exact squared distance of rows prow (A tile) and pcol (B tile) from the hi / lo operand tiles in shared memory (x = hi + lo
// exactly; 128-byte swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)), in the canonical order of d2_block32 (eight
// independent chunk sums + tree: also eight independent dependency chains instead of one of 32 fmas)
template <int DP>
__device__ __forceinline__ float exact_d2_smem(const uint8_t* ops, int prow, int pcol) {
    constexpr int kKB = DP / 32, kTile = 128 * 128;
    float acc = 0.f;
#pragma unroll
    for (int kb = 0; kb < kKB; ++kb) {
        const uint8_t* a_hi = ops + (0 * kKB + kb) * kTile + prow * 128;
        const uint8_t* a_lo = ops + (1 * kKB + kb) * kTile + prow * 128;
        const uint8_t* b_hi = ops + (2 * kKB + kb) * kTile + pcol * 128;
        const uint8_t* b_lo = ops + (3 * kKB + kb) * kTile + pcol * 128;
        float p[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int ca = (c ^ (prow & 7)) << 4, cb = (c ^ (pcol & 7)) << 4;
            const float4 ah = *reinterpret_cast<const float4*>(a_hi + ca), al = *reinterpret_cast<const float4*>(a_lo + ca);
            const float4 bh = *reinterpret_cast<const float4*>(b_hi + cb), bl = *reinterpret_cast<const float4*>(b_lo + cb);
            p[c] = d2_chunk(make_float4(ah.x + al.x, ah.y + al.y, ah.z + al.z, ah.w + al.w),
                            make_float4(bh.x + bl.x, bh.y + bl.y, bh.z + bl.z, bh.w + bl.w));
        }
        const float t = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), __fadd_rn(p[2], p[3])), __fadd_rn(__fadd_rn(p[4], p[5]), __fadd_rn(p[6], p[7])));
        acc = kb == 0 ? t : __fadd_rn(acc, t);
    }
    return acc;
}

// Persistent: CTA b takes tiles lin0 + (b + k gridDim.x) lin_step, k = 0, 1, ... of this rank's n_tiles.  Per tile: TMA of the
// four operand tiles -> 12 kKB MMAs -> eight warps classify the 128 x 128 accumulators; pairs near the bracket are re-evaluated
// exactly OUT OF THE OPERAND TILES IN SHARED MEMORY (no global loads).  Counters stay in registers and the bracket histogram in
// shared memory across the CTA's tiles; both are flushed once.
template <int DP>
__global__ void __launch_bounds__(kCntThreads, CountCfg<DP>::kCtasPerSm)
k_count_tc(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const float* __restrict__ r,
           const float* __restrict__ ssum, long long n, int d, int T, long long lin0, long long lin_step, long long n_tiles,
           SelHdr* __restrict__ hdr, float* __restrict__ cand, long long cand_cap, unsigned* __restrict__ ghist, int dbg) {
    using Cfg = CountCfg<DP>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint16_t* lists = reinterpret_cast<uint16_t*>(gen + Cfg::kOperandBytes);
    float* vals = reinterpret_cast<float*>(gen + Cfg::kOperandBytes + Cfg::kListBytes);
    unsigned* shist = reinterpret_cast<unsigned*>(gen + Cfg::kOperandBytes + Cfg::kListBytes + Cfg::kValBytes);
    float2* cj2_s = reinterpret_cast<float2*>(gen + Cfg::kOperandBytes + Cfg::kListBytes + Cfg::kValBytes + Cfg::kHistBytes);
    float* sj_s = reinterpret_cast<float*>(cj2_s + 128);
    float* si_s = sj_s + 128;
    constexpr int kMisc = Cfg::kOperandBytes + Cfg::kListBytes + Cfg::kValBytes + Cfg::kHistBytes + 4 * 128 * 4;
    const uint32_t bars = base + kMisc;                                    // full, tmem_full, done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + kMisc + 32);
    const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
    const uint32_t full_bar = bars, tmem_full_bar = bars + 8, done_bar = bars + 16;
    const long long my_tiles = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0 && lane == 0) {
        mbar_init(full_bar, 1);
        mbar_init(tmem_full_bar, 1);
        mbar_init(done_bar, 8);                                            // one arrival per classification warp
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (elect_one_sync()) {
            for (long long it = 0; it < my_tiles; ++it) {
                int ti, tj;
                tri_tile(lin0 + (blockIdx.x + it * gridDim.x) * lin_step, T, &ti, &tj);
                if (it > 0) mbar_wait(done_bar, (uint32_t)((it - 1) & 1));  // operand tiles (and TMEM) of the previous tile are no longer read
                mbar_expect_tx(full_bar, Cfg::kOperandBytes);
#pragma unroll
                for (int kb = 0; kb < Cfg::kKB; ++kb) {
                    tma_load_2d(base + (0 * Cfg::kKB + kb) * Cfg::kTile, &tm_hi, full_bar, kb * 32, ti * 128);
                    tma_load_2d(base + (1 * Cfg::kKB + kb) * Cfg::kTile, &tm_lo, full_bar, kb * 32, ti * 128);
                    tma_load_2d(base + (2 * Cfg::kKB + kb) * Cfg::kTile, &tm_hi, full_bar, kb * 32, tj * 128);
                    tma_load_2d(base + (3 * Cfg::kKB + kb) * Cfg::kTile, &tm_lo, full_bar, kb * 32, tj * 128);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc_tf32(128, 128);
            for (long long it = 0; it < my_tiles; ++it) {
                mbar_wait(full_bar, (uint32_t)(it & 1));                   // implies done(it - 1): the TMA of this tile waited for it
                tc_fence_after();
                if (!(dbg & 4))
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
        }
    } else {
        // ---- classification: warp cw owns TMEM lane quarter q (rows 32 q .. 32 q + 31 of the tile) x column half ch ----
        const int cw = warp - 2, q = warp & 3, ch = cw >> 2;
        const int t256 = threadIdx.x - 64;
        for (int i = t256; i < kSelBins; i += 256) shist[i] = 0u;
        const int row = 32 * q + lane;
        const float lo_b = hdr->lo, hi_b = hdr->hi;
        const float scale = hi_b > lo_b ? (float)kSelBins / (hi_b - lo_b) : 0.f;
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
        for (long long it = 0; it < my_tiles; ++it) {
            int ti, tj;
            tri_tile(lin0 + (blockIdx.x + it * gridDim.x) * lin_step, T, &ti, &tj);
            const long long i0 = (long long)ti * 128, j0 = (long long)tj * 128;
            named_bar_sync(1, 256);                    // the previous tile's r / s rows are no longer read
            // per column: C_j = r_j - 2 eps s_j + D eps^2 and the margin share kCntMargin r_j; per row the counterparts in
            // registers, so that the classification value is ONE fma: w~ = max(-2 S_ij + (R_i + C_j), 0)
            const float eps2 = 2e-6f, deps2 = (float)d * 1e-12f;
            if (t256 < 128) {
                const long long j = j0 + t256;
                const float rj = j < n ? r[j] : 0.f, sj = j < n ? ssum[j] : 0.f;
                cj2_s[t256] = make_float2(rj - eps2 * sj + deps2, kCntMargin * rj);
                sj_s[t256] = sj;
            } else {
                const long long i = i0 + (t256 - 128);
                si_s[t256 - 128] = i < n ? ssum[i] : 0.f;
            }
            const long long gi = i0 + row;
            const float ri = gi < n ? r[gi] : 0.f;
            named_bar_sync(1, 256);
            const float Ri = ri + eps2 * si_s[row], mRi = kCntMargin * ri;
            const bool interior = ti != tj && i0 + 128 <= n && j0 + 128 <= n;     // no bounds / j > i tests needed
            mbar_wait(tmem_full_bar, (uint32_t)(it & 1));
            tc_fence_after();
#pragma unroll 1
            for (int chunk = 0; chunk < ((dbg & 1) ? 0 : 2); ++chunk) {
                const int col0 = 64 * ch + 32 * chunk;
                float acc[32];
                tmem_ld32(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, acc);
                unsigned fmask = 0;                       // bit e: pair (row, col0 + e) needs the exact evaluation
                if (interior) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float2 c = cj2_s[col0 + e];
                        const float w = fmaxf(fmaf(-2.f, acc[e], Ri + c.x), 0.f), mg = mRi + c.y;
                        const bool is_below = w > mg && w < lo_b - mg, is_above = w > hi_b + mg;
                        below += is_below ? 1u : 0u;
                        if (!(is_below || is_above)) fmask |= 1u << e;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int col = col0 + e;
                        const float2 c = cj2_s[col];
                        const float w = fmaxf(fmaf(-2.f, acc[e], Ri + c.x), 0.f), mg = mRi + c.y;
                        const bool valid = gi < n && j0 + col < n && j0 + col > gi;
                        const bool is_below = w > mg && w < lo_b - mg, is_above = w > hi_b + mg;
                        below += (valid && is_below) ? 1u : 0u;
                        if (valid && !(is_below || is_above)) fmask |= 1u << e;
                    }
                }
                // compaction of the flagged pairs into the warp's list: one prefix sum per chunk
                const unsigned mine = (unsigned)__popc(fmask);
                unsigned incl = mine;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned y = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += y;
                }
                const unsigned cnt = __shfl_sync(0xffffffffu, incl, 31);     // warp-uniform
                unsigned pos = incl - mine;
                while (fmask) {
                    const int e = __ffs(fmask) - 1;
                    fmask &= fmask - 1;
                    my_list[pos++] = (uint16_t)((row << 8) | (col0 + e));
                }
                __syncwarp();
                // exact re-evaluation of the flagged pairs, one pair per lane (eight lanes per pair with conflict-free row reads
                // and a shuffle butterfly was measured 25 % SLOWER: eight times the loop trips)
                for (unsigned b0 = 0; b0 < ((dbg & 2) ? 0u : cnt); b0 += 32) {
                    const unsigned idx = b0 + lane;
                    bool is_c = false;
                    float v = 0.f;
                    if (idx < cnt) {
                        const unsigned pc = my_list[idx];
                        const int prow = pc >> 8, pcol = pc & 255;
                        v = pair_w(exact_d2_smem<DP>(gen, prow, pcol), si_s[prow], sj_s[pcol], d);
                        zeros += v == 0.f ? 1u : 0u;
                        below += (v > 0.f && v < lo_b) ? 1u : 0u;
                        is_c = v > 0.f && v >= lo_b && v <= hi_b;
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, is_c);
                    if (m) {
                        if (is_c) {
                            my_vals[vcnt + __popc(m & lt_mask)] = v;
                            atomicAdd(&shist[sel_lbin(v, lo_b, scale)], 1u);
                        }
                        vcnt += __popc(m);
                        __syncwarp();
                        if (vcnt > kCntValCap - 32) flush_vals();
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(done_bar);
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
        named_bar_sync(1, 256);
        for (int i = t256; i < kSelBins; i += 256) {
            const unsigned c = shist[i];
            if (c) atomicAdd(&ghist[i], c);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 128);
}

// =========================================================================================================================
// K3b flash kernel
// =========================================================================================================================
constexpr int kFlashThreads = 320;        // warp 0: TMA, warp 1: TMEM + MMA issue, warps 2..9: SIMT (all eight on one work item)
constexpr int kFlashBJ = 64;              // columns (particles j) per block
constexpr int kFlashChainBlocks = 16;     // column blocks accumulated into one TMEM accumulator (1024 columns = 384 MMAs, see header)

template <int DP, int TILES>
struct FlashCfg {
    static constexpr int NV = 2 * DP;                              // [G | Xc] columns = N of UMMA #2
    static constexpr int kKB1 = DP / 32;                           // K blocks of UMMA #1
    static constexpr int kKB2 = kFlashBJ / 32;                     // K blocks of UMMA #2
    static constexpr int kXjBytes = kFlashBJ * 128;                // one K block of the Xc_j tile: 64 rows x 32 floats
    static constexpr int kVtBytes = NV * 128;                      // one K block of the [G | Xc]^T tile: NV rows x 32 floats
    static constexpr int kStageBytes = 2 * kKB1 * kXjBytes + 2 * kKB2 * kVtBytes;
    static constexpr int kStages = DP == 32 ? 4 : 2;
    static constexpr int kNumBars = 2 * kStages + 7;               // full[], empty[], s_full[2], k_full[2], xi_full, o_full, o_empty
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
    static constexpr int kColS0 = 0;                               // buffer b: S / K_hi at 128 b, K_lo at 128 b + 64
    static constexpr int kColO = 256;                              // O of tile t at + t NV
    static constexpr int kColXi = 256 + TILES * NV;                // Xc_i hi of tile t at + 2 DP t, lo at + DP
    static_assert(kColXi + TILES * 2 * DP <= 512, "TMEM budget");
};

// Work items w = (column block j, row tile t), t fastest; item w uses S/K buffer w & 1.  The MMA warp keeps S TWO items ahead:
// S(0), S(1), then per item [wait K(w)] O(w) S(w + 2) - S(w + 2) reuses buffer w & 1 behind O(w) in the (in-order) tensor pipe -
// so the eight SIMT warps find S(w + 1) complete when they finish K(w) and never wait in steady state.  SIMT warp (q, half):
// TMEM lanes 32 q .. 32 q + 31 (rows), columns 32 half .. 32 half + 31 of the 64-column block: 32 exponentials per thread.
template <int DP, int TILES>
__global__ void __launch_bounds__(kFlashThreads, 1)
k_phi_flash(const __grid_constant__ CUtensorMap tm_xj_hi, const __grid_constant__ CUtensorMap tm_xj_lo,
            const __grid_constant__ CUtensorMap tm_vt_hi, const __grid_constant__ CUtensorMap tm_vt_lo,
            const float* __restrict__ xc_hi, const float* __restrict__ xc_lo, const float* __restrict__ r, const float* __restrict__ ncr,
            const float* __restrict__ hptr, long long n, long long row_begin, long long n_rows, float* __restrict__ o,
            float* __restrict__ rs_part, int nj_total, int jb_per_split, int dbg) {
    using Cfg = FlashCfg<DP, TILES>;
    constexpr int NV = Cfg::NV;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + Cfg::kStages * Cfg::kStageBytes;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + Cfg::kStages * Cfg::kStageBytes + 8 * Cfg::kNumBars);
    const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * 128 * TILES;       // first local row of this CTA
    const int jb0 = blockIdx.y * jb_per_split;
    const int nj = min(nj_total, jb0 + jb_per_split) - jb0;         // >= 1 (host)
    const int W = nj * TILES;                                       // work items: (column block, row tile)

    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::kStages + s); };
    auto s_full = [&](int b) { return bars + 8u * (2 * Cfg::kStages + b); };
    auto k_full = [&](int b) { return bars + 8u * (2 * Cfg::kStages + 2 + b); };
    const uint32_t xi_full = bars + 8u * (2 * Cfg::kStages + 4), o_full = bars + 8u * (2 * Cfg::kStages + 5), o_empty = bars + 8u * (2 * Cfg::kStages + 6);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(k_full(b), 8); }
        mbar_init(xi_full, 8);
        mbar_init(o_full, 1);
        mbar_init(o_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer: per column block the Xc_j tile (hi, lo) and the [G | Xc]^T tile (hi, lo) ----
        if (elect_one_sync()) {
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
        if (elect_one_sync()) {
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
                if (!(dbg & 4))
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
                const int jc = j % kFlashChainBlocks;                  // position inside the accumulation chain
                if (jc == 0 && t == 0 && j > 0) {                      // a new chain overwrites O: the SIMT warps must have drained the last one
                    mbar_wait(o_empty, ((j / kFlashChainBlocks) - 1) & 1);
                    tc_fence_after();
                }
                const uint32_t vb = base + s * Cfg::kStageBytes + 2 * Cfg::kKB1 * Cfg::kXjBytes;
                const uint32_t d_o = tmem + Cfg::kColO + NV * t;
                const uint32_t a_hi = tmem + Cfg::kColS0 + 128 * b, a_lo = a_hi + kFlashBJ;
                if (!(dbg & 2))
#pragma unroll
                for (int kb = 0; kb < Cfg::kKB2; ++kb) {
                    const uint64_t b_hi = make_kmajor_sw128_desc(vb + kb * Cfg::kVtBytes);
                    const uint64_t b_lo = make_kmajor_sw128_desc(vb + (Cfg::kKB2 + kb) * Cfg::kVtBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);
                        const uint32_t ka = (uint32_t)(kb * 32 + k * kUmmaK);
                        umma_tf32_ts(d_o, a_hi + ka, b_hi + adv, idesc2, (jc | kb | k) != 0);
                        umma_tf32_ts(d_o, a_hi + ka, b_lo + adv, idesc2, 1);
                        umma_tf32_ts(d_o, a_lo + ka, b_hi + adv, idesc2, 1);
                    }
                }
                if (t == TILES - 1) {
                    umma_commit(empty_bar(s));                         // the stage's tiles are no longer read
                    if (jc == kFlashChainBlocks - 1 || j == nj - 1) umma_commit(o_full);      // chain complete
                }
            };
            mbar_wait(xi_full, 0);
            tc_fence_after();
            issue_s(0);
            if (W > 1) issue_s(1);
            for (int w = 0; w < W; ++w) {
                issue_o(w);
                if (w + 2 < W) issue_s(w + 2);
            }
        }
    } else {
        // ---- SIMT warps ----
        const int half = (warp - 2) >> 2, q = warp & 3;
        const int row = 32 * q + lane;
        const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
        const float c = 1.4426950408889634f / __ldg(hptr), c2 = 2.f * c;
        // Xc rows of the CTA's tile(s) -> TMEM (A operand of UMMA #1 for the whole kernel): TILES == 2: warp half h stores tile h
        // (hi and lo); TILES == 1: half 0 stores hi, half 1 lo
        {
            const int t = TILES == 2 ? half : 0;
            const long long lr = m0 + 128 * t + row;
            const bool ok = lr < n_rows;
            const long long gi = row_begin + lr;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                if (TILES == 1 && part != half) continue;
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
        float nri[TILES], rs[TILES];
        long long my_gi[TILES];
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
            const long long lr = m0 + 128 * t + row;
            my_gi[t] = row_begin + lr;
            nri[t] = lr < n_rows ? -c * __ldg(r + my_gi[t]) : 0.f;
            rs[t] = 0.f;
        }
        // O of this thread's row accumulates over the chains in registers (round-to-nearest adds): TILES == 2: warp half h owns
        // tile h, all NV columns; TILES == 1: half h owns columns [h NV/2, (h + 1) NV/2)
        constexpr int NC = TILES == 2 ? NV : NV / 2;
        constexpr int kChainItems = kFlashChainBlocks * TILES;
        const int et = TILES == 2 ? half : 0, c_begin = TILES == 2 ? 0 : half * (NV / 2);
        float oacc[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) oacc[i] = 0.f;
        int chains_done = 0;
        auto drain = [&]() {
            mbar_wait(o_full, chains_done & 1);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < NC; cc += 32) {
                float acc[32];
                tmem_ld32(tmem + lane_addr + (uint32_t)(Cfg::kColO + NV * et + c_begin + cc), acc);
#pragma unroll
                for (int i = 0; i < 32; ++i) oacc[cc + i] += acc[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
            ++chains_done;
        };
        const uint32_t col_off = 32u * half;
        for (int w = 0; w < W; ++w) {
            const int j = w / TILES, t = w % TILES, b = w & 1;
            // a finished chain is drained ONE ITEM LATE: its last O has retired by then (no wait), and the MMA warp holds the next
            // chain's first O until the drain is done
            if (w > 1 && (w - 1) % kChainItems == 0) drain();
            const long long jx = (long long)(jb0 + j) * kFlashBJ + col_off;     // first column of this warp's 32
            const long long tile_gi0 = row_begin + m0 + 128 * t;
            const bool diag = jx < tile_gi0 + 128 && jx + 32 > tile_gi0;      // the diagonal crosses this warp's block
            const float my_nri = TILES == 2 ? (t ? nri[TILES - 1] : nri[0]) : nri[0];
            const long long gi = TILES == 2 ? (t ? my_gi[TILES - 1] : my_gi[0]) : my_gi[0];
            mbar_wait(s_full(b), (w >> 1) & 1);
            tc_fence_after();
            const uint32_t t_s = tmem + lane_addr + (uint32_t)(Cfg::kColS0 + 128 * b) + col_off;
            float v[32], lo[32];
            if (dbg & 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(k_full(b));
                continue;
            }
            if (dbg & 8) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            } else tmem_ld32(t_s, v);
            const float4* nc4 = reinterpret_cast<const float4*>(ncr + jx);
            float rsum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 nc = __ldg(nc4 + k);
                const float ncv[4] = {nc.x, nc.y, nc.z, nc.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float arg = fminf(fmaf(v[4 * k + e], c2, my_nri + ncv[e]), 0.f);
                    if (diag && jx + 4 * k + e == gi) arg = 0.f;              // K_ii = 1 exactly
                    const float kv = ex2_fast(arg);
                    const float hi = __uint_as_float(__float_as_uint(kv) & 0xFFFFE000u);
                    rsum += kv;
                    v[4 * k + e] = hi;
                    lo[4 * k + e] = kv - hi;
                }
            }
            if (!(dbg & 16)) {
                tmem_st32(t_s, v);
                tmem_st32(t_s + kFlashBJ, lo);
            }
            if (TILES == 2 && t) rs[TILES - 1] += rsum; else rs[0] += rsum;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(k_full(b));
        }
        if (W > 1 && (W - 1) % kChainItems == 0) drain();                     // the chain that ended one item before the last
        drain();                                                              // the last chain
        // ---- epilogue: O and the partial row sums of this column split ----
        const long long e_lr = m0 + 128 * et + row;
        if (e_lr < n_rows) {
            float* dst = o + ((long long)blockIdx.y * n_rows + e_lr) * NV + c_begin;
#pragma unroll
            for (int i = 0; i < NC; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(oacc[i], oacc[i + 1], oacc[i + 2], oacc[i + 3]);
        }
        // partial row sums: slot 2 split + half, per tile
#pragma unroll
        for (int t = 0; t < TILES; ++t) {
            const long long lr = m0 + 128 * t + row;
            if (lr < n_rows) rs_part[(2ll * blockIdx.y + half) * n_rows + lr] = rs[t];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace ocbnn
