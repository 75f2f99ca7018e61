// svgd_tc.cu — tensor-core path of the SVGD update (K3b): the two dense contractions
//     S = Xc Xc^T                 (n x n x D)      -> K_ij = exp(-(r_i + r_j - 2 S_ij)/h)
//     O = K [G | Xc]              (n x 2D x n)     -> phi_i = (O_G - (2/h)(O_X - x_i rowsum_i)) / n
// run on tcgen05 with TMA-staged operands in 3xTF32 precision (tc_gemm.cuh); Xc = X - mean(X) (phi is translation
// invariant; centring removes the cancellation in the Gram form of the distance).  The exact fp32 SIMT arm in svgd.cu is
// the parity reference of this path.  Also exports the GEMM building block itself for testing.
#include <cstdlib>
#include <type_traits>
#include <cudaTypedefs.h>
#include "tc_gemm.cuh"

namespace ocbnn {
namespace tc {

// ---- tensor maps --------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }();
    return fn;
}

// row-major fp32 matrix [rows x cols] (cols % 4 == 0), box = box_rows x 32 columns, 128-byte swizzle, OOB reads as zero
static int make_tmap(CUtensorMap* m, const float* ptr, long long rows, long long cols, int box_rows) {
    auto fn = encode_fn();
    OCBNN_REQUIRE(fn, OCBNN_ENODEVICE, "cuTensorMapEncodeTiled is not available from this driver");
    OCBNN_REQUIRE(cols % 4 == 0 && ((uintptr_t)ptr & 15) == 0, OCBNN_EINVAL, "TMA operand needs a 16-byte aligned base and row pitch");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    OCBNN_REQUIRE(r == CUDA_SUCCESS, OCBNN_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return OCBNN_OK;
}

// ---- epilogues ----------------------------------------------------------------------------------------------------------
struct EpiStore {          // C[split][row][col] = acc
    float* c;
    long long ldc, split_stride;
    int m, n;
    __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32], int split) const {
        if (row >= m) return;
        float* dst = c + split * split_stride + (long long)row * ldc + col0;
        if (col0 + 32 <= n && (ldc & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < n) dst[i] = acc[i];
        }
    }
};

// Gram epilogue: d2[row][col] = max(r_i + r_j - 2 S, 0) for the local rows of a rank (i = row_begin + row), zero in the n..n4 padding
struct EpiGramRows {
    float* d2;
    const float* r;
    long long n, n4, row_begin, n_rows;
    __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32], int) const {
        if (row >= n_rows) return;
        const float ri = r[row_begin + row];
        float* dst = d2 + (long long)row * n4 + col0;
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (col0 + i < n4) {
                // the diagonal is exactly 0 (the Gram form would leave the tensor core's accumulation bias of S_ii there, and
                // K_ii = 1 is the largest term of every row)
                const bool off_diag = col0 + i < n && col0 + i != row_begin + row;
                dst[i] = off_diag ? fmaxf(fmaf(-2.f, acc[i], ri + r[col0 + i]), 0.f) : 0.f;
            }
    }
};

template <int BN, class Epi>
static int launch_gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo, const Epi& epi,
                       long long m, long long n, long long k, int splits, cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    auto kern = k_gemm_tf32x3<BN, Epi>;
    OCBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    const int num_kb = (int)((k + kBK - 1) / kBK);
    const int per = (num_kb + splits - 1) / splits;
    dim3 grid((unsigned)((m + kBM - 1) / kBM), (unsigned)((n + BN - 1) / BN), (unsigned)splits);
    kern<<<grid, kThreads, Cfg::kSmemBytes, st>>>(a_hi, a_lo, b_hi, b_lo, epi, num_kb, per);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

static unsigned grid_for_ll(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// hi = x with the 13 low mantissa bits cleared (exact tf32), lo = x - hi (exact in fp32); optional column padding with zeros
__global__ void __launch_bounds__(256)
k_split_tf32(const float* __restrict__ x, long long rows, long long cols, long long ld_out, float* __restrict__ hi, float* __restrict__ lo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = rows * ld_out, stride = (long long)gridDim.x * blockDim.x;
    for (; i < tot; i += stride) {
        const long long r = i / ld_out, c = i - r * ld_out;
        const float v = c < cols ? x[r * cols + c] : 0.f;
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}


// =========================================================================================================================
// SVGD tensor pipeline.  Workspace layout (floats unless noted), n4 = n rounded up to 4, dp = D rounded up to 32:
//   hdr (256 B) | mean[dp] | r[n] | s[n] | xc[n x dp] | xc_hi[n x dp] | xc_lo[n x dp] | d2[n_rows x n4] | k_hi, k_lo[n_rows x n4] |
//   vt_hi, vt_lo[2dp x n4] | rowsum[n_rows] | o[splits x n_rows x 2dp] | select header | select candidates[n^2/32]
// =========================================================================================================================
struct TcWs {
    long long n, d, n4, dp, n_rows, row_begin;
    int splits, bn;
    float *mean, *r, *s, *xc, *xc_hi, *xc_lo, *d2, *k_hi, *k_lo, *vt_hi, *vt_lo, *rowsum, *o;
    float *cand, *list2, *samples;   // bracket select: candidates, short list (cand_cap floats each), pair sample
    unsigned* ghist;                  // bracket select: linear histogram over the bracket
    unsigned long long* red64;        // multi-rank bracket select: [zeros, below, candidates, overflow | ghist as 2048 u64] summed over ranks
    unsigned long long* sel; // bracket-select header: see SelHdr
    long long cand_cap;
    bool flash;                       // D <= 64: svgd_flash.cuh kernels, no d2 / K in the workspace
    int fl_tiles, fl_splits, fl_per, fl_nj;   // flash phi: row tiles per CTA, column splits, column blocks per split / in total
    float* ncr;                       // flash phi: -r_j log2(e) / h per column (-inf beyond n)
    unsigned* hist;
    unsigned long long* state;
    long long bytes;
};

static int pick_splits(long long n_rows, long long n4, int two_dp, int bn) {
    const long long tiles = ((n_rows + kBM - 1) / kBM) * ((two_dp + bn - 1) / bn);
    long long s = (148 + tiles - 1) / tiles;
    const long long kb = (n4 + kBK - 1) / kBK;
    if (s > 16) s = 16;
    if (s > kb) s = kb;
    return (int)(s < 1 ? 1 : s);
}

long long tc_workspace_bytes(long long n, long long d, long long n_rows);      // K splits of the fused phi kernel (defined with it)

constexpr int kSelSamplesHost = 32768;      // = kSelSamples (pair sample of the bracket select, defined with its kernels)
constexpr int kSelSamplesMaxHost = 1 << 17;
// work split of the flash phi kernel: `tiles` 128-row tiles per CTA (2 when D <= 32 and there are enough rows to fill the
// SMs that way), column blocks of 64 particles, at most 8 of them (512 columns) accumulated into one TMEM accumulator
static void flash_plan(long long n, long long n_rows, int dp, int* tiles, int* splits, int* per, int* nj_total) {
    const int nj = (int)((n + 63) / 64);
    const int t = (dp == 32 && n_rows > 128) ? 2 : 1;
    const long long row_blocks = (n_rows + 128 * t - 1) / (128 * t);
    // Column splits: one CTA per SM, so the kernel time is (waves of 148 CTAs) x (column blocks per CTA).  Up to n = 4096 one wave
    // is best (a 149th CTA would double the time); for more rows than SMs several waves of shorter CTAs waste less of the last wave
    // (n = 32768: 128 row blocks x 8 splits = 7 waves x 64 blocks instead of 1 wave x 512 on 128 of the 148 SMs).  The accumulation
    // chain limit is handled inside the kernel (accumulators drained every kFlashChainBlocks blocks).
    int best_s = 1;
    long long best_cost = -1;
    for (int sp = 1; sp <= 32 && sp <= nj; ++sp) {
        const long long waves = (row_blocks * sp + 147) / 148, per_cta = (nj + sp - 1) / sp;
        // a CTA costs ~5 column blocks on top of its own (TMEM allocation, the row tile's X into TMEM, the drain): measured at
        // n = 16384 (16 splits, 7 waves: 313 us against 280 us for 2 splits in one wave) and n = 32768 (8 splits: 966 against 1025 us)
        const long long cost = waves * (per_cta + 5) * 64 + sp;    // (+ sp: fewer partial outputs on ties)
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = sp; }
    }
    int p = (nj + best_s - 1) / best_s;
    if (p < 1) p = 1;
    *tiles = t;
    *per = p;
    *splits = (nj + p - 1) / p;
    *nj_total = nj;
}

static TcWs carve(void* workspace, long long n, long long d, long long row_begin, long long n_rows) {
    TcWs w{};
    w.n = n; w.d = d; w.n4 = (n + 3) / 4 * 4; w.dp = (d + 31) / 32 * 32; w.n_rows = n_rows; w.row_begin = row_begin;
    w.bn = 2 * w.dp <= 64 ? 64 : 128;
    w.splits = pick_splits(n_rows, w.n4, (int)(2 * w.dp), w.bn);
    char* p = reinterpret_cast<char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    auto take = [&](long long floats) { float* q = reinterpret_cast<float*>(p); p += (floats * 4 + 255) / 256 * 256; return q; };
    w.hist = reinterpret_cast<unsigned*>(take(3 * 2048));
    w.state = reinterpret_cast<unsigned long long*>(take(16));
    w.mean = take(w.dp);
    w.r = take(n);
    w.s = take(n);
    w.xc = take(n * w.dp);
    w.xc_hi = take(n * w.dp);
    w.xc_lo = take(n * w.dp);
    w.flash = w.dp <= 64;
    if (w.flash) {
        // D <= 64: neither the distances nor K are ever materialised (svgd_flash.cuh); what is left are the per-split partial
        // outputs of the flash kernel
        flash_plan(n, n_rows, (int)w.dp, &w.fl_tiles, &w.fl_splits, &w.fl_per, &w.fl_nj);
        w.d2 = nullptr;
        w.k_lo = nullptr;
        w.k_hi = take((long long)w.fl_splits * 2 * n_rows);                        // partial row sums (rs_part)
        w.ncr = take((long long)w.fl_nj * 64);
    } else {
        w.fl_tiles = w.fl_splits = w.fl_per = w.fl_nj = 0;
        w.d2 = take(n_rows * w.n4);                                                // distance rows of this rank (Gram GEMM)
        w.k_hi = take(n_rows * w.n4);                                              // K = exp(-d2/h) as tf32 hi / lo matrices
        w.k_lo = take(n_rows * w.n4);
        w.ncr = nullptr;
    }
    w.vt_hi = take(2 * w.dp * w.n4);
    w.vt_lo = take(2 * w.dp * w.n4);
    w.rowsum = take(n_rows);
    w.o = take((long long)(w.flash ? w.fl_splits : w.splits) * n_rows * 2 * w.dp);
    w.sel = reinterpret_cast<unsigned long long*>(take(64));
    // candidates: ~2x the pairs the bracket keeps (3.3 % of the pairs up to n = 4096, narrower above: see sel_samples)
    w.cand_cap = n_rows * n / 16 > (1 << 16) ? n_rows * n / 16 : (1 << 16);
    w.cand = take(w.cand_cap);
    w.list2 = take(w.cand_cap);
    w.samples = take(kSelSamplesMaxHost);
    w.ghist = reinterpret_cast<unsigned*>(take(4096));
    w.red64 = reinterpret_cast<unsigned long long*>(take(2 * (4 + 2048)));
    w.bytes = p - reinterpret_cast<char*>(workspace);
    return w;
}

// column means of x (n x d): one block per column
__global__ void __launch_bounds__(256) k_colmean(const float* __restrict__ x, long long n, int d, int dp, float* __restrict__ mean) {
    const int c = blockIdx.x;
    double s = 0.0;
    if (c < d)
        for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i * d + c];
    __shared__ double red[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        mean[c] = c < d ? (float)(s / (double)n) : 0.f;
    }
}

// one warp per row: xc = x - mean split into tf32 hi/lo (zero padded to dp), r = ||xc||^2, s = sum xc
__global__ void __launch_bounds__(256)
k_center_split(const float* __restrict__ x, const float* __restrict__ mean, long long n, int d, int dp, float* __restrict__ xc, float* __restrict__ hi,
               float* __restrict__ lo, float* __restrict__ r, float* __restrict__ ssum) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double rr = 0.0, ss = 0.0;
    for (int c = lane; c < dp; c += 32) {
        const float v = c < d ? x[row * d + c] - mean[c] : 0.f;
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        xc[row * dp + c] = v;
        hi[row * dp + c] = h;
        lo[row * dp + c] = v - h;
        rr += (double)v * (double)v;
        ss += (double)v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        rr += __shfl_xor_sync(0xffffffffu, rr, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if (lane == 0) { r[row] = (float)rr; ssum[row] = (float)ss; }
}

__device__ __forceinline__ void hist_add_warp(unsigned* sh, unsigned bin) {
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin != 0xFFFFFFFFu && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&sh[bin], (unsigned)__popc(peers));
}

// radix-select histogram pass over the eps-perturbed distances reconstructed from d2:
//   ||x_i - x_j + eps||^2 = d2 + 2 eps (s_i - s_j) + D eps^2 ,  pairs i < j only, exact zeros dropped.
// Equal bins inside a warp are combined with __match_any_sync before touching shared memory.
__global__ void __launch_bounds__(256)
k_hist_d2(const float* __restrict__ d2, const float* __restrict__ ssum, long long n, long long n4, int d, long long row_begin, long long n_rows,
          int pass, unsigned* __restrict__ hist, const unsigned long long* __restrict__ state) {
    __shared__ unsigned sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const unsigned prefix = pass > 0 ? (unsigned)state[0] : 0u;
    const float eps = 1e-6f, deps2 = (float)d * 1e-12f;
    for (long long t = blockIdx.x; t < n_rows; t += gridDim.x) {
        const long long i = row_begin + t;
        const float si = ssum[i];
        for (long long j0 = (i + 1) / 256 * 256; j0 < n; j0 += 256) {
            const long long j = j0 + threadIdx.x;
            unsigned bin = 0xFFFFFFFFu;
            if (j > i && j < n) {
                const float v = sqrtf(fmaxf(d2[t * n4 + j] + 2.f * eps * (si - ssum[j]) + deps2, 0.f));
                const unsigned b = __float_as_uint(v);
                if (b != 0u) {
                    if (pass == 0) bin = b >> 21;
                    else if (pass == 1) { if ((b >> 21) == (prefix >> 21)) bin = (b >> 10) & 2047u; }
                    else if ((b >> 10) == (prefix >> 10)) bin = b & 1023u;
                }
            }
            hist_add_warp(sh, bin);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2048; b += blockDim.x)
        if (sh[b]) atomicAdd(&hist[pass * 2048 + b], sh[b]);
}

// ---- K3a, single-rank fast path: exact lower median by bracket + count + compact ---------------------------------------
// The radix-select histograms above pay for contention: the n(n-1)/2 distances of a particle cloud cluster in a handful of
// top-bit bins.  With all rows on one GPU the median is found without a contended histogram:
//   k_sel_sample   32768 pseudo-random pairs (i < j), their distances read from d2;
//   k_sel_bracket  one CTA histograms the sample linearly between its min and max and takes the bin edges of the sample
//                  quantiles 0.5 -/+ 0.0155 (5.6 sigma of the sample-rank error) as a bracket [lo, hi] that holds the true
//                  median and ~3.3 % of all distances;
//   k_sel_count    one pass over the upper triangle of d2: counts exact zeros and values below lo in registers, appends the
//                  values inside the bracket to a candidate list (warp-private shared-memory staging, one global atomic per
//                  ~200 candidates) and counts them in a 4096-bin LINEAR histogram over the bracket (uniform occupancy, so
//                  the global atomics do not collide);
//   k_sel_gather   every CTA locates the histogram bin that holds the wanted rank and the candidates of that bin are
//                  appended to a second, short list; the LAST CTA to finish radix-selects (11 + 11 + 10 bits) inside the
//                  short list and emits h = med^2 / ln n.  If the bracket missed (never observed; probability ~1e-8 per
//                  call) or a list overflowed, that CTA falls back to the 3-pass radix select over d2.  Exact either way.
struct SelHdr {                 // 64 bytes in the workspace
    float lo, hi;
    unsigned ncand, overflow;
    unsigned long long zeros, below;
    unsigned n2, bin;           // short list length, hit bin
    unsigned long long resid;   // rank inside the hit bin
    unsigned long long ncand_tot;   // candidates over all ranks (multi-rank flow; = ncand on one rank)
    unsigned ok, done;              // multi-rank flow: bracket hit (set by k_sel_gather); single rank: CTAs of k_sel_gather that have finished
};
// The whole select works on w = d^2 (the eps-perturbed SQUARED distance): sqrt is monotone, so the k-th smallest distance is
// the root of the k-th smallest w, exact zeros are the same pairs, and no pair needs a square root - only the answer does.

__device__ __forceinline__ float pair_w(float d2v, float si, float sj, int d) {        // ||x_i - x_j + eps||^2 from the centred form
    const float eps = 1e-6f, deps2 = (float)d * 1e-12f;
    return fmaxf(d2v + 2.f * eps * (si - sj) + deps2, 0.f);
}
__device__ __forceinline__ float pair_dist(const float* __restrict__ d2, const float* __restrict__ ssum, long long n4, int d, long long i, long long j) {
    return pair_w(d2[i * n4 + j], ssum[i], ssum[j], d);
}
// THE exact squared distance of the flash path (D <= 64), one definition for every evaluator (bracket sample, counting kernel,
// fallback select) so that they agree bit for bit: per 32-float block, eight partial sums - one per 16-byte chunk, e0^2 then
// three sequential fmas - combined by the xor-butterfly tree ((p0+p1)+(p2+p3))+((p4+p5)+(p6+p7)); blocks are added in order.
// The tree is what eight cooperating lanes compute with three shuffles (fp32 addition is commutative, so every lane of the
// butterfly holds the same bits), which lets the counting kernel read its operand rows conflict-free.
__device__ __forceinline__ float d2_chunk(const float4 a, const float4 b) {
    float e = a.x - b.x;
    float p = __fmul_rn(e, e);
    e = a.y - b.y; p = fmaf(e, e, p);
    e = a.z - b.z; p = fmaf(e, e, p);
    e = a.w - b.w; p = fmaf(e, e, p);
    return p;
}
__device__ __forceinline__ float d2_block32(const float4* __restrict__ a4, const float4* __restrict__ b4) {
    float p[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) p[c] = d2_chunk(__ldg(a4 + c), __ldg(b4 + c));
    return __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), __fadd_rn(p[2], p[3])), __fadd_rn(__fadd_rn(p[4], p[5]), __fadd_rn(p[6], p[7])));
}
__device__ __forceinline__ float d2_exact_rows(const float* __restrict__ xc, int dp, long long i, long long j) {
    const float4* a4 = reinterpret_cast<const float4*>(xc + i * dp);
    const float4* b4 = reinterpret_cast<const float4*>(xc + j * dp);
    float acc = d2_block32(a4, b4);
    for (int kb = 1; kb < dp / 32; ++kb) acc = __fadd_rn(acc, d2_block32(a4 + 8 * kb, b4 + 8 * kb));
    return acc;
}

// distance source of the select's fallback: the materialised matrix (D > 64) or, on the flash path, the centred particles
// themselves (exact fp32 differences, the arithmetic of exact_d2 in svgd_flash.cuh)
struct DistSrc {
    const float* d2;       // rows [row_off, ...) x n4, or NULL
    const float* xc;       // n x dp (used when d2 == NULL)
    long long n4, row_off;
    int dp;
    __device__ __forceinline__ float operator()(const float* __restrict__ ssum, int d, long long i, long long j) const {
        if (d2) return pair_w(d2[(i - row_off) * n4 + j], ssum[i], ssum[j], d);
        const float acc = d2_exact_rows(xc, dp, i, j);
        return pair_w(acc, ssum[i], ssum[j], d);
    }
};
__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

constexpr int kSelSamples = kSelSamplesHost, kSelBins = 4096;

// block-wide: index of the bin holding 0-based rank `rank` in bins[0..NT*BPT) (shared or global), residual rank out.
// NT threads, BPT consecutive bins per thread.  Contains __syncthreads.
template <int NT, int BPT>
__device__ __forceinline__ int block_find_bin(const unsigned* bins, unsigned long long rank, unsigned long long* resid, unsigned* wsum, int* s_bin,
                                              unsigned long long* s_res) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned loc[BPT], mine = 0;
#pragma unroll
    for (int k = 0; k < BPT; ++k) { loc[k] = bins[t * BPT + k]; mine += loc[k]; }
    unsigned incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += y;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    if (t == 0) { *s_bin = NT * BPT - 1; *s_res = 0; }
    __syncthreads();
    unsigned long long cum = 0;
    for (int w = 0; w < warp; ++w) cum += wsum[w];
    cum += incl - mine;
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        if (loc[k] && rank >= cum && rank < cum + loc[k]) { *s_bin = t * BPT + k; *s_res = rank - cum; }
        cum += loc[k];
    }
    __syncthreads();
    *resid = *s_res;
    return *s_bin;
}

__global__ void __launch_bounds__(1024)
k_sel_bracket(const float* __restrict__ samples, int ns, SelHdr* __restrict__ hdr, unsigned* __restrict__ ghist, int sabotage) {
    __shared__ unsigned sh[kSelBins];
    __shared__ float red_min[32], red_max[32];
    __shared__ unsigned red_cnt[32];
    __shared__ unsigned wsum[32];
    __shared__ int s_bin;
    __shared__ unsigned long long s_res;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < kSelBins; i += 1024) { sh[i] = 0u; ghist[i] = 0u; }
    float mn = INFINITY, mx = 0.f;
    unsigned cnt = 0;
    for (int k = t; k < ns; k += 1024) {
        const float x = samples[k];
        if (x != 0.f) { mn = fminf(mn, x); mx = fmaxf(mx, x); ++cnt; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if (lane == 0) { red_min[warp] = mn; red_max[warp] = mx; red_cnt[warp] = cnt; }
    __syncthreads();
    mn = INFINITY; mx = 0.f; cnt = 0;
    for (int w = 0; w < 32; ++w) { mn = fminf(mn, red_min[w]); mx = fmaxf(mx, red_max[w]); cnt += red_cnt[w]; }
    const float scale = mx > mn ? (float)kSelBins / (mx - mn) : 0.f;
    for (int k = t; k < ns; k += 1024) {
        const float x = samples[k];
        if (x != 0.f) atomicAdd(&sh[min(kSelBins - 1, (int)((x - mn) * scale))], 1u);
    }
    __syncthreads();
    // half-width of the quantile bracket: 5.6 sigma of the sample-rank error (0.0155 at 32768 samples)
    const float delta = 2.8f * rsqrtf((float)ns), width = scale > 0.f ? 1.f / scale : 0.f;
    const unsigned r_lo = (unsigned)fmaxf(0.f, floorf((0.5f - delta) * (float)cnt)), r_hi = min(cnt ? cnt - 1 : 0u, (unsigned)ceilf((0.5f + delta) * (float)cnt));
    unsigned long long dummy;
    const int b_lo = block_find_bin<1024, 4>(sh, r_lo, &dummy, wsum, &s_bin, &s_res);
    const int b_hi = block_find_bin<1024, 4>(sh, r_hi, &dummy, wsum, &s_bin, &s_res);
    if (t == 0) {
        float lo = fmaxf(mn + ((float)b_lo - 1.f) * width, 0.f), hi = mn + ((float)b_hi + 2.f) * width;      // one bin of margin each side
        if (cnt == 0) { lo = 0.f; hi = 0.f; }
        if (sabotage) { lo = mx * 4.f + 1.f; hi = mx * 8.f + 2.f; }
        hdr->lo = lo; hdr->hi = hi;
        hdr->ncand = 0u; hdr->overflow = 0u; hdr->zeros = 0ull; hdr->below = 0ull; hdr->n2 = 0u; hdr->bin = 0u; hdr->resid = 0ull; hdr->done = 0u;
    }
}

__device__ __forceinline__ int sel_lbin(float v, float lo, float scale) { return min(kSelBins - 1, max(0, (int)((v - lo) * scale))); }

__global__ void __launch_bounds__(256)
k_sel_count(const float* __restrict__ d2, const float* __restrict__ ssum, long long n, long long n4, int d, long long row_begin, long long n_rows,
            SelHdr* __restrict__ hdr, float* __restrict__ cand, long long cand_cap, unsigned* __restrict__ ghist) {
    __shared__ float stage[8][256];
    __shared__ unsigned long long red[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float lo = hdr->lo, hi = hdr->hi;
    const float scale = hi > lo ? (float)kSelBins / (hi - lo) : 0.f;
    unsigned zeros = 0, below = 0, staged = 0;      // staged: warp-uniform
    auto flush = [&]() {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&hdr->ncand, staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((long long)base + staged > cand_cap) { if (lane == 0) hdr->overflow = 1u; }
        else for (unsigned k = lane; k < staged; k += 32) cand[base + k] = stage[warp][k];
        __syncwarp();
        staged = 0;
    };
    for (long long t = blockIdx.x; t < n_rows; t += gridDim.x) {          // d2 holds rows [row_begin, row_begin + n_rows)
        const long long i = row_begin + t;
        const float si = ssum[i];
        for (long long j0 = (i + 1) / 256 * 256; j0 < n; j0 += 256) {
            const long long j = j0 + threadIdx.x;
            bool is_c = false;
            float v = 0.f;
            if (j > i && j < n) {
                v = pair_w(d2[t * n4 + j], si, ssum[j], d);
                if (v == 0.f) ++zeros;
                else if (v < lo) ++below;
                else is_c = v <= hi;
            }
            const unsigned m = __ballot_sync(0xffffffffu, is_c);
            if (m) {
                if (is_c) {
                    stage[warp][staged + __popc(m & ((1u << lane) - 1u))] = v;
                    atomicAdd(&ghist[sel_lbin(v, lo, scale)], 1u);
                }
                staged += __popc(m);
                __syncwarp();
                if (staged > 256 - 32) flush();
            }
        }
    }
    if (staged) flush();
    unsigned long long z = zeros, b = below;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        z += __shfl_xor_sync(0xffffffffu, z, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if (lane == 0) { red[0][warp] = z; red[1][warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        z = 0; b = 0;
        for (int w = 0; w < 8; ++w) { z += red[0][w]; b += red[1][w]; }
        if (z) atomicAdd(&hdr->zeros, z);
        if (b) atomicAdd(&hdr->below, b);
    }
}

__device__ __forceinline__ bool sel_ok(const SelHdr* hdr, long long n, unsigned long long* rank_out, unsigned long long ncand_tot) {
    const unsigned long long pairs = (unsigned long long)n * (unsigned long long)(n - 1) / 2ull;
    const unsigned long long nz = pairs - hdr->zeros;
    const unsigned long long rank = nz ? (nz - 1) / 2 : 0;                      // LOWER median (torch.median)
    *rank_out = rank;
    return !hdr->overflow && nz > 0 && rank >= hdr->below && rank < hdr->below + ncand_tot;
}

// last stage of the select, run by ONE CTA of NT threads: radix select (11 + 11 + 10 bits) inside the short list of the hit
// bin - or, if the bracket missed, over the whole upper triangle of d2 - and h = med^2 / ln n.  Fields that other CTAs of the
// same launch wrote (n2, resid) are read past the L1 (__ldcg).
template <int NT>
__device__ __forceinline__ void sel_final_body(const DistSrc dist, const float* __restrict__ ssum, long long n, int d,
                                               const SelHdr* hdr, const float* list2, long long cap2, float* __restrict__ out_h,
                                               unsigned long long* __restrict__ dbg, unsigned* sh, unsigned* wsum, int* s_bin, unsigned long long* s_res) {
    const int t = threadIdx.x;
    unsigned long long rank;
    const unsigned n2 = __ldcg(&hdr->n2);
    const bool ok = sel_ok(hdr, n, &rank, hdr->ncand) && (long long)n2 <= cap2;
    const unsigned long long pairs = (unsigned long long)n * (unsigned long long)(n - 1) / 2ull;
    float med = nanf("");                               // torch: median of an empty tensor
    if (pairs != hdr->zeros) {
        unsigned prefix = 0;
        unsigned long long r = ok ? __ldcg(&hdr->resid) : rank;
        if (ok && n2 <= 512u) {
            // short list (the usual case: ~n^2 / 270 000 candidates in the hit bin): the value of rank r by direct counting - one pass
            // instead of three radix passes.  Squared distances are positive floats: their order is the order of the bit patterns.
            for (unsigned i = t; i < n2; i += NT) sh[i] = __float_as_uint(__ldcg(list2 + i));
            if (t == 0) *s_res = 0ull;
            __syncthreads();
            for (unsigned i = t; i < n2; i += NT) {
                const unsigned v = sh[i];
                unsigned lt = 0, le = 0;
                for (unsigned j = 0; j < n2; ++j) {
                    const unsigned x = sh[j];
                    lt += x < v ? 1u : 0u;
                    le += x <= v ? 1u : 0u;
                }
                if ((unsigned long long)lt <= r && r < (unsigned long long)le) *s_res = (unsigned long long)v;     // (ties write the same value)
            }
            __syncthreads();
            prefix = (unsigned)*s_res;
        } else
        for (int pass = 0; pass < 3; ++pass) {
            __syncthreads();
            for (int i = t; i < 2048; i += NT) sh[i] = 0u;
            __syncthreads();
            auto add = [&](unsigned bits) {
                if (pass == 0) atomicAdd(&sh[bits >> 21], 1u);
                else if (pass == 1) { if ((bits >> 21) == (prefix >> 21)) atomicAdd(&sh[(bits >> 10) & 2047u], 1u); }
                else if ((bits >> 10) == (prefix >> 10)) atomicAdd(&sh[bits & 1023u], 1u);
            };
            if (ok) {
                for (unsigned i = t; i < n2; i += NT) add(__float_as_uint(__ldcg(list2 + i)));
            } else {
                // fallback: radix select over the upper triangle of d2 (one CTA; never taken with a sane bracket; 32-bit bin
                // counts, i.e. n <= 92681)
                for (long long i = 0; i < n; ++i)
                    for (long long j = i + 1 + t; j < n; j += NT) {
                        const unsigned bits = __float_as_uint(dist(ssum, d, i, j));
                        if (bits != 0u) add(bits);
                    }
            }
            __syncthreads();
            const int b = block_find_bin<NT, 2048 / NT>(sh, r, &r, wsum, s_bin, s_res);
            prefix |= pass == 0 ? ((unsigned)b << 21) : pass == 1 ? ((unsigned)b << 10) : (unsigned)b;
        }
        med = sqrtf(__uint_as_float(prefix));           // the select ran on squared distances
    }
    if (t == 0) {
        if (out_h) *out_h = __fdiv_rn(__fmul_rn(med, med), (float)log((double)n));
        if (dbg) { dbg[0] = ok ? 1ull : 0ull; dbg[1] = hdr->ncand; dbg[2] = hdr->below; dbg[3] = hdr->zeros; }
    }
}

// Single rank (multi = 0): gather + final in one launch - the last CTA to finish (ticket in hdr->done) runs the final select.
__global__ void __launch_bounds__(256)
k_sel_gather(long long n, SelHdr* hdr, const unsigned* __restrict__ ghist, const float* __restrict__ cand, float* list2, long long cap2, int multi,
             const DistSrc dist, const float* __restrict__ ssum, int d, float* __restrict__ out_h,
             unsigned long long* __restrict__ dbg) {
    __shared__ unsigned sh[2048];
    __shared__ unsigned wsum[8];
    __shared__ int s_bin;
    __shared__ unsigned long long s_res;
    __shared__ bool s_last;
    unsigned long long rank, resid;
    const bool ok = sel_ok(hdr, n, &rank, multi ? hdr->ncand_tot : hdr->ncand);
    if (multi && blockIdx.x == 0 && threadIdx.x == 0) hdr->ok = ok ? 1u : 0u;
    if (ok) {                                           // uniform over the grid; otherwise the final stage takes the fallback
        const int b1 = block_find_bin<256, 16>(ghist, rank - hdr->below, &resid, wsum, &s_bin, &s_res);
        if (blockIdx.x == 0 && threadIdx.x == 0) { hdr->bin = (unsigned)b1; hdr->resid = resid; }
        const float lo = hdr->lo, hi = hdr->hi;
        const float scale = hi > lo ? (float)kSelBins / (hi - lo) : 0.f;
        const unsigned ncand = hdr->ncand;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < ncand; i += gridDim.x * blockDim.x) {
            const float v = cand[i];
            if (sel_lbin(v, lo, scale) == b1) {
                const unsigned pos = atomicAdd(&hdr->n2, 1u);
                if ((long long)pos < cap2) list2[pos] = v;
            }
        }
    }
    if (multi) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&hdr->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    sel_final_body<256>(dist, ssum, n, d, hdr, list2, cap2, out_h, dbg, sh, wsum, &s_bin, &s_res);
}


// ---- bracket select across ranks ---------------------------------------------------------------------------------------
// Every rank holds all particles (all-gather) and the distances of ITS rows.  The bracket is sampled from X, so every rank
// derives the same [lo, hi] with no communication; each rank counts / compacts the pairs of its rows (k_sel_count); ONE
// all-reduce sums the counters and the linear histogram (k_sel_pack / k_sel_unpack around it); every rank finds the same
// hit bin and gathers its own candidates of that bin (k_sel_gather); the final radix select runs over those short lists with
// the 2048-bin histogram all-reduced between the three passes (k_hist_sel / k_select_update_sel) - or, if the bracket
// missed, over the ranks' distance rows like the legacy select.  All decisions are taken on the device.
__global__ void __launch_bounds__(256)
k_sel_pack(const SelHdr* __restrict__ hdr, const unsigned* __restrict__ ghist, unsigned long long* __restrict__ red64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2048) red64[4 + i] = (unsigned long long)ghist[2 * i] | ((unsigned long long)ghist[2 * i + 1] << 32);
    if (i == 0) { red64[0] = hdr->zeros; red64[1] = hdr->below; red64[2] = hdr->ncand; red64[3] = hdr->overflow; }
}
__global__ void __launch_bounds__(256)
k_sel_unpack(SelHdr* __restrict__ hdr, unsigned* __restrict__ ghist, const unsigned long long* __restrict__ red64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2048) { ghist[2 * i] = (unsigned)red64[4 + i]; ghist[2 * i + 1] = (unsigned)(red64[4 + i] >> 32); }
    if (i == 0) { hdr->zeros = red64[0]; hdr->below = red64[1]; hdr->ncand_tot = red64[2]; hdr->overflow = red64[3] ? 1u : 0u; }
}
// radix-select histogram pass of the final stage: over this rank's short list (bracket hit) or over its distance rows (miss)
__global__ void __launch_bounds__(256)
k_hist_sel(const DistSrc dist, const float* __restrict__ ssum, long long n, int d, long long row_start, long long row_step, long long n_rows,
           const SelHdr* __restrict__ hdr, const float* __restrict__ list2, long long cap2, int pass, unsigned* __restrict__ hist,
           const unsigned long long* __restrict__ state) {
    __shared__ unsigned sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const unsigned prefix = pass > 0 ? (unsigned)state[0] : 0u;
    auto add = [&](unsigned bits) {
        unsigned bin = 0xFFFFFFFFu;
        if (bits != 0u) {
            if (pass == 0) bin = bits >> 21;
            else if (pass == 1) { if ((bits >> 21) == (prefix >> 21)) bin = (bits >> 10) & 2047u; }
            else if ((bits >> 10) == (prefix >> 10)) bin = bits & 1023u;
        }
        hist_add_warp(sh, bin);
    };
    if (hdr->ok) {
        const long long n2 = min((long long)hdr->n2, cap2);
        for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < n2; i0 += (long long)gridDim.x * blockDim.x) {
            const long long i = i0 + threadIdx.x;
            add(i < n2 ? __float_as_uint(list2[i]) : 0u);
        }
    } else {
        for (long long t = blockIdx.x; t < n_rows; t += gridDim.x) {       // bracket missed: this rank's rows, every pair i < j
            const long long i = row_start + t * row_step;
            for (long long j0 = (i + 1) / 256 * 256; j0 < n; j0 += 256) {
                const long long j = j0 + threadIdx.x;
                add((j > i && j < n) ? __float_as_uint(dist(ssum, d, i, j)) : 0u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2048; b += blockDim.x)
        if (sh[b]) atomicAdd(&hist[pass * 2048 + b], sh[b]);
}
// after the all-reduce of pass p: locate the bin of the wanted rank; after pass 2 emit h (one CTA of 1024 threads)
__global__ void __launch_bounds__(1024)
k_select_update_sel(int pass, const unsigned* __restrict__ hist, unsigned long long* __restrict__ state, long long n, const SelHdr* __restrict__ hdr,
                    float* __restrict__ out_h) {
    __shared__ unsigned wsum[32];
    __shared__ int s_bin;
    __shared__ unsigned long long s_res;
    const unsigned long long pairs = (unsigned long long)n * (unsigned long long)(n - 1) / 2ull;
    const unsigned long long nz = pairs - hdr->zeros;
    unsigned long long rank = pass == 0 ? (hdr->ok ? hdr->resid : (nz ? (nz - 1) / 2 : 0)) : state[1];
    __syncthreads();
    unsigned long long resid;
    const int b = block_find_bin<1024, 2>(hist + pass * 2048, rank, &resid, wsum, &s_bin, &s_res);
    if (threadIdx.x == 0) {
        unsigned prefix = pass == 0 ? 0u : (unsigned)state[0];
        prefix |= pass == 0 ? ((unsigned)b << 21) : pass == 1 ? ((unsigned)b << 10) : (unsigned)b;
        state[0] = prefix;
        state[1] = resid;
        if (pass == 2 && out_h) {
            float med = sqrtf(__uint_as_float(prefix));                    // the select ran on squared distances
            if (nz == 0) med = nanf("");
            *out_h = __fdiv_rn(__fmul_rn(med, med), (float)log((double)n));
        }
    }
}

// ---- pair sample of the bracket select, drawn straight from the centred particles ---------------------------------------
template <int DP>
__global__ void __launch_bounds__(256)
k_sel_sample_x(const float* __restrict__ xc, const float* __restrict__ ssum, long long n, int d, float* __restrict__ samples, int ns) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ns) return;
    const unsigned a = mix32((unsigned)idx * 2u + 0x9e3779b9u), b = mix32(a ^ 0x85ebca6bu);
    long long i = (long long)(a % (unsigned)n), j = (long long)(b % (unsigned)n);
    if (i == j) j = (j + 1) % n;
    if (i > j) { const long long tmp = i; i = j; j = tmp; }
    const float acc = d2_exact_rows(xc, DP, i, j);
    samples[idx] = pair_w(acc, ssum[i], ssum[j], d);
}

// pair sample for D > 64 (one warp per pair; the distance is only a bracket estimate)
__global__ void __launch_bounds__(256)
k_sel_sample_wide(const float* __restrict__ xc, const float* __restrict__ ssum, long long n, int d, int dp, float* __restrict__ samples, int ns) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (idx >= ns) return;
    const unsigned a = mix32((unsigned)idx * 2u + 0x9e3779b9u), b = mix32(a ^ 0x85ebca6bu);
    long long i = (long long)(a % (unsigned)n), j = (long long)(b % (unsigned)n);
    if (i == j) j = (j + 1) % n;
    if (i > j) { const long long tmp = i; i = j; j = tmp; }
    float acc = 0.f;
    for (int c = lane; c < dp; c += 32) {
        const float e = xc[i * dp + c] - xc[j * dp + c];
        acc = fmaf(e, e, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) samples[idx] = pair_w(acc, ssum[i], ssum[j], d);
}

// K = exp(-d2/h) split into tf32 hi/lo, row sums; one block per row
__global__ void __launch_bounds__(256)
k_kmat_split(const float* __restrict__ d2, const float* __restrict__ hptr, long long n, long long n4, long long n_rows, float* __restrict__ k_hi,
             float* __restrict__ k_lo, float* __restrict__ rowsum) {
    const float nih = -1.f / *hptr;
    for (long long t = blockIdx.x; t < n_rows; t += gridDim.x) {
        float rs = 0.f;
        for (long long j = threadIdx.x * 4; j < n4; j += 256 * 4) {
            const float4 v = *reinterpret_cast<const float4*>(d2 + t * n4 + j);
            float k[4] = {expf(v.x * nih), expf(v.y * nih), expf(v.z * nih), expf(v.w * nih)};
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (j + e >= n) k[e] = 0.f;
                h[e] = __uint_as_float(__float_as_uint(k[e]) & 0xFFFFE000u);
                l[e] = k[e] - h[e];
                rs += k[e];
            }
            *reinterpret_cast<float4*>(k_hi + t * n4 + j) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(k_lo + t * n4 + j) = make_float4(l[0], l[1], l[2], l[3]);
        }
        __shared__ float red[8];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = rs;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += red[w];
            rowsum[t] = s;
        }
    }
}

// Vt[c][j] : c < dp -> g[j][c] ; c >= dp -> xc[j][c - dp] ; split hi/lo, zero padded.  32x32 shared-memory transpose.
__global__ void __launch_bounds__(256)
k_vt_split(const float* __restrict__ g, const float* __restrict__ xc_hi, const float* __restrict__ xc_lo, long long n, long long n4, int d, int dp,
           float* __restrict__ vt_hi, float* __restrict__ vt_lo, const float* __restrict__ r, const float* __restrict__ hptr, float* __restrict__ ncr,
           long long ncr_len) {
    __shared__ float tile[32][33];
    const long long j0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 8 rows per pass
    if (ncr && blockIdx.y == 0 && ty == 0) {
        // flash phi: exponent offset of column j, -r_j log2(e) / h; -inf beyond n switches the padding columns off (K = 0)
        const float c = 1.4426950408889634f / __ldg(hptr);
        const long long end = blockIdx.x == gridDim.x - 1 ? ncr_len : j0 + 32;
        for (long long j = j0 + tx; j < end; j += 32) ncr[j] = j < n ? -c * r[j] : -INFINITY;
    }
    for (int rr = ty; rr < 32; rr += 8) {
        const long long j = j0 + rr;
        const int c = c0 + tx;
        float v = 0.f;
        if (j < n) {
            if (c < dp) v = c < d ? g[j * d + c] : 0.f;
            else v = xc_hi[j * dp + (c - dp)] + xc_lo[j * dp + (c - dp)];
        }
        tile[rr][tx] = v;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        const int c = c0 + rr;
        const long long j = j0 + tx;
        if (j < n4) {
            const float v = tile[tx][rr];
            const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            vt_hi[(long long)c * n4 + j] = h;
            vt_lo[(long long)c * n4 + j] = v - h;
        }
    }
}

// phi[t][c] = (O_G - (2/h)(O_X - xc_i[c] rowsum_t)) / n, summing the K splits
__global__ void __launch_bounds__(256)
k_phi_finalize(const float* __restrict__ o, int splits, const float* __restrict__ rowsum, int rs_splits, const float* __restrict__ xc_hi,
               const float* __restrict__ xc_lo, const float* __restrict__ hptr, long long n, int d, int dp, long long row_begin, long long n_rows,
               float* __restrict__ phi, float* __restrict__ ad_x, float* __restrict__ ad_acc, float ad_lr) {
    const float c2 = 2.f / *hptr, invn = 1.f / (float)n;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = n_rows * d, stride = (long long)gridDim.x * blockDim.x;
    for (; idx < tot; idx += stride) {
        const long long t = idx / d;
        const int c = (int)(idx - t * d);
        float og = 0.f, ox = 0.f;
        for (int sp = 0; sp < splits; ++sp) {
            const float* base = o + ((long long)sp * n_rows + t) * 2 * dp;
            og += base[c];
            ox += base[dp + c];
        }
        const long long i = row_begin + t;
        const float xi = xc_hi[i * dp + c] + xc_lo[i * dp + c];
        float rsum = 0.f;                       // rs_splits partial row sums (fused path) or one finished sum
        for (int sp = 0; sp < rs_splits; ++sp) rsum += rowsum[(long long)sp * n_rows + t];
        const float v = (og - c2 * (ox - xi * rsum)) * invn;
        phi[idx] = v;
        if (ad_x) {
            // K3c fused (same arithmetic as k_adagrad with grad = -phi, inference.py:280): one launch less on the epoch's chain
            const float gr = -v;
            const float a = __fadd_rn(ad_acc[idx], __fmul_rn(gr, gr));
            ad_acc[idx] = a;
            const float sd = __fadd_rn(sqrtf(a), 1e-10f);
            ad_x[idx] = __fadd_rn(ad_x[idx], __fmul_rn(-ad_lr, __fdiv_rn(gr, sd)));
        }
    }
}

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace tc
}  // namespace ocbnn
#include "svgd_flash.cuh"
namespace ocbnn {
namespace tc {

// pairs sampled for the bracket of the median select: 8192 up to n = 4096, then proportional to the pair count, capped at 2^17.
// The bracket holds 5.6 sigma of the sample-rank error = 2.8 / sqrt(samples) of all pairs on each side (6 % of the pairs at 8192
// samples, 1.5 % at 2^17); every pair inside it is re-evaluated exactly out of shared memory by the counting kernel, which is
// cheap - the one-CTA bracket kernel over a large sample was not (165 us for 2^20 samples at n = 32768).
static int sel_samples(long long n) {
    long long s = n * n / 2048;
    if (s < 8192) s = 8192;
    if (s > kSelSamplesMaxHost) s = kSelSamplesMaxHost;
    return (int)(s / 1024 * 1024);
}

static DistSrc dist_src(const TcWs& w) { return DistSrc{w.d2, w.xc, w.n4, w.row_begin, (int)w.dp}; }

// centre + split X (every rank holds all particles); D > 64: also the squared distances of rows [row_begin, row_begin + n_rows)
// by the Gram GEMM on the tensor cores.  D <= 64: nothing else - the select and phi recompute their tiles (svgd_flash.cuh).
static int tc_prepare(const TcWs& w, const float* x, cudaStream_t st) {
    // centre = column means of the first min(n, 1024) particles: phi is translation invariant and the centre only has to sit inside
    // the cloud (it removes the cancellation in r_i + r_j - 2 S); a full-column mean would be a 24 us latency chain at n = 32768
    k_colmean<<<(unsigned)w.dp, 256, 0, st>>>(x, w.n < 1024 ? w.n : 1024, (int)w.d, (int)w.dp, w.mean);
    OCBNN_LAUNCH_CHECK();
    k_center_split<<<(unsigned)((w.n * 32 + 255) / 256), 256, 0, st>>>(x, w.mean, w.n, (int)w.d, (int)w.dp, w.xc, w.xc_hi, w.xc_lo, w.r, w.s);
    OCBNN_LAUNCH_CHECK();
    if (w.flash || w.n_rows == 0) return OCBNN_OK;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    int rc;
    if ((rc = make_tmap(&ta_hi, w.xc_hi + w.row_begin * w.dp, w.n_rows, w.dp, kBM))) return rc;
    if ((rc = make_tmap(&ta_lo, w.xc_lo + w.row_begin * w.dp, w.n_rows, w.dp, kBM))) return rc;
    if ((rc = make_tmap(&tb_hi, w.xc_hi, w.n, w.dp, 128))) return rc;
    if ((rc = make_tmap(&tb_lo, w.xc_lo, w.n, w.dp, 128))) return rc;
    EpiGramRows epi{w.d2, w.r, w.n, w.n4, w.row_begin, w.n_rows};
    return launch_gemm<128>(ta_hi, ta_lo, tb_hi, tb_lo, epi, w.n_rows, w.n, w.dp, 1, st);
}

// bracket of the median select, sampled from the centred particles (identical on every rank: no communication)
static int sel_bracket(const TcWs& w, int mode, cudaStream_t st) {
    const int ns = sel_samples(w.n);
    SelHdr* hdr = reinterpret_cast<SelHdr*>(w.sel);
    if (w.dp == 32) k_sel_sample_x<32><<<ns / 256, 256, 0, st>>>(w.xc, w.s, w.n, (int)w.d, w.samples, ns);
    else if (w.dp == 64) k_sel_sample_x<64><<<ns / 256, 256, 0, st>>>(w.xc, w.s, w.n, (int)w.d, w.samples, ns);
    else k_sel_sample_wide<<<ns / 8, 256, 0, st>>>(w.xc, w.s, w.n, (int)w.d, (int)w.dp, w.samples, ns);
    OCBNN_LAUNCH_CHECK();
    k_sel_bracket<<<1, 1024, 0, st>>>(w.samples, ns, hdr, w.ghist, mode == 1 ? 0 : 1);      // modes 0 / 2: sabotaged -> full select
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

static int count_dbg() {        // OCBNN_COUNT_DBG: timing probes of k_count_tc (1 no classification, 2 no exact re-evaluation, 4 no MMAs)
    static const int v = [] { const char* e = getenv("OCBNN_COUNT_DBG"); return e ? atoi(e) : 0; }();
    return v;
}

// counting pass of this rank's share of the pairs.  Flash path: the 128 x 128 tiles on or above the diagonal, dealt out
// cyclically (tile index % world == rank); D > 64: the rows [row_begin, row_begin + n_rows) of the materialised distances.
static int sel_count(const TcWs& w, int rank, int world, cudaStream_t st) {
    SelHdr* hdr = reinterpret_cast<SelHdr*>(w.sel);
    if (!w.flash) {
        if (w.n_rows > 0) {
            k_sel_count<<<(unsigned)(w.n_rows < 148 * 8 ? w.n_rows : 148 * 8), 256, 0, st>>>(w.d2, w.s, w.n, w.n4, (int)w.d, w.row_begin, w.n_rows, hdr,
                                                                                         w.cand, w.cand_cap, w.ghist);
            OCBNN_LAUNCH_CHECK();
        }
        return OCBNN_OK;
    }
    const int T = (int)((w.n + 127) / 128);
    const long long tiles = (long long)T * (T + 1) / 2;
    const long long mine = tiles > rank ? (tiles - rank + world - 1) / world : 0;
    if (mine == 0) return OCBNN_OK;
    CUtensorMap tm_hi, tm_lo;
    int rc;
    if ((rc = make_tmap(&tm_hi, w.xc_hi, w.n, w.dp, 128))) return rc;
    if ((rc = make_tmap(&tm_lo, w.xc_lo, w.n, w.dp, 128))) return rc;
    if (w.dp == 32) {
        auto kern = k_count_tc<32>;
        OCBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CountCfg<32>::kSmemBytes));
        const long long grid = mine < 148 * CountCfg<32>::kCtasPerSm ? mine : 148 * CountCfg<32>::kCtasPerSm;
        kern<<<(unsigned)grid, kCntThreads, CountCfg<32>::kSmemBytes, st>>>(tm_hi, tm_lo, w.r, w.s, w.n, (int)w.d, T, rank, world, mine, hdr, w.cand,
                                                                          w.cand_cap, w.ghist, count_dbg());
    } else {
        auto kern = k_count_tc<64>;
        OCBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CountCfg<64>::kSmemBytes));
        const long long grid = mine < 148 * CountCfg<64>::kCtasPerSm ? mine : 148 * CountCfg<64>::kCtasPerSm;
        kern<<<(unsigned)grid, kCntThreads, CountCfg<64>::kSmemBytes, st>>>(tm_hi, tm_lo, w.r, w.s, w.n, (int)w.d, T, rank, world, mine, hdr, w.cand,
                                                                          w.cand_cap, w.ghist, count_dbg());
    }
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

// experiment switch (OCBNN_FLASH_DBG, bit mask; 0 = the real kernel): 1 SIMT warps only hand the buffers on, 2 no O MMAs, 4 no S
// MMAs, 8 no TMEM load, 16 no TMEM store - timing probes for the phase costs, results are garbage
static int flash_dbg() {
    static const int v = [] { const char* e = getenv("OCBNN_FLASH_DBG"); return e ? atoi(e) : 0; }();
    return v;
}

template <int DP, int TILES>
static int launch_flash(const TcWs& w, const float* h, cudaStream_t st) {
    using Cfg = FlashCfg<DP, TILES>;
    CUtensorMap xj_hi, xj_lo, vt_hi, vt_lo;
    int rc;
    if ((rc = make_tmap(&xj_hi, w.xc_hi, w.n, w.dp, kFlashBJ))) return rc;
    if ((rc = make_tmap(&xj_lo, w.xc_lo, w.n, w.dp, kFlashBJ))) return rc;
    if ((rc = make_tmap(&vt_hi, w.vt_hi, 2 * w.dp, w.n4, Cfg::NV))) return rc;
    if ((rc = make_tmap(&vt_lo, w.vt_lo, 2 * w.dp, w.n4, Cfg::NV))) return rc;
    auto kern = k_phi_flash<DP, TILES>;
    OCBNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    dim3 grid((unsigned)((w.n_rows + 128 * TILES - 1) / (128 * TILES)), (unsigned)w.fl_splits);
    kern<<<grid, kFlashThreads, Cfg::kSmemBytes, st>>>(xj_hi, xj_lo, vt_hi, vt_lo, w.xc_hi, w.xc_lo, w.r, w.ncr, h, w.n, w.row_begin, w.n_rows, w.o,
                                                     w.k_hi, w.fl_nj, w.fl_per, flash_dbg());
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

static int tc_phi(const TcWs& w, const float* g, const float* h, float* phi, float* ad_x, float* ad_acc, float ad_lr, cudaStream_t st) {
    dim3 gv((unsigned)((w.n4 + 31) / 32), (unsigned)(2 * w.dp / 32));
    k_vt_split<<<gv, 256, 0, st>>>(g, w.xc_hi, w.xc_lo, w.n, w.n4, (int)w.d, (int)w.dp, w.vt_hi, w.vt_lo, w.r, h, w.ncr, (long long)w.fl_nj * 64);
    OCBNN_LAUNCH_CHECK();
    int rc;
    if (w.flash) {
        if (w.dp == 32) rc = w.fl_tiles == 2 ? launch_flash<32, 2>(w, h, st) : launch_flash<32, 1>(w, h, st);
        else rc = launch_flash<64, 1>(w, h, st);
        if (rc) return rc;
        k_phi_finalize<<<grid_for_ll(w.n_rows * w.d), 256, 0, st>>>(w.o, w.fl_splits, w.k_hi, 2 * w.fl_splits, w.xc_hi, w.xc_lo, h,
                                                                  w.n, (int)w.d, (int)w.dp, w.row_begin, w.n_rows, phi, ad_x, ad_acc, ad_lr);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    }
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if ((rc = make_tmap(&tb_hi, w.vt_hi, 2 * w.dp, w.n4, w.bn))) return rc;
    if ((rc = make_tmap(&tb_lo, w.vt_lo, 2 * w.dp, w.n4, w.bn))) return rc;
    k_kmat_split<<<(unsigned)(w.n_rows < 148 * 8 ? w.n_rows : 148 * 8), 256, 0, st>>>(w.d2, h, w.n, w.n4, w.n_rows, w.k_hi, w.k_lo, w.rowsum);
    OCBNN_LAUNCH_CHECK();
    if ((rc = make_tmap(&ta_hi, w.k_hi, w.n_rows, w.n4, kBM))) return rc;
    if ((rc = make_tmap(&ta_lo, w.k_lo, w.n_rows, w.n4, kBM))) return rc;
    EpiStore epi{w.o, 2 * w.dp, w.n_rows * 2 * w.dp, (int)w.n_rows, (int)(2 * w.dp)};
    rc = w.bn == 64 ? launch_gemm<64>(ta_hi, ta_lo, tb_hi, tb_lo, epi, w.n_rows, 2 * w.dp, w.n4, w.splits, st)
                    : launch_gemm<128>(ta_hi, ta_lo, tb_hi, tb_lo, epi, w.n_rows, 2 * w.dp, w.n4, w.splits, st);
    if (rc) return rc;
    k_phi_finalize<<<grid_for_ll(w.n_rows * w.d), 256, 0, st>>>(w.o, w.splits, w.rowsum, 1, w.xc_hi, w.xc_lo, h, w.n, (int)w.d, (int)w.dp, w.row_begin,
                                                              w.n_rows, phi, ad_x, ad_acc, ad_lr);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}

long long tc_workspace_bytes(long long n, long long d, long long n_rows) { return carve(nullptr, n, d, 0, n_rows).bytes + 512; }

int tc_svgd_prepare(const float* x, long long n, long long d, long long row_begin, long long n_rows, void* ws, long long ws_bytes, cudaStream_t st) {
    OCBNN_REQUIRE(ws && ws_bytes >= tc_workspace_bytes(n, d, n_rows), OCBNN_EINVAL, "SVGD tensor path: workspace too small");
    return tc_prepare(carve(ws, n, d, row_begin, n_rows), x, st);
}
// legacy 3-pass radix select over the materialised distance rows (D > 64 only)
int tc_svgd_hist(long long n, long long d, long long row_begin, long long n_rows, int pass, unsigned* hist, const unsigned long long* state, void* ws,
                 cudaStream_t st) {
    TcWs w = carve(ws, n, d, row_begin, n_rows);
    OCBNN_REQUIRE(!w.flash, OCBNN_EUNSUPPORTED, "tc_svgd_hist: D <= 64 keeps no distance matrix");
    k_hist_d2<<<(unsigned)(n_rows < 148 * 8 ? n_rows : 148 * 8), 256, 0, st>>>(w.d2, w.s, n, w.n4, (int)d, row_begin, n_rows, pass, hist, state);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
bool tc_is_flash(long long n, long long d) { return carve(nullptr, n, d, 0, n).flash; }

// Exact lower median of the perturbed pair distances -> h = med^2 / ln n, all pairs on this rank (after tc_svgd_prepare with
// all rows).  mode 1: bracket select; 0 / 2: the full radix select (one CTA, exact distances recomputed on the flash path) -
// the fallback the bracket select takes in-kernel when the bracket misses.  dbg: optional 4 x u64 (bracket hit, candidates,
// below, zeros).
int tc_svgd_select_local(long long n, long long d, void* ws, float* out_h, int mode, unsigned long long* dbg, cudaStream_t st) {
    TcWs w = carve(ws, n, d, 0, n);
    SelHdr* hdr = reinterpret_cast<SelHdr*>(w.sel);
    int rc;
    if ((rc = sel_bracket(w, mode, st))) return rc;
    if (mode == 1 && (rc = sel_count(w, 0, 1, st))) return rc;
    if (count_dbg()) return OCBNN_OK;          // timing probe: the counters are garbage, do not let the select fall back on them
    k_sel_gather<<<148, 256, 0, st>>>(n, hdr, w.ghist, w.cand, w.list2, w.cand_cap, 0, dist_src(w), w.s, (int)d, out_h, dbg);      // gather + final
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
// ---- the same select across ranks --------------------------------------------------------------------------------------
// stage 1: bracket from X, counting pass over this rank's share of the pairs, counters packed into the reduce buffer
// (*red64_out, 4 + 2048 u64: the caller all-reduces it as int64 sums)
int tc_svgd_sel_stage1(long long n, long long d, long long row_begin, long long n_rows, int rank, int world, void* ws, int mode,
                       unsigned long long** red64_out, cudaStream_t st) {
    TcWs w = carve(ws, n, d, row_begin, n_rows);
    SelHdr* hdr = reinterpret_cast<SelHdr*>(w.sel);
    if (red64_out) *red64_out = w.red64;
    int rc;
    if ((rc = sel_bracket(w, mode, st))) return rc;
    if (mode == 1 && (rc = sel_count(w, rank, world, st))) return rc;
    k_sel_pack<<<8, 256, 0, st>>>(hdr, w.ghist, w.red64);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
// stage 2 (after the all-reduce of the reduce buffer): global counters in, this rank's candidates of the hit bin gathered
int tc_svgd_sel_stage2(long long n, long long d, long long row_begin, long long n_rows, void* ws, cudaStream_t st) {
    TcWs w = carve(ws, n, d, row_begin, n_rows);
    SelHdr* hdr = reinterpret_cast<SelHdr*>(w.sel);
    k_sel_unpack<<<8, 256, 0, st>>>(hdr, w.ghist, w.red64);
    OCBNN_LAUNCH_CHECK();
    k_sel_gather<<<148, 256, 0, st>>>(n, hdr, w.ghist, w.cand, w.list2, w.cand_cap, 1, dist_src(w), w.s, (int)d, nullptr, nullptr);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
__global__ void k_zero_words(unsigned* p, int nwords) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) p[i] = 0u;
}
// stage 3, per radix pass: local histogram over this rank's short list - or, if the bracket missed, over its share of ALL pairs
// (flash: rows rank, rank + world, ...; D > 64: its distance rows).  *hist_out = the 2048 u32 the caller all-reduces.
int tc_svgd_sel_hist(long long n, long long d, long long row_begin, long long n_rows, int rank, int world, int pass, void* ws, unsigned** hist_out,
                     cudaStream_t st) {
    TcWs w = carve(ws, n, d, row_begin, n_rows);
    if (pass == 0) {
        k_zero_words<<<(3 * 2048 + 255) / 256, 256, 0, st>>>(w.hist, 3 * 2048);
        OCBNN_LAUNCH_CHECK();
    }
    const long long rs = w.flash ? rank : row_begin, rstep = w.flash ? world : 1;
    const long long rows = w.flash ? (n > rank ? (n - rank + world - 1) / world : 0) : n_rows;
    k_hist_sel<<<148, 256, 0, st>>>(dist_src(w), w.s, n, (int)d, rs, rstep, rows, reinterpret_cast<const SelHdr*>(w.sel), w.list2, w.cand_cap, pass, w.hist,
                                    w.state);
    OCBNN_LAUNCH_CHECK();
    if (hist_out) *hist_out = w.hist + pass * 2048;
    return OCBNN_OK;
}
// ... and the rank bookkeeping / final h
int tc_svgd_sel_update(long long n, long long d, long long row_begin, long long n_rows, int pass, void* ws, float* out_h, cudaStream_t st) {
    TcWs w = carve(ws, n, d, row_begin, n_rows);
    k_select_update_sel<<<1, 1024, 0, st>>>(pass, w.hist, w.state, n, reinterpret_cast<const SelHdr*>(w.sel), out_h);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
// ad_x / ad_acc (n_rows x d, or NULL): the owned particle rows and their Adagrad accumulator, updated with grad = -phi by the finalize kernel
int tc_svgd_phi(const float* g, const float* h, long long n, long long d, long long row_begin, long long n_rows, float* phi, void* ws, cudaStream_t st,
                float* ad_x, float* ad_acc, float ad_lr) {
    return tc_phi(carve(ws, n, d, row_begin, n_rows), g, h, phi, ad_x, ad_acc, ad_lr, st);
}

}  // namespace tc
}  // namespace ocbnn

using namespace ocbnn;
using namespace ocbnn::tc;

static unsigned grid_for(long long n) {   // (used by the exported GEMM)
    long long b = (n + 255) / 256;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

extern "C" OCBNN_API int64_t ocbnn_gemm_tf32x3_workspace_bytes(int64_t m, int64_t n, int64_t k) {
    const int64_t kp = (k + 31) / 32 * 32;
    return 2 * (m + n) * kp * (int64_t)sizeof(float) + 256;
}

// C (m x n, row-major) = A (m x k) * B (n x k)^T, 3xTF32 on tcgen05.  Device pointers; workspace from the function above.
extern "C" OCBNN_API int ocbnn_gemm_tf32x3(const float* a, const float* b, float* c, int64_t m, int64_t n, int64_t k, void* workspace,
                                           int64_t workspace_bytes, void* stream) {
    OCBNN_REQUIRE(a && b && c && m >= 1 && n >= 1 && k >= 1, OCBNN_EINVAL, "ocbnn_gemm_tf32x3: bad argument");
    OCBNN_REQUIRE(workspace && workspace_bytes >= ocbnn_gemm_tf32x3_workspace_bytes(m, n, k), OCBNN_EINVAL, "ocbnn_gemm_tf32x3: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const long long kp = (k + 31) / 32 * 32;
    float* a_hi = reinterpret_cast<float*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* a_lo = a_hi + m * kp;
    float* b_hi = a_lo + m * kp;
    float* b_lo = b_hi + n * kp;
    k_split_tf32<<<grid_for(m * kp), 256, 0, st>>>(a, m, k, kp, a_hi, a_lo);
    OCBNN_LAUNCH_CHECK();
    k_split_tf32<<<grid_for(n * kp), 256, 0, st>>>(b, n, k, kp, b_hi, b_lo);
    OCBNN_LAUNCH_CHECK();
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    const bool wide = n > 64;
    int rc;
    if ((rc = make_tmap(&ta_hi, a_hi, m, kp, kBM))) return rc;
    if ((rc = make_tmap(&ta_lo, a_lo, m, kp, kBM))) return rc;
    if ((rc = make_tmap(&tb_hi, b_hi, n, kp, wide ? 128 : 64))) return rc;
    if ((rc = make_tmap(&tb_lo, b_lo, n, kp, wide ? 128 : 64))) return rc;
    EpiStore epi{c, n, 0, (int)m, (int)n};
    return wide ? launch_gemm<128>(ta_hi, ta_lo, tb_hi, tb_lo, epi, m, n, k, 1, st) : launch_gemm<64>(ta_hi, ta_lo, tb_hi, tb_lo, epi, m, n, k, 1, st);
}
