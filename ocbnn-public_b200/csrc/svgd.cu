#include <atomic>
#include <cstdlib>
// svgd.cu — SVGD kernels, exact fp32 SIMT path.
//   K3a  median-heuristic bandwidth: exact LOWER median of ||x_i - x_j + 1e-6||, i<j, zeros dropped, by a
//        3-pass radix select on the fp32 bit patterns (11+11+10 bits); h = med^2 / ln n   (inference.py:208-217)
//   K3b  phi_i = (1/n) sum_j K_ij [ g_j - (2/h)(x_j - x_i) ],  K_ij = exp(-||x_i - x_j||^2 / h)
//        (closed form of the n^2 autograd loop, inference.py:219-259)
//   K3c  Adagrad update (torch.optim.Adagrad defaults, inference.py:272-281)
// The tcgen05 path for the same contraction lives in svgd_tc.cu; this file is the exact reference arm on the GPU
// and the path used when D is small.
#include "common.cuh"

namespace ocbnn {

constexpr int kLpr = 4;           // lanes per particle row
constexpr int kRowsPerCta = 32;   // 128 threads
constexpr int kTj = 64;           // staged j-tile
constexpr int kBins = 2048;

// select_state: [0] prefix bits, [1] rank still to skip inside the candidate set, [2] candidates, [3] scratch
__device__ __forceinline__ bool bin_of(float d, int pass, uint32_t prefix, uint32_t& bin) {
    uint32_t b = __float_as_uint(d);
    if (b == 0u) return false;                       // exact zeros are dropped (fkernel.nonzero())
    if (pass == 0) { bin = b >> 21; return true; }
    if (pass == 1) { if ((b >> 21) != (prefix >> 21)) return false; bin = (b >> 10) & 2047u; return true; }
    if ((b >> 10) != (prefix >> 10)) return false;
    bin = b & 1023u;
    return true;
}

template <int DL>
__global__ void __launch_bounds__(kRowsPerCta* kLpr)
k_dist_hist(const float* __restrict__ x, long long n, int d, long long row_start, long long row_step, long long n_rows, int pass,
            uint32_t* __restrict__ hist, const unsigned long long* __restrict__ state) {
    __shared__ uint32_t sh[kBins];
    __shared__ __align__(16) float xs[kTj][kLpr * DL];
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0u;
    const uint32_t prefix = pass > 0 ? (uint32_t)state[0] : 0u;
    const int r = threadIdx.x % kLpr;
    const long long t = (long long)blockIdx.x * kRowsPerCta + threadIdx.x / kLpr;
    const bool rowok = t < n_rows;
    const long long i = rowok ? row_start + t * row_step : n;     // n: no pairs
    float xi[DL];
#pragma unroll
    for (int k = 0; k < DL; ++k) {
        int dim = r * DL + k;
        xi[k] = (rowok && dim < d) ? __ldg(x + i * d + dim) : 0.f;
    }
    // smallest row of this CTA decides where the j sweep starts
    const long long t0 = (long long)blockIdx.x * kRowsPerCta;
    const long long imin = row_start + t0 * row_step;
    for (long long j0 = (imin + 1) / kTj * kTj; j0 < n; j0 += kTj) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < kTj * kLpr * DL; idx += blockDim.x) {
            int jj = idx / (kLpr * DL), dim = idx % (kLpr * DL);
            long long j = j0 + jj;
            xs[jj][dim] = (j < n && dim < d) ? __ldg(x + j * d + dim) : 0.f;
        }
        __syncthreads();
        for (int jj = 0; jj < kTj; ++jj) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < DL; ++k) {
                float df = (xi[k] - xs[jj][r * DL + k]) + 1e-6f;      // PairwiseDistance: ||x1 - x2 + eps||
                df = (r * DL + k < d) ? df : 0.f;
                s = fmaf(df, df, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            long long j = j0 + jj;
            if (r == 0 && j > i && j < n) {
                uint32_t bin;
                if (bin_of(sqrtf(s), pass, prefix, bin)) atomicAdd(&sh[bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kBins; b += blockDim.x)
        if (sh[b]) atomicAdd(&hist[pass * kBins + b], sh[b]);
}

// generic-D distance histogram: one warp per (i, j-range); used when D > 64
__global__ void __launch_bounds__(256)
k_dist_hist_wide(const float* __restrict__ x, long long n, int d, long long row_start, long long row_step, long long n_rows, int pass,
                 uint32_t* __restrict__ hist, const unsigned long long* __restrict__ state) {
    __shared__ uint32_t sh[kBins];
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const uint32_t prefix = pass > 0 ? (uint32_t)state[0] : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long t = blockIdx.x;
    if (t < n_rows) {
        const long long i = row_start + t * row_step;
        for (long long j = i + 1 + warp; j < n; j += nwarp) {
            float s = 0.f;
            for (int k = lane; k < d; k += 32) {
                float df = (__ldg(x + i * d + k) - __ldg(x + j * d + k)) + 1e-6f;
                s = fmaf(df, df, s);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            uint32_t bin;
            if (lane == 0 && bin_of(sqrtf(s), pass, prefix, bin)) atomicAdd(&sh[bin], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kBins; b += blockDim.x)
        if (sh[b]) atomicAdd(&hist[pass * kBins + b], sh[b]);
}

__global__ void k_zero_u32(uint32_t* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
}

// after pass p: locate the bin holding the wanted rank; after pass 2 emit h.  One block of 256 threads: 8 bins per thread,
// block-wide exclusive scan, the thread whose range contains the rank writes the new state.
__global__ void __launch_bounds__(256)
k_select_update(int pass, const uint32_t* __restrict__ hist, unsigned long long* __restrict__ state, long long n,
                float* __restrict__ out_h) {
    __shared__ unsigned long long wsum[8];
    __shared__ unsigned long long s_tot;
    const uint32_t* h = hist + pass * kBins;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned long long loc[8], mine = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { loc[k] = h[t * 8 + k]; mine += loc[k]; }
    unsigned long long incl = mine;                         // inclusive scan over threads
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        unsigned long long v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long base = 0;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    if (t == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        s_tot = tot;
    }
    __syncthreads();
    const unsigned long long tot = s_tot;
    const unsigned long long rank = pass == 0 ? (tot ? (tot - 1) / 2 : 0) : state[1];     // LOWER median (torch.median)
    const unsigned long long excl = base + incl - mine;
    const bool owner = tot == 0 ? t == 0 : (rank >= excl && rank < excl + mine);
    __syncthreads();                                        // every thread has read state[1] before the owner overwrites it
    if (owner) {
        unsigned long long cum = excl;
        int b = t * 8;
        for (int k = 0; k < 8; ++k, ++b) {
            if (cum + loc[k] > rank) break;
            cum += loc[k];
        }
        if (b >= kBins) b = kBins - 1;
        uint32_t prefix = pass == 0 ? 0u : (uint32_t)state[0];
        prefix |= pass == 0 ? ((uint32_t)b << 21) : pass == 1 ? ((uint32_t)b << 10) : (uint32_t)b;
        if (pass == 0) state[2] = tot;
        const unsigned long long cands = pass == 0 ? tot : state[2];
        state[0] = prefix;
        state[1] = rank - cum;
        if (pass == 2 && out_h) {
            float med = __uint_as_float(prefix);
            if (cands == 0) med = nanf("");                // torch: median of an empty tensor
            *out_h = __fdiv_rn(__fmul_rn(med, med), (float)log((double)n));
        }
    }
}

// ---- K3b: fused exact phi for D <= 64 -------------------------------------------------------------------------------
template <int DL>
__global__ void __launch_bounds__(kRowsPerCta* kLpr)
k_svgd_phi_small(const float* __restrict__ x, const float* __restrict__ g, long long n, int d, long long row_begin, long long n_rows,
                 const float* __restrict__ hptr, float* __restrict__ phi) {
    __shared__ __align__(16) float xs[kTj][kLpr * DL];
    __shared__ __align__(16) float gs[kTj][kLpr * DL];
    const float h = *hptr;
    const float inv_h = 1.f / h;
    const int r = threadIdx.x % kLpr;
    const long long t = (long long)blockIdx.x * kRowsPerCta + threadIdx.x / kLpr;
    const bool rowok = t < n_rows;
    const long long i = rowok ? row_begin + t : 0;
    float xi[DL], og[DL], ox[DL];
#pragma unroll
    for (int k = 0; k < DL; ++k) {
        int dim = r * DL + k;
        xi[k] = (dim < d) ? __ldg(x + i * d + dim) : 0.f;
        og[k] = 0.f;
        ox[k] = 0.f;
    }
    float rowsum = 0.f;
    for (long long j0 = 0; j0 < n; j0 += kTj) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < kTj * kLpr * DL; idx += blockDim.x) {
            int jj = idx / (kLpr * DL), dim = idx % (kLpr * DL);
            long long j = j0 + jj;
            bool ok = j < n && dim < d;
            xs[jj][dim] = ok ? __ldg(x + j * d + dim) : 0.f;
            gs[jj][dim] = ok ? __ldg(g + j * d + dim) : 0.f;
        }
        __syncthreads();
        const int jn = (int)min((long long)kTj, n - j0);
        for (int jj = 0; jj < jn; ++jj) {
            float xv[DL];
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < DL; ++k) {
                xv[k] = xs[jj][r * DL + k];
                float df = xv[k] - xi[k];
                s = fmaf(df, df, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const float kij = expf(-s * inv_h);
            rowsum += kij;
#pragma unroll
            for (int k = 0; k < DL; ++k) {
                og[k] = fmaf(kij, gs[jj][r * DL + k], og[k]);
                ox[k] = fmaf(kij, xv[k], ox[k]);
            }
        }
    }
    if (rowok) {
        const float c = 2.f * inv_h, invn = 1.f / (float)n;
#pragma unroll
        for (int k = 0; k < DL; ++k) {
            int dim = r * DL + k;
            if (dim < d) phi[t * d + dim] = (og[k] - c * (ox[k] - xi[k] * rowsum)) * invn;
        }
    }
}

// ---- K3b wide-D exact path: kernel matrix, then a tiled fp32 GEMM  phi = (K [G | X] ...) ---------------------------------
// kmat[t][j] = exp(-||x_i - x_j||^2/h), i = row_begin + t ; one warp per (t, j)
__global__ void __launch_bounds__(256)
k_kernel_matrix_wide(const float* __restrict__ x, long long n, int d, long long row_begin, long long n_rows, const float* __restrict__ hptr,
                     float* __restrict__ kmat, float* __restrict__ rowsum) {
    const float inv_h = 1.f / *hptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const long long t = blockIdx.x;
    if (t >= n_rows) return;
    const long long i = row_begin + t;
    float rs = 0.f;
    for (long long j = warp; j < n; j += nwarp) {
        float s = 0.f;
        for (int k = lane; k < d; k += 32) {
            float df = __ldg(x + i * d + k) - __ldg(x + j * d + k);
            s = fmaf(df, df, s);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        float kij = expf(-s * inv_h);
        if (lane == 0) kmat[t * n + j] = kij;
        rs += kij;
    }
    __shared__ float red[8];
    if (lane == 0) red[warp] = rs;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += red[w];
        rowsum[t] = s;
    }
}

// phi[t][c] = ( sum_j K[t][j] (g[j][c] - (2/h) x[j][c]) + (2/h) x_i[c] rowsum[t] ) / n ; 64x64 tiles, 4x4 per thread
__global__ void __launch_bounds__(256)
k_phi_gemm_wide(const float* __restrict__ kmat, const float* __restrict__ rowsum, const float* __restrict__ x, const float* __restrict__ g,
                long long n, int d, long long row_begin, long long n_rows, const float* __restrict__ hptr, float* __restrict__ phi) {
    __shared__ __align__(16) float ks[16][64 + 4];   // [j][t]
    __shared__ __align__(16) float vs[16][64 + 4];   // [j][c]
    const float c2 = 2.f / *hptr;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const long long t0 = (long long)blockIdx.y * 64;
    const int c0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (long long j0 = 0; j0 < n; j0 += 16) {
        for (int idx = threadIdx.x; idx < 16 * 64; idx += 256) {
            int jj = idx % 16, tt = idx / 16;
            long long t = t0 + tt, j = j0 + jj;
            ks[jj][tt] = (t < n_rows && j < n) ? __ldg(kmat + t * n + j) : 0.f;
            int cc = idx % 64, j2 = idx / 64;
            long long jb = j0 + j2;
            int c = c0 + cc;
            vs[j2][cc] = (jb < n && c < d) ? (__ldg(g + jb * d + c) - c2 * __ldg(x + jb * d + c)) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            float4 a = *reinterpret_cast<const float4*>(&ks[jj][4 * ty]);
            float4 b = *reinterpret_cast<const float4*>(&vs[jj][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
        }
        __syncthreads();
    }
    const float invn = 1.f / (float)n;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        long long t = t0 + 4 * ty + p;
        if (t >= n_rows) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int c = c0 + 4 * tx + q;
            if (c < d) phi[t * d + c] = (acc[p][q] + c2 * __ldg(x + (row_begin + t) * d + c) * rowsum[t]) * invn;
        }
    }
}

// ---- K3c ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_adagrad(float* __restrict__ x, float* __restrict__ acc, const float* __restrict__ dir, long long n, float lr, float sign) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float gr = sign * dir[i];
        float a = __fadd_rn(acc[i], __fmul_rn(gr, gr));                     // state_sum.addcmul_(grad, grad)
        acc[i] = a;
        float sd = __fadd_rn(sqrtf(a), 1e-10f);                             // std = sqrt(state_sum) + eps
        x[i] = __fadd_rn(x[i], __fmul_rn(-lr, __fdiv_rn(gr, sd)));         // param.addcdiv_(grad, std, value=-lr)
    }
}

}  // namespace ocbnn

using namespace ocbnn;

// tensor path (svgd_tc.cu), collectives (comm.cu)
#include "svgd_tc.h"
#include "comm.h"

// use_tensor: 0 exact fp32 SIMT kernels ; 1 tcgen05 path ; -1 choose (tensor path from 1024 particles up)
static bool want_tensor(int32_t use_tensor, int64_t n) { return use_tensor > 0 || (use_tensor < 0 && n >= 1024); }
static const int64_t kHdrBytes = 3 * kBins * sizeof(uint32_t) + 256;     // select histogram + select state at the start of every workspace

extern "C" OCBNN_API int64_t ocbnn_svgd_workspace_bytes(int64_t n, int64_t d, int64_t n_rows, int32_t use_tensor) {
    int64_t b = kHdrBytes;
    if (want_tensor(use_tensor, n)) return b + tc::tc_workspace_bytes(n, d, n_rows);
    if (d > 64) b += (n_rows * n + n_rows) * (int64_t)sizeof(float);      // kernel matrix + row sums (exact wide path)
    return b;
}

// 1 bracket select (default), 0 full radix select, 2 bracket select with a sabotaged bracket (takes the in-kernel fallback);
// initial value from OCBNN_SVGD_SELECT = fast | legacy | fallback
static std::atomic<int> g_select_mode{[] { const char* e = getenv("OCBNN_SVGD_SELECT"); return !e ? 1 : e[0] == 'l' ? 0 : (e[0] == 'f' && e[1] == 'a' && e[2] == 'l') ? 2 : 1; }()};
namespace ocbnn { int svgd_select_mode_value() { return g_select_mode.load(); } }
extern "C" OCBNN_API int ocbnn_svgd_select_mode(int32_t mode) {
    const int prev = g_select_mode.load();
    if (mode >= 0 && mode <= 2) g_select_mode.store(mode);
    return prev;
}

namespace ocbnn {
// SIMT arm: one radix-select histogram pass over rows row_start + t row_step (pairs i < j), and the rank bookkeeping after it
int simt_hist_pass(const float* x, long long n, long long d, long long row_start, long long row_step, long long n_rows, int pass, uint32_t* hist,
                   const unsigned long long* state, cudaStream_t st) {
    if (pass == 0) {
        k_zero_u32<<<(3 * kBins + 255) / 256, 256, 0, st>>>(hist, 3 * kBins);
        OCBNN_LAUNCH_CHECK();
    }
    if (n_rows <= 0) return OCBNN_OK;
    if (d <= 64) {
        unsigned blocks = (unsigned)((n_rows + kRowsPerCta - 1) / kRowsPerCta);
        if (d <= 32) k_dist_hist<8><<<blocks, kRowsPerCta * kLpr, 0, st>>>(x, n, (int)d, row_start, row_step, n_rows, pass, hist, state);
        else k_dist_hist<16><<<blocks, kRowsPerCta * kLpr, 0, st>>>(x, n, (int)d, row_start, row_step, n_rows, pass, hist, state);
    } else {
        k_dist_hist_wide<<<(unsigned)n_rows, 256, 0, st>>>(x, n, (int)d, row_start, row_step, n_rows, pass, hist, state);
    }
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
int simt_select_update(int pass, uint32_t* hist, unsigned long long* state, long long n, float* out_h, cudaStream_t st) {
    k_select_update<<<1, 256, 0, st>>>(pass, hist, state, n, out_h);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
// SIMT arm of phi for rows [row_begin, row_begin + n_rows); `wide` (D > 64): n_rows x n kernel matrix + n_rows row sums
int simt_phi(const float* x, const float* g, long long n, long long d, long long row_begin, long long n_rows, const float* h, float* out_phi, float* wide,
             cudaStream_t st) {
    if (d <= 64) {
        unsigned blocks = (unsigned)((n_rows + kRowsPerCta - 1) / kRowsPerCta);
        if (d <= 32) k_svgd_phi_small<8><<<blocks, kRowsPerCta * kLpr, 0, st>>>(x, g, n, (int)d, row_begin, n_rows, h, out_phi);
        else k_svgd_phi_small<16><<<blocks, kRowsPerCta * kLpr, 0, st>>>(x, g, n, (int)d, row_begin, n_rows, h, out_phi);
        OCBNN_LAUNCH_CHECK();
        return OCBNN_OK;
    }
    float* kmat = wide;
    float* rowsum = kmat + n_rows * n;
    k_kernel_matrix_wide<<<(unsigned)n_rows, 256, 0, st>>>(x, n, (int)d, row_begin, n_rows, h, kmat, rowsum);
    OCBNN_LAUNCH_CHECK();
    dim3 grid((unsigned)((d + 63) / 64), (unsigned)((n_rows + 63) / 64));
    k_phi_gemm_wide<<<grid, 256, 0, st>>>(kmat, rowsum, x, g, n, (int)d, row_begin, n_rows, h, out_phi);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
int adagrad_launch(float* x, float* acc, const float* dir, long long n_elem, float lr, float grad_sign, cudaStream_t st) {
    if (n_elem == 0) return OCBNN_OK;
    long long blocks = (n_elem + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_adagrad<<<(unsigned)blocks, 256, 0, st>>>(x, acc, dir, n_elem, lr, grad_sign);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
}  // namespace ocbnn

// K3a, all particles on this GPU
extern "C" OCBNN_API int ocbnn_svgd_bandwidth(const float* x, int64_t n, int64_t d, float* out_h, int32_t use_tensor, void* workspace,
                                              int64_t workspace_bytes, void* stream) {
    OCBNN_REQUIRE(x && out_h && n >= 2 && d >= 1, OCBNN_EINVAL, "ocbnn_svgd_bandwidth: bad argument");
    OCBNN_REQUIRE(workspace && workspace_bytes >= ocbnn_svgd_workspace_bytes(n, d, n, use_tensor), OCBNN_EINVAL,
                  "ocbnn_svgd_bandwidth: workspace too small (see ocbnn_svgd_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* hist = reinterpret_cast<uint32_t*>(workspace);
    unsigned long long* state = reinterpret_cast<unsigned long long*>(hist + 3 * kBins);
    void* tcws = reinterpret_cast<char*>(workspace) + kHdrBytes;
    int rc;
    if (want_tensor(use_tensor, n)) {
        if ((rc = tc::tc_svgd_prepare(x, n, d, 0, n, tcws, workspace_bytes - kHdrBytes, st))) return rc;
        const int mode = g_select_mode.load();
        if (mode == 0 && !tc::tc_is_flash(n, d)) {          // 3-pass radix select over the materialised distances (D > 64)
            for (int pass = 0; pass < 3; ++pass) {
                if (pass == 0) {
                    k_zero_u32<<<(3 * kBins + 255) / 256, 256, 0, st>>>(hist, 3 * kBins);
                    OCBNN_LAUNCH_CHECK();
                }
                if ((rc = tc::tc_svgd_hist(n, d, 0, n, pass, hist, state, tcws, st))) return rc;
                if ((rc = simt_select_update(pass, hist, state, n, pass == 2 ? out_h : nullptr, st))) return rc;
            }
            return OCBNN_OK;
        }
        return tc::tc_svgd_select_local(n, d, tcws, out_h, mode, state, st);
    }
    for (int pass = 0; pass < 3; ++pass) {
        if ((rc = simt_hist_pass(x, n, d, 0, 1, n, pass, hist, state, st))) return rc;
        if ((rc = simt_select_update(pass, hist, state, n, pass == 2 ? out_h : nullptr, st))) return rc;
    }
    return OCBNN_OK;
}

// use_tensor: 0 SIMT ; 1 tensor path (centres/splits X itself) ; 2 tensor path, the workspace was prepared for this X and these
// rows by ocbnn_svgd_bandwidth ; -1 choose between 0 and 1
extern "C" OCBNN_API int ocbnn_svgd_phi(const float* x, const float* g, int64_t n, int64_t d, int64_t row_begin, int64_t n_rows, const float* h,
                              float* out_phi, int32_t use_tensor, void* workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    OCBNN_REQUIRE(x && g && h && out_phi && n >= 1 && d >= 1 && row_begin >= 0 && row_begin + n_rows <= n, OCBNN_EINVAL,
                  "ocbnn_svgd_phi: bad argument");
    if (n_rows == 0) return OCBNN_OK;
    if (want_tensor(use_tensor, n)) {
        OCBNN_REQUIRE(workspace && workspace_bytes >= ocbnn_svgd_workspace_bytes(n, d, n_rows, 1), OCBNN_EINVAL,
                      "ocbnn_svgd_phi: workspace too small for the tensor path (see ocbnn_svgd_workspace_bytes)");
        void* tcws = reinterpret_cast<char*>(workspace) + kHdrBytes;
        if (use_tensor != 2) {
            int rc = tc::tc_svgd_prepare(x, n, d, row_begin, n_rows, tcws, workspace_bytes - kHdrBytes, st);
            if (rc) return rc;
        }
        return tc::tc_svgd_phi(g, h, n, d, row_begin, n_rows, out_phi, tcws, st);
    }
    OCBNN_REQUIRE(d <= 64 || (workspace && workspace_bytes >= kHdrBytes + (n_rows * n + n_rows) * (int64_t)sizeof(float)), OCBNN_EINVAL,
                  "ocbnn_svgd_phi: workspace too small for the wide exact path (see ocbnn_svgd_workspace_bytes)");
    return simt_phi(x, g, n, d, row_begin, n_rows, h, out_phi, d > 64 ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kHdrBytes) : nullptr, st);
}

// One SVGD iteration, particles sharded over the ranks of `comm` (NULL: one rank).  See include/ocbnn_b200.h.
// Stream plan: K1 of the owned rows runs on the side stream while the main stream centres X and does the (compute-heavy, X-only)
// first stage of the select; every collective is issued on the main stream after the join, in the same order on every rank.
extern "C" OCBNN_API int ocbnn_svgd_step(ocbnn_problem* ph, float* x_all, float* g_all, float* acc, float* phi, float* h, int64_t n, int64_t row_begin,
                                         int64_t n_rows, float lr, int32_t batch_offset, int32_t batch_stride, int32_t with_data, int32_t use_tensor,
                                         void* workspace, int64_t workspace_bytes, ocbnn_comm* comm_h, void* stream) {
    OCBNN_REQUIRE(ph && x_all && g_all && acc && phi && h && n >= 2, OCBNN_EINVAL, "ocbnn_svgd_step: bad argument");
    Comm* comm = reinterpret_cast<Comm*>(comm_h);
    const int rank = comm_rank(comm), world = comm_world(comm);
    long long lo, cnt;
    shard_of(n, rank, world, &lo, &cnt);
    OCBNN_REQUIRE(row_begin == lo && n_rows == cnt, OCBNN_EINVAL, "ocbnn_svgd_step: rank %d of %d owns rows [%lld, %lld) of %lld, not [%lld, %lld)", rank, world,
                  lo, lo + cnt, (long long)n, (long long)row_begin, (long long)(row_begin + n_rows));
    OCBNN_REQUIRE(batch_stride <= 1 || (batch_offset >= 0 && batch_offset < batch_stride), OCBNN_EINVAL, "batch_offset must be in [0, stride)");
    const Problem& p = *problem_of(ph);
    const long long d = p.d;
    const bool tensor = want_tensor(use_tensor, n);
    OCBNN_REQUIRE(workspace && workspace_bytes >= ocbnn_svgd_workspace_bytes(n, d, n_rows, tensor ? 1 : 0), OCBNN_EINVAL,
                  "ocbnn_svgd_step: workspace too small (see ocbnn_svgd_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* hist = reinterpret_cast<uint32_t*>(workspace);
    unsigned long long* state = reinterpret_cast<unsigned long long*>(hist + 3 * kBins);
    void* tcws = reinterpret_cast<char*>(workspace) + kHdrBytes;
    const long long tcws_bytes = workspace_bytes - kHdrBytes;
    int rc;

    // ---- fork: gradients of the owned rows (K1) on the side stream ----
    SideStream* side;
    if ((rc = side_stream(&side))) return rc;
    OCBNN_CUDA(cudaEventRecord(side->fork, st));
    OCBNN_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    if (n_rows > 0) {
        const EvalArgs a = eval_args(p, batch_offset, batch_stride, with_data);
        if ((rc = logpost_any(p, a, x_all + row_begin * d, n_rows, nullptr, g_all + row_begin * d, 0, side->stream))) return rc;
    }
    OCBNN_CUDA(cudaEventRecord(side->join, side->stream));

    // ---- main: bandwidth (K3a) ----
    const int mode = g_select_mode.load();
    const bool legacy = tensor && mode == 0 && !tc::tc_is_flash(n, d);       // plain radix select over materialised distances
    unsigned long long* red64 = nullptr;
    if (tensor) {
        if ((rc = tc::tc_svgd_prepare(x_all, n, d, row_begin, n_rows, tcws, tcws_bytes, st))) return rc;
        if (world == 1 && !legacy) {
            if ((rc = tc::tc_svgd_select_local(n, d, tcws, h, mode, state, st))) return rc;
        } else if (!legacy) {
            if ((rc = tc::tc_svgd_sel_stage1(n, d, row_begin, n_rows, rank, world, tcws, mode, &red64, st))) return rc;
        }
    }
    OCBNN_CUDA(cudaStreamWaitEvent(st, side->join, 0));                       // join
    if ((rc = comm_all_gather_rows(comm, g_all, n, d, st))) return rc;
    if (tensor && legacy) {
        for (int pass = 0; pass < 3; ++pass) {
            if (pass == 0) {
                k_zero_u32<<<(3 * kBins + 255) / 256, 256, 0, st>>>(hist, 3 * kBins);
                OCBNN_LAUNCH_CHECK();
            }
            if (n_rows > 0 && (rc = tc::tc_svgd_hist(n, d, row_begin, n_rows, pass, hist, state, tcws, st))) return rc;
            if ((rc = comm_all_sum(comm, hist + pass * kBins, kBins, kCommU32, st))) return rc;
            if ((rc = simt_select_update(pass, hist, state, n, pass == 2 ? h : nullptr, st))) return rc;
        }
    } else if (tensor && world > 1) {
        if ((rc = comm_all_sum(comm, red64, 4 + 2048, kCommU64, st))) return rc;
        if ((rc = tc::tc_svgd_sel_stage2(n, d, row_begin, n_rows, tcws, st))) return rc;
        for (int pass = 0; pass < 3; ++pass) {
            unsigned* hp = nullptr;
            if ((rc = tc::tc_svgd_sel_hist(n, d, row_begin, n_rows, rank, world, pass, tcws, &hp, st))) return rc;
            if ((rc = comm_all_sum(comm, hp, 2048, kCommU32, st))) return rc;
            if ((rc = tc::tc_svgd_sel_update(n, d, row_begin, n_rows, pass, tcws, h, st))) return rc;
        }
    } else if (!tensor) {
        // exact SIMT arm: rows dealt out cyclically (balances the i < j triangle), histograms summed over the ranks
        const long long my_rows = n > rank ? (n - rank + world - 1) / world : 0;
        for (int pass = 0; pass < 3; ++pass) {
            if ((rc = simt_hist_pass(x_all, n, d, rank, world, my_rows, pass, hist, state, st))) return rc;
            if ((rc = comm_all_sum(comm, hist + pass * kBins, kBins, kCommU32, st))) return rc;
            if ((rc = simt_select_update(pass, hist, state, n, pass == 2 ? h : nullptr, st))) return rc;
        }
    }

    // ---- phi (K3b) and Adagrad (K3c) on the owned rows, then the particles go round ----
    if (n_rows > 0) {
        // particles.grad = -phi (inference.py:280); on the tensor path the Adagrad update rides in the finalize kernel of phi
        if (tensor) rc = tc::tc_svgd_phi(g_all, h, n, d, row_begin, n_rows, phi, tcws, st, x_all + row_begin * d, acc, lr);
        else rc = simt_phi(x_all, g_all, n, d, row_begin, n_rows, h, phi, d > 64 ? reinterpret_cast<float*>(tcws) : nullptr, st);
        if (rc) return rc;
        if (!tensor && (rc = adagrad_launch(x_all + row_begin * d, acc, phi, n_rows * d, lr, -1.f, st))) return rc;
    }
    return comm_all_gather_rows(comm, x_all, n, d, st);
}

extern "C" OCBNN_API int ocbnn_adagrad_step(float* x, float* acc, const float* dir, int64_t n_elem, float lr, float grad_sign, void* stream) {
    OCBNN_REQUIRE(x && acc && dir && n_elem >= 0, OCBNN_EINVAL, "ocbnn_adagrad_step: bad argument");
    return adagrad_launch(x, acc, dir, n_elem, lr, grad_sign, (cudaStream_t)stream);
}
