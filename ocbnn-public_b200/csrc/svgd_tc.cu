// svgd_tc.cu — tensor-core path of the SVGD update (K3b): the two dense contractions
//     S = Xc Xc^T                 (n x n x D)      -> K_ij = exp(-(r_i + r_j - 2 S_ij)/h)
//     O = K [G | Xc]              (n x 2D x n)     -> phi_i = (O_G - (2/h)(O_X - x_i rowsum_i)) / n
// run on tcgen05 with TMA-staged operands in 3xTF32 precision (tc_gemm.cuh); Xc = X - mean(X) (phi is translation
// invariant; centring removes the cancellation in the Gram form of the distance).  The exact fp32 SIMT arm in svgd.cu is
// the parity reference of this path.  Also exports the GEMM building block itself for testing.
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

struct EpiGram {           // D2[row][col] = max(r_row + r_col - 2 S, 0)
    float* d2;
    const float* r;
    long long n;
    __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[32], int) const {
        if (row >= n) return;
        const float ri = r[row];
        float* dst = d2 + (long long)row * n + col0;
        if (col0 + 32 <= n && (n & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 rj = *reinterpret_cast<const float4*>(r + col0 + i);
                *reinterpret_cast<float4*>(dst + i) =
                    make_float4(fmaxf(fmaf(-2.f, acc[i], ri + rj.x), 0.f), fmaxf(fmaf(-2.f, acc[i + 1], ri + rj.y), 0.f),
                                fmaxf(fmaf(-2.f, acc[i + 2], ri + rj.z), 0.f), fmaxf(fmaf(-2.f, acc[i + 3], ri + rj.w), 0.f));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < n) dst[i] = fmaxf(fmaf(-2.f, acc[i], ri + r[col0 + i]), 0.f);
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

}  // namespace tc
}  // namespace ocbnn

using namespace ocbnn;
using namespace ocbnn::tc;

static unsigned grid_for(long long n) {
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
