// tc_gemm.cuh — tcgen05 / TMEM / TMA building block: C[M x N] = A[M x K] * B[N x K]^T in "3xTF32" precision.
//
// tcgen05 has no fp32 MMA; kind::tf32 keeps 10 mantissa bits (5e-4 relative error), far from the 1e-5 parity bar.
// Every fp32 operand is therefore split on the fly into hi = tf32(x) and lo = x - hi (both exact in fp32), the tiles of
// all four parts are staged by TMA (128B swizzle, K-major) and each K step issues hi*hi + hi*lo + lo*hi into the same
// TMEM accumulator (the dropped lo*lo term is 2^-22 relative).
//
// One CTA = one 128 x BN output tile (x one K split): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer
// (one elected lane), warps 2..5 = epilogue (one TMEM lane quarter each).  smem full/empty mbarrier ring between
// producer and MMA, tcgen05.commit to release stages and to signal the epilogue.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace ocbnn {
namespace tc {

constexpr int kBM = 128;          // UMMA M (cta_group::1)
constexpr int kBK = 32;           // fp32 elements per K block = one 128-byte swizzle row
constexpr int kUmmaK = 8;         // tf32 elements per MMA K step (32 bytes)
constexpr int kThreads = 192;     // 6 warps
constexpr uint32_t kSpinLimit = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp.  TMA / tcgen05 instructions take their operands from uniform registers; issued under a
// `lane == 0` branch the compiler cannot prove the region is single-threaded and wraps EVERY such instruction in an
// ELECT / branch "waterfall" loop (measured: ~100 clk per tcgen05.mma instead of the 32-64 clk the tensor pipe needs).
// elect.sync tells it.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// warp index, known to be warp-uniform by the compiler
__device__ __forceinline__ int warp_idx_sync() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded spin: a protocol bug traps (CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread t <-> TMEM lane base+t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// The tensor core adds each MMA into its fp32 accumulator with truncation, so a chain of m MMAs into one TMEM accumulator
// carries a (one-sided) bias of ~m 2^-24 of the partial sums - measured 1.4e-5 relative on the distances of a K = 11232 Gram
// (351 K blocks x 12 MMAs).  Accumulation chains are therefore cut after kChainKB K blocks (192 MMAs): two TMEM accumulators
// ping-pong, the epilogue warps drain every finished chunk into fp32 registers (round-to-nearest adds) while the tensor pipe
// runs the next one.
constexpr int kChainKB = 16;

template <int BN>
struct GemmCfg {
    static constexpr int kStages = BN >= 128 ? 3 : 4;
    static constexpr int kTileA = kBM * kBK * 4;                 // 16 KB
    static constexpr int kTileB = BN * kBK * 4;
    static constexpr int kStageBytes = 2 * kTileA + 2 * kTileB;  // A_hi, A_lo, B_hi, B_lo
    static constexpr int kNumBars = 2 * kStages + 4;             // full[], empty[], acc_full[2], acc_empty[2]
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int kAccCols = BN < 32 ? 32 : BN;           // one accumulator
    static constexpr int kTmemCols = 2 * kAccCols;               // power of two >= 32
};

// Epilogue functor interface: void operator()(int row, int col0, const float (&acc)[32], int split) const
// called by the thread that owns accumulator row `row` of the C tile for columns [col0, col0 + 32).
template <int BN, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_tf32x3(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
              const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, Epi epi, int num_kb, int kb_per_split) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + Cfg::kStages * Cfg::kStageBytes;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + Cfg::kStages * Cfg::kStageBytes + 8 * Cfg::kNumBars);
    const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
    const int kb_begin = blockIdx.z * kb_per_split;
    const int kb_end = min(num_kb, kb_begin + kb_per_split);
    const int nkb = kb_end - kb_begin;
    const int nchunks = (nkb + kChainKB - 1) / kChainKB;

    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (Cfg::kStages + s); };
    auto acc_full = [&](int b) { return bars + 8u * (2 * Cfg::kStages + b); };
    auto acc_empty = [&](int b) { return bars + 8u * (2 * Cfg::kStages + 2 + b); };

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), 4);                    // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        if (elect_one_sync()) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % Cfg::kStages;
                const uint32_t ph = (i / Cfg::kStages) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                const uint32_t st = base + s * Cfg::kStageBytes;
                mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
                const int kx = (kb_begin + i) * kBK;
                tma_load_2d(st, &tm_a_hi, full_bar(s), kx, m0);
                tma_load_2d(st + Cfg::kTileA, &tm_a_lo, full_bar(s), kx, m0);
                tma_load_2d(st + 2 * Cfg::kTileA, &tm_b_hi, full_bar(s), kx, n0);
                tma_load_2d(st + 2 * Cfg::kTileA + Cfg::kTileB, &tm_b_lo, full_bar(s), kx, n0);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = make_idesc_tf32(kBM, BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % Cfg::kStages;
                const uint32_t ph = (i / Cfg::kStages) & 1;
                const int chunk = i / kChainKB, b = chunk & 1, ic = i - chunk * kChainKB;
                if (ic == 0) {                             // a fresh chain: the accumulator must have been drained
                    mbar_wait(acc_empty(b), ((chunk >> 1) & 1) ^ 1);
                    tc_fence_after();
                }
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t st = base + s * Cfg::kStageBytes;
                const uint32_t d_acc = tmem_d + (uint32_t)(b * Cfg::kAccCols);
                const uint64_t a_hi = make_kmajor_sw128_desc(st), a_lo = make_kmajor_sw128_desc(st + Cfg::kTileA);
                const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * Cfg::kTileA), b_lo = make_kmajor_sw128_desc(st + 2 * Cfg::kTileA + Cfg::kTileB);
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k) {
                    const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);      // +32 B per K step inside the swizzle atom
                    umma_tf32(d_acc, a_hi + adv, b_hi + adv, idesc, (ic | k) != 0);
                    umma_tf32(d_acc, a_hi + adv, b_lo + adv, idesc, 1);
                    umma_tf32(d_acc, a_lo + adv, b_hi + adv, idesc, 1);
                }
                umma_commit(empty_bar(s));                 // frees the stage when these MMAs retire
                if (ic == kChainKB - 1 || i == nkb - 1) umma_commit(acc_full(b));      // chain complete
            }
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4; running sums of the chunks in registers
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        float sum[BN];
#pragma unroll
        for (int i = 0; i < BN; ++i) sum[i] = 0.f;
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int b = chunk & 1;
            mbar_wait(acc_full(b), (chunk >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < BN; c += 32) {
                float acc[32];
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * Cfg::kAccCols + c), acc);
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[c + i] += acc[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(b));
        }
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = sum[c + i];
            epi(row, n0 + c, acc, (int)blockIdx.z);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, Cfg::kTmemCols);
}

}  // namespace tc
}  // namespace ocbnn
