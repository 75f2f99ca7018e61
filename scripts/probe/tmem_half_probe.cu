// Probe (GPU box): (1) an M = 64 tcgen05.mma accumulator placed at TMEM lane offset 16 (upper half of every subpartition);
// (2) the thread <-> element mapping of tcgen05.ld.16x256b.x1 at lane offsets 0 and 16.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe/tmem_half_probe scripts/probe/tmem_half_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ inline int ct_off(int r, int c, int RG, int CG) { return ((r >> 3) * RG + (c >> 3) * CG + (r & 7) * 16 + (c & 7) * 2) >> 1; }

// A: 64 x 16 (K-major, core-tiled), B: 16 x 16 (K-major)  ->  D (64 x 16) at lane offset `lane_off`
__global__ void __launch_bounds__(128) k_probe(const __nv_bfloat16* __restrict__ abuf, const __nv_bfloat16* __restrict__ bbuf, int lane_off,
                                               float* __restrict__ out32 /* 128 x 16: 32x32b readout */, float* __restrict__ out16 /* 128 thr x 4 regs x 2 offsets */) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(sm);
    __nv_bfloat16* B = A + 64 * 16;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < 64 * 16; i += 128) A[i] = abuf[i];
    for (int i = threadIdx.x; i < 16 * 16; i += 128) B[i] = bbuf[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < 32; c0 += 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0), "r"(0x7fc00000u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
        uint64_t da = make_desc(smem_u32(A), 8 * 128, 128);      // A: row groups contiguous (128 B), column cores 8 groups x 128 B apart
        uint64_t db = make_desc(smem_u32(B), 2 * 128, 128);      // B: 16 rows = 2 groups
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + ((uint32_t)lane_off << 16)),
                     "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    {
        uint32_t done = 0;
        for (uint32_t spin = 0; spin < (1u << 26) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
        if (!done) __trap();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 16; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out32[threadIdx.x * 16 + c0 + i] = __uint_as_float(r[i]);
    }
    for (int off = 0; off < 2; ++off) {
        uint32_t r[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(tmem + ((uint32_t)(warp * 32 + off * 16) << 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out16[(threadIdx.x * 2 + off) * 4 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

int main() {
    std::vector<float> a(64 * 16), b(16 * 16);
    std::vector<__nv_bfloat16> abuf(64 * 16), bbuf(16 * 16);
    // A[i][k] = (k == 0 ? i : 0) ... D[i][j] = sum_k A[i][k] B[j][k]; choose B[j][0] = 1, B[j][1] = j, A[i][0] = i, A[i][1] = 1/... -> D[i][j] = i + 100 j
    for (int i = 0; i < 64; ++i) for (int k = 0; k < 16; ++k) { a[i * 16 + k] = k == 0 ? (float)i : (k == 1 ? 1.f : 0.f); abuf[ct_off(i, k, 128, 8 * 128)] = __float2bfloat16(a[i * 16 + k]); }
    for (int j = 0; j < 16; ++j) for (int k = 0; k < 16; ++k) { b[j * 16 + k] = k == 0 ? 1.f : (k == 1 ? 100.f * j : 0.f); bbuf[ct_off(j, k, 128, 2 * 128)] = __float2bfloat16(b[j * 16 + k]); }
    __nv_bfloat16 *da, *db; float *d32, *d16;
    cudaMalloc(&da, abuf.size() * 2); cudaMalloc(&db, bbuf.size() * 2); cudaMalloc(&d32, 128 * 16 * 4); cudaMalloc(&d16, 128 * 2 * 4 * 4);
    cudaMemcpy(da, abuf.data(), abuf.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bbuf.data(), bbuf.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    for (int lane_off : {0, 16}) {
        k_probe<<<1, 128, 16384>>>(da, db, lane_off, d32, d16);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lane_off %d: CUDA error %s\n", lane_off, cudaGetErrorString(e)); return 1; }
        std::vector<float> o32(128 * 16), o16(128 * 8);
        cudaMemcpy(o32.data(), d32, o32.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(o16.data(), d16, o16.size() * 4, cudaMemcpyDeviceToHost);
        printf("=== D (value = row + 100 col) with the MMA's D address at lane offset %d; 32x32b readout: thread -> value of column 0 and 1\n", lane_off);
        for (int t = 0; t < 128; ++t) { printf("%s%3d:%5.0f,%5.0f ", t % 8 == 0 ? "\n  " : "", t, o32[t * 16], o32[t * 16 + 1]); }
        printf("\n--- 16x256b.x1 readout (warp 0 and warp 1 shown): thread -> 4 regs at ld lane offset 0 | 16\n");
        for (int t = 0; t < 64; ++t) {
            printf("  t%2d: [%5.0f %5.0f %5.0f %5.0f] | [%5.0f %5.0f %5.0f %5.0f]%s", t, o16[(t * 2) * 4], o16[(t * 2) * 4 + 1], o16[(t * 2) * 4 + 2], o16[(t * 2) * 4 + 3],
                   o16[(t * 2 + 1) * 4], o16[(t * 2 + 1) * 4 + 1], o16[(t * 2 + 1) * 4 + 2], o16[(t * 2 + 1) * 4 + 3], t % 2 ? "\n" : "");
        }
    }
    return 0;
}
