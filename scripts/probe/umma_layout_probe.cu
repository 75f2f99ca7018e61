// Probe (GPU box): which no-swizzle shared-memory descriptor conventions tcgen05.mma kind::tf32 accepts for K-major and MN-major
// operands stored as "core-tiled" arrays (8 rows x 16 bytes per core matrix), and the TMEM row mapping of M = 64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe/umma_layout_probe scripts/probe/umma_layout_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;     // layout type 0 = no swizzle
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// core-tiled element offset (floats): rows x cols array, row-group stride RG bytes, col-group stride CG bytes
__host__ __device__ inline int ct_off(int r, int c, int RG, int CG) { return ((r >> 3) * RG + (c >> 2) * CG + (r & 7) * 16 + (c & 3) * 4) >> 2; }

struct Case {
    int m, n, k;          // MMA M, N and total K
    int a_mn, b_mn;       // majors
    int a_lbo, a_sbo, a_step;   // bytes; step = start-address advance per K = 8
    int b_lbo, b_sbo, b_step;
};

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ abuf, const float* __restrict__ bbuf, int afloats, int bfloats, Case cs,
                                               float* __restrict__ out /* 128 lanes x 256 cols */) {
    extern __shared__ __align__(1024) uint8_t sm[];
    float* A = reinterpret_cast<float*>(sm);
    float* B = A + afloats;
    uint64_t* bar = reinterpret_cast<uint64_t*>(B + bfloats);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < afloats; i += 128) A[i] = abuf[i];
    for (int i = threadIdx.x; i < bfloats; i += 128) B[i] = bbuf[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    // zero the accumulator region first through a dummy? not needed: first MMA overwrites (acc = 0)
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(cs.m, cs.n, cs.a_mn, cs.b_mn);
        for (int ks = 0; ks < cs.k / 8; ++ks) {
            uint64_t da = make_desc(smem_u32(A) + ks * cs.a_step, cs.a_lbo, cs.a_sbo);
            uint64_t db = make_desc(smem_u32(B) + ks * cs.b_step, cs.b_lbo, cs.b_sbo);
            umma(tmem, da, db, idesc, ks != 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    {
        uint32_t done = 0;
        for (uint32_t spin = 0; spin < (1u << 26) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
        if (!done) __trap();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < cs.n; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[threadIdx.x * 256 + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

static float val(int i, int j, int salt) {
    uint32_t h = (uint32_t)i * 1315423911u ^ (uint32_t)j * 2654435761u ^ (uint32_t)salt * 0x9E3779B9u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return (float)((int)(h % 15u) - 7);
}

// run one case; X is the logical A (m x k) and Y the logical B (n x k); storage arrays given
static const char* g_only = getenv("PROBE_ONLY");
static double run(const char* name, Case cs, const std::vector<float>& abuf, const std::vector<float>& bbuf, const std::vector<float>& X,
                  const std::vector<float>& Y, bool m64_rows) {
    if (g_only && !strstr(name, g_only)) return 0;
    float *da, *db, *dout;
    cudaMalloc(&da, abuf.size() * 4); cudaMalloc(&db, bbuf.size() * 4); cudaMalloc(&dout, 128 * 256 * 4);
    cudaMemcpy(da, abuf.data(), abuf.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bbuf.data(), bbuf.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 128 * 256 * 4);
    size_t smem = 200 * 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, 128, smem>>>(da, db, (int)abuf.size(), (int)bbuf.size(), cs, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> out(128 * 256);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad = 0, badrow = 0, badcol = 0, shown = 0;
    std::vector<int> rb(cs.m, 0), cb(cs.n, 0);
    for (int i = 0; i < cs.m; ++i)
        for (int j = 0; j < cs.n; ++j) {
            double ref = 0;
            for (int k = 0; k < cs.k; ++k) ref += (double)X[i * cs.k + k] * Y[j * cs.k + k];
            int lane = (cs.m == 64 && m64_rows) ? (i % 16) + 32 * (i / 16) : i;
            double d = fabs(ref - out[lane * 256 + j]);
            if (d > worst) worst = d;
            if (d != 0) { ++bad; rb[i] = 1; cb[j] = 1; }
        }
    for (int v : rb) badrow += v;
    for (int v : cb) badcol += v;
    printf("%-72s max |err| = %g %s (bad %d of %d; rows %d/%d cols %d/%d)\n", name, worst, worst == 0 ? "OK" : "MISMATCH", bad, cs.m * cs.n, badrow, cs.m, badcol, cs.n);
    if (bad && getenv("PROBE_VERBOSE") && (badrow < cs.m || badcol < cs.n)) {
        printf("   bad rows:"); for (int i = 0; i < cs.m; ++i) if (rb[i]) printf(" %d", i); printf("\n   bad cols:"); for (int j = 0; j < cs.n; ++j) if (cb[j]) printf(" %d", j); printf("\n");
    }
    (void)shown;
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return worst;
}

int main() {
    // logical operands: P (128 points) x F features "act" array stored core-tiled [p][f]; W (KW x NW) stored core-tiled [k][n]
    const int P = 128, F = 16, KW = 16, NW = 64;
    const int ARG = 128, ACG = (P / 8) * 128;       // act: p-groups contiguous, f-groups 2048 B apart
    const int WRG = 128, WCG = (KW / 8) * 128;      // W: k-groups contiguous, n-groups (KW/8)*128 B apart
    std::vector<float> act(P * F), w(KW * NW), actbuf(P * F), wbuf(KW * NW);
    for (int p = 0; p < P; ++p) for (int f = 0; f < F; ++f) { act[p * F + f] = val(p, f, 1); actbuf[ct_off(p, f, ARG, ACG)] = act[p * F + f]; }
    for (int k = 0; k < KW; ++k) for (int n = 0; n < NW; ++n) { w[k * NW + n] = val(k, n, 4); wbuf[ct_off(k, n, WRG, WCG)] = w[k * NW + n]; }
    for (int swapA = 0; swapA < 2; ++swapA)
        for (int swapB = 0; swapB < 2; ++swapB) {
            char nm[128];
            // GEMM1  Z[p][n] = sum_f act[p][f] W[f][n] : A = act K-major (MN = p, K = f), B = W MN-major (MN = n, K = k)
            {
                Case c{128, NW, F, 0, 1, swapA ? ARG : ACG, swapA ? ACG : ARG, 2 * ACG, swapB ? WCG : WRG, swapB ? WRG : WCG, WRG};
                std::vector<float> X(128 * F), Y(NW * F);
                for (int p = 0; p < 128; ++p) for (int f = 0; f < F; ++f) X[p * F + f] = act[p * F + f];
                for (int n = 0; n < NW; ++n) for (int k = 0; k < F; ++k) Y[n * F + k] = w[k * NW + n];
                snprintf(nm, sizeof nm, "GEMM1 A K-major(lbo=%s) B MN-major(lbo=%s)", swapA ? "rowgrp" : "colgrp", swapB ? "colgrp" : "rowgrp");
                run(nm, c, actbuf, wbuf, X, Y, true);
            }
            // GEMM2  dA[p][k] = sum_n dZ[p][n] W[k][n] : A = dZ K-major (use act as dZ with F = n... here: K = NW needs a 128 x 64 array)
        }
    // second set with a 128 x 64 "dZ" array
    {
        const int F2 = 64, ACG2 = (P / 8) * 128;
        std::vector<float> dz(P * F2), dzbuf(P * F2);
        for (int p = 0; p < P; ++p) for (int f = 0; f < F2; ++f) { dz[p * F2 + f] = val(p, f, 2); dzbuf[ct_off(p, f, ARG, ACG2)] = dz[p * F2 + f]; }
        for (int swapA = 0; swapA < 2; ++swapA)
            for (int swapB = 0; swapB < 2; ++swapB) {
                char nm[128];
                // GEMM2: A = dZ K-major (MN = p, K = n), B = W K-major (MN = k (N = 16), K = n)
                Case c{128, KW, NW, 0, 0, swapA ? ARG : ACG2, swapA ? ACG2 : ARG, 2 * ACG2, swapB ? WRG : WCG, swapB ? WCG : WRG, 2 * WCG};
                std::vector<float> X(128 * NW), Y(KW * NW);
                for (int p = 0; p < 128; ++p) for (int n = 0; n < NW; ++n) X[p * NW + n] = dz[p * F2 + n];
                for (int k = 0; k < KW; ++k) for (int n = 0; n < NW; ++n) Y[k * NW + n] = w[k * NW + n];
                snprintf(nm, sizeof nm, "GEMM2 A K-major(lbo=%s) B K-major(lbo=%s)", swapA ? "rowgrp" : "colgrp", swapB ? "rowgrp" : "colgrp");
                run(nm, c, dzbuf, wbuf, X, Y, true);
            }
        // GEMM3: dW[f][n] = sum_p act2[p][f] dZ[p][n] : A = act2 MN-major (MN = f, K = p), B = dZ MN-major (MN = n, K = p); M = 64 and M = 128
        for (int M : {64, 128}) {
            const int FA = M;                    // act2 has M features
            std::vector<float> a2(P * FA), a2buf(P * FA);
            for (int p = 0; p < P; ++p) for (int f = 0; f < FA; ++f) { a2[p * FA + f] = val(p, f, 3); a2buf[ct_off(p, f, ARG, ACG2)] = a2[p * FA + f]; }
            for (int swapA = 0; swapA < 2; ++swapA)
                for (int swapB = 0; swapB < 2; ++swapB)
                    for (int rows = 0; rows < (M == 64 ? 2 : 1); ++rows) {
                        char nm[160];
                        Case c{M, F2, P, 1, 1, swapA ? ACG2 : ARG, swapA ? ARG : ACG2, ARG, swapB ? ACG2 : ARG, swapB ? ARG : ACG2, ARG};
                        std::vector<float> X(M * P), Y(F2 * P);
                        for (int f = 0; f < M; ++f) for (int p = 0; p < P; ++p) X[f * P + p] = a2[p * FA + f];
                        for (int n = 0; n < F2; ++n) for (int p = 0; p < P; ++p) Y[n * P + p] = dz[p * F2 + n];
                        snprintf(nm, sizeof nm, "GEMM3 M=%d A MN-major(lbo=%s) B MN-major(lbo=%s) rows=%s", M, swapA ? "colgrp" : "rowgrp", swapB ? "colgrp" : "rowgrp",
                                 rows || M == 128 ? "16-per-subpartition" : "linear");
                        run(nm, c, a2buf, dzbuf, X, Y, rows != 0);
                    }
        }
    }
    return 0;
}
