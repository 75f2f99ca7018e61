// probe.cu — instruction-roofline probes measured in the same run as the benchmark (MEASURED_PEAKS.json carries only
// HBM and bf16 tensor peaks; the toy-shape kernels are bound by the fp32 FMA and MUFU pipes).
#include "common.cuh"

namespace ocbnn {

// mode 0: FFMA (scalar fp32 fma) ; mode 1: FFMA2 (packed f32x2 fma, sm_100) ; mode 2: MUFU.EX2
template <int MODE>
__global__ void __launch_bounds__(256) k_probe(int iters, float* sink) {
    const float s = 1.0f + 1e-7f * threadIdx.x;
    if (MODE == 0) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], s, 0.25f);
        }
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += a[i];
        if (r == 12345.678f) sink[0] = r;
    } else if (MODE == 1) {
        float2 a[8];
        const float2 s2 = make_float2(s, s), c2 = make_float2(0.25f, 0.125f);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = make_float2(0.001f * (threadIdx.x + i), 0.002f * (threadIdx.x + i));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], s2, c2);
        }
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
        if (r == 12345.678f) sink[0] = r;
    } else {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = exp2f(-a[i]);     // MUFU.EX2 (+ the negate folded into the operand)
        }
        float r = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) r += a[i];
        if (r == 12345.678f) sink[0] = r;
    }
}

}  // namespace ocbnn

using namespace ocbnn;

// Runs probe `mode` and returns operations per second in *out_rate: fp32 FLOP/s (fma = 2) for modes 0/1, MUFU op/s for mode 2.
extern "C" OCBNN_API int ocbnn_probe_peak(int32_t mode, double* out_rate) {
    OCBNN_REQUIRE(out_rate && mode >= 0 && mode <= 2, OCBNN_EINVAL, "ocbnn_probe_peak: bad argument");
    int dev = 0, sms = 0;
    OCBNN_CUDA(cudaGetDevice(&dev));
    OCBNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* sink = nullptr;
    OCBNN_CUDA(cudaMalloc(&sink, sizeof(float)));
    cudaEvent_t e0, e1;
    OCBNN_CUDA(cudaEventCreate(&e0));
    OCBNN_CUDA(cudaEventCreate(&e1));
    const int iters = 20000, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        OCBNN_CUDA(cudaEventRecord(e0));
        if (mode == 0) k_probe<0><<<blocks, threads>>>(iters, sink);
        else if (mode == 1) k_probe<1><<<blocks, threads>>>(iters, sink);
        else k_probe<2><<<blocks, threads>>>(iters, sink);
        OCBNN_LAUNCH_CHECK();
        OCBNN_CUDA(cudaEventRecord(e1));
        OCBNN_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        OCBNN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double ops = (double)blocks * threads * iters * 8.0 * (mode == 0 ? 2.0 : mode == 1 ? 4.0 : 1.0);
        double rate = ops / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *out_rate = best;
    return OCBNN_OK;
}
