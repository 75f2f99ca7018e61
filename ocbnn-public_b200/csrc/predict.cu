// predict.cu — output heads of predict() and the posterior-sample statistics built on them, in the library.
//
//   ocbnn_predict          forward + softmax / sigmoid / argmax / mean over samples
//                          (RegressorMixin.predict base.py:267-277, ClassifierMixin.predict :446-461, BinaryClassifierMixin.predict :544-559)
//   ocbnn_violation_flags  per posterior sample: does the prediction enter (lo, hi) anywhere on the domain
//                          (the rejection count of run_toys.py:229-250, Figure 3c)
//
// HBM-bound elementwise passes over the (M, P, s_K) forward output: one read + one write of it, the mean one more read.
#include "common.cuh"

namespace ocbnn {

// one thread per (sample, point) row of s_K outputs; in place: raw -> probabilities; classes written beside them
template <bool BINARY>
__global__ void __launch_bounds__(256)
k_predict_head(float* __restrict__ f, long long rows, int sk, int write_probs, int* __restrict__ out_class) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        float* v = f + r * sk;
        if (BINARY) {
            // sigmoid, branch-stable on both tails
            float z = v[0];
            float e = expf(-fabsf(z));
            float p = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
            if (write_probs) v[0] = p;
            if (out_class) out_class[r] = p >= 0.5f;
        } else {
            float mx = v[0];
            int arg = 0;
            for (int k = 1; k < sk; ++k) {
                float t = v[k];
                if (t > mx) { mx = t; arg = k; }          // first maximum wins, as torch.argmax
            }
            if (write_probs) {
                float s = 0.f;
                for (int k = 0; k < sk; ++k) s += expf(v[k] - mx);
                float inv = 1.f / s;
                for (int k = 0; k < sk; ++k) v[k] = expf(v[k] - mx) * inv;
            }
            if (out_class) out_class[r] = arg;
        }
    }
}

// mean over the M samples of f[m][c], c < cols (= P * s_K): consecutive threads read consecutive columns
__global__ void __launch_bounds__(256)
k_predict_mean(const float* __restrict__ f, long long m, long long cols, float* __restrict__ out) {
    long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    long long i = 0;
    for (; i + 4 <= m; i += 4) {
        s0 += f[(i + 0) * cols + c];
        s1 += f[(i + 1) * cols + c];
        s2 += f[(i + 2) * cols + c];
        s3 += f[(i + 3) * cols + c];
    }
    for (; i < m; ++i) s0 += f[i * cols + c];
    out[c] = ((s0 + s1) + (s2 + s3)) / (float)m;
}

// one warp per sample: any point whose first output lies strictly inside (lo, hi)
__global__ void __launch_bounds__(256)
k_violation_flags(const float* __restrict__ f, long long m, long long npts, int sk, float lo, float hi, uint8_t* __restrict__ flags) {
    long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= m) return;
    const float* row = f + w * npts * sk;
    int hit = 0;
    for (long long p = lane; p < npts; p += 32) {
        float v = row[p * sk];
        hit |= (v > lo && v < hi);
    }
    hit = __any_sync(0xffffffffu, hit);
    if (lane == 0) flags[w] = (uint8_t)hit;
}

static int forward_any(const Problem& p, const float* w, long long m, const float* x, long long npts, float* out, cudaStream_t st) {
    if (small_net_supported(p)) return launch_forward_small(p, w, m, x, npts, out, st);
    return launch_forward_generic(p, w, m, x, npts, out, st);
}

}  // namespace ocbnn

using namespace ocbnn;

extern "C" OCBNN_API int ocbnn_predict(ocbnn_problem* h, const float* w, int64_t m, const float* x, int64_t npoints, int32_t head,
                                       float* out, int32_t* out_class, float* out_mean, void* stream) {
    OCBNN_REQUIRE(h && w && x && out && m >= 0 && npoints >= 0, OCBNN_EINVAL, "ocbnn_predict: bad argument");
    OCBNN_REQUIRE(head == OCBNN_PREDICT_RAW || head == OCBNN_PREDICT_PROBS || head == OCBNN_PREDICT_CLASS, OCBNN_EINVAL, "ocbnn_predict: unknown head");
    if (m == 0 || npoints == 0) return OCBNN_OK;
    const Problem& p = *problem_of(h);
    const int sk = p.desc.layer_sizes[p.desc.n_layers];
    const int lik = p.desc.lik_head;
    OCBNN_REQUIRE(head == OCBNN_PREDICT_RAW || lik == OCBNN_HEAD_LIK_SOFTMAX || lik == OCBNN_HEAD_LIK_BERNOULLI, OCBNN_EINVAL,
                  "ocbnn_predict: probabilities / classes need a classifier problem");
    OCBNN_REQUIRE(head != OCBNN_PREDICT_CLASS || out_class, OCBNN_EINVAL, "ocbnn_predict: OCBNN_PREDICT_CLASS needs out_class");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = forward_any(p, w, m, x, npoints, out, st);
    if (rc) return rc;
    const long long rows = (long long)m * npoints;
    if (head != OCBNN_PREDICT_RAW) {
        const int write_probs = head == OCBNN_PREDICT_PROBS;
        unsigned grid = (unsigned)((rows + 255) / 256 < (long long)p.sm_count * 16 ? (rows + 255) / 256 : (long long)p.sm_count * 16);
        if (lik == OCBNN_HEAD_LIK_BERNOULLI) k_predict_head<true><<<grid, 256, 0, st>>>(out, rows, sk, write_probs, out_class);
        else k_predict_head<false><<<grid, 256, 0, st>>>(out, rows, sk, write_probs, out_class);
        OCBNN_LAUNCH_CHECK();
    }
    if (out_mean) {
        OCBNN_REQUIRE(head != OCBNN_PREDICT_CLASS, OCBNN_EINVAL, "ocbnn_predict: out_mean is the mean of the outputs or probabilities, not of classes");
        const long long cols = npoints * sk;
        k_predict_mean<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>(out, m, cols, out_mean);
        OCBNN_LAUNCH_CHECK();
    }
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_violation_flags(ocbnn_problem* h, const float* w, int64_t m, const float* x, int64_t npoints, float lo, float hi,
                                               float* scratch, uint8_t* out_flags, void* stream) {
    OCBNN_REQUIRE(h && w && (x || npoints == 0) && scratch && out_flags && m >= 0 && npoints >= 0, OCBNN_EINVAL, "ocbnn_violation_flags: bad argument");
    if (m == 0) return OCBNN_OK;
    const Problem& p = *problem_of(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (npoints == 0) {
        OCBNN_CUDA(cudaMemsetAsync(out_flags, 0, (size_t)m, st));
        return OCBNN_OK;
    }
    int rc = forward_any(p, w, m, x, npoints, scratch, st);
    if (rc) return rc;
    k_violation_flags<<<(unsigned)((m * 32 + 255) / 256), 256, 0, st>>>(scratch, m, npoints, p.desc.layer_sizes[p.desc.n_layers], lo, hi, out_flags);
    OCBNN_LAUNCH_CHECK();
    return OCBNN_OK;
}
