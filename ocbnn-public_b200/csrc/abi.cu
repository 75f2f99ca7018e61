// abi.cu — the C-ABI of libocbnn_b200.so (see include/ocbnn_b200.h): problem handles, argument checks and
// dispatch to the kernels.  No torch, no C++ types across the boundary.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <vector>
#include "common.cuh"

namespace ocbnn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return OCBNN_ECUDA + (int)e;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct ProblemImpl : Problem {
    float* scratch_f = nullptr;   // device scratch for per-step scalars (SGLD step sizes)
    int* scratch_i = nullptr;
    size_t scratch_n = 0;
};

static int ensure_scratch(ProblemImpl* p, size_t n) {
    if (n <= p->scratch_n) return OCBNN_OK;
    if (p->scratch_f) cudaFree(p->scratch_f);
    if (p->scratch_i) cudaFree(p->scratch_i);
    p->scratch_f = nullptr;
    p->scratch_i = nullptr;
    OCBNN_CUDA(cudaMalloc(&p->scratch_f, n * sizeof(float)));
    OCBNN_CUDA(cudaMalloc(&p->scratch_i, n * sizeof(int)));
    p->scratch_n = n;
    return OCBNN_OK;
}

// ---- generic-net HMC / SGLD building blocks (elementwise; the heavy part is launch_logpost_generic) ---------------------
__global__ void k_hmc_begin(const float* __restrict__ p0, float* __restrict__ p, const float* __restrict__ lp, const float* __restrict__ grad,
                            const float* __restrict__ q, float* __restrict__ q0, float* __restrict__ hbuf, long long m, int d, float eps_half) {
    // one warp per chain: p = p0 ; K0 ; U0 ; q0 = q ; p += eps/2 * g
    const long long chain = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chain >= m) return;
    double ks = 0.0;
    for (int i = lane; i < d; i += 32) {
        float pv = p0[chain * d + i];
        ks += (double)(pv * pv);
        q0[chain * d + i] = q[chain * d + i];
        p[chain * d + i] = __fadd_rn(pv, __fmul_rn(eps_half, grad[chain * d + i]));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ks += __shfl_xor_sync(0xffffffffu, ks, off);
    if (lane == 0) {
        hbuf[chain * 4 + 0] = -lp[chain];
        hbuf[chain * 4 + 1] = 0.5f * (float)ks;
    }
}
__global__ void k_axpy_rn(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = __fadd_rn(y[i], __fmul_rn(a, x[i]));
}
__global__ void k_hmc_finish(const float* __restrict__ p, const float* __restrict__ lp, float* __restrict__ q, const float* __restrict__ q0,
                             float* __restrict__ hbuf, const double* __restrict__ u, long long m, int d, int restore,
                             uint8_t* __restrict__ out_acc, float* __restrict__ out_h, int* __restrict__ flags,
                             float* __restrict__ sample_dst) {
    const long long chain = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (chain >= m) return;
    double ks = 0.0;
    bool isn = false;
    for (int i = lane; i < d; i += 32) {
        float pv = p[chain * d + i];
        ks += (double)(pv * pv);
        isn |= isnan(q[chain * d + i]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ks += __shfl_xor_sync(0xffffffffu, ks, off);
    isn = __any_sync(0xffffffffu, isn);
    const float U0 = hbuf[chain * 4 + 0], K0 = hbuf[chain * 4 + 1];
    const float U1 = -lp[chain], K1 = 0.5f * (float)ks;
    const float dH = __fsub_rn(__fadd_rn(__fsub_rn(U0, U1), K0), K1);
    const bool acc = u[chain] < (double)expf(dH);
    if (lane == 0) {
        if (out_acc) out_acc[chain] = acc ? 1 : 0;
        if (out_h) reinterpret_cast<float4*>(out_h)[chain] = make_float4(U0, K0, U1, K1);
        if (flags && isn) flags[chain] |= 1;
    }
    if (restore && !acc)
        for (int i = lane; i < d; i += 32) q[chain * d + i] = q0[chain * d + i];
    if (sample_dst)
        for (int i = lane; i < d; i += 32) sample_dst[chain * d + i] = q[chain * d + i];
}
__global__ void k_sgld_update(float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ noise, float he, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) w[i] = __fadd_rn(w[i], __fadd_rn(__fmul_rn(g[i], he), noise[i]));
}
__global__ void k_zero_i32(int* p, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

static unsigned ew_blocks(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

EvalArgs eval_args(const Problem& p, int offset, int stride, int with_data) {
    EvalArgs a = p.base;
    a.batch_offset = stride > 1 ? offset : 0;
    a.batch_stride = stride > 1 ? stride : 1;
    a.with_data = with_data ? 1 : 0;
    return a;
}

int logpost_any(const Problem& p, const EvalArgs& a, const float* w, long long m, float* lp, float* grad, int lanes, cudaStream_t st) {
    if (small_net_supported(p)) {
        int pts = (a.with_data ? batch_count(a.n_train, 0, a.batch_stride) : 0) + a.n_cpoints;
        if (lanes <= 0) lanes = choose_lanes(p, m, pts);
        return launch_logpost_small(p, a, w, m, lp, grad, lanes, st);
    }
    return launch_logpost_generic(p, a, w, m, lp, grad, st);
}

const Problem* problem_of(const ocbnn_problem* h) { return reinterpret_cast<const ProblemImpl*>(h); }

}  // namespace ocbnn

using namespace ocbnn;

extern "C" OCBNN_API const char* ocbnn_last_error(void) { return g_err; }
extern "C" OCBNN_API int ocbnn_abi_version(void) { return OCBNN_ABI_VERSION; }
extern "C" OCBNN_API int64_t ocbnn_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" OCBNN_API int ocbnn_problem_create(const ocbnn_desc* desc, ocbnn_problem** out) {
    OCBNN_REQUIRE(desc && out, OCBNN_EINVAL, "ocbnn_problem_create: NULL argument");
    OCBNN_REQUIRE(desc->abi_version == OCBNN_ABI_VERSION, OCBNN_EINVAL, "ABI version mismatch: header %d, caller %d", OCBNN_ABI_VERSION,
                  desc->abi_version);
    OCBNN_REQUIRE(desc->n_layers >= 1 && desc->n_layers <= OCBNN_MAX_LAYERS, OCBNN_EINVAL, "n_layers must be in [1,%d]", OCBNN_MAX_LAYERS);
    for (int l = 0; l <= desc->n_layers; ++l) OCBNN_REQUIRE(desc->layer_sizes[l] >= 1, OCBNN_EINVAL, "layer_sizes[%d] must be >= 1", l);
    OCBNN_REQUIRE(desc->n_train >= 0 && desc->n_cpoints >= 0, OCBNN_EINVAL, "negative point count");
    OCBNN_REQUIRE(desc->n_train == 0 || (desc->x_train && desc->y_train && desc->y_width >= 1), OCBNN_EINVAL, "training data missing");
    OCBNN_REQUIRE(desc->n_cpoints == 0 || (desc->cx && desc->ckind && desc->cweight && (desc->cparam || desc->cparam_stride == 0)),
                  OCBNN_EINVAL, "constraint tables missing");
    OCBNN_REQUIRE((desc->prior_mean_vec == nullptr) == (desc->prior_std_vec == nullptr), OCBNN_EINVAL,
                  "prior_mean_vec and prior_std_vec must be given together");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: the OC-BNN engine has no CPU fallback");
        return OCBNN_ENODEVICE;
    }
    int dev = 0;
    OCBNN_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    OCBNN_CUDA(cudaGetDeviceProperties(&prop, dev));
    OCBNN_REQUIRE(prop.major == 10, OCBNN_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, prop.major,
                  prop.minor);

    ProblemImpl* p = new ProblemImpl();
    p->desc = *desc;
    p->sm_count = prop.multiProcessorCount;
    p->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    const int s0 = desc->layer_sizes[0];
    int d = 0;
    for (int l = 1; l <= desc->n_layers; ++l) d += (desc->layer_sizes[l - 1] + 1) * desc->layer_sizes[l];
    p->d = d;
    int np = desc->y_width > desc->cparam_stride + 1 ? desc->y_width : desc->cparam_stride + 1;   // +1: internal NEG_EXP layout
    if (np < 2) np = 2;
    int rs = s0 + 2 + np;
    if ((rs & 1) == 0) ++rs;
    p->rs = rs;
    p->np = np;

    const long long rows = desc->n_train + desc->n_cpoints;
    std::vector<float> pts((size_t)(rows > 0 ? rows : 1) * rs, 0.f);
    for (long long n = 0; n < desc->n_train; ++n) {
        float* r = &pts[(size_t)n * rs];
        for (int k = 0; k < s0; ++k) r[k] = desc->x_train[n * s0 + k];
        int kind = desc->lik_head;
        memcpy(&r[s0], &kind, sizeof(int));
        r[s0 + 1] = 1.f;
        for (int k = 0; k < desc->y_width; ++k) r[s0 + 2 + k] = desc->y_train[n * desc->y_width + k];
    }
    // constraint rows grouped by head kind (positive | negative | Dirichlet) so that the kernels run one compile-time
    // head per segment; the order inside a group is the caller's
    int seg_n[3] = {0, 0, 0};
    int neg_n1 = -2, neg_n2 = -2;       // -2: no negative row seen yet, -1: rows differ
    {
        const int order[3] = {OCBNN_HEAD_POS_GAUSS, OCBNN_HEAD_NEG_EXP, OCBNN_HEAD_DIRICHLET};
        long long out_row = desc->n_train;
        for (int sgi = 0; sgi < 3; ++sgi) {
            for (long long c = 0; c < desc->n_cpoints; ++c) {
                if (desc->ckind[c] != order[sgi]) continue;
                float* r = &pts[(size_t)out_row * rs];
                for (int k = 0; k < s0; ++k) r[k] = desc->cx[c * s0 + k];
                int kind = desc->ckind[c];
                memcpy(&r[s0], &kind, sizeof(int));
                r[s0 + 1] = desc->cweight[c];
                const float* src = desc->cparam_stride ? desc->cparam + c * desc->cparam_stride : nullptr;
                if (kind == OCBNN_HEAD_NEG_EXP) {
                    // ABI: ngaps, (a,b)...  ->  internal: n1, n2, one-sided (s c, s) terms Phi(s f - s c), two-sided (a, b) terms
                    const int ngaps = src ? (int)src[0] : 0;
                    if (ngaps < 0 || 1 + 2 * ngaps > desc->cparam_stride) {
                        delete p;
                        set_error("constraint point %lld: gap count %d does not fit cparam_stride %d", c, ngaps, desc->cparam_stride);
                        return OCBNN_EINVAL;
                    }
                    float* dst = &r[s0 + 2];
                    int n1 = 0, n2 = 0;
                    float* w1 = dst + 2;
                    for (int k = 0; k < ngaps; ++k) {
                        const float a = src[1 + 2 * k], b = src[2 + 2 * k];
                        if (std::isinf(a) && a < 0) { w1[2 * n1] = b; w1[2 * n1 + 1] = 1.f; ++n1; }         // Phi(f - b)  = Phi(s f - s c), (s c, s) = (b, 1)
                        else if (std::isinf(b) && b > 0) { w1[2 * n1] = -a; w1[2 * n1 + 1] = -1.f; ++n1; }  // Phi(a - f)  = Phi(s f - s c), (s c, s) = (-a, -1)
                    }
                    float* w2 = dst + 2 + 2 * n1;
                    for (int k = 0; k < ngaps; ++k) {
                        const float a = src[1 + 2 * k], b = src[2 + 2 * k];
                        if (!(std::isinf(a) && a < 0) && !(std::isinf(b) && b > 0)) { w2[2 * n2] = a; w2[2 * n2 + 1] = b; ++n2; }
                    }
                    memcpy(&dst[0], &n1, sizeof(int));
                    memcpy(&dst[1], &n2, sizeof(int));
                    if (neg_n1 == -2) { neg_n1 = n1; neg_n2 = n2; }
                    else if (neg_n1 != n1 || neg_n2 != n2) neg_n1 = neg_n2 = -1;
                } else {
                    for (int k = 0; k < desc->cparam_stride; ++k) r[s0 + 2 + k] = src[k];
                }
                ++out_row;
                ++seg_n[sgi];
            }
        }
        if (out_row != rows) {
            delete p;
            set_error("ckind must be OCBNN_HEAD_POS_GAUSS, OCBNN_HEAD_NEG_EXP or OCBNN_HEAD_DIRICHLET for every constraint point");
            return OCBNN_EINVAL;
        }
    }
    std::vector<float2> pri((size_t)d);
    for (int i = 0; i < d; ++i) {
        float mu = desc->prior_mean_vec ? desc->prior_mean_vec[i] : desc->prior_mean;
        float sd = desc->prior_std_vec ? desc->prior_std_vec[i] : desc->prior_std;
        pri[i] = make_float2(mu, 1.f / (sd * sd));
    }
    cudaError_t e = cudaMalloc(&p->pts, pts.size() * sizeof(float));
    // the prior table is followed by a second one of zeros (mean 0, 1/var 0): what a likelihood-only evaluation points at
    if (e == cudaSuccess) e = cudaMalloc(&p->pri, 2 * pri.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemset(p->pri, 0, 2 * pri.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMemcpy(p->pts, pts.data(), pts.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->pri, pri.data(), pri.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (p->pts) cudaFree(p->pts);
        if (p->pri) cudaFree(p->pri);
        delete p;
        return cuda_fail(e, "problem upload", __FILE__, __LINE__);
    }
    // host pointers are not kept
    p->desc.prior_mean_vec = p->desc.prior_std_vec = nullptr;
    p->desc.x_train = p->desc.y_train = p->desc.cx = p->desc.cweight = p->desc.cparam = nullptr;
    p->desc.ckind = nullptr;

    EvalArgs& a = p->base;
    a.pts = p->pts;
    a.pri = p->pri;
    a.n_train = desc->n_train;
    a.n_cpoints = desc->n_cpoints;
    a.n_pos = seg_n[0];
    a.n_neg = seg_n[1];
    a.n_dir = seg_n[2];
    a.neg_n1 = neg_n1 < 0 ? -1 : neg_n1;
    a.neg_n2 = neg_n2 < 0 ? -1 : neg_n2;
    a.lik_head = desc->lik_head;
    a.rs = rs;
    a.s0 = s0;
    a.batch_offset = 0;
    a.batch_stride = 1;
    a.with_data = 1;
    a.const_prior = desc->prior_const;
    a.const_lik = desc->lik_const;
    a.lik_scale = desc->lik_scale;
    a.tau0 = desc->neg_tau0;
    a.tau1 = desc->neg_tau1;
    a.pos_inv_var = desc->pos_inv_var;
    *out = reinterpret_cast<ocbnn_problem*>(p);
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_problem_destroy(ocbnn_problem* h) {
    if (!h) return OCBNN_OK;
    ProblemImpl* p = reinterpret_cast<ProblemImpl*>(h);
    if (p->pts) cudaFree(p->pts);
    if (p->pri) cudaFree(p->pri);
    if (p->scratch_f) cudaFree(p->scratch_f);
    if (p->scratch_i) cudaFree(p->scratch_i);
    delete p;
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_problem_nweights(const ocbnn_problem* h, int64_t* out_d) {
    OCBNN_REQUIRE(h && out_d, OCBNN_EINVAL, "ocbnn_problem_nweights: NULL argument");
    *out_d = reinterpret_cast<const ProblemImpl*>(h)->d;
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_logpost_grad(ocbnn_problem* h, const float* w, int64_t m, int32_t batch_offset, int32_t batch_stride, int32_t with_data,
                                  float* out_lp, float* out_grad, int32_t lanes, void* stream) {
    OCBNN_REQUIRE(h && w && m >= 0, OCBNN_EINVAL, "ocbnn_logpost_grad: bad argument");
    if (m == 0) return OCBNN_OK;
    const ProblemImpl& p = *reinterpret_cast<ProblemImpl*>(h);
    OCBNN_REQUIRE(batch_stride <= 1 || (batch_offset >= 0 && batch_offset < batch_stride), OCBNN_EINVAL, "batch_offset must be in [0, stride)");
    EvalArgs a = eval_args(p, batch_offset, batch_stride, with_data);
    if (with_data == 2) {       // likelihood only: no constraint rows, no prior constants, the all-zero prior table
        a.n_cpoints = a.n_pos = a.n_neg = a.n_dir = 0;
        a.const_prior = 0.0;
        a.pri = p.pri + p.d;
    }
    return logpost_any(p, a, w, m, out_lp, out_grad, lanes, (cudaStream_t)stream);
}

extern "C" OCBNN_API int ocbnn_forward(ocbnn_problem* h, const float* w, int64_t m, const float* x, int64_t npoints, float* out, void* stream) {
    OCBNN_REQUIRE(h && w && x && out && m >= 0 && npoints >= 0, OCBNN_EINVAL, "ocbnn_forward: bad argument");
    if (m == 0 || npoints == 0) return OCBNN_OK;
    const ProblemImpl& p = *reinterpret_cast<ProblemImpl*>(h);
    if (small_net_supported(p)) return launch_forward_small(p, w, m, x, npoints, out, (cudaStream_t)stream);
    return launch_forward_generic(p, w, m, x, npoints, out, (cudaStream_t)stream);
}

extern "C" OCBNN_API int ocbnn_hmc_iterate(ocbnn_problem* h, float* q, const float* p0, const double* u, int64_t m, int32_t n_it, double eps,
                                 int32_t leapfrog_l, int32_t with_data, int32_t restore_on_reject, uint8_t* out_accept, float* out_h,
                                 int32_t* out_flags, float* out_samples, int32_t thin, int32_t lanes, void* stream) {
    OCBNN_REQUIRE(h && q && p0 && u && m >= 0 && n_it >= 0 && leapfrog_l >= 1, OCBNN_EINVAL, "ocbnn_hmc_iterate: bad argument");
    if (m == 0 || n_it == 0) return OCBNN_OK;
    ProblemImpl& p = *reinterpret_cast<ProblemImpl*>(h);
    cudaStream_t st = (cudaStream_t)stream;
    EvalArgs a = eval_args(p, 0, 1, with_data);
    const float e = (float)eps, eh = (float)(eps / 2);
    if (small_net_supported(p)) {
        int pts = (a.with_data ? (int)a.n_train : 0) + a.n_cpoints;
        if (lanes <= 0) lanes = choose_lanes_hmc(p, m, pts);
        return launch_hmc_small(p, a, q, p0, u, m, n_it, e, eh, leapfrog_l, restore_on_reject, out_accept, out_h, out_flags, out_samples,
                                thin, lanes, st);
    }
    // generic nets: K1 (one CTA per chain) + elementwise leapfrog kernels, driven from the host on the caller's stream
    const int d = p.d;
    const long long md = m * d;
    // one stream-ordered allocation for every temporary (nothing to unwind if it fails): p | grad | q0 | lp | (U0,K0,U1,K1)
    float* tmp = nullptr;
    const long long mda = (md + 63) & ~63LL, ma = (m + 63) & ~63LL;      // every segment starts 256-byte aligned
    OCBNN_CUDA(cudaMallocAsync(&tmp, (size_t)(3 * mda + ma + 4 * m) * sizeof(float), st));
    float *pbuf = tmp, *gbuf = tmp + mda, *q0 = tmp + 2 * mda, *lp = tmp + 3 * mda, *hbuf = tmp + 3 * mda + ma;
    const unsigned wb = (unsigned)((m * 32 + 255) / 256);
    int rc = OCBNN_OK;
    if (out_flags) { k_zero_i32<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(out_flags, m); count_launch(); }
    rc = launch_logpost_generic(p, a, q, m, lp, gbuf, st);
    for (int it = 0; it < n_it && rc == OCBNN_OK; ++it) {
        if (restore_on_reject && it > 0) rc = launch_logpost_generic(p, a, q, m, lp, gbuf, st);   // state may have been restored
        if (rc) break;
        k_hmc_begin<<<wb, 256, 0, st>>>(p0 + (long long)it * md, pbuf, lp, gbuf, q, q0, hbuf, m, d, eh);
        count_launch();
        for (int l = 1; l <= leapfrog_l && rc == OCBNN_OK; ++l) {
            k_axpy_rn<<<ew_blocks(md), 256, 0, st>>>(q, pbuf, e, md);
            count_launch();
            rc = launch_logpost_generic(p, a, q, m, lp, gbuf, st);
            if (rc) break;
            k_axpy_rn<<<ew_blocks(md), 256, 0, st>>>(pbuf, gbuf, l < leapfrog_l ? e : eh, md);
            count_launch();
        }
        if (rc) break;
        float* dst = (out_samples && thin > 0 && (it + 1) % thin == 0) ? out_samples + (long long)((it + 1) / thin - 1) * md : nullptr;
        k_hmc_finish<<<wb, 256, 0, st>>>(pbuf, lp, q, q0, hbuf, u + (long long)it * m, m, d, restore_on_reject,
                                         out_accept ? out_accept + (long long)it * m : nullptr, out_h ? out_h + (long long)it * m * 4 : nullptr,
                                         out_flags, dst);
        count_launch();
    }
    cudaFreeAsync(tmp, st);
    if (rc) return rc;
    OCBNN_CUDA(cudaGetLastError());
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_sgld_iterate(ocbnn_problem* h, float* w, const float* noise, const float* half_eps, const int32_t* batch_offsets,
                                  int32_t batch_stride, int64_t m, int32_t t_steps, int32_t with_data, float* out_samples, int32_t thin,
                                  int32_t lanes, void* stream) {
    OCBNN_REQUIRE(h && w && noise && half_eps && m >= 0 && t_steps >= 0, OCBNN_EINVAL, "ocbnn_sgld_iterate: bad argument");
    OCBNN_REQUIRE(batch_stride <= 1 || batch_offsets, OCBNN_EINVAL, "ocbnn_sgld_iterate: batch_offsets required when batch_stride > 1");
    if (m == 0 || t_steps == 0) return OCBNN_OK;
    if (batch_stride > 1)
        for (int t = 0; t < t_steps; ++t)
            OCBNN_REQUIRE(batch_offsets[t] >= 0 && batch_offsets[t] < batch_stride, OCBNN_EINVAL,
                          "ocbnn_sgld_iterate: batch_offsets[%d] = %d is outside [0, %d)", t, (int)batch_offsets[t], (int)batch_stride);
    ProblemImpl& p = *reinterpret_cast<ProblemImpl*>(h);
    cudaStream_t st = (cudaStream_t)stream;
    EvalArgs a = eval_args(p, 0, batch_stride, with_data);
    if (small_net_supported(p)) {
        int rc = ensure_scratch(&p, (size_t)t_steps);
        if (rc) return rc;
        OCBNN_CUDA(cudaMemcpyAsync(p.scratch_f, half_eps, t_steps * sizeof(float), cudaMemcpyHostToDevice, st));
        if (batch_stride > 1)
            OCBNN_CUDA(cudaMemcpyAsync(p.scratch_i, batch_offsets, t_steps * sizeof(int), cudaMemcpyHostToDevice, st));
        int pts = (a.with_data ? batch_count(a.n_train, 0, a.batch_stride) : 0) + a.n_cpoints;
        if (lanes <= 0) lanes = choose_lanes(p, m, pts);
        return launch_sgld_small(p, a, w, noise, p.scratch_f, p.scratch_i, m, t_steps, out_samples, thin, lanes, st);
    }
    const long long md = m * p.d;
    float* gbuf = nullptr;
    OCBNN_CUDA(cudaMallocAsync(&gbuf, md * sizeof(float), st));
    int rc = OCBNN_OK;
    for (int t = 0; t < t_steps; ++t) {
        EvalArgs at = eval_args(p, batch_stride > 1 ? batch_offsets[t] : 0, batch_stride, with_data);
        rc = launch_logpost_generic(p, at, w, m, nullptr, gbuf, st);
        if (rc) break;
        k_sgld_update<<<ew_blocks(md), 256, 0, st>>>(w, gbuf, noise + (long long)t * md, half_eps[t], md);
        count_launch();
        if (out_samples && thin > 0 && (t + 1) % thin == 0)
            OCBNN_CUDA(cudaMemcpyAsync(out_samples + (long long)((t + 1) / thin - 1) * md, w, md * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    cudaFreeAsync(gbuf, st);
    if (rc) return rc;
    OCBNN_CUDA(cudaGetLastError());
    return OCBNN_OK;
}

extern "C" OCBNN_API int ocbnn_hmc_iterate_host(ocbnn_problem* h, float* q_host, const float* p0_host, const double* u_host, int64_t m, int32_t n_it,
                                      double eps, int32_t leapfrog_l, int32_t with_data, int32_t restore_on_reject, uint8_t* accept_host,
                                      int32_t lanes) {
    OCBNN_REQUIRE(h && q_host && p0_host && u_host && m >= 1 && n_it >= 1, OCBNN_EINVAL, "ocbnn_hmc_iterate_host: bad argument");
    ProblemImpl& p = *reinterpret_cast<ProblemImpl*>(h);
    const long long md = m * p.d;
    float *q = nullptr, *p0 = nullptr;
    double* u = nullptr;
    uint8_t* acc = nullptr;
    cudaStream_t st = nullptr;
    OCBNN_CUDA(cudaMalloc(&q, md * sizeof(float)));
    OCBNN_CUDA(cudaMalloc(&p0, (size_t)n_it * md * sizeof(float)));
    OCBNN_CUDA(cudaMalloc(&u, (size_t)n_it * m * sizeof(double)));
    OCBNN_CUDA(cudaMalloc(&acc, (size_t)n_it * m));
    OCBNN_CUDA(cudaMemcpyAsync(q, q_host, md * sizeof(float), cudaMemcpyHostToDevice, st));
    OCBNN_CUDA(cudaMemcpyAsync(p0, p0_host, (size_t)n_it * md * sizeof(float), cudaMemcpyHostToDevice, st));
    OCBNN_CUDA(cudaMemcpyAsync(u, u_host, (size_t)n_it * m * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = ocbnn_hmc_iterate(h, q, p0, u, m, n_it, eps, leapfrog_l, with_data, restore_on_reject, acc, nullptr, nullptr, nullptr, 0, lanes, st);
    if (rc == OCBNN_OK) {
        cudaError_t e = cudaMemcpyAsync(q_host, q, md * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && accept_host) e = cudaMemcpyAsync(accept_host, acc, (size_t)n_it * m, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "hmc_iterate_host copy-back", __FILE__, __LINE__);
    }
    cudaFree(q);
    cudaFree(p0);
    cudaFree(u);
    cudaFree(acc);
    return rc;
}
