// small_dispatch.cu — routes a problem to the register-resident kernels of its net shape (if instantiated).
#include <cstdlib>
#include "common.cuh"

namespace ocbnn {

#define OCBNN_DECLARE_SMALL_NET(NAME)                                                                                         \
    int NAME##_logpost(const Problem&, const EvalArgs&, const float*, long long, float*, float*, int, cudaStream_t);           \
    int NAME##_hmc(const Problem&, const EvalArgs&, float*, const float*, const double*, long long, int, float, float, int,    \
                   int, uint8_t*, float*, int*, float*, int, int, cudaStream_t);                                              \
    int NAME##_sgld(const Problem&, const EvalArgs&, float*, const float*, const float*, const int*, long long, int, float*,   \
                    int, int, cudaStream_t);                                                                                  \
    int NAME##_forward(const float*, long long, const float*, long long, float*, cudaStream_t);

OCBNN_DECLARE_SMALL_NET(net_reg1_10_1)
OCBNN_DECLARE_SMALL_NET(net_cls2_10_3)
OCBNN_DECLARE_SMALL_NET(net_bin2_10_1)

struct SmallEntry {
    int s0, h, sk, act;
    decltype(&net_reg1_10_1_logpost) logpost;
    decltype(&net_reg1_10_1_hmc) hmc;
    decltype(&net_reg1_10_1_sgld) sgld;
    decltype(&net_reg1_10_1_forward) forward;
};

#define OCBNN_ENTRY(NAME, S0, H, SK, ACT) {S0, H, SK, ACT, NAME##_logpost, NAME##_hmc, NAME##_sgld, NAME##_forward}
static const SmallEntry kSmall[] = {
    OCBNN_ENTRY(net_reg1_10_1, 1, 10, 1, OCBNN_ACT_RBF),
    OCBNN_ENTRY(net_cls2_10_3, 2, 10, 3, OCBNN_ACT_RBF),
    OCBNN_ENTRY(net_bin2_10_1, 2, 10, 1, OCBNN_ACT_RBF),
};

static const SmallEntry* find_small(const Problem& p) {
    if (p.desc.n_layers != 2) return nullptr;
    for (const SmallEntry& e : kSmall)
        if (e.s0 == p.desc.layer_sizes[0] && e.h == p.desc.layer_sizes[1] && e.sk == p.desc.layer_sizes[2] && e.act == p.desc.activation)
            return &e;
    return nullptr;
}

int small_net_supported(const Problem& p) { return find_small(p) != nullptr; }

// lanes per chain: the smallest power of two that puts >= ~8 warps on every SM, capped by the number of points
int choose_lanes(const Problem& p, long long m, int points) {
    const long long want = (long long)p.sm_count * 6 * 32;      // measured: 4096 chains run best with 8 lanes (6.1e8 vs 5.5e8 at 16)
    int g = 1;
    while (g < 32 && m * g < want && 2 * g <= (points > 0 ? points : 1)) g *= 2;
    return g;
}

// HMC: nets served by the slot-ownership kernel (k_hmc_own: one input, one output, rbf, <= 16 slots) run 16 lanes where the rule above
// says 8 (measured, W2, 4096 chains: 0.89 G steps/s at 16 lanes with ownership, 0.83 G at 8 lanes replicated, 0.72 G at 16 replicated)
int choose_lanes_hmc(const Problem& p, long long m, int points) {
    int g = choose_lanes(p, m, points);
    const SmallEntry* e = find_small(p);
    static const bool want_own = [] { const char* v = getenv("OCBNN_HMC_OWN"); return v ? atoi(v) != 0 : true; }();
    if (want_own && g == 8 && e && e->s0 == 1 && e->sk == 1 && e->act == OCBNN_ACT_RBF && e->h <= 10 && points >= 32) g = 16;
    return g;
}

int launch_logpost_small(const Problem& p, const EvalArgs& a, const float* w, long long m, float* lp, float* grad, int lanes,
                         cudaStream_t st) {
    const SmallEntry* e = find_small(p);
    OCBNN_REQUIRE(e, OCBNN_EUNSUPPORTED, "no register-resident kernel for this net shape");
    return e->logpost(p, a, w, m, lp, grad, lanes, st);
}
int launch_hmc_small(const Problem& p, const EvalArgs& a, float* q, const float* p0, const double* u, long long m, int n_it, float eps,
                     float eps_half, int L, int restore, uint8_t* acc, float* h, int* flags, float* samples, int thin, int lanes,
                     cudaStream_t st) {
    const SmallEntry* e = find_small(p);
    OCBNN_REQUIRE(e, OCBNN_EUNSUPPORTED, "no register-resident kernel for this net shape");
    return e->hmc(p, a, q, p0, u, m, n_it, eps, eps_half, L, restore, acc, h, flags, samples, thin, lanes, st);
}
int launch_sgld_small(const Problem& p, const EvalArgs& a, float* w, const float* noise, const float* he, const int* offs, long long m,
                      int t_steps, float* samples, int thin, int lanes, cudaStream_t st) {
    const SmallEntry* e = find_small(p);
    OCBNN_REQUIRE(e, OCBNN_EUNSUPPORTED, "no register-resident kernel for this net shape");
    return e->sgld(p, a, w, noise, he, offs, m, t_steps, samples, thin, lanes, st);
}
int launch_forward_small(const Problem& p, const float* w, long long m, const float* x, long long npts, float* out, cudaStream_t st) {
    const SmallEntry* e = find_small(p);
    OCBNN_REQUIRE(e, OCBNN_EUNSUPPORTED, "no register-resident kernel for this net shape");
    return e->forward(w, m, x, npts, out, st);
}

}  // namespace ocbnn
