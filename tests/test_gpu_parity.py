"""Parity tests proper (run with -m gpu on a B200): the CUDA engine, called through the C-ABI (ctypes), against
(a) the golden vectors produced by the unmodified reference and (b) the numpy oracle in fp32 and float64.

Tolerance, as BASELINE.json states it: log-posterior and gradients within 1e-5 relative (norm-wise per weight vector).
Rows that are ill-conditioned in fp32 — where the fp32 and float64 oracles themselves disagree by more than that,
e.g. saturated tanh penalties — get max(1e-5, 4 x that disagreement) against the reference, and are always held to the
float64 oracle at 1e-5 + the same slack (the engine accumulates values in float64)."""
import numpy as np
import pytest
import torch

import cases
import ocbnn_oracle as orc
from helpers import bound, golden_config, hold, load_golden, make_model, rel_err, spec_from_golden
from ocbnn_b200 import engine as E

pytestmark = pytest.mark.gpu

ALL = list(cases.CASES)
SMALL_NETS = ["w1_reg_vanilla", "w2_reg_neg", "f3b_reg_neg2", "reg_multi_neg", "f6_reg_pos", "reg_pos_neg", "cls_vanilla", "w3_cls_dir",
              "bin_vanilla", "f4_bin_aocp", "f3a_reg_aocp"]
TOL = 1e-5


def dev():
    return torch.device("cuda", 0)


def t2d(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dev(), dtype).contiguous()


def check(x, ref, x32, x64, what, floor=TOL):
    """what = (case, quantity...).  Against the float64 oracle the engine is held to 1e-5 + 2 cond, cond = the fp32-vs-float64 gap
    of the oracle itself for that row (0 where fp32 is well conditioned).  Against the reference's fp32 golden vectors: 1e-5 plus,
    on ill-conditioned rows only, the reference's own distance from the float64 truth (triangle inequality)."""
    cond = rel_err(x32, x64)
    slack_ref = np.maximum(4 * cond, rel_err(ref, x64) + 2 * cond)
    key = "/".join(str(w) for w in what[1:])
    hold("k1", what[0], key + ":vs_f64_oracle", rel_err(x, x64), floor, 2 * cond)
    hold("k1", what[0], key + ":vs_reference", rel_err(x, ref), floor, slack_ref)


@pytest.mark.parametrize("name", ALL)
def test_k1_logpost_and_grad(name):
    g = load_golden(name)
    m = make_model(name, g)
    spec = spec_from_golden(name, g)
    W = g["W"]
    prob = m.engine_problem()
    lanes_list = (0, 1, 2, 4, 8, 16, 32) if name in SMALL_NETS else (0,)
    for with_data, key in ((True, "post"), (False, "prior")):
        f = (lambda dt: orc.log_posterior(spec, W, None, True, dtype=dt)) if with_data else (lambda dt: orc.log_prior(spec, W, dtype=dt))
        (lp32, g32), (lp64, g64) = f(np.float32), f(np.float64)
        for lanes in lanes_list:
            lp, gr = prob.logpost_grad(t2d(W), (0, 1), with_data, lanes=lanes)
            check(lp.cpu().numpy(), g["lp_" + key], lp32, lp64, (name, key, "value"))
            check(gr.cpu().numpy(), g["g_" + key], g32, g64, (name, key, "grad"))
    nb = golden_config(g).get("nbatches")
    if nb:
        for off in (1, nb - 1):
            bi = orc.batch_indices(spec.N_train, off, nb)
            (lp32, g32), (lp64, g64) = orc.log_posterior(spec, W, bi, True), orc.log_posterior(spec, W, bi, True, dtype=np.float64)
            lp, gr = prob.logpost_grad(t2d(W), (off, nb), True)
            check(lp.cpu().numpy(), g[f"lp_post_b{off}"], lp32, lp64, (name, "batch", off))
            check(gr.cpu().numpy(), g[f"g_post_b{off}"], g32, g64, (name, "batch grad", off))


@pytest.mark.parametrize("name", ALL)
def test_host_api_values(name):
    """The reference-facing methods (log_prior/log_posterior/log_likelihood/forward/predict) return the reference's values."""
    g = load_golden(name)
    m = make_model(name, g)
    m.weights = torch.tensor(g["W"][0]).requires_grad_(True)
    assert abs(float(m.log_prior()) - g["lp_prior"][0]) <= 2e-5 * abs(g["lp_prior"][0])
    assert abs(float(m.log_posterior()) - g["lp_post"][0]) <= 2e-5 * abs(g["lp_post"][0])
    assert abs(float(m.log_likelihood()) - g["lp_lik"][0]) <= 2e-5 * max(abs(g["lp_lik"][0]), abs(g["lp_post"][0]))
    dom = torch.tensor(g["predict_domain"])
    F = torch.stack([m.forward(dom, weights=torch.tensor(w)) for w in g["W"][:2]])
    scale = max(1.0, np.abs(g["forward_out"]).max())
    assert np.abs(F.numpy() - g["forward_out"][:2]).max() < 2e-5 * scale
    Fb = m.forward(dom, weights=torch.tensor(g["W"]))                 # batched over weight vectors
    assert np.abs(Fb.numpy() - g["forward_out"]).max() < 2e-5 * scale
    if "predict_probs" in g:
        P = m.predict(torch.tensor(g["W"]), dom, return_probs=True)
        assert np.abs(P.numpy() - g["predict_probs"]).max() < 1e-5
        cls = m.predict(torch.tensor(g["W"]), dom)
        assert cls.shape == tuple(g["predict_probs"].shape[:2])
    else:
        P = m.predict(torch.tensor(g["W"]), dom)
        assert np.abs(P.numpy() - g["forward_out"]).max() < 2e-5 * scale


def _hmc_tags(g):
    return [t for t in ("hmc", "hmc_prior") if t + "_q" in g]


@pytest.mark.parametrize("name", cases.SAMPLER_CASES)
def test_k2_hmc_trajectory(name):
    """Injected momenta and uniforms: same accept decisions, same (U0,K0,U1,K1), same end state as the reference."""
    g = load_golden(name)
    m = make_model(name, g)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    prob = m.engine_problem()
    for tag in _hmc_tags(g):
        L = int(g[tag + "_L"])
        wd = tag == "hmc"
        for it in range(g[tag + "_p0"].shape[0]):
            q_start = (g[tag + "_q0"] if it == 0 else g[tag + "_q"][it - 1])[None]
            args = (g[tag + "_p0"][it][None, None], g[tag + "_u"][it][None, None], cfg["hmc_epsilon"], L)
            q32, a32, H32 = orc.hmc_iterate(spec, q_start, *args, with_data=wd)
            q64, a64, H64 = orc.hmc_iterate(spec, q_start, *args, with_data=wd, dtype=np.float64)
            for lanes in ((1, 4, 16, 32) if name in SMALL_NETS else (0,)):
                q = t2d(q_start)
                acc, H, flags = prob.hmc_iterate(q, t2d(args[0]), t2d(args[1], torch.float64), cfg["hmc_epsilon"], L, wd, want_h=True, lanes=lanes)
                assert int(flags[0]) == 0
                ref_H = g[tag + "_H"][it]
                Hn = H.cpu().numpy()[0, 0]
                alpha_ref = np.exp(np.float32(ref_H[0] - ref_H[2] + ref_H[1] - ref_H[3]))
                u = float(g[tag + "_u"][it])
                # acceptance probability alpha = exp(U0 - U1 + K0 - K1): the engine's against the reference's (differences of
                # O(10..1e4) energies, so alpha carries |H| x 1e-7 of relative noise); outside that band the decisions are identical
                alpha = np.exp(np.float32(Hn[0] - Hn[2] + Hn[1] - Hn[3]))
                if np.isfinite(alpha_ref) and alpha_ref > 1e-30:
                    hold("k2_hmc", name, f"{tag}:accept_probability", abs(alpha - alpha_ref) / alpha_ref, 5e-3)
                if abs(u - alpha_ref) > bound("k2_hmc", name, f"{tag}:accept_probability", 5e-3) * max(alpha_ref, 1e-30):
                    assert bool(acc[0, 0]) == bool(g[tag + "_acc"][it]), (name, tag, it, lanes)
                hold("k2_hmc", name, f"{tag}:energies", np.abs(Hn - ref_H).max() / max(1.0, np.abs(ref_H).max()), 5e-5)
                # end state of an L-step trajectory: the fp32-vs-float64 gap of the oracle itself (the reference's own
                # arithmetic noise over L steps) is the conditioning slack, as in K1
                cond = rel_err(q32, q64).max()
                hold("k2_hmc", name, f"{tag}:end_state_vs_reference", rel_err(q.cpu().numpy(), g[tag + "_q"][it][None]).max(), 1e-4, 4 * cond)
                hold("k2_hmc", name, f"{tag}:end_state_vs_f64_oracle", rel_err(q.cpu().numpy(), q64).max(), 1e-4, 4 * cond)


def test_k2_chains_are_independent_and_restore_works():
    """Replicated chains give bit-identical results whatever their position in the launch; restore_on_reject=1 restores q0."""
    name = "w2_reg_neg"
    g = load_golden(name)
    m = make_model(name, g)
    cfg = golden_config(g)
    prob = m.engine_problem()
    M, n_it, L = 333, 3, 7
    q0 = np.repeat(g["hmc_q0"][None], M, 0)
    p0 = np.repeat(g["hmc_p0"][:n_it, None, :], M, 1)
    u = np.repeat(g["hmc_u"][:n_it, None], M, 1)
    u[u == np.inf] = 0.5
    for lanes in (1, 8, 16, 32):
        q = t2d(q0)
        acc, H, _ = prob.hmc_iterate(q, t2d(p0), t2d(u, torch.float64), cfg["hmc_epsilon"], L, True, want_h=True, lanes=lanes)
        qc = q.cpu().numpy()
        assert (qc == qc[0]).all() and (acc.cpu().numpy() == acc.cpu().numpy()[:, :1]).all()
        single = t2d(q0[:1])
        prob.hmc_iterate(single, t2d(p0[:, :1]), t2d(u[:, :1], torch.float64), cfg["hmc_epsilon"], L, True, lanes=lanes)
        assert (single.cpu().numpy() == qc[:1]).all()
    # forced rejects: the reference-faithful mode keeps the proposal, restore mode returns to q0 exactly
    u_rej = np.full_like(u, np.inf)
    q_keep, q_rest = t2d(q0[:5]), t2d(q0[:5])
    a1, _, _ = prob.hmc_iterate(q_keep, t2d(p0[:, :5]), t2d(u_rej[:, :5], torch.float64), cfg["hmc_epsilon"], L, True, restore_on_reject=False)
    a2, _, _ = prob.hmc_iterate(q_rest, t2d(p0[:, :5]), t2d(u_rej[:, :5], torch.float64), cfg["hmc_epsilon"], L, True, restore_on_reject=True)
    assert int(a1.sum()) == 0 and int(a2.sum()) == 0
    assert (q_rest.cpu().numpy() == q0[:5]).all()
    assert np.abs(q_keep.cpu().numpy() - q0[:5]).max() > 0


def test_hmc_infer_driver_semantics():
    """infer(): burn-in, thinning, sample packing, naming and history file as in inference.py:80-124."""
    import os
    name = "w1_reg_vanilla"
    g = load_golden(name)
    m = make_model(name, g, hmc_nburnin=2, infer_nsamples=3, hmc_ninterval=2, hmc_l=5)
    m.weights = torch.tensor(g["hmcdrv_q0"]).requires_grad_(True)
    p_all, u_all, pos = g["hmcdrv_p0"], g["hmcdrv_u"], [0]

    def draw(n_it, M, device, lo=0, hi=None):
        p = torch.tensor(p_all[pos[0]:pos[0] + n_it])[:, None, :].to(device)
        u = torch.tensor(u_all[pos[0]:pos[0] + n_it])[:, None].to(device)
        pos[0] += n_it
        return p.contiguous(), u.contiguous()
    m._hmc_draw = draw
    m.infer(verbose=False)
    samples, nm = m.all_bayes_samples[-1]
    assert nm == str(g["hmcdrv_name"]) and samples.shape == g["hmcdrv_samples"].shape and samples.dtype == torch.float32
    hold("hmc_infer", name, "samples", rel_err(samples.numpy(), g["hmcdrv_samples"]).max(), 1e-4)
    assert os.path.exists(f"history/{m.uid}_hmc1.pt")
    assert torch.equal(torch.load(f"history/{m.uid}_hmc1.pt"), samples)
    assert pos[0] == 2 + 3 * 2


def test_hmc_multichain_infer():
    g = load_golden("w1_reg_vanilla")
    m = make_model("w1_reg_vanilla", g, hmc_nburnin=3, infer_nsamples=10, hmc_ninterval=2, hmc_l=5, nchains=4, rng="device")
    m.infer(verbose=False)
    s, _ = m.all_bayes_samples[-1]
    assert s.shape == (10, m.nweights) and torch.isfinite(s).all()
    assert m.chain_states.shape == (4, m.nweights)
    assert 0 <= m.accepts <= 3 * 2 * 4


@pytest.mark.parametrize("name", cases.SAMPLER_CASES + cases.BATCHED_SAMPLER_CASES)
def test_k4_sgld(name):
    g = load_golden(name)
    epa = float(g["sgld_epa"])
    m = make_model(name, g, sampler="SGLD", sgld_nburnin=3, infer_nsamples=2, sgld_ninterval=2, sgld_epa=epa)
    m.weights = torch.tensor(g["sgld_w0"]).requires_grad_(True)
    z, pos = g["sgld_z"], [0]

    def noise(ts, M, device, lo=0, hi=None):
        sd = np.array([np.float32(m.get_stepsize(t) ** 0.5) for t in ts], dtype=np.float32)
        out = torch.tensor(z[pos[0]:pos[0] + len(ts)] * sd[:, None])[:, None, :].to(device)
        pos[0] += len(ts)
        return out.contiguous()
    m._sgld_noise = noise
    m.infer(verbose=False)
    samples, nm = m.all_bayes_samples[-1]
    assert nm.startswith("sgld_") and samples.shape == g["sgld_samples"].shape
    hold("k4_sgld", name, "samples", rel_err(samples.numpy(), g["sgld_samples"]).max(), 1e-4)
    hold("k4_sgld", name, "final", rel_err(m.weights.detach().numpy()[None], g["sgld_final"][None]).max(), 1e-4)


@pytest.mark.parametrize("name", cases.SAMPLER_CASES + cases.BATCHED_SAMPLER_CASES)
def test_k4_sgld_repro_stepsize(name):
    """SGLD at the step size the repro configs use (sgld_epa = 0.05 -> eps_1 = 6.4e-3; inference.py:300-302,331-340): the
    reference's own infer() for two injected-noise steps from the warm state (one where the second leaves fp32 range)."""
    g = load_golden(name)
    burn = int(g["sgldr_nburnin"])
    m = make_model(name, g, sampler="SGLD", sgld_nburnin=burn, infer_nsamples=1, sgld_ninterval=1)
    assert float(m.sgld_epa) == float(g["sgldr_epa"]) == 0.05
    m.weights = torch.tensor(g["sgldr_w0"]).requires_grad_(True)
    z, pos = g["sgldr_z"], [0]

    def noise(ts, M, device, lo=0, hi=None):
        sd = np.array([np.float32(m.get_stepsize(t) ** 0.5) for t in ts], dtype=np.float32)
        out = torch.tensor(z[pos[0]:pos[0] + len(ts)] * sd[:, None])[:, None, :].to(device)
        pos[0] += len(ts)
        return out.contiguous()
    m._sgld_noise = noise
    m.infer(verbose=False)
    samples, _ = m.all_bayes_samples[-1]
    assert pos[0] == burn + 1 and samples.shape == g["sgldr_samples"].shape
    # the step multiplies the gradient's relative error by |eps/2 grad| / |w| (up to ~10 on the stiff posteriors)
    hold("k4_sgld_repro", name, "samples", rel_err(samples.numpy(), g["sgldr_samples"]).max(), 1e-4)
    hold("k4_sgld_repro", name, "final", rel_err(m.weights.detach().numpy()[None], g["sgldr_final"][None]).max(), 1e-4)


@pytest.mark.parametrize("tensor", [0, 1], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("name", cases.SAMPLER_CASES + cases.BATCHED_SAMPLER_CASES)
def test_k3_svgd(name, tensor):
    """Three SVGD epochs against the reference's golden vectors (n = 6-10 particles), on the exact SIMT arm and on the
    tcgen05 path (forced: the default picks it from 1024 particles up)."""
    g = load_golden(name)
    n = g["svgd_x0"].shape[0]
    m = make_model(name, g, sampler="SVGD", svgd_epochs=3, infer_nsamples=n)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    st = m.svgd_setup(torch.tensor(g["svgd_x0"]), use_tensor=tensor)
    assert bool(st["tensor"]) == bool(tensor)
    # bandwidth: exact lower median -> h equal to the reference's to fp32 rounding
    h = E.svgd_bandwidth(st["X"], st["ws"], use_tensor=tensor)
    arm = "tcgen05" if tensor else "simt"
    h_tol = 2e-6 if (tensor == 0 or m.nweights <= 64) else 1e-5
    hold("k3_svgd", name, f"{arm}:h_vs_reference", abs(float(h) - g["svgd_h"][0]) / g["svgd_h"][0], h_tol)
    hold("k3_svgd", name, f"{arm}:h_vs_oracle", abs(float(h) - orc.svgd_rbf_h(g["svgd_x0"])) / g["svgd_h"][0], h_tol)
    for ep in range(1, 4):
        m.svgd_epoch(st, ep, True, use_tensor=tensor)
        if ep == 1:
            hold("k3_svgd", name, f"{arm}:phi_vs_reference", rel_err(st["phi"].cpu().numpy(), g["svgd_phi"][0]).max(), 2e-5)
        hold("k3_svgd", name, f"{arm}:h_epoch{ep}", abs(float(st["h"]) - g["svgd_h"][ep - 1]) / g["svgd_h"][ep - 1], 1e-4)
    # three Adagrad steps: the first is exactly -lr sign(phi) (acc = phi^2), so a component of phi whose magnitude is at the
    # fp32 noise floor can flip a coordinate by 2 lr; the bound is on the whole particle matrix
    hold("k3_svgd", name, f"{arm}:final_particles", np.abs(st["X"].cpu().numpy() - g["svgd_final"]).max() / max(1.0, np.abs(g["svgd_final"]).max()), 5e-3)


def test_k3_median_select_and_phi_at_scale():
    """n = 1500, D = 31 / 63 / 200: radix-select median equals the sorted lower median; phi equals the oracle's closed form."""
    gen = torch.Generator().manual_seed(5)
    for D in (31, 63, 200):
        n = 1500 if D < 100 else 300
        X = torch.randn(n, D, generator=gen) * 0.8
        G = torch.randn(n, D, generator=gen)
        Xd, Gd = X.to(dev()), G.to(dev())
        ws = E.svgd_workspace(n, D, dev())
        h = E.svgd_bandwidth(Xd, ws)
        h_ref = orc.svgd_rbf_h(X.numpy())
        assert abs(float(h) - float(h_ref)) <= 2e-6 * float(h_ref), (D, float(h), float(h_ref))
        phi = E.svgd_phi(Xd, Gd, h, 0, n, ws, use_tensor=0)
        ref = orc.svgd_phi(X.numpy().astype(np.float64), G.numpy().astype(np.float64), float(h), dtype=np.float64)
        assert rel_err(phi.cpu().numpy(), ref).max() < 1e-5
        part = E.svgd_phi(Xd, Gd, h, 100, 57, ws, use_tensor=0)          # a rank's row share
        assert torch.equal(part, phi[100:157])


@pytest.mark.parametrize("name", cases.SAMPLER_CASES + cases.BATCHED_SAMPLER_CASES)
def test_k5_bbb(name):
    g = load_golden(name)
    cfg = golden_config(g)
    m = make_model(name, g, sampler="BBB", bbb_epochs=3, bbb_esamples=5, infer_nsamples=4)
    spec = spec_from_golden(name, g)
    D = m.nweights
    qp0 = np.concatenate([cfg["bbb_init_mean"] + g["bbb_init_noise"], cfg["bbb_init_std"] * np.ones((D, 1), np.float32)], 1).astype(np.float32)
    qp, acc = t2d(qp0), torch.zeros(D, 2, device=dev())
    nb = cfg.get("nbatches") or 0
    batch1 = orc.batch_indices(spec.N_train, 1 % nb, nb) if nb else None
    elbo_o, grad_o = orc.bbb_elbo_grad(spec, qp0, g["bbb_eps"][0], batch1, dtype=np.float64)
    W = E.bbb_sample(qp, t2d(g["bbb_eps"][0]))
    lp, G = m.engine_problem().logpost_grad(W, (1 % nb, nb) if nb else (0, 1), True)
    gq, elbo = E.bbb_reduce(qp, t2d(g["bbb_eps"][0]), lp, G, 5)
    hold("k5_bbb", name, "elbo_vs_reference", abs(float(elbo) - g["bbb_elbos"][0]) / abs(g["bbb_elbos"][0]), 2e-5)
    hold("k5_bbb", name, "elbo_vs_f64_oracle", abs(float(elbo) - elbo_o) / abs(elbo_o), 1e-5)
    hold("k5_bbb", name, "grad_vs_f64_oracle", rel_err(gq.cpu().numpy().T, grad_o.T).max(), 2e-5)
    state = {}
    elbos = [float(m.bbb_epoch(qp, acc, t2d(g["bbb_eps"][ep - 1]), ep, True, k_total=5, state=state)) for ep in range(1, 4)]
    hold("k5_bbb", name, "step_elbo_vs_reference", abs(elbos[0] - g["bbb_elbos"][0]) / abs(g["bbb_elbos"][0]), 2e-5)
    hold("k5_bbb", name, "q_params", np.abs(qp.cpu().numpy() - g["bbb_qparams"]).max() / max(1.0, np.abs(g["bbb_qparams"]).max()), 5e-3)


def test_bbb_infer_matches_reference_stream():
    """With rng='host' a seeded infer() consumes the same torch.randn stream as the reference: same q_params and samples."""
    name = "cls_vanilla"
    g = load_golden(name)
    m = make_model(name, g, sampler="BBB", bbb_epochs=3, bbb_esamples=5, infer_nsamples=4)
    torch.manual_seed(21)
    m.infer(verbose=False)
    hold("bbb_infer", name, "q_params", np.abs(m.q_params.numpy() - g["bbb_qparams"]).max() / max(1.0, np.abs(g["bbb_qparams"]).max()), 5e-3)
    s, nm = m.all_bayes_samples[-1]
    assert nm.startswith("bbb_") and s.shape == g["bbb_samples"].shape
    hold("bbb_infer", name, "samples", np.abs(s.numpy() - g["bbb_samples"]).max() / max(1.0, np.abs(g["bbb_samples"]).max()), 2e-2)
    hold("bbb_infer", name, "elbo", abs(m.elbos[0] - g["bbb_elbos"][0]) / abs(g["bbb_elbos"][0]), 2e-5)


def test_full_size_properties():
    """At the bench scale (M = 65536 chains, W2 shapes) through size-independent properties: replicated chains agree
    bit-for-bit across the launch, lanes variants agree to fp32 reassociation, energy error of a short trajectory is small."""
    name = "w2_reg_neg"
    g = load_golden(name)
    m = make_model(name, g)
    cfg = golden_config(g)
    prob = m.engine_problem()
    M = 65536
    gen = torch.Generator(device="cuda").manual_seed(3)
    base = t2d(g["hmc_q0"])
    q0 = base[None].repeat(M, 1) + 0.01 * torch.randn(M, base.numel(), device=dev(), generator=gen)
    q0[M // 2:] = q0[:M // 2]                       # second half replicates the first half
    p0 = torch.randn(2, M, base.numel(), device=dev(), generator=gen)
    p0[:, M // 2:] = p0[:, :M // 2]
    u = torch.rand(2, M, device=dev(), dtype=torch.float64, generator=gen)
    u[:, M // 2:] = u[:, :M // 2]
    outs = {}
    for lanes in (1, 2):
        q = q0.clone()
        acc, H, flags = prob.hmc_iterate(q, p0, u, cfg["hmc_epsilon"], 10, True, want_h=True, lanes=lanes)
        assert int(flags.max()) == 0 and torch.isfinite(q).all()
        assert torch.equal(q[M // 2:], q[:M // 2]) and torch.equal(acc[:, M // 2:], acc[:, :M // 2])
        outs[lanes] = (q, H)
    rel = (outs[1][0] - outs[2][0]).norm(dim=1) / outs[1][0].norm(dim=1)
    assert float(rel.median()) < 1e-5
    lp, _ = prob.logpost_grad(q0, (0, 1), True, want_grad=False)
    assert torch.allclose(-lp, outs[1][1][0, :, 0], rtol=1e-6, atol=1e-3)      # U0 of iteration 0 = -log posterior(q0)


@pytest.mark.parametrize("name", ["f4_bin_aocp", "f3a_reg_aocp"])
def test_k6_aocp_objective_and_learning(name):
    """AOCP objective, its gradient (vs the reference's autograd) and three epochs of learn_gaussian_aocp."""
    from ocbnn_b200.aocp import gaussian_aocp_objective
    g = load_golden(name)
    m = make_model(name, g)
    m._aocp_params = torch.tensor(g["aocp_params"])
    loss, grad = gaussian_aocp_objective(m)
    hold("k6_aocp", name, "loss_vs_reference", abs(float(loss) - float(g["aocp_loss"])) / max(abs(float(g["aocp_loss"])), 1e-3), 5e-5)
    hold("k6_aocp", name, "grad_vs_reference", rel_err(grad.cpu().numpy().T, g["aocp_grad"].T).max(), 5e-4)
    assert abs(float(m.gaussian_aocp_objective()) - float(g["aocp_loss"])) <= 5e-5 * max(abs(float(g["aocp_loss"])), 1e-3)
    m.update_config(aocp_nepochs=3)
    torch.manual_seed(11)
    m.learn_gaussian_aocp(verbose=False)
    hold("k6_aocp", name, "learned_params", np.abs(m._aocp_params.numpy() - g["aocp_learn_params"]).max() / max(1.0, np.abs(g["aocp_learn_params"]).max()), 5e-3)
    hold("k6_aocp", name, "learned_mean", np.abs(m.aocp_mean.numpy() - g["aocp_learn_mean"]).max() / max(1.0, np.abs(g["aocp_learn_mean"]).max()), 5e-3)
    import os
    assert os.path.exists(f"history/{m.uid}_aocp.pt")


def test_k6_aocp_at_recid_scale():
    """D = 11201, 6 constraint points (the reference needs a 500 MB dense matrix per point here): float64 oracle parity."""
    from ocbnn_b200.aocp import gaussian_aocp_objective
    name = "recid_small"
    g = load_golden(name)
    m = make_model(name, g)
    spec = spec_from_golden(name, g)
    m._aocp_params = torch.tensor(g["aocp_params"])
    loss, grad = gaussian_aocp_objective(m)
    xs = g["cr_cr_prob_xsamples"]
    lo, go = orc.aocp_objective_binary(spec, g["aocp_params"], prob=dict(xs=xs, p=xs[:, 1]), dtype=np.float64)
    hold("k6_aocp", name, "loss_vs_f64_oracle", abs(float(loss) - lo) / abs(lo), 2e-5)
    hold("k6_aocp", name, "grad_vs_f64_oracle", rel_err(grad.cpu().numpy().T, go.T).max(), 1e-4)


def test_hmc_run_host_samples_and_svgd_graph_equivalence():
    """hmc_run(samples_host=...) delivers the same samples as the device tensor it returns; the CUDA-graph replay of the SVGD
    epoch is bit-identical to the eager epoch."""
    import os, tempfile
    from ocbnn_b200 import workloads
    os.chdir(tempfile.mkdtemp())
    m = workloads.build("W2", seed=0)
    m.update_config(rng="device", hmc_l=5)
    M, n_it = 257, 6
    q = (0.5 * torch.randn(M, m.nweights, generator=torch.Generator().manual_seed(2))).cuda()
    host = torch.empty(n_it // 2, M, m.nweights).pin_memory()
    torch.manual_seed(11)
    acc, smp = m.hmc_run(q, n_it, True, collect_every=2, chunk=2, samples_host=host)
    torch.cuda.synchronize()
    assert smp.shape == host.shape and torch.equal(smp.cpu(), host)
    assert int(acc.sum()) > 0

    ms = workloads.build("W4", seed=0, infer_nsamples=1100)
    X0 = torch.randn(1100, ms.nweights, generator=torch.Generator().manual_seed(3))
    st_g = ms.svgd_setup(X0)
    ms.update_config(svgd_graph=False)
    st_e = ms.svgd_setup(X0)
    for ep in range(1, 6):                      # graph path: eager, capture + replay, replay, ...
        ms.update_config(svgd_graph=True)
        ms.svgd_epoch(st_g, ep, True)
        ms.update_config(svgd_graph=False)
        ms.svgd_epoch(st_e, ep, True)
    torch.cuda.synchronize()
    assert "graphs" in st_g and any(not isinstance(v, str) for v in st_g["graphs"].values())
    assert torch.equal(st_g["X"], st_e["X"]) and torch.equal(st_g["h"], st_e["h"])


def test_full_size_properties_hmc_262144_chains_and_svgd_4096_particles():
    """BASELINE.json's full sizes through size-independent properties.
    HMC (W2, 262144 chains, L = 50): chains are independent, so 512 distinct chains replicated 512 times each give
    bit-identical replicas wherever they sit in the launch, and every accept flag equals the reference's rule applied to the
    returned energies.  SVGD (W4, 4096 particles): the update is equivariant under a permutation of the particles (bandwidth
    invariant, phi rows permuted)."""
    import os, tempfile
    from ocbnn_b200 import workloads
    os.chdir(tempfile.mkdtemp())
    m = workloads.build("W2", seed=0)
    prob = m.engine_problem()
    base, reps, L = 512, 512, 50
    g = torch.Generator().manual_seed(5)
    q0 = 0.5 * torch.randn(base, m.nweights, generator=g)
    p0 = torch.randn(1, base, m.nweights, generator=g)
    u = torch.rand(1, base, generator=g, dtype=torch.float64)
    perm = torch.randperm(base * reps, generator=g)
    src = (torch.arange(base * reps) % base)[perm]                    # chain slot -> which distinct chain it replicates
    q = q0[src].cuda().contiguous()
    acc, H, flags = prob.hmc_iterate(q, p0[:, src].cuda().contiguous(), u[:, src].cuda().contiguous(), m.hmc_epsilon, L, True, want_h=True)
    assert int(flags.max()) == 0
    qc, ac, Hc = q.cpu(), acc.cpu()[0], H.cpu()[0]
    first = torch.full((base,), -1, dtype=torch.long)
    first[src.flip(0)] = torch.arange(base * reps - 1, -1, -1)         # first slot of every distinct chain
    assert torch.equal(qc, qc[first[src]]) and torch.equal(ac, ac[first[src]]) and torch.equal(Hc, Hc[first[src]])
    # the Metropolis-Hastings decision is the reference's: u < exp(U0 - U1 + K0 - K1) evaluated in fp32 (inference.py:74)
    dH = ((Hc[:, 0] - Hc[:, 2]) + Hc[:, 1]) - Hc[:, 3]
    assert torch.isfinite(dH).all()
    assert torch.equal(ac.bool(), u[0, src] < torch.exp(dH).double())

    ms = workloads.build("W4", seed=0, infer_nsamples=4096)
    X = torch.randn(4096, ms.nweights, generator=g)
    G = 5 * torch.randn(4096, ms.nweights, generator=g)
    pm = torch.randperm(4096, generator=g)
    out = []
    for Xi, Gi in ((X, G), (X[pm], G[pm])):
        Xd, Gd = Xi.cuda().contiguous(), Gi.cuda().contiguous()
        ws = E.svgd_workspace(4096, ms.nweights, Xd.device, use_tensor=1)
        h = E.svgd_bandwidth(Xd, ws, use_tensor=1)
        out.append((float(h), E.svgd_phi(Xd, Gd, h, 0, 4096, ws, use_tensor=2).cpu()))
    assert abs(out[0][0] - out[1][0]) <= 2e-6 * out[0][0]
    rel = (out[0][1][pm] - out[1][1]).norm(dim=1) / out[0][1][pm].norm(dim=1)
    assert float(rel.max()) < 1e-5, float(rel.max())


@pytest.mark.parametrize("env, select", [
    # small generic nets run the SIMT variant by default: force the 3xTF32 tensor variant on them
    (dict(OCBNN_GEN_TC="0", OCBNN_GEN_MMA="1", OCBNN_GEN_MINB="2"), "test_k1_logpost_and_grad and (odd or wide1 or deep)"),
    # mma.sync 3xTF32 variant on the nets the tcgen05 kernel (gen_tc.cu) serves by default
    (dict(OCBNN_GEN_TC="0", OCBNN_GEN_MMA="1"), "test_k1_logpost_and_grad and (credit or recid)"),
    # credit / recid run the tensor variant by default: force the SIMT register-tile variant, both backward schedules
    (dict(OCBNN_GEN_TC="0", OCBNN_GEN_MMA="0", OCBNN_GEN_SCHED="1"), "test_k1_logpost_and_grad and (credit or recid)"),
    (dict(OCBNN_GEN_TC="0", OCBNN_GEN_MMA="0", OCBNN_GEN_SCHED="0", OCBNN_GEN_TP="16"), "test_k1_logpost_and_grad and (credit or odd)"),
    # tcgen05 kernel (gen_tc.cu): 64-point tiles (MMA M = 64) and the in-place dZ layout on nets that default to 128-point tiles /
    # their own dZ array
    (dict(OCBNN_GEN_TC_TP="64"), "test_k1_logpost_and_grad and (credit or odd or deep or wide1)"),
    (dict(OCBNN_GEN_TC_TP="128", OCBNN_GEN_TC_OVERLAP="0"), "test_k1_logpost_and_grad and (credit or odd or deep)"),
], ids=["tensor_path_forced", "mma_sync_path_forced", "simt_path_forced", "simt_16_point_tiles", "tcgen05_64_point_tiles", "tcgen05_dz_in_place"])
def test_generic_net_kernel_variants(env, select):
    """The generic-net K1 picks its variant (tile size, threads, SIMT / tensor path) from the net's shared-memory footprint; the
    variant knobs are read once per process, so each combination runs the K1 parity test in a child process."""
    import os
    import subprocess
    import sys
    e = dict(os.environ, **env)
    select += " and not kernel_variants"                     # never this test itself (-k also matches parameter ids)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k", select,
                        "-p", "no:cacheprovider"], env=e, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, r.stdout[-500:]
