"""Pins the numpy oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz, made by oracle/make_golden.py).

Tolerances: the reference computes in fp32 with its own summation order, so the fp32 oracle is compared at 2e-5
norm-wise relative error (a few 1e-6 typical) and the float64 oracle bounds the reference's own rounding noise."""
import numpy as np
import pytest

import cases
import ocbnn_oracle as orc
from helpers import load_golden, rel_err, spec_from_golden, golden_config

ALL = list(cases.CASES)
TOL = 2e-5


def check_rows(x32, x64, ref, what):
    """fp32 oracle and float64 oracle against the reference, row by row.  The disagreement between the two oracle
    precisions measures how ill-conditioned a row is in fp32 (saturated tanh terms cancel catastrophically in the
    reference too), so the tolerance of a row is max(TOL, 4 x that disagreement)."""
    cond = rel_err(x32, x64)
    tol = np.maximum(TOL, 4 * cond)
    e32, e64 = rel_err(x32, ref), rel_err(x64, ref)
    assert (e32 <= tol).all(), (what, "fp32 oracle vs reference", e32, tol)
    assert (e64 <= tol).all(), (what, "fp64 oracle vs reference", e64, tol)
    assert (np.minimum(e32, e64) <= 10 * TOL).sum() >= max(1, len(e32) - 2), (what, "too few well-conditioned rows", e32, e64)


@pytest.mark.parametrize("name", ALL)
def test_logpost_value_and_grad(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    W = g["W"]
    r32 = orc.log_prior(spec, W, dtype=np.float32), orc.log_likelihood(spec, W, None, dtype=np.float32), orc.log_posterior(spec, W, None, True, dtype=np.float32)
    r64 = orc.log_prior(spec, W, dtype=np.float64), orc.log_likelihood(spec, W, None, dtype=np.float64), orc.log_posterior(spec, W, None, True, dtype=np.float64)
    for i, key in enumerate(("prior", "lik", "post")):
        check_rows(r32[i][0], r64[i][0], g["lp_" + key], (name, key))
        check_rows(r32[i][1], r64[i][1], g["g_" + key], (name, key + " grad"))
    base = spec_from_golden(name, g, use_ocbnn=False)
    lp, G = orc.log_prior(base, W)
    assert rel_err(lp, g["lp_prior_base"]).max() < TOL and rel_err(G, g["g_prior_base"]).max() < TOL


@pytest.mark.parametrize("name", [n for n in ALL if golden_config(load_golden(n)).get("nbatches")])
def test_minibatch_posterior(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    nb = golden_config(g)["nbatches"]
    for off in (1, nb - 1):
        bi = orc.batch_indices(spec.N_train, off, nb)
        lp, G = orc.log_posterior(spec, g["W"], bi, True)
        lp64, G64 = orc.log_posterior(spec, g["W"], bi, True, dtype=np.float64)
        check_rows(lp, lp64, g[f"lp_post_b{off}"], (name, off))
        check_rows(G, G64, g[f"g_post_b{off}"], (name, off, "grad"))


@pytest.mark.parametrize("name", ALL)
def test_forward_and_predict(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    F = orc.forward(spec, g["W"], g["predict_domain"])
    assert np.abs(F - g["forward_out"]).max() < 1e-4 * max(1.0, np.abs(g["forward_out"]).max())
    if "predict_probs" in g:
        P = orc.predict(spec, g["W"], g["predict_domain"], return_probs=True)
        assert np.abs(P - g["predict_probs"]).max() < 1e-5


def _hmc_tags(g):
    return [t for t in ("hmc", "hmc_prior") if t + "_q" in g]


@pytest.mark.parametrize("name", cases.SAMPLER_CASES)
def test_hmc_trajectory(name):
    """Injected momenta/uniforms: identical accept flags, same trajectory (fp32 reassociation noise), no restore on reject."""
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    for tag in _hmc_tags(g):
        L = int(g[tag + "_L"])
        for it in range(g[tag + "_p0"].shape[0]):
            # every iteration restarts from the REFERENCE's previous state: one trajectory of L steps is compared at a time
            q = (g[tag + "_q0"] if it == 0 else g[tag + "_q"][it - 1])[None].copy()
            args = (g[tag + "_p0"][it][None, None], g[tag + "_u"][it][None, None], cfg["hmc_epsilon"], L)
            q64, _, _ = orc.hmc_iterate(spec, q, *args, with_data=(tag == "hmc"), dtype=np.float64)
            q, acc, H = orc.hmc_iterate(spec, q, *args, with_data=(tag == "hmc"))
            assert bool(acc[0, 0]) == bool(g[tag + "_acc"][it]), (name, tag, it)
            assert rel_err(q, g[tag + "_q"][it][None]).max() < max(1e-4, 4 * rel_err(q, q64).max())
            ref = g[tag + "_H"][it]
            assert abs(H[0, 0, 0] - ref[0]) <= 2e-5 * max(1.0, abs(ref[0])) and abs(H[0, 0, 2] - ref[2]) <= 5e-5 * max(1.0, abs(ref[2]))
            assert abs(H[0, 0, 1] - ref[1]) <= 2e-5 * max(1.0, abs(ref[1]))
        assert not g[tag + "_acc"][2]          # u = 2 forces a reject at iteration 2 ...
        assert np.abs(g[tag + "_q"][2] - g[tag + "_q"][1]).max() > 0   # ... and the reference still moved (no restore)


@pytest.mark.parametrize("name", cases.SAMPLER_CASES + cases.BATCHED_SAMPLER_CASES)
def test_sgld_steps(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    epa = float(g["sgld_epa"]) if "sgld_epa" in g else float(cfg["sgld_epa"])
    nb = cfg.get("nbatches") or 0
    w = g["sgld_w0"][None].copy()
    samples = []
    for i, t in enumerate(range(1, 8)):
        eps = orc.sgld_stepsize(epa, cfg["sgld_epb"], cfg["sgld_epgamma"], t)
        noise = g["sgld_z"][i] * np.float32(eps ** 0.5)
        batch = orc.batch_indices(spec.N_train, t % nb, nb) if nb else None
        w = orc.sgld_step(spec, w, noise[None], eps, batch)
        if t > 3 and (t - 3) % 2 == 0:
            samples.append(w[0].copy())
    assert rel_err(np.stack(samples), g["sgld_samples"]).max() < 2e-4
    assert rel_err(w, g["sgld_final"][None]).max() < 2e-4


@pytest.mark.parametrize("name", cases.SAMPLER_CASES + cases.BATCHED_SAMPLER_CASES)
def test_svgd_epochs(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    nb = cfg.get("nbatches") or 0
    X = g["svgd_x0"].copy()
    acc = np.zeros_like(X)
    for ep in range(1, 4):
        batch = orc.batch_indices(spec.N_train, ep % nb, nb) if nb else None
        X, acc, phi, h = orc.svgd_step(spec, X, acc, cfg["svgd_init_lr"], batch)
        assert abs(h - g["svgd_h"][ep - 1]) <= 1e-5 * g["svgd_h"][ep - 1], (name, ep)
        if ep == 1:
            assert rel_err(phi, g["svgd_phi"][0]).max() < 5e-5
    # Adagrad's first step is +-lr whatever the gradient: later epochs amplify 1-ulp differences, so compare loosely
    assert np.abs(X - g["svgd_final"]).max() < 5e-3 * max(1.0, np.abs(g["svgd_final"]).max())


@pytest.mark.parametrize("name", cases.SAMPLER_CASES + cases.BATCHED_SAMPLER_CASES)
def test_bbb_epochs(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    nb = cfg.get("nbatches") or 0
    D = spec.nweights
    qp = np.concatenate([cfg["bbb_init_mean"] + g["bbb_init_noise"], cfg["bbb_init_std"] * np.ones((D, 1), np.float32)], 1).astype(np.float32)
    acc = np.zeros_like(qp)
    for ep in range(1, 4):
        batch = orc.batch_indices(spec.N_train, ep % nb, nb) if nb else None
        qp_new, acc_new, elbo = orc.bbb_step(spec, qp, acc, g["bbb_eps"][ep - 1], cfg["bbb_init_lr"], batch)
        if ep == 1:
            assert abs(elbo - g["bbb_elbos"][0]) <= 2e-5 * abs(g["bbb_elbos"][0])
        qp, acc = qp_new, acc_new
    assert np.abs(qp - g["bbb_qparams"]).max() < 5e-3 * max(1.0, np.abs(g["bbb_qparams"]).max())


def test_aocp_objective_binary():
    g = load_golden("f4_bin_aocp")
    spec = spec_from_golden("f4_bin_aocp", g)
    loss, grad = orc.aocp_objective_binary(spec, g["aocp_params"], prob=dict(xs=g["cr_cr_prob_xsamples"], p=g["aocp_prob_targets"]))
    assert abs(loss - g["aocp_loss"]) <= 2e-5 * abs(g["aocp_loss"])
    assert rel_err(grad[None], g["aocp_grad"][None]).max() < 5e-5


def test_aocp_objective_regression():
    import torch
    g = load_golden("f3a_reg_aocp")
    spec = spec_from_golden("f3a_reg_aocp", g)
    xs = g["cr_cr_aocp_xsamples"]
    iv = cases.if_f3a(torch.tensor(xs))
    loss, grad = orc.aocp_objective_regression(spec, g["aocp_params"], xs, iv)
    assert abs(loss - g["aocp_loss"]) <= 5e-5 * max(abs(g["aocp_loss"]), 1e-3)
    assert rel_err(grad[None], g["aocp_grad"][None]).max() < 5e-4
