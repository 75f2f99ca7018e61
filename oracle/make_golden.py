"""
Golden-vector generator: runs the UNMODIFIED reference (imported from
/root/reference, PyTorch CPU, autograd) on the seeded cases of tests/cases.py
and writes tests/golden/<case>.npz.  Run in the build container only:

    python oracle/make_golden.py [case ...]

The reference cannot travel to the GPU box, the fixtures can.  RNG injection
(momenta, uniforms, Langevin noise, initial particles) is done by patching the
names the reference looks up (`bnn.inference.Normal`, `numpy.random.uniform`);
the reference's own source is never modified.  TEST INFRASTRUCTURE ONLY.
"""

import json
import os
import sys
import tempfile
from unittest import mock

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bnn as refbnn                      # noqa: E402  (the reference package)
import bnn.inference as refinf            # noqa: E402
from data import dataloader as refdata    # noqa: E402
import cases                              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
M_EVAL = 6


def np32(t):
    return t.detach().cpu().numpy()


def dataset_for(case):
    name = case["data"]
    if name.startswith("toy"):
        return getattr(refdata, name)()
    return getattr(cases, name)()


def config_for(case):
    with open(os.path.join(REF, "repro", case["yaml"] + ".yaml")) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    cfg.update(case.get("overrides", {}))
    return cfg


def eval_points(m, case_name, case, gen):
    """In-distribution weight vectors: rows of a pretrained .pt when the shape matches, else N(0, .7)."""
    D = m.nweights
    for cand in (case["yaml"] + "_hmc2", case["yaml"] + "_hmc1", case["yaml"] + "_svgd2", case["yaml"] + "_svgd1"):
        p = os.path.join(REF, "repro", cand + ".pt")
        if os.path.exists(p):
            t = torch.load(p)
            if t.ndim == 2 and t.shape[1] == D and t.shape[0] >= M_EVAL:
                idx = torch.linspace(0, t.shape[0] - 1, M_EVAL - 2).long()
                W = torch.cat([t[idx].float(), 0.7 * torch.randn(2, D, generator=gen)], 0)
                return W
    if m.aocp_mean is not None and case.get("aocp") == "random":
        return m.aocp_mean[None] + m.aocp_std[None] * torch.randn(M_EVAL, D, generator=gen)
    return 0.7 * torch.randn(M_EVAL, D, generator=gen)


def value_and_grad(m, fn, W):
    vals, grads = [], []
    for w in W:
        m.weights = torch.autograd.Variable(w.clone(), requires_grad=True)
        v = fn()
        v.backward()
        vals.append(v.item())
        grads.append(m.weights.grad.clone())
    return np.array(vals, dtype=np.float32), np32(torch.stack(grads))


class FakeNormal:
    """Stands in for torch.distributions.Normal inside bnn.inference: .sample() pops injected draws."""
    queue = []
    scaled = False

    def __init__(self, loc, scale):
        self.scale = scale

    def sample(self, size=None):
        z = FakeNormal.queue.pop(0)
        if FakeNormal.scaled:                        # SGLD: Normal(0, sqrt(eps)).sample()
            return z * self.scale
        return z.clone()


def run_case(name):
    case = cases.CASES[name]
    cfg = config_for(case)
    out = {"config_json": np.array(json.dumps(cfg))}
    work = tempfile.mkdtemp(prefix="golden_")
    os.chdir(work)
    cfg_path = os.path.join(work, "cfg.yaml")
    with open(cfg_path, "w") as f:
        yaml.dump(cfg, f)
    ds = dataset_for(case)
    for k in ("X_train", "Y_train"):
        out["ds_" + k] = np32(ds[k])
    out["ds_X_train_min"] = np.array(ds["X_train_min"], dtype=np.float64)
    out["ds_X_train_max"] = np.array(ds["X_train_max"], dtype=np.float64)

    gen = torch.Generator().manual_seed(1234)

    def fresh(sampler):
        m = cases.build_model(name, refbnn, cfg_path, ds, seed=0, sampler=sampler)
        if case.get("aocp"):
            if case["aocp"] == "random":
                g2 = torch.Generator().manual_seed(7)
                params = torch.stack([0.3 * torch.randn(m.nweights, generator=g2),
                                      -1.0 + 0.3 * torch.randn(m.nweights, generator=g2)], 1)
                pth = os.path.join(work, "aocp.pt")
                torch.save(params, pth)
            else:
                pth = os.path.join(REF, "repro", case["aocp"] + ".pt")
            m.load_gaussian_aocp_parameters(pth)
            out["aocp_params"] = np32(torch.load(pth))
        m.update_config(use_ocbnn=bool(case.get("use_ocbnn", False)))
        return m

    m = fresh("HMC")
    out["init_weights"] = np32(m.weights)
    for attr in ("_cr_neg_xsamples", "_cr_pos_xsamples", "_cr_pos_ysamples", "_cr_xsamples",
                 "_cr_aocp_xsamples", "_cr_det_xsamples", "_cr_prob_xsamples"):
        if hasattr(m, attr):
            out["cr" + attr] = np32(getattr(m, attr))
    if hasattr(m, "_cr_ylens"):
        out["cr_ylens"] = np.array(m._cr_ylens, dtype=np.int64)
    if m.aocp_mean is not None:
        out["aocp_mean"], out["aocp_std"] = np32(m.aocp_mean), np32(m.aocp_std)

    # ---- K1: log-prior / log-likelihood values and autograd gradients --------------------------
    W = eval_points(m, name, case, gen)
    out["W"] = np32(W)
    out["lp_prior"], out["g_prior"] = value_and_grad(m, m.log_prior, W)
    out["lp_lik"], out["g_lik"] = value_and_grad(m, m.log_likelihood, W)
    out["lp_post"], out["g_post"] = value_and_grad(m, m.log_posterior, W)
    if cfg.get("nbatches"):
        nb = cfg["nbatches"]
        for off in (1, nb - 1):
            bi = torch.arange(off, m.N_train, nb)
            v, g = value_and_grad(m, lambda: m.log_posterior(bi), W)
            out[f"lp_post_b{off}"], out[f"g_post_b{off}"] = v, g
    # baseline prior as well (use_ocbnn False)
    m.update_config(use_ocbnn=False)
    out["lp_prior_base"], out["g_prior_base"] = value_and_grad(m, m.log_prior, W)
    m.update_config(use_ocbnn=bool(case.get("use_ocbnn", False)))

    # ---- predict ----------------------------------------------------------------------------------
    dom = torch.stack([torch.linspace(-2.0, 2.0, 7) + 0.1 * d for d in range(m.Xdim)], 1)
    out["predict_domain"] = np32(dom)
    out["forward_out"] = np32(torch.stack([m.forward(dom, weights=w) for w in W]))
    if case["cls"] != "Regressor":
        out["predict_probs"] = np32(m.predict(W, dom, return_probs=True))

    # ---- AOCP objective and its autograd gradient ----------------------------------------------
    if case.get("aocp") and name != "recid_small":
        # torch>=1.8 validates Distribution arguments; the regression objective passes Python floats to
        # Normal.cdf (fine under the pinned torch 1.5).  Switch the check off (a torch global, not a
        # change to the reference) for this block only.
        torch.distributions.Distribution.set_default_validate_args(False)
        params = torch.tensor(out["aocp_params"]).clone()
        m._aocp_params = torch.autograd.Variable(params, requires_grad=True)
        loss = m.gaussian_aocp_objective()
        loss.backward()
        out["aocp_loss"] = np.array(loss.item(), dtype=np.float32)
        out["aocp_grad"] = np32(m._aocp_params.grad)
        if case["cls"] == "BinaryClassifier":
            out["aocp_prob_targets"] = np32(cases.pf_x2(m._cr_prob_xsamples))
        # three epochs of the learning loop from a seeded init
        m2 = fresh("HMC")
        m2.update_config(aocp_nepochs=3)
        torch.manual_seed(11)
        out["aocp_learn_init_noise"] = np32(torch.randn(m2.nweights, 1))
        torch.manual_seed(11)
        m2.learn_gaussian_aocp(verbose=False)
        out["aocp_learn_params"] = np32(m2._aocp_params)
        out["aocp_learn_mean"], out["aocp_learn_std"] = np32(m2.aocp_mean), np32(m2.aocp_std)
        torch.distributions.Distribution.set_default_validate_args(True)

    small = name in cases.SAMPLER_CASES
    batched = name in cases.BATCHED_SAMPLER_CASES

    POSTERIOR_FILES = {"w1_reg_vanilla": "example_hmc1", "w2_reg_neg": "F1b_hmc2", "f3b_reg_neg2": "F3b_svgd2",
                       "f6_reg_pos": "F6_hmc2", "cls_vanilla": "cbakeoff_hmc1", "w3_cls_dir": "F2b_hmc2",
                       "bin_vanilla": "F4_hmc1", "f4_bin_aocp": "F4_hmc2"}

    def warm_state(with_data):
        if with_data and name in POSTERIOR_FILES:      # a pretrained POSTERIOR sample of the reference itself
            t = torch.load(os.path.join(REF, "repro", POSTERIOR_FILES[name] + ".pt"))
            return t[t.shape[0] // 2].float().clone()
        return _warm_state(with_data)

    def _warm_state(with_data):
        """In-distribution start state for the sampler trajectories: 40 short HMC iterations of the reference itself
        from W[0] (a stiff, far-from-equilibrium start amplifies fp32 reassociation noise chaotically)."""
        mw = fresh("HMC")
        mw.update_config(hmc_l=10)
        mw.weights = torch.autograd.Variable(W[0].clone(), requires_grad=True)
        mw.accepts = mw.rejects = 0
        torch.manual_seed(99)
        np.random.seed(99)
        for _ in range(150):
            mw.single(with_data)
        return mw.weights.data.clone()

    warm = {}
    if small:
        warm[True] = warm_state(True)
        if name in ("w2_reg_neg", "w3_cls_dir"):
            warm[False] = warm_state(False)

    # ---- K2: HMC trajectories with injected momenta and uniforms --------------------------------
    if small:
        for with_data in ((True, False) if name in ("w2_reg_neg", "w3_cls_dir") else (True,)):
            tag = "hmc" if with_data else "hmc_prior"
            n_it = 4
            L = cfg["hmc_l"] if name in ("w1_reg_vanilla", "w2_reg_neg") else 9
            m = fresh("HMC")
            m.update_config(hmc_l=L)
            q0 = warm[with_data].clone()
            m.weights = torch.autograd.Variable(q0.clone(), requires_grad=True)
            p0 = torch.randn(n_it, m.nweights, generator=gen)
            u = torch.rand(n_it, generator=gen, dtype=torch.float64)
            # make iteration 2 a certain reject to pin the no-restore quirk
            u[2] = float("inf")
            origU, origK = m._compute_potential, m._compute_kinetic
            qs, accs, Hs = [], [], []
            for it in range(n_it):
                FakeNormal.queue, FakeNormal.scaled = [p0[it]], False
                Ulog, Klog = [], []

                def potU(wd, _o=origU, _l=Ulog):
                    v = _o(wd)
                    _l.append(v.item())
                    return v

                def kinK(p, _o=origK, _l=Klog):
                    v = _o(p)
                    _l.append(v.item())
                    return v
                m._compute_potential, m._compute_kinetic = potU, kinK
                m.accepts = m.rejects = 0
                with mock.patch.object(refinf, "Normal", FakeNormal), \
                        mock.patch("numpy.random.uniform", lambda: float(u[it])):
                    m.single(with_data)
                qs.append(np32(m.weights.data).copy())
                accs.append(m.accepts == 1)
                Hs.append([Ulog[0], Klog[0], Ulog[-1], Klog[-1]])
            out[tag + "_q0"], out[tag + "_p0"], out[tag + "_u"] = np32(q0), np32(p0), u.numpy()
            out[tag + "_q"] = np.stack(qs)
            out[tag + "_acc"] = np.array(accs)
            out[tag + "_H"] = np.array(Hs, dtype=np.float32)
            out[tag + "_L"] = np.array(L)

    # ---- HMC driver semantics (burn-in, thinning, naming) on the tutorial case ------------------
    if name == "w1_reg_vanilla":
        m = fresh("HMC")
        m.update_config(hmc_nburnin=2, infer_nsamples=3, hmc_ninterval=2, hmc_l=5)
        q0 = W[1].clone()
        m.weights = torch.autograd.Variable(q0.clone(), requires_grad=True)
        n_it = 2 + 3 * 2
        p0 = torch.randn(n_it, m.nweights, generator=gen)
        u = torch.rand(n_it, generator=gen, dtype=torch.float64)
        FakeNormal.queue, FakeNormal.scaled = list(p0), False
        uq = list(u.numpy())
        with mock.patch.object(refinf, "Normal", FakeNormal), \
                mock.patch("numpy.random.uniform", lambda: uq.pop(0)):
            m.infer(verbose=False)
        out["hmcdrv_q0"], out["hmcdrv_p0"], out["hmcdrv_u"] = np32(q0), np32(p0), u.numpy()
        out["hmcdrv_samples"] = np32(m.all_bayes_samples[-1][0])
        out["hmcdrv_name"] = np.array(m.all_bayes_samples[-1][1])

    # ---- K4: SGLD through the reference's own infer() ---------------------------------------------
    if small or batched:
        m = fresh("SGLD")
        m.update_config(sgld_nburnin=3, infer_nsamples=2, sgld_ninterval=2)
        # the repro step size (eps ~ 6e-3) is beyond the stability limit of these stiff posteriors (curvature ~ N/sigma_noise^2,
        # 1/aocp_std^2): trajectories amplify 1-ulp differences by orders of magnitude per step and go to NaN in the AOCP
        # cases.  The fixture pins the ARITHMETIC of a step, so it uses a step size inside the stable region.
        m.update_config(sgld_epa=1e-5 if m.aocp_mean is not None else 1e-4)
        out["sgld_epa"] = np.array(float(m.sgld_epa))
        w0 = warm[True].clone() if small else W[0].clone()
        m.weights = torch.autograd.Variable(w0.clone(), requires_grad=True)
        T = 3 + 2 * 2
        z = torch.randn(T, m.nweights, generator=gen)
        FakeNormal.queue, FakeNormal.scaled = list(z), True
        with mock.patch.object(refinf, "Normal", FakeNormal):
            m.infer(verbose=False)
        out["sgld_w0"], out["sgld_z"] = np32(w0), np32(z)
        out["sgld_samples"] = np32(m.all_bayes_samples[-1][0])
        out["sgld_final"] = np32(m.weights.data)

    # ---- K3: SVGD through infer() with injected initial particles --------------------------------
    if small or batched:
        n = 10 if small else 6
        m = fresh("SVGD")
        m.update_config(svgd_epochs=3, infer_nsamples=n)
        if m.aocp_mean is not None and batched:
            X0 = m.aocp_mean[None] + 3 * m.aocp_std[None] * torch.randn(n, m.nweights, generator=gen)
        else:
            X0 = torch.cat([W[:min(M_EVAL, n)], 0.7 * torch.randn(n - min(M_EVAL, n), m.nweights, generator=gen)], 0)
        hs, phis = [], []
        orig_single = m.single

        def single_rec(with_data, batch_indices=None):
            hs.append(m._compute_rbf_h().item())
            upd = orig_single(with_data, batch_indices)
            phis.append(np32(upd).copy())
            return upd
        m.single = single_rec
        FakeNormal.queue, FakeNormal.scaled = [X0], False
        with mock.patch.object(refinf, "Normal", FakeNormal):
            m.infer(verbose=False)
        out["svgd_x0"] = np32(X0)
        out["svgd_h"] = np.array(hs, dtype=np.float32)
        out["svgd_phi"] = np.stack(phis)
        out["svgd_final"] = np32(m.all_bayes_samples[-1][0])

    # ---- K5: BBB through infer(); eps regenerated from the same seed -----------------------------
    if small or batched:
        k, ns, ep = 5, 4, 3
        m = fresh("BBB")
        m.update_config(bbb_epochs=ep, bbb_esamples=k, infer_nsamples=ns)
        torch.manual_seed(21)
        init_noise = torch.randn(m.nweights, 1)
        eps = torch.stack([torch.randn(k, m.nweights) for _ in range(ep)])
        final_eps = torch.randn(ns, m.nweights)
        torch.manual_seed(21)
        m.infer(verbose=False)
        out["bbb_init_noise"], out["bbb_eps"], out["bbb_final_eps"] = np32(init_noise), np32(eps), np32(final_eps)
        out["bbb_qparams"] = np32(m.q_params)
        out["bbb_elbos"] = np.array(m.elbos, dtype=np.float32)
        out["bbb_samples"] = np32(m.all_bayes_samples[-1][0])

    # ---- K4 at the repro step size: two injected-noise SGLD steps of the reference's own infer() from the warm state with
    # sgld_epa as the YAML / config.yaml has it (0.05 -> eps_1 = 6.4e-3).  Two steps stay finite on every case (the long
    # fixture above needs the small step size because seven steps at 6e-3 amplify 1-ulp differences chaotically); the draws
    # come after every other fixture so that the older arrays keep their random streams.
    if small or batched:
        z = torch.randn(2, D_ := m.nweights, generator=gen)
        for burn in (1, 0):                      # two steps; ONE where the second leaves fp32 range (recid_small: curvature 1/aocp_std^2 ~ 1e4)
            m = fresh("SGLD")
            m.update_config(sgld_nburnin=burn, infer_nsamples=1, sgld_ninterval=1)
            w0 = warm[True].clone() if small else W[0].clone()
            m.weights = torch.autograd.Variable(w0.clone(), requires_grad=True)
            FakeNormal.queue, FakeNormal.scaled = list(z[:burn + 1]), True
            with mock.patch.object(refinf, "Normal", FakeNormal):
                m.infer(verbose=False)
            if np.isfinite(np32(m.weights.data)).all():
                break
        out["sgldr_epa"] = np.array(float(m.sgld_epa))
        out["sgldr_nburnin"] = np.array(burn)
        out["sgldr_w0"], out["sgldr_z"] = np32(w0), np32(z[:burn + 1])
        out["sgldr_samples"] = np32(m.all_bayes_samples[-1][0])
        out["sgldr_final"] = np32(m.weights.data)
        assert np.isfinite(out["sgldr_final"]).all(), name

    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    sz = os.path.getsize(os.path.join(GOLD, name + ".npz"))
    print(f"[golden] {name}: {len(out)} arrays, {sz / 1024:.0f} KiB")


if __name__ == "__main__":
    import logging
    logging.disable(logging.CRITICAL)
    names = sys.argv[1:] or list(cases.CASES)
    for nm in names:
        run_case(nm)
