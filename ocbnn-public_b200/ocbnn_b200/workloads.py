"""
Workload definitions of the measurement plan: the repro configs' shapes with synthetic or toy data.

The toy data sets are the tiny fixed point sets of the reference's figures (data/dataloader.py:27-115: toy1 = 6 points
of -x^4+3x^2+1, toy2 = 24 points / 3 classes, toy4 = 4 points, toy5 = 20 points on a 2x10 grid); the application-sized
sets (credit, recid) are synthetic draws of the same shape (the credit CSV is not distributed with the reference).
"""

import math

import numpy as np
import torch
import yaml

# flat config of the reference (config.yaml:4-51) with the values every repro YAML shares
BASE_CONFIG = dict(
    architecture=[10], sigma_w=1, activation="rbf", sigma_noise=0.1, nbatches=0, infer_nsamples=1000,
    hmc_nburnin=10000, hmc_ninterval=10, hmc_epsilon=0.003, hmc_l=50,
    bbb_epochs=10000, bbb_init_mean=0, bbb_init_std=-0.5, bbb_init_lr=0.01, bbb_esamples=5,
    svgd_epochs=1000, svgd_init_lr=0.5,
    sgld_nburnin=15000, sgld_ninterval=10, sgld_epa=0.05, sgld_epb=60, sgld_epgamma=0.5,
    use_ocbnn=False, ocp_nsamples=100, cocp_expo_gamma=10000, cocp_expo_tau=[15, 2.0],
    cocp_dirichlet_gamma=1, cocp_dirichlet_alpha=0.1, cocp_gaussian_sigma_c=1.5,
    aocp_nepochs=50, aocp_init_mean=0, aocp_init_std=0, aocp_init_lr=0.1, aocp_std_multiplier=1)


def write_config(path, **overrides):
    cfg = dict(BASE_CONFIG)
    cfg.update(overrides)
    with open(path, "w") as f:
        yaml.dump(cfg, f)
    return path


def toy1():
    x = torch.tensor([-2, -1.8, -1, 1, 1.8, 2]).unsqueeze(1)
    return dict(dataset_name="toy1", X_train=x, Y_train=-(x ** 4) + 3 * (x ** 2) + 1, X_train_min=[-3.0], X_train_max=[3.0])


def toy2():
    pts = [[-0.5568, -2.8603], [-2.8323, 0.9413], [2.0402, 2.6044], [-0.3121, -3.2628], [-3.2933, 0.8421], [2.2045, 2.6692],
           [-0.0165, -3.1644], [-2.9325, 0.8448], [2.3423, 2.8354], [-0.2595, -2.9736], [-2.9650, 0.7760], [2.2982, 2.6264],
           [-0.0605, -2.8752], [-2.4977, 1.2420], [1.8175, 3.1518], [-0.2743, -3.0461], [-2.7137, 1.7121], [1.9846, 3.5413],
           [0.3204, -2.8159], [-2.6090, 1.1718], [2.0794, 3.1475], [0.1280, -2.9005], [-2.9434, 1.1424], [1.9929, 2.9455]]
    return dict(dataset_name="toy2", X_train=torch.tensor(pts), Y_train=torch.arange(3).repeat(8),
                X_train_min=[-3.0, -3.0], X_train_max=[3.0, 3.0])


def toy4():
    return dict(dataset_name="toy4", X_train=torch.tensor([-3.5, -2, 2, 3.5]).unsqueeze(1),
                Y_train=torch.tensor([-1, 0.25, 3.5, 4.5]).unsqueeze(1), X_train_min=[-4.0], X_train_max=[4.0])


def toy5():
    x = torch.stack((torch.arange(2.0).repeat_interleave(10), torch.arange(0.1, 1.1, 0.1).repeat(2))).t()
    y = torch.tensor([0] * 2 + [1] * 8 + [0] * 8 + [1] * 2)
    return dict(dataset_name="toy5", X_train=x, Y_train=y, X_train_min=[0.0, 0.0], X_train_max=[1.0, 1.0])


def synth_credit(n=131072, seed=0):
    """credit.yaml shape: 10 features, 2 classes, positive rate 1/3 after upsampling (dataloader.py:187)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 10, generator=g)
    y = (torch.rand(n, generator=g) < 1.0 / 3.0).long()
    y[0], y[1] = 0, 1
    return dict(dataset_name="synth_credit", X_train=x, Y_train=y, X_train_min=[-3.0] * 10, X_train_max=[3.0] * 10)


def synth_recid(n=6172, seed=0):
    """recid.yaml shape: COMPAS has 6172 rows x 9 features, 18.5 % positives; feature 1 is binary."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 9, generator=g)
    x[:, 1] = (torch.rand(n, generator=g) < 0.5).float()
    y = (torch.rand(n, generator=g) < 0.185).long()
    return dict(dataset_name="synth_recid", X_train=x, Y_train=y, X_train_min=[-3.0] * 9, X_train_max=[3.0] * 9)


# ---- constraint callbacks of the repro scripts ----------------------------------------------------------------------
def ifunc_f1a(X):          # run_toys.py:37 — y must lie in (2.5, 3) for x in (-0.3, 0.3)
    return [[(2.5, 3.0)] for _ in range(len(X))]


def ifunc_f3b(X):          # run_toys.py:192 — y must avoid (1, 2.5) for x in (-1, 1)
    return [[(-np.inf, 1.0), (2.5, np.inf)] for _ in range(len(X))]


def build(workload, config_path="bench_cfg.yaml", sampler=None, seed=0, **overrides):
    """Model for a named workload of SURVEY §8(d): W1 (example.yaml), W2 (F1a.yaml + negative COCP), W3 (F2a.yaml + Dirichlet
    COCP), W4 (F3b.yaml SVGD + negative COCP), W4P (F4.yaml shape: binary classifier + AOCP prior), W6 (credit.yaml shape), W7 (recid.yaml
    shape, AOCP prior)."""
    from . import models as M
    torch.manual_seed(seed)
    w = workload.upper()
    if w in ("W1", "W2"):
        cls = getattr(M, f"BNN{sampler or 'HMC'}Regressor")
        m = cls(uid=w, configfile=write_config(config_path, **overrides))
        m.load(**toy1())
        if w == "W2":
            m.add_deterministic_constraint(constrained_domain=(-0.3, 0.3), interval_func=ifunc_f1a, prior_type="negative_exponential_cocp")
            m.update_config(use_ocbnn=True)
    elif w == "W3":
        cls = getattr(M, f"BNN{sampler or 'HMC'}Classifier")
        m = cls(uid=w, configfile=write_config(config_path, cocp_dirichlet_alpha=0.85, cocp_dirichlet_gamma=10, **overrides))
        m.load(**toy2())
        m.add_deterministic_constraint(constrained_domain=(1.0, 3.0, -2.0, 0.0), forbidden_classes=[0, 2])
        m.update_config(use_ocbnn=True)
    elif w == "W4":
        cls = getattr(M, f"BNN{sampler or 'SVGD'}Regressor")
        m = cls(uid=w, configfile=write_config(config_path, **overrides))
        m.load(**toy4())
        m.add_deterministic_constraint(constrained_domain=(-1.0, 1.0), interval_func=ifunc_f3b, prior_type="negative_exponential_cocp")
        m.update_config(use_ocbnn=True)
    elif w in ("W4P", "W4'"):                     # F4-shaped: binary classifier on toy5 with an AOCP diagonal-Gaussian prior (run_toys.py:267-286)
        cls = getattr(M, f"BNN{sampler or 'HMC'}BinaryClassifier")
        kw = dict(ocp_nsamples=1000, aocp_std_multiplier=30)
        kw.update(overrides)
        m = cls(uid="W4P", configfile=write_config(config_path, **kw))
        m.load(**toy5())
        g = torch.Generator().manual_seed(7)
        params = torch.stack([0.5 * torch.randn(m.nweights, generator=g), -0.5 + 0.3 * torch.randn(m.nweights, generator=g)], 1)
        m._set_aocp(params)                       # synthetic stand-in for repro/F4_aocp.pt (same shape and scale)
        m.update_config(use_ocbnn=True)
    elif w == "W6":
        cls = getattr(M, f"BNN{sampler or 'BBB'}Classifier")
        kw = dict(architecture=[50, 50], nbatches=60, ocp_nsamples=500, bbb_init_lr=0.1, bbb_init_std=1, cocp_dirichlet_alpha=0.05,
                  cocp_dirichlet_gamma=10)
        kw.update(overrides)
        m = cls(uid=w, configfile=write_config(config_path, **kw))
        m.load(**synth_credit())
        box = (0.25, 3.0, -3.0, -1.0) + (-0.2, 0.2) * 8
        m.add_deterministic_constraint(constrained_domain=box, forbidden_classes=[1])
        m.update_config(use_ocbnn=True)
    elif w == "W7":
        cls = getattr(M, f"BNN{sampler or 'SVGD'}BinaryClassifier")
        kw = dict(architecture=[100, 100], nbatches=100, ocp_nsamples=30, aocp_std_multiplier=30, infer_nsamples=50)
        kw.update(overrides)
        m = cls(uid=w, configfile=write_config(config_path, **kw))
        m.load(**synth_recid())
        g = torch.Generator().manual_seed(7)
        params = torch.stack([0.5 * torch.randn(m.nweights, generator=g), -0.5 + 0.3 * torch.randn(m.nweights, generator=g)], 1)
        m._set_aocp(params)                       # synthetic stand-in for repro/recid_aocp.pt (same shape and scale)
        m.update_config(use_ocbnn=True)
    else:
        raise ValueError(f"unknown workload {workload}")
    return m


# ---- algorithmic work per unit, used by the roofline report ---------------------------------------------
def flops_per_point(sizes):
    """forward + dW + dA, no input gradient for layer 1; FMA = 2."""
    return 6 * sum(a * b for a, b in zip(sizes[:-1], sizes[1:])) - 2 * sizes[0] * sizes[1]


def hmc_step_work(model, with_data=True):
    """(FLOP, MUFU ops, HBM bytes) of one chain-leapfrog-step."""
    cp = model.engine_problem().cp if hasattr(model, "_problem") and model._problem is not None else None
    sizes = model.layer_sizes
    npts = (model.N_train if with_data else 0) + (cp.cx.shape[0] if cp is not None else 0)
    D = model.nweights
    hidden = sum(sizes[1:-1])
    mufu = npts * hidden if model.activation == "rbf" else 0
    if cp is not None and cp.cx.shape[0]:
        neg = int((cp.ckind == 1).sum())
        mufu += neg * 5                                   # 2 Phi terms x 2 exp + one reciprocal for the pair (phi_fast2_1r)
        mufu += int((cp.ckind == 3).sum()) * (sizes[-1] + 2)
    return npts * flops_per_point(sizes) + 8 * D, mufu, 12.0 * D / float(model.hmc_l)
