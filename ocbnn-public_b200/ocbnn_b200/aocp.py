"""
Gaussian AOCP (amortised output-constrained prior): objective, gradient and the Adagrad learning loop of
bnn/base.py:164-188, 390-419, 624-672.  The user callbacks (interval_func / prob_func) are evaluated ONCE on the fixed
constraint points; the objective and its closed-form gradient run in the CUDA engine (K6, csrc/aocp.cu).
"""

import logging
import math

import numpy as np
import torch

from . import engine as E


def _targets(model, dev):
    """List of (mode, cx, targets, max_iv, count) blocks of the model's AOCP constraints."""
    blocks = []
    S = model.ocp_nsamples
    if model._task == "binary":
        if "gaussian_aocp" in model.dconstraints:
            xs = model._cr_det_xsamples
            cls = torch.cat([torch.full((S,), float(dclass)) for _, dclass in model.dconstraints["gaussian_aocp"]])
            blocks.append((0, xs, cls, 0))
        if "gaussian_aocp" in model.pconstraints:
            xs = model._cr_prob_xsamples
            ps = torch.cat([torch.as_tensor(pfunc(xs[i * S:(i + 1) * S, :])).float().reshape(-1)
                            for i, (_, pfunc) in enumerate(model.pconstraints["gaussian_aocp"])])
            blocks.append((1, xs, ps, 0))
    elif model._task == "regression":
        if model.Xdim != 1:
            raise NotImplementedError("the regression AOCP objective of the reference is defined for Xdim = 1 (base.py:401)")
        xs = model._cr_aocp_xsamples
        ivs = []
        for i, (_, ifunc) in enumerate(model.dconstraints["gaussian_aocp"]):
            ivs += ifunc(xs[i * S:(i + 1) * S, :])
        max_iv = max(len(iv) for iv in ivs)
        t = torch.full((len(ivs), 2 * max_iv), float("nan"))
        for c, iv in enumerate(ivs):
            for k, (lb, ub) in enumerate(iv):
                t[c, 2 * k], t[c, 2 * k + 1] = float(lb), float(ub)
        blocks.append((2, xs, t, max_iv))
    else:
        raise NotImplementedError("the AOCP is defined for regression and binary classification")
    return [(mode, xs.to(dev, torch.float32).contiguous(), t.to(dev, torch.float32).contiguous(), miv) for mode, xs, t, miv in blocks]


def gaussian_aocp_objective(model, params=None, blocks=None):
    """(loss (1,), grad (D,2)) device tensors of the AOCP objective at `params` (default: model._aocp_params)."""
    dev = model._device()
    prob = model.engine_problem()
    P = (model._aocp_params if params is None else params).detach().to(dev, torch.float32).contiguous()
    blocks = _targets(model, dev) if blocks is None else blocks
    denom = sum(b[1].shape[0] for b in blocks)
    loss = torch.zeros(1, device=dev)
    grad = torch.zeros(model.nweights, 2, device=dev)
    for i, (mode, xs, t, miv) in enumerate(blocks):
        prob.aocp_objective(P, xs, t, mode, miv, float(model.sigma_noise), 1.0 / denom, loss, grad, accumulate=i > 0)
    return loss, grad


def learn_gaussian_aocp(model, verbose=True):
    """Adagrad on the AOCP objective (base.py:164-188); saves history/{uid}_aocp.pt and sets aocp_mean / aocp_std."""
    dev = model._device()
    init_means = model.aocp_init_mean + torch.randn(model.nweights, 1)
    init_log_stds = model.aocp_init_std * torch.ones(model.nweights, 1)
    P = torch.cat([init_means, init_log_stds], dim=1).to(dev).contiguous()
    acc = torch.zeros_like(P)
    blocks = _targets(model, dev)
    for epoch in range(1, model.aocp_nepochs + 1):
        loss, grad = gaussian_aocp_objective(model, P, blocks)
        E.adagrad_step(P, acc, grad, model.aocp_init_lr, 1.0)
        if verbose and epoch % 10 == 0:
            logging.info(f"[{model.uid}] Epoch {epoch}: {float(loss):.2f}")
    model._aocp_params = P.cpu()
    torch.save(model._aocp_params, f"history/{model.uid}_aocp.pt")
    model._set_aocp(model._aocp_params)
    logging.info(f"[{model.uid}] AOCP parameters learnt. Also saved as `history/{model.uid}_aocp.pt`.")
