"""K1 of the application nets at BASELINE.json's full sizes: the tcgen05 kernel (gen_tc.cu) against the float64
oracle on the synthetic credit / recid shapes - minibatch and full-data evaluations, i.e. 21 .. 1028 point tiles per weight vector,
which exercises the TMEM accumulation chains of the weight gradient and their drains (the golden fixtures have <= 6 tiles)."""
import os
import tempfile

import numpy as np
import pytest
import torch

import ocbnn_oracle as orc
from helpers import hold, rel_err
from ocbnn_b200 import workloads

pytestmark = pytest.mark.gpu


def _spec(m, wl):
    common = dict(sigma_w=m.sigma_w, sigma_noise=m.sigma_noise, use_ocbnn=True, naive_sigmoid=False)
    if wl == "W6":
        return orc.Spec(m.layer_sizes, m.activation, "classifier", m.X_train.numpy(), m.Y_train.numpy(),
                        dir=dict(xs=m._cr_xsamples.numpy(), forbidden=[[1]], S=m.ocp_nsamples, gamma_d=m.cocp_dirichlet_gamma,
                                 alpha=m.cocp_dirichlet_alpha), **common)
    return orc.Spec(m.layer_sizes, m.activation, "binary", m.X_train.numpy(), m.Y_train.numpy(), aocp_mean=m.aocp_mean.numpy(),
                    aocp_std=m.aocp_std.numpy(), **common)


@pytest.mark.parametrize("wl,full_rows", [("W6", 20000), ("W7", None)])
def test_k1_application_shape_full_size(wl, full_rows):
    os.chdir(tempfile.mkdtemp())
    m = workloads.build(wl, seed=0)
    if full_rows:                                   # W6: the full-batch leg on the first 20000 rows keeps the float64 oracle in seconds
        m.load(**{**workloads.synth_credit(n=full_rows), "dataset_name": "synth_credit_20000"})
        m.add_deterministic_constraint(constrained_domain=(0.25, 3.0, -3.0, -1.0) + (-0.2, 0.2) * 8, forbidden_classes=[1])
        m.dconstraints["positive_dirichlet_cocp"] = m.dconstraints["positive_dirichlet_cocp"][-1:]
        m._cr_xsamples = m._cr_xsamples[-m.ocp_nsamples:]
    spec = _spec(m, wl)
    prob = m.engine_problem()
    g = torch.Generator().manual_seed(11)
    W = (0.3 * torch.randn(5, m.nweights, generator=g)).numpy()
    Wd = torch.tensor(W).cuda().contiguous()
    nb = m.nbatches
    for tag, batch, bi in (("minibatch", (1, nb), orc.batch_indices(spec.N_train, 1, nb)), ("full", (0, 1), None)):
        lp, gr = prob.logpost_grad(Wd, batch, True)
        lp64, g64 = orc.log_posterior(spec, W, bi, True, dtype=np.float64)
        lp32, g32 = orc.log_posterior(spec, W, bi, True, dtype=np.float32)
        assert np.isfinite(lp64).all()
        hold("k1_fullsize", wl, tag + "/grad", rel_err(gr.cpu().numpy(), g64), 1e-5, 2 * rel_err(g32, g64))
        hold("k1_fullsize", wl, tag + "/value", np.abs(lp.cpu().numpy() - lp64) / np.abs(lp64), 1e-5, 2 * np.abs(lp32 - lp64) / np.abs(lp64))
    # forward / predict on the same machinery (k_forward_gen_tc): 3000 points, 24 tiles of 128 (47 of 64)
    Xf = m.X_train[:3000]
    F = prob.forward(Wd, Xf.cuda().contiguous()).cpu().numpy()
    F64 = orc.forward(spec, W.astype(np.float64), Xf.numpy().astype(np.float64), dtype=np.float64)
    hold("k1_fullsize", wl, "forward", np.abs(F - F64).max() / max(1.0, np.abs(F64).max()), 1e-5)
    lp, gr = prob.logpost_grad(Wd, (0, 1), False)                 # prior only: no points
    lp64, g64 = orc.log_posterior(spec, W, None, False, dtype=np.float64)
    hold("k1_fullsize", wl, "prior/grad", rel_err(gr.cpu().numpy(), g64), 1e-5)
