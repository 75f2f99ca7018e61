"""The C restatement (oracle/ocbnn_oracle.c, the multi-core CPU baseline) against the golden vectors of the reference and
the numpy oracle."""
import numpy as np
import pytest

import c_port
import cases
import ocbnn_oracle as orc
from helpers import golden_config, load_golden, rel_err, spec_from_golden

CASES = [n for n in cases.CASES if n not in ("recid_small", "credit_small")]


@pytest.mark.parametrize("name", CASES)
def test_c_port_logpost(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    cs = c_port.CSpec(spec)
    for with_data, key in ((True, "post"), (False, "prior")):
        lp, gr = cs.logpost_grad(g["W"], with_data)
        o32 = orc.log_posterior(spec, g["W"], None, True) if with_data else orc.log_prior(spec, g["W"])
        o64 = orc.log_posterior(spec, g["W"], None, True, dtype=np.float64) if with_data else orc.log_prior(spec, g["W"], dtype=np.float64)
        tol = np.maximum(3e-5, 4 * rel_err(o32[1], o64[1]))
        assert (rel_err(lp, g["lp_" + key]) <= 3e-5).all(), (name, key)
        assert (rel_err(gr, g["g_" + key]) <= tol).all(), (name, key, rel_err(gr, g["g_" + key]), tol)


@pytest.mark.parametrize("name", ["w1_reg_vanilla", "w2_reg_neg", "cls_vanilla"])
def test_c_port_hmc(name):
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    cs = c_port.CSpec(spec)
    L = int(g["hmc_L"])
    for it in range(g["hmc_p0"].shape[0]):
        q0 = (g["hmc_q0"] if it == 0 else g["hmc_q"][it - 1])[None]
        q, acc = cs.hmc_iterate(q0, g["hmc_p0"][it][None, None], g["hmc_u"][it][None, None], cfg["hmc_epsilon"], L)
        q64, _, _ = orc.hmc_iterate(spec, q0, g["hmc_p0"][it][None, None], g["hmc_u"][it][None, None], cfg["hmc_epsilon"], L, dtype=np.float64)
        q32, _, _ = orc.hmc_iterate(spec, q0, g["hmc_p0"][it][None, None], g["hmc_u"][it][None, None], cfg["hmc_epsilon"], L)
        assert bool(acc[0, 0]) == bool(g["hmc_acc"][it])
        assert rel_err(q, g["hmc_q"][it][None]).max() < max(1e-4, 4 * rel_err(q32, q64).max())


@pytest.mark.parametrize("name", ["f3b_reg_neg2", "w1_reg_vanilla", "bin_vanilla"])
def test_c_port_svgd(name):
    """The C SVGD epoch (bench.py's multi-core CPU baseline for the SVGD metric) against the reference's golden vectors."""
    g = load_golden(name)
    spec = spec_from_golden(name, g)
    cfg = golden_config(g)
    cs = c_port.CSpec(spec)
    X, acc = g["svgd_x0"].copy(), np.zeros_like(g["svgd_x0"])
    for threads in (1, 3):
        h = c_port.svgd_rbf_h(g["svgd_x0"], threads)
        assert abs(h - g["svgd_h"][0]) <= 2e-6 * g["svgd_h"][0]
    Xn, accn, phi, h = c_port.svgd_step(cs, X, acc, cfg["svgd_init_lr"], True, threads=2)
    assert rel_err(phi, g["svgd_phi"][0]).max() < 2e-5
    Xo, _, phi_o, h_o = orc.svgd_step(spec, X, acc, cfg["svgd_init_lr"], None, True)
    assert abs(h - h_o) <= 2e-6 * h_o and rel_err(phi, phi_o).max() < 2e-5
