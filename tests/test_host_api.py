"""Host logic without a GPU: the API mirror draws the same random streams as the reference, registers constraints the
same way, and the constraint compiler's tables reproduce the reference's log prior / log posterior values."""
import numpy as np
import pytest
import torch

import cases
import ocbnn_b200
from helpers import golden_config, load_golden, make_model, rel_err
from ocbnn_b200 import engine
from ocbnn_b200.compiler import batch_spec, compile_problem, forbidden_gaps
from table_eval import logpost_value

ALL = list(cases.CASES)


@pytest.mark.parametrize("name", ALL)
def test_same_rng_stream_and_constraint_points(name):
    g = load_golden(name)
    m = make_model(name, g)
    assert m.nweights == g["W"].shape[1]
    # (initial weights are drawn before any constraint point, base.py:68)
    np.testing.assert_array_equal(m.weights.detach().numpy(), g["init_weights"])
    for attr in ("_cr_neg_xsamples", "_cr_pos_xsamples", "_cr_pos_ysamples", "_cr_xsamples", "_cr_aocp_xsamples",
                 "_cr_det_xsamples", "_cr_prob_xsamples"):
        if "cr" + attr in g:
            np.testing.assert_array_equal(getattr(m, attr).numpy(), g["cr" + attr], err_msg=attr)
    if "cr_ylens" in g:
        assert list(m._cr_ylens) == list(g["cr_ylens"])
    if "aocp_mean" in g:
        np.testing.assert_allclose(m.aocp_mean.numpy(), g["aocp_mean"], rtol=0, atol=0)
        np.testing.assert_allclose(m.aocp_std.numpy(), g["aocp_std"], rtol=1e-7)


@pytest.mark.parametrize("name", [n for n in ALL if n != "recid_small"])
def test_compiled_tables_reproduce_reference_values(name):
    g = load_golden(name)
    m = make_model(name, g)
    cp = compile_problem(m)
    W = g["W"][:3]
    tol = 3e-5
    assert rel_err(logpost_value(cp, W, with_data=False), g["lp_prior"][:3]).max() < tol
    assert rel_err(logpost_value(cp, W), g["lp_post"][:3]).max() < tol
    nb = golden_config(g).get("nbatches")
    if nb:
        assert rel_err(logpost_value(cp, W, batch=(1, nb)), g["lp_post_b1"][:3]).max() < tol
    m.update_config(use_ocbnn=False)
    cp0 = compile_problem(m)
    assert cp0.cx.shape[0] == 0 and cp0.prior_mean_vec is None
    assert rel_err(logpost_value(cp0, W, with_data=False), g["lp_prior_base"][:3]).max() < tol


def test_forbidden_gaps_joint_decision():
    S = 5
    X = torch.linspace(-1, 1, S).unsqueeze(1)
    g1 = forbidden_gaps(cases.if_f1a(X))          # permitted (2.5,3): gaps (-inf,2.5) and (3,inf)
    assert g1.shape == (2, S, 2) and np.isneginf(g1[0, :, 0]).all() and (g1[0, :, 1] == 2.5).all()
    assert (g1[1, :, 0] == 3.0).all() and np.isposinf(g1[1, :, 1]).all()
    g2 = forbidden_gaps(cases.if_f3b(X))          # permitted (-inf,1) U (2.5,inf): one gap (1,2.5)
    assert g2.shape == (1, S, 2) and (g2[0] == np.array([1.0, 2.5], dtype=np.float32)).all()
    g3 = forbidden_gaps(cases.if_touch(X))        # intervals that touch: only the final gap (2,inf) survives
    assert g3.shape == (1, S, 2) and (g3[0, :, 0] == 2.0).all()
    # joint: one point whose second interval starts below the first one's end removes that gap for ALL points
    def mixed(X):
        return [[(-np.inf, 0.0), (1.0 if i else -1.0, np.inf)] for i in range(len(X))]
    assert forbidden_gaps(mixed(X)).shape[0] == 0


def test_batch_spec():
    N = 23
    assert batch_spec(None, N) == (0, 1)
    for nb in (2, 5, 23, 40):
        for t in (0, 1, nb - 1):
            if t % nb >= N:
                continue
            bi = torch.arange(t % nb, N, nb)
            off, stride = batch_spec(bi, N)
            assert torch.equal(torch.arange(off, N, stride), bi) and (stride == 1 or off < stride)
    with pytest.raises(NotImplementedError):
        batch_spec(torch.tensor([0, 1, 5]), N)


def test_api_surface_matches_reference():
    names = {f"BNN{s}{t}" for s in ("HMC", "BBB", "SVGD", "SGLD") for t in ("Regressor", "Classifier", "BinaryClassifier")}
    assert set(ocbnn_b200.__all__) == names
    import bnn as alias
    assert alias.BNNHMCRegressor is ocbnn_b200.BNNHMCRegressor
    g = load_golden("w1_reg_vanilla")
    m = make_model("w1_reg_vanilla", g)
    for meth in ("load", "forward", "log_prior", "log_posterior", "log_likelihood", "predict", "add_deterministic_constraint",
                 "learn_gaussian_aocp", "load_gaussian_aocp_parameters", "load_bayes_samples", "update_config", "clear_all_samples",
                 "debug_mode", "switch_off_debug_mode", "infer"):
        assert callable(getattr(m, meth)), meth
    assert m.all_bayes_samples == [] and m.dconstraints == {} and m.pconstraints == {}
    m.debug_mode()
    assert (m.infer_nsamples, m.hmc_nburnin, m.bbb_epochs, m.svgd_epochs, m.sgld_nburnin, m.aocp_nepochs) == (3, 5, 100, 10, 5, 1)
    m.switch_off_debug_mode()
    assert m.infer_nsamples == golden_config(g)["infer_nsamples"]
    with pytest.raises(Exception):
        m.switch_off_debug_mode()
    import os
    assert os.path.exists(f"history/{m.uid}.yaml")


def test_classifier_rejects_bad_forbidden_classes():
    g = load_golden("cls_vanilla")
    m = make_model("cls_vanilla", g)
    with pytest.raises(Exception):
        m.add_deterministic_constraint(constrained_domain=(0, 1, 0, 1), forbidden_classes=[0, 1, 2])
    with pytest.raises(Exception):
        m.add_deterministic_constraint(constrained_domain=(0, 1, 0, 1), forbidden_classes=[])


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly (never fall back to PyTorch/CPU)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = load_golden("w1_reg_vanilla")
    m = make_model("w1_reg_vanilla", g)
    for call in (m.log_prior, m.log_posterior, lambda: m.forward(torch.zeros(2, 1)), lambda: m.infer(verbose=False),
                 lambda: m.predict(torch.zeros(2, m.nweights), torch.zeros(3, 1))):
        with pytest.raises(engine.EngineError):
            call()
