"""evaluation metrics and standardisation (bnn/utils.py:217-232, run_apps.py:31-41,130-143,
data/dataloader.py:134-176).  CPU part: the host arithmetic against sklearn / numpy restatements of the reference lines.
GPU part: the metrics through the CUDA predictive against the numpy oracle's predict."""
import numpy as np
import pytest
import torch
from sklearn.metrics import accuracy_score, precision_recall_fscore_support

from ocbnn_b200 import utils as U


def test_binary_f1_matches_sklearn():
    rng = np.random.default_rng(0)
    for n, p in ((50, 0.5), (400, 0.1), (10, 0.0), (10, 1.0)):
        y = rng.random(n) < 0.4
        yhat = rng.random(n) < p
        _, _, f1, _ = precision_recall_fscore_support(y, yhat, average="binary", zero_division=0)
        assert abs(U._binary_f1(y, yhat) - f1) < 1e-12


def test_standardisation_matches_reference_lines():
    g = torch.Generator().manual_seed(1)
    X = torch.randn(200, 9, generator=g) * 3 + 1
    ref = X.clone()
    mean = [ref[:, i].mean().item() for i in range(9)]
    std = [ref[:, i].std().item() for i in range(9)]                          # dataloader.py:134-138 (torch: unbiased)
    for j in (0, 2, 3):
        ref[:, j] = (ref[:, j] - mean[j]) / std[j]                            # :140-142
    out = U.standardize_compas(X.clone())
    assert torch.equal(out["X_train"], ref) and out["X_train_mean"] == mean and out["X_train_std"] == std
    assert out["X_train_min"] == [ref[:, i].min().item() for i in range(9)]
    tr, te = np.random.default_rng(2).normal(2, 5, (300, 10)), np.random.default_rng(3).normal(2, 5, (40, 10))
    m = [tr[:, i].mean() for i in range(10)]
    s = [tr[:, i].std() for i in range(10)]                                   # :167-170 (numpy: population)
    tr_ref, te_ref = tr.copy(), te.copy()
    for j in (1, 4, 5):
        te_ref[:, j] = (te_ref[:, j] - m[j]) / s[j]                           # :173-175
        tr_ref[:, j] = (tr_ref[:, j] - m[j]) / s[j]
    out = U.standardize_credit(tr, te)
    assert torch.equal(out["X_train"], torch.tensor(tr_ref).float()) and torch.equal(out["X_test"], torch.tensor(te_ref).float())


def test_bnn_utils_alias_exports_the_reference_names():
    import bnn.utils as BU
    for name in ("eval_accuracy_and_f1_score", "recid_evaluate_high_risk", "credit_evaluate_recourse"):
        assert callable(getattr(BU, name))


@pytest.mark.gpu
def test_metrics_through_the_cuda_predictive():
    import os, tempfile
    import ocbnn_oracle as orc
    from helpers import load_golden, make_model, spec_from_golden
    os.chdir(tempfile.mkdtemp())
    g = torch.Generator().manual_seed(4)
    # binary classifier (recid shape): accuracy / F1 and the high-risk counts
    name = "recid_small"
    gold = load_golden(name)
    m, spec = make_model(name, gold), spec_from_golden(name, gold)
    samples = torch.tensor(gold["W"])
    X = m.X_train[:300]
    Y = (torch.rand(300, generator=g) < 0.3).long()
    m.all_bayes_samples.append((samples, "svgd_ocbnn"))
    probs = orc.predict(spec, gold["W"], X.numpy(), return_probs=True, dtype=np.float64).mean(axis=0)
    pm = U.posterior_mean_probs(m, samples, X)
    assert np.abs(pm.numpy() - probs).max() < 1e-5
    preds = probs >= 0.5
    acc, f1 = U.eval_accuracy_and_f1_score(m, is_binary=True, X_eval=X, Y_eval=Y)
    _, _, f1_ref, _ = precision_recall_fscore_support(Y.numpy(), preds, average="binary", zero_division=0)
    assert abs(acc - accuracy_score(Y.numpy(), preds)) < 1e-6 and abs(f1 - f1_ref) < 1e-6      # bnn/utils.py:229-231
    m.X_train_race = (torch.rand(m.X_train.shape[0], generator=g) < 0.5).float()
    full = orc.predict(spec, gold["W"], m.X_train.numpy(), return_probs=True, dtype=np.float64).mean(axis=0) >= 0.5
    race = m.X_train_race.numpy()
    want = (int(((race == 1) & full).sum()), int((race == 1).sum()), int(((race == 0) & full).sum()), int((race == 0).sum()))
    assert U.recid_evaluate_high_risk(m) == want                                                # run_apps.py:31-41
    # 2-class classifier (credit shape): accuracy / F1 and the recourse means
    name = "credit_small"
    gold = load_golden(name)
    m, spec = make_model(name, gold), spec_from_golden(name, gold)
    samples = torch.tensor(gold["W"])
    m.all_bayes_samples.append((samples, "bbb_ocbnn"))
    m.X_test = torch.randn(400, m.Xdim, generator=g)
    m.Y_test = (torch.rand(400, generator=g) < 0.3).long()
    m.X_train_mean, m.X_train_std = [0.0, 40.0] + [0.0] * 8, [1.0, 12.0] + [1.0] * 8
    probs = orc.predict(spec, gold["W"], m.X_test.numpy(), return_probs=True, dtype=np.float64).mean(axis=0)
    preds = probs.argmax(axis=1)
    acc, f1 = U.eval_accuracy_and_f1_score(m)
    _, _, f1_ref, _ = precision_recall_fscore_support(m.Y_test.numpy(), preds, average="binary", zero_division=0)
    assert abs(acc - accuracy_score(m.Y_test.numpy(), preds)) < 1e-6 and abs(f1 - f1_ref) < 1e-6
    Xn, Yn = m.X_test.numpy(), m.Y_test.numpy()
    young = Xn[:, 1] * 12.0 + 40.0 < 35                                                         # run_apps.py:133
    want = [Xn[young][Yn[young] == 0, 0].mean(), Xn[young][Yn[young] == 1, 0].mean(),
            Xn[young][preds[young] == 0, 0].mean(), Xn[young][preds[young] == 1, 0].mean()]
    got = [float(v) for v in U.credit_evaluate_recourse(m)]
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["w3_cls_dir", "credit_small", "recid_small", "bin_vanilla"])
def test_predict_heads_run_in_the_library(name):
    """ocbnn_predict: softmax / sigmoid probabilities, the decisions and the mean over samples against the reference's golden
    probabilities (base.py:446-461, 544-559) and the float64 oracle."""
    import os, tempfile
    import ocbnn_oracle as orc
    from ocbnn_b200 import engine as E
    from helpers import load_golden, make_model, spec_from_golden
    os.chdir(tempfile.mkdtemp())
    gold = load_golden(name)
    m, spec = make_model(name, gold), spec_from_golden(name, gold)
    W, dom = torch.tensor(gold["W"]), torch.tensor(gold["predict_domain"])
    P = m.predict(W, dom, return_probs=True)
    assert P.shape == tuple(gold["predict_probs"].shape) and np.abs(P.numpy() - gold["predict_probs"]).max() < 1e-5
    P64 = orc.predict(spec, gold["W"], gold["predict_domain"], return_probs=True, dtype=np.float64)
    cls = m.predict(W, dom)
    want = orc.predict(spec, gold["W"], gold["predict_domain"], dtype=np.float64)
    margin = np.abs(P64 - 0.5) if spec.task == "binary" else np.sort(P64, axis=2)[:, :, -1] - np.sort(P64, axis=2)[:, :, -2]
    clear = margin > 1e-5                                       # decisions may differ only where fp32 cannot tell
    assert cls.shape == want.shape and cls.dtype == (torch.bool if spec.task == "binary" else torch.int64)
    assert np.array_equal(cls.numpy()[clear], want[clear])
    prob = m.engine_problem()
    Wd, Xd = W.cuda().contiguous(), dom.cuda().float().contiguous()
    out, both, mean = prob.predict(Wd, Xd, E.PREDICT_PROBS, want_class=True, want_mean=True)
    assert np.abs(mean.cpu().numpy().reshape(P64.mean(axis=0).shape) - P64.mean(axis=0)).max() < 1e-5
    assert np.array_equal(both.cpu().numpy()[clear], want.astype(np.int32)[clear])
    with pytest.raises(E.EngineError):                          # a domain of the wrong width is refused, not read out of bounds
        prob.predict(Wd, torch.zeros(3, m.Xdim + 1, device="cuda"))


@pytest.mark.gpu
def test_constraint_violation_fraction_f3c():
    """Figure 3c's metric (run_toys.py:229-233): fraction of posterior samples that enter (1.0, 2.5) anywhere on
    arange(-1, 1, 0.05), against the reference's own loop over the oracle's predictions; edge cases: no samples inside,
    all inside, empty domain, a single sample."""
    import os, tempfile
    import ocbnn_oracle as orc
    from ocbnn_b200 import engine as E
    from helpers import load_golden, make_model, spec_from_golden
    os.chdir(tempfile.mkdtemp())
    name = "w2_reg_neg"
    gold = load_golden(name)
    m, spec = make_model(name, gold), spec_from_golden(name, gold)
    g = torch.Generator().manual_seed(7)
    samples = torch.cat([torch.tensor(gold["W"]), 0.8 * torch.randn(300, m.nweights, generator=g)])
    domain = torch.arange(-1.0, 1.0, 0.05).unsqueeze(dim=1)
    results = orc.predict(spec, samples.numpy(), domain.numpy(), dtype=np.float64)[:, :, 0]
    clear = ~((np.abs(results - 1.0) < 1e-5) | (np.abs(results - 2.5) < 1e-5)).any(axis=1)
    violation_count = 0
    for line in results[clear]:                                 # the reference's loop, run_toys.py:230-232
        violation_count += any([(point > 1.0 and point < 2.5) for point in line])
    want = violation_count / int(clear.sum())
    got = U.constraint_violation_fraction(m, samples[torch.tensor(clear)], domain, (1.0, 2.5))
    assert 0.0 < want < 1.0 and got == want
    assert U.constraint_violation_fraction(m, samples, domain, (1e6, 2e6)) == 0.0
    assert U.constraint_violation_fraction(m, samples, domain, (-1e6, 1e6)) == 1.0
    assert U.constraint_violation_fraction(m, samples[:1], domain, (-1e6, 1e6)) == 1.0
    assert U.constraint_violation_fraction(m, samples, domain[:0], (-1e6, 1e6)) == 0.0
