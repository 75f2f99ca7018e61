"""
Shared case table for the golden-vector generator (oracle/make_golden.py, run
against the unmodified reference in the build container) and for the parity
tests (run against this repo's host API, the oracle and the CUDA engine).

A case names: the model class, the config (a reference repro YAML + overrides,
the merged dict is stored in the fixture), the data set (stored in the
fixture), the constraint calls (Python callbacks, defined HERE so both sides
run the very same functions) and which sampler artefacts to record.
"""

import numpy as np

INF = np.inf


# ---- constraint callbacks (mirroring the run scripts of the reference) -------

def if_f1a(X):            # run_toys.py:37  permitted [(2.5, 3.0)]
    return [[(2.5, 3.0)] for _ in range(len(X))]


def if_f3b(X):            # run_toys.py:192 permitted (-inf,1) U (2.5,inf)
    return [[(-INF, 1.0), (2.5, INF)] for _ in range(len(X))]


def if_f6(X):             # run_toys.py:294 positive constraint y = 5cos(x/1.7)
    return [[(5 * np.cos(x.item() / 1.7), 5 * np.cos(x.item() / 1.7))] for x in X]


def if_slab(X):           # docstring example base.py:297, x-dependent bounds
    return [[(-INF, 1 + x.item()), (3 + x.item(), INF)] for x in X]


def if_touch(X):          # second permitted interval starts where the first ends -> the
    return [[(-INF, 0.5), (0.5, 2.0)] for _ in X]   # joint all(marker<lb) test FAILS for gap 2


def if_f3a(X):            # run_toys.py:160
    return [[(-INF, 0) if x.item() < 0 else (0, INF)] for x in X]


def pf_x2(X):             # run_toys.py:265 / run_apps.py:74
    return X[:, 1]


# ---- synthetic data sets of the application shapes ---------------------------

def synth_credit(n=240, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, 10, generator=g)
    Y = (torch.rand(n, generator=g) < (1.0 / 3.0)).long()
    Y[0], Y[1] = 0, 1
    return {"dataset_name": "synth_credit", "X_train": X, "Y_train": Y,
            "X_train_min": [-3.0] * 10, "X_train_max": [3.0] * 10}


def synth_recid(n=310, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, 9, generator=g)
    X[:, 1] = (torch.rand(n, generator=g) < 0.5).float()
    Y = (torch.rand(n, generator=g) < 0.185).long()
    return {"dataset_name": "synth_recid", "X_train": X, "Y_train": Y,
            "X_train_min": [-3.0] * 9, "X_train_max": [3.0] * 9}


def synth_reg_y2(n=12, seed=0):
    """Two-output regression (the MVN branch of RegressorMixin.log_likelihood, base.py:265)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    X = 3.0 * torch.rand(n, 2, generator=g) - 1.5
    Y = torch.stack([torch.sin(2 * X[:, 0]) + 0.3 * X[:, 1], X[:, 0] * X[:, 1]], 1) + 0.05 * torch.randn(n, 2, generator=g)
    return {"dataset_name": "synth_reg_y2", "X_train": X, "Y_train": Y,
            "X_train_min": [-1.5] * 2, "X_train_max": [1.5] * 2}


CREDIT_BOX = (0.25, 3.0, -3.0, -1.0) + (-0.2, 0.2) * 8
RECID_BOX = (-0.2, 0.2, 0.0, 1.0) + (-0.2, 0.2) * 7


# ---- the case table ---------------------------------------------------------
# cls: suffix of the BNN<Sampler><cls> class name; yaml: reference repro file;
# data: toyN (reference dataloader) or a synth_* function above;
# constraints: list of (method, kwargs); aocp: repro .pt to load or 'random'.

CASES = {
    "w1_reg_vanilla": dict(cls="Regressor", yaml="example", data="toy1"),
    "w2_reg_neg": dict(cls="Regressor", yaml="F1a", data="toy1", use_ocbnn=True, constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(-0.3, 0.3), interval_func=if_f1a,
                                              prior_type="negative_exponential_cocp"))]),
    "f3b_reg_neg2": dict(cls="Regressor", yaml="F3b", data="toy4", use_ocbnn=True, constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(-1.0, 1.0), interval_func=if_f3b,
                                              prior_type="negative_exponential_cocp"))]),
    "reg_multi_neg": dict(cls="Regressor", yaml="F1a", data="toy1", use_ocbnn=True,
                          overrides=dict(ocp_nsamples=37), constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(-0.3, 0.3), interval_func=if_slab,
                                              prior_type="negative_exponential_cocp")),
        ("add_deterministic_constraint", dict(constrained_domain=(-INF, -2.5), interval_func=if_touch,
                                              prior_type="negative_exponential_cocp")),
        ("add_deterministic_constraint", dict(constrained_domain=(2.2, INF), interval_func=if_f1a,
                                              prior_type="negative_exponential_cocp"))]),
    "f6_reg_pos": dict(cls="Regressor", yaml="F6", data="toy6", use_ocbnn=True, constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(s - 0.5, s + 0.5), interval_func=if_f6,
                                              prior_type="positive_gaussian_cocp")) for s in (3.89, -2.01, 15.03)]),
    "reg_pos_neg": dict(cls="Regressor", yaml="F6", data="toy6", use_ocbnn=True,
                        overrides=dict(ocp_nsamples=20), constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(3.39, 4.39), interval_func=if_f6,
                                              prior_type="positive_gaussian_cocp")),
        ("add_deterministic_constraint", dict(constrained_domain=(-1.0, 1.0), interval_func=if_f3b,
                                              prior_type="negative_exponential_cocp"))]),
    "reg_y2": dict(cls="Regressor", yaml="example", data="synth_reg_y2", overrides=dict(architecture=[12])),
    "reg_relu": dict(cls="Regressor", yaml="example", data="toy1", overrides=dict(activation="relu")),
    "reg_linear": dict(cls="Regressor", yaml="example", data="toy1", overrides=dict(activation="none")),
    "reg_deep": dict(cls="Regressor", yaml="F1a", data="toy1", use_ocbnn=True,
                     overrides=dict(architecture=[8, 6], ocp_nsamples=25), constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(-0.3, 0.3), interval_func=if_f1a,
                                              prior_type="negative_exponential_cocp"))]),
    # generic-net kernel, shapes that are not multiples of the 4 x 4 register tiles / 4 x 4 tile blocks: three hidden layers,
    # a short list next to a long one, one wide hidden layer (several hand-out rounds)
    "reg_odd3": dict(cls="Regressor", yaml="F1a", data="toy1", use_ocbnn=True,
                     overrides=dict(architecture=[21, 13, 9], ocp_nsamples=40), constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(-0.3, 0.3), interval_func=if_f1a,
                                              prior_type="negative_exponential_cocp"))]),
    "cls_odd2": dict(cls="Classifier", yaml="F2a", data="toy2", use_ocbnn=True,
                     overrides=dict(architecture=[35, 18], ocp_nsamples=30), constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(1.0, 3.0, -2.0, 0.0), forbidden_classes=[0, 2],
                                              prior_type="positive_dirichlet_cocp"))]),
    "bin_wide1": dict(cls="BinaryClassifier", yaml="F4", data="toy5", overrides=dict(architecture=[70])),
    "cls_vanilla": dict(cls="Classifier", yaml="cbakeoff", data="toy2"),
    "w3_cls_dir": dict(cls="Classifier", yaml="F2a", data="toy2", use_ocbnn=True, constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(1.0, 3.0, -2.0, 0.0), forbidden_classes=[0, 2],
                                              prior_type="positive_dirichlet_cocp"))]),
    "bin_vanilla": dict(cls="BinaryClassifier", yaml="F4", data="toy5"),
    "f4_bin_aocp": dict(cls="BinaryClassifier", yaml="F4", data="toy5", use_ocbnn=True, aocp="F4_aocp",
                        overrides=dict(ocp_nsamples=40), constraints=[
        ("add_probabilistic_constraint", dict(constrained_domain=(-0.5, 1.5, 0.0, 1.0), prob_func=pf_x2,
                                              prior_type="gaussian_aocp"))]),
    "f3a_reg_aocp": dict(cls="Regressor", yaml="F3a", data="toy3", use_ocbnn=True, aocp="F3a_aocp",
                         overrides=dict(ocp_nsamples=30), constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=(-3.0, 3.0), interval_func=if_f3a,
                                              prior_type="gaussian_aocp"))]),
    "credit_small": dict(cls="Classifier", yaml="credit", data="synth_credit", use_ocbnn=True,
                         overrides=dict(nbatches=6, ocp_nsamples=50), constraints=[
        ("add_deterministic_constraint", dict(constrained_domain=CREDIT_BOX, forbidden_classes=[1],
                                              prior_type="positive_dirichlet_cocp"))]),
    "recid_small": dict(cls="BinaryClassifier", yaml="recid", data="synth_recid", use_ocbnn=True, aocp="random",
                        overrides=dict(nbatches=5, ocp_nsamples=6), constraints=[
        ("add_probabilistic_constraint", dict(constrained_domain=RECID_BOX, prob_func=pf_x2,
                                              prior_type="gaussian_aocp"))]),
}

# cases small enough for sampler trajectories from the (slow) reference
SAMPLER_CASES = ["w1_reg_vanilla", "w2_reg_neg", "f3b_reg_neg2", "f6_reg_pos", "cls_vanilla", "w3_cls_dir",
                 "bin_vanilla", "f4_bin_aocp", "reg_deep", "reg_y2"]
BATCHED_SAMPLER_CASES = ["credit_small", "recid_small"]


def build_model(case_name, api_module, config_path, dataset, seed=0, sampler="HMC", uid=None):
    """Construct + load + register constraints exactly the way the run scripts do.

    `api_module` is either the reference `bnn` package or this repo's host
    package; both expose BNN<Sampler><cls>.  RNG: torch.manual_seed(seed) right
    before load(), so the initial weights and every constraint point are drawn
    from the same stream on both sides.
    """
    import torch
    case = CASES[case_name]
    klass = getattr(api_module, f"BNN{sampler}{case['cls']}")
    torch.manual_seed(seed)
    m = klass(uid=uid or f"{case_name}-{sampler}", configfile=config_path)
    m.load(**dataset)
    for meth, kw in case.get("constraints", []):
        getattr(m, meth)(**kw)
    return m
