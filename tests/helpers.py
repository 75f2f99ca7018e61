"""Shared helpers of the parity tests: golden loading, oracle Spec construction, model construction."""
import json
import os

import numpy as np
import torch
import yaml

import cases
import ocbnn_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_config(g):
    return json.loads(str(g["config_json"]))


def golden_dataset(g):
    Y = torch.tensor(g["ds_Y_train"])
    return {"dataset_name": "golden", "X_train": torch.tensor(g["ds_X_train"]), "Y_train": Y,
            "X_train_min": [float(v) for v in g["ds_X_train_min"]], "X_train_max": [float(v) for v in g["ds_X_train_max"]]}


TASK = {"Regressor": "regression", "Classifier": "classifier", "BinaryClassifier": "binary"}


def spec_from_golden(name, g, use_ocbnn=None, naive_sigmoid=True):
    """Oracle Spec built from the fixture arrays and the case's own callbacks (independent of the host package)."""
    case = cases.CASES[name]
    cfg = golden_config(g)
    X, Y = g["ds_X_train"], g["ds_Y_train"]
    task = TASK[case["cls"]]
    sK = {"regression": Y.shape[1] if Y.ndim == 2 else 1, "classifier": int(Y.max()) + 1, "binary": 1}[task]
    sizes = [X.shape[1]] + list(cfg["architecture"]) + [sK]
    S = cfg["ocp_nsamples"]
    neg = pos = dr = None
    cons = case.get("constraints", [])
    negc = [kw for _, kw in cons if kw.get("prior_type") == "negative_exponential_cocp"]
    posc = [kw for _, kw in cons if kw.get("prior_type") == "positive_gaussian_cocp"]
    dirc = [kw for _, kw in cons if kw.get("prior_type") == "positive_dirichlet_cocp"]
    if negc:
        xs = g["cr_cr_neg_xsamples"]
        iv = []
        for c, kw in enumerate(negc):
            lst = kw["interval_func"](torch.tensor(xs[c * S:(c + 1) * S]))
            iv.append(np.array([[(float(a), float(b)) for a, b in row] for row in lst], dtype=np.float64))
        neg = dict(xs=xs, intervals=iv, S=S, gamma=cfg["cocp_expo_gamma"], tau=cfg["cocp_expo_tau"])
    if posc:
        pos = dict(xs=g["cr_cr_pos_xsamples"], ys=g["cr_cr_pos_ysamples"], ylens=list(g["cr_ylens"]), S=S,
                   sigma_c=cfg["cocp_gaussian_sigma_c"])
    if dirc:
        dr = dict(xs=g["cr_cr_xsamples"], forbidden=[kw["forbidden_classes"] for kw in dirc], S=S,
                  gamma_d=cfg["cocp_dirichlet_gamma"], alpha=cfg["cocp_dirichlet_alpha"])
    uo = bool(case.get("use_ocbnn", False)) if use_ocbnn is None else use_ocbnn
    return orc.Spec(sizes, cfg["activation"], task, X, Y, sigma_w=cfg["sigma_w"], sigma_noise=cfg["sigma_noise"], use_ocbnn=uo,
                    aocp_mean=g.get("aocp_mean"), aocp_std=g.get("aocp_std"), neg=neg, pos=pos, dir=dr, naive_sigmoid=naive_sigmoid)


def write_config(g, path="cfg.yaml", **overrides):
    cfg = golden_config(g)
    cfg.update(overrides)
    with open(path, "w") as f:
        yaml.dump(cfg, f)
    return path


def make_model(name, g, sampler="HMC", **overrides):
    """This repo's host model for a golden case, built exactly like the generator built the reference model."""
    import ocbnn_b200
    case = cases.CASES[name]
    path = write_config(g, **overrides)
    m = cases.build_model(name, ocbnn_b200, path, golden_dataset(g), seed=0, sampler=sampler)
    if case.get("aocp"):
        torch.save(torch.tensor(g["aocp_params"]), "aocp_params.pt")
        m.load_gaussian_aocp_parameters("aocp_params.pt")
    m.update_config(use_ocbnn=bool(case.get("use_ocbnn", False)))
    return m


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (rows = weight vectors)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.ndim <= 1:
        return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    ax = tuple(range(1, a.ndim))
    return np.sqrt(((a - b) ** 2).sum(axis=ax)) / np.maximum(np.sqrt((b ** 2).sum(axis=ax)), 1e-30)
