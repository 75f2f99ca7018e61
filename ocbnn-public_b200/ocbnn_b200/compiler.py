"""
Constraint compiler: freezes a model (config + data + registered constraints + the constraint points already
sampled by add_*_constraint) into flat tables the CUDA engine consumes (ocbnn_desc, include/ocbnn_b200.h).

The reference re-invokes the user's Python callbacks (`interval_func`, `prob_func`) inside EVERY prior
evaluation (bnn/base.py:373); here they are evaluated once, on the fixed constraint points, and the result is
stored per point.  What is reproduced from the reference:

  * log_prior dispatch (base.py:106-121): with use_ocbnn and AOCP parameters loaded, the prior is ONLY
    N(aocp_mean, aocp_std); otherwise isotropic + every registered COCP.
  * negative exponential COCP (base.py:364-388): the forbidden gaps between the permitted intervals; the
    test `all(marker < lbs)` is a JOINT decision over the S points of one constraint; the total is divided by
    S * n_constraints and multiplied by -gamma.
  * positive Gaussian COCP (base.py:344-358): sigma_c enters as a VARIANCE; uniform mixture over the K targets.
  * positive Dirichlet COCP (base.py:498-511): concentration gamma for allowed, gamma*alpha for forbidden classes.
  * likelihoods (base.py:252-265, 432-444, 528-542); multi-output regression uses variance = sigma_noise.
"""

import math
from dataclasses import dataclass, field

import numpy as np
import torch

from . import engine as E

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


@dataclass
class CompiledProblem:
    layer_sizes: list
    activation: int
    lik_head: int
    lik_scale: float
    lik_const: float
    prior_mean: float
    prior_std: float
    prior_mean_vec: object
    prior_std_vec: object
    prior_const: float
    X: np.ndarray
    Y: np.ndarray
    cx: np.ndarray
    ckind: np.ndarray
    cweight: np.ndarray
    cparam: np.ndarray
    tau: tuple = (0.0, 0.0)
    pos_inv_var: float = 0.0
    info: dict = field(default_factory=dict)

    @property
    def nweights(self):
        return sum((m + 1) * n for m, n in zip(self.layer_sizes[:-1], self.layer_sizes[1:]))


def activation_code(name):
    return {"rbf": E.ACT_RBF, "relu": E.ACT_RELU}.get(name, E.ACT_IDENTITY)


def _np(t, dtype=np.float32):
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def forbidden_gaps(intervals):
    """Gaps (a, b) penalised by the negative COCP for ONE constraint.

    intervals: interval_func(xs) = list (S) of lists (K) of (lb, ub).  Returns (ngaps, S, 2) float32.
    Walks base.py:374-384: marker = -inf; for interval k: if all(marker < lb_k) -> gap (marker, lb_k);
    marker = ub_k; finally if all(marker < +inf) -> gap (marker, +inf).
    """
    S = len(intervals)
    K = len(intervals[0])
    iv = np.array([[(float(lb), float(ub)) for lb, ub in row] for row in intervals], dtype=np.float32).reshape(S, K, 2)
    marker = np.full((S,), -np.inf, dtype=np.float32)
    gaps = []
    for k in range(K):
        lbs, ubs = iv[:, k, 0], iv[:, k, 1]
        if np.all(marker < lbs):
            gaps.append(np.stack([marker, lbs], axis=1))
        marker = ubs
    if np.all(marker < np.inf):
        gaps.append(np.stack([marker, np.full((S,), np.inf, dtype=np.float32)], axis=1))
    return np.stack(gaps, axis=0) if gaps else np.zeros((0, S, 2), dtype=np.float32)


def compile_problem(m):
    """Build the CompiledProblem of model `m` for its CURRENT config/data/constraints."""
    sizes = [int(s) for s in m.layer_sizes]
    D = int(m.nweights)
    sK = sizes[-1]
    X = _np(m.X_train)
    N = X.shape[0]
    task = m._task

    # ---- likelihood -----------------------------------------------------------------------------------------
    if task == "regression":
        Y = _np(m.Y_train).reshape(N, -1)
        sn = float(m.sigma_noise)
        if sK == 1:
            lik_scale = 1.0 / (2.0 * sn * sn)
            lik_const = N * (-math.log(sn) - LOG_SQRT_2PI)
        else:
            lik_scale = 1.0 / (2.0 * sn)
            lik_const = N * (-0.5 * sK * (math.log(2.0 * math.pi) + math.log(sn)))
        lik_head = E.HEAD_LIK_GAUSS
    elif task == "classifier":
        Y = _np(m.Y_train).reshape(N, 1)
        lik_scale, lik_const, lik_head = 0.0, 0.0, E.HEAD_LIK_SOFTMAX
    else:
        Y = _np(m.Y_train).reshape(N, 1)
        lik_scale, lik_const, lik_head = 0.0, 0.0, E.HEAD_LIK_BERNOULLI

    # ---- prior over the weights -------------------------------------------------------------------------------
    cx, ckind, cweight, cparams = [], [], [], []
    pmv = psv = None
    info = {}
    tau = (0.0, 0.0)
    pos_inv_var = 0.0
    if m.use_ocbnn and m.aocp_mean is not None:
        pmv, psv = _np(m.aocp_mean), _np(m.aocp_std)
        prior_const = float(-np.log(psv.astype(np.float64)).sum() - D * LOG_SQRT_2PI)
        pmean, pstd = 0.0, 1.0
        info["prior"] = "aocp"
    else:
        pmean, pstd = 0.0, float(m.sigma_w)
        prior_const = D * (-math.log(pstd) - LOG_SQRT_2PI)
        info["prior"] = "isotropic"
        if m.use_ocbnn:
            S = int(m.ocp_nsamples)
            if "positive_gaussian_cocp" in m.dconstraints:
                sc = float(m.cocp_gaussian_sigma_c)
                pos_inv_var = 1.0 / sc
                xs, ys = _np(m._cr_pos_xsamples), _np(m._cr_pos_ysamples)
                idx = 0
                for K in m._cr_ylens:
                    K = int(K)
                    for s in range(S):
                        cx.append(xs[idx + s])
                        ckind.append(E.HEAD_POS_GAUSS)
                        cweight.append(1.0)
                        cparams.append([float(K)] + [float(ys[idx + k * S + s, 0]) for k in range(K)])
                    prior_const += S * (-0.5 * math.log(2.0 * math.pi * sc) + math.log(1.0 / K))
                    idx += K * S
                info["pos"] = len(m._cr_ylens)
            if "negative_exponential_cocp" in m.dconstraints:
                cons = m.dconstraints["negative_exponential_cocp"]
                nc = len(cons)
                tau = (float(m.cocp_expo_tau[0]), float(m.cocp_expo_tau[1]))
                wgt = -float(m.cocp_expo_gamma) / float(S * nc)
                xs_t = m._cr_neg_xsamples
                for c, (_, ifunc) in enumerate(cons):
                    xc = xs_t[c * S:(c + 1) * S, :]
                    gaps = forbidden_gaps(ifunc(xc))            # the user callback runs ONCE, here
                    xcn = _np(xc)
                    for s in range(S):
                        cx.append(xcn[s])
                        ckind.append(E.HEAD_NEG_EXP)
                        cweight.append(wgt)
                        cparams.append([float(gaps.shape[0])] + [float(v) for g in gaps[:, s, :] for v in g])
                info["neg"] = nc
            if "positive_dirichlet_cocp" in m.dconstraints:
                cons = m.dconstraints["positive_dirichlet_cocp"]
                xs = _np(m._cr_xsamples)
                gd, al = float(m.cocp_dirichlet_gamma), float(m.cocp_dirichlet_alpha)
                for c, (_, fclasses) in enumerate(cons):
                    cnt = np.bincount(np.asarray(fclasses, dtype=np.int64), minlength=sK).astype(np.float64)
                    conc = gd - (gd * cnt) * (1.0 - al)
                    norm = math.lgamma(conc.sum()) - sum(math.lgamma(a) for a in conc)
                    for s in range(S):
                        cx.append(xs[c * S + s])
                        ckind.append(E.HEAD_DIRICHLET)
                        cweight.append(1.0)
                        cparams.append([float(a - 1.0) for a in conc])
                    prior_const += S * norm
                info["dir"] = len(cons)

    C = len(cx)
    stride = max([len(p) for p in cparams], default=0)
    cparam = np.zeros((C, stride), dtype=np.float32)
    for i, p in enumerate(cparams):
        cparam[i, :len(p)] = p
    return CompiledProblem(
        layer_sizes=sizes, activation=activation_code(m.activation), lik_head=lik_head, lik_scale=float(lik_scale),
        lik_const=float(lik_const), prior_mean=pmean, prior_std=pstd, prior_mean_vec=pmv, prior_std_vec=psv,
        prior_const=float(prior_const), X=X, Y=Y,
        cx=np.asarray(cx, dtype=np.float32).reshape(C, sizes[0]), ckind=np.asarray(ckind, dtype=np.int32),
        cweight=np.asarray(cweight, dtype=np.float32), cparam=cparam, tau=tau, pos_inv_var=float(pos_inv_var), info=info)


def batch_spec(batch_indices, N):
    """(offset, stride) of a `torch.arange(t % nbatches, N, nbatches)` index tensor; None -> full batch."""
    if batch_indices is None:
        return (0, 1)
    bi = torch.as_tensor(batch_indices).long().cpu()
    if bi.numel() == 0:
        raise ValueError("empty batch")
    off = int(bi[0])
    stride = int(bi[1] - bi[0]) if bi.numel() > 1 else max(N - off, off + 1, 2)
    if stride < 1 or not torch.equal(bi, torch.arange(off, N, stride)):
        raise NotImplementedError("the engine evaluates strided batches arange(offset, N, stride) only (inference.py:180,276,334)")
    if stride == 1:
        if off != 0:
            raise NotImplementedError("contiguous batches must start at 0")
        return (0, 1)
    return (off, stride)
