"""Float64 numpy evaluation of a CompiledProblem's VALUE (the formula in include/ocbnn_b200.h) — test infrastructure that
checks the constraint compiler's tables on a machine without a GPU.  Never imported by the product."""
import numpy as np

import ocbnn_oracle as orc


def _spec(cp):
    act = {0: "none", 1: "rbf", 2: "relu"}[cp.activation]
    return orc.Spec(cp.layer_sizes, act, "regression", cp.X, cp.Y)


def _head(kind, f, prm, cp):
    t0, t1 = cp.tau
    if kind == 0:
        r = f - prm[:len(f)]
        return float(-(r * r).sum() * cp.lik_scale)
    if kind == 1:
        n = int(prm[0])
        tot = 0.0
        phi = lambda z: 0.25 * (np.tanh(-t0 * z) + 1) * (np.tanh(-t1 * z) + 1)
        for k in range(n):
            a, b = float(prm[1 + 2 * k]), float(prm[2 + 2 * k])
            tot += phi(a - f[0]) * phi(f[0] - b)
        return float(tot)
    if kind == 2:
        K = int(prm[0])
        t = np.array([-0.5 * (f[0] - float(prm[1 + k])) ** 2 * cp.pos_inv_var for k in range(K)])
        return float(t.max() + np.log(np.exp(t - t.max()).sum()))
    logp = f - f.max() - np.log(np.exp(f - f.max()).sum())
    if kind == 3:
        return float((prm[:len(f)] * logp).sum())
    if kind == 4:
        return float(logp[int(prm[0])])
    z = float(f[0])
    return float(prm[0] * z - np.logaddexp(0.0, z))


def logpost_value(cp, W, batch=(0, 1), with_data=True):
    W = np.asarray(W, dtype=np.float64)
    spec = _spec(cp)
    out = np.zeros(W.shape[0])
    mean = cp.prior_mean_vec if cp.prior_mean_vec is not None else cp.prior_mean
    std = cp.prior_std_vec if cp.prior_std_vec is not None else cp.prior_std
    out += cp.prior_const - (((W - mean) ** 2) / (2 * np.asarray(std, dtype=np.float64) ** 2)).sum(axis=1)
    if cp.cx.shape[0]:
        F = orc.forward(spec, W, cp.cx.astype(np.float64), dtype=np.float64)
        for m in range(W.shape[0]):
            for c in range(cp.cx.shape[0]):
                out[m] += float(cp.cweight[c]) * _head(int(cp.ckind[c]), F[m, c], cp.cparam[c].astype(np.float64), cp)
    if with_data:
        N = cp.X.shape[0]
        idx = np.arange(batch[0], N, batch[1]) if batch[1] > 1 else np.arange(N)
        F = orc.forward(spec, W, cp.X[idx].astype(np.float64), dtype=np.float64)
        mult = N / len(idx)
        for m in range(W.shape[0]):
            s = sum(_head(cp.lik_head, F[m, j], cp.Y[i].astype(np.float64), cp) for j, i in enumerate(idx))
            out[m] += mult * s + cp.lik_const
    return out
