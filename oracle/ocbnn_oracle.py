"""
ocbnn_oracle — CPU restatement (numpy) of the OC-BNN posterior-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module.  It is used by `tests/`, by `__graft_entry__.smoke()` and by
`bench.py`'s `cpu_baseline` / `--impl reference` legs as the *checker* and as
the CPU baseline, never as the thing that is measured as the product or
shipped.

What it restates (all citations are into the reference tree, dtak/ocbnn-public):

  forward                    bnn/base.py:72-104
  isotropic / AOCP prior     bnn/base.py:106-133
  Gaussian likelihood        bnn/base.py:252-265
  softmax likelihood         bnn/base.py:432-444
  Bernoulli likelihood       bnn/base.py:524-542
  positive Gaussian COCP     bnn/base.py:344-358
  negative exponential COCP  bnn/base.py:360-388
  positive Dirichlet COCP    bnn/base.py:498-511
  AOCP objectives            bnn/base.py:390-419 (regression), 624-672 (binary)
  HMC leapfrog + MH          bnn/inference.py:29-78
  BBB reparameterised ELBO   bnn/inference.py:132-153
  SVGD bandwidth/kernel/phi  bnn/inference.py:208-259 (+ Adagrad :272-281)
  SGLD step                  bnn/inference.py:300-340

The reference differentiates with autograd; this file uses closed-form
gradients instead (no autograd anywhere), evaluated for M weight vectors at a
time.  Third-party arithmetic that the reference takes from PyTorch (pinned
`torch==1.5.0` in requirements.txt; semantics identical in the torch 2.11 used
to generate the fixtures) is restated from its published definition:
Normal/MVN/Dirichlet.log_prob, LogSoftmax, PairwiseDistance(p=2, eps=1e-6),
Tensor.median (lower median), optim.Adagrad (eps=1e-10, lr_decay=0).

Parity pin: the reference has no tests and no golden vectors of its own
(SURVEY.md §4).  This oracle is pinned against outputs of the unmodified
reference run in the build container: `oracle/make_golden.py` imports
`/root/reference`, runs it on seeded inputs and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function here against those files.

`dtype` selects the arithmetic type: np.float32 mirrors the reference, while
np.float64 is the "more exact than the reference" arm used to bound the
reference's own rounding noise in the parity tests.
"""

import math
import numpy as np

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


# --------------------------------------------------------------------------
# problem specification
# --------------------------------------------------------------------------

class Spec:
    """Plain container describing one model + prior + data set.

    layer_sizes : [s0, ..., sK]
    activation  : 'rbf' | 'relu' | anything else = identity   (base.py:79-87)
    task        : 'regression' | 'classifier' | 'binary'
    X_train (N,s0) float ; Y_train (N,Ydim) float or (N,) int
    use_ocbnn, aocp_mean, aocp_std                            (base.py:106-121)
    neg : dict(xs (nc*S,s0), intervals list[nc] of (S,K,2), S, gamma, tau)
    pos : dict(xs (tot,s0), ys (tot,1), ylens list, S, sigma_c)
    dir : dict(xs (nc*S,s0), forbidden list[nc] of lists, S, gamma_d, alpha)
    naive_sigmoid : restate base.py:524-526 literally (NaN hazards included)
    """

    def __init__(self, layer_sizes, activation, task, X_train, Y_train,
                 sigma_w=1.0, sigma_noise=0.1, use_ocbnn=False,
                 aocp_mean=None, aocp_std=None, neg=None, pos=None, dir=None,
                 naive_sigmoid=True):
        self.layer_sizes = [int(s) for s in layer_sizes]
        self.activation = activation
        self.task = task
        self.X_train = np.asarray(X_train)
        self.Y_train = np.asarray(Y_train)
        self.sigma_w = float(sigma_w)
        self.sigma_noise = float(sigma_noise)
        self.use_ocbnn = bool(use_ocbnn)
        self.aocp_mean = None if aocp_mean is None else np.asarray(aocp_mean)
        self.aocp_std = None if aocp_std is None else np.asarray(aocp_std)
        self.neg, self.pos, self.dir = neg, pos, dir
        self.naive_sigmoid = naive_sigmoid

    @property
    def layer_shapes(self):
        return list(zip(self.layer_sizes[:-1], self.layer_sizes[1:]))

    @property
    def nweights(self):
        return sum((m + 1) * n for m, n in self.layer_shapes)

    @property
    def N_train(self):
        return self.Y_train.shape[0]


# --------------------------------------------------------------------------
# MLP forward / vector-Jacobian product            (base.py:72-104)
# --------------------------------------------------------------------------

def _unpack(spec, W):
    """Packed layout: per layer W_l (m,n) row-major then b_l (n)  (base.py:72-77)."""
    M = W.shape[0]
    off = 0
    for m, n in spec.layer_shapes:
        Wl = W[:, off:off + m * n].reshape(M, m, n)
        bl = W[:, off + m * n:off + m * n + n].reshape(M, 1, n)
        yield Wl, bl, off
        off += (m + 1) * n


def _act(spec, z):
    if spec.activation == "rbf":
        return np.exp(-(z * z))
    if spec.activation == "relu":
        return np.maximum(z, 0)
    return z


def _dact(spec, z, a):
    """d activation / d z given pre-activation z and activation a."""
    if spec.activation == "rbf":
        return a * (-2 * z)
    if spec.activation == "relu":
        return (z > 0).astype(z.dtype)
    return np.ones_like(z)


def forward(spec, W, X, dtype=np.float32, keep=False):
    """f(X; W) for M packed weight vectors: (M,D),(P,s0) -> (M,P,sK)."""
    W = np.asarray(W, dtype=dtype)
    if W.ndim == 1:
        W = W[None]
    A = np.broadcast_to(np.asarray(X, dtype=dtype), (W.shape[0],) + X.shape)
    cache = []
    nl = len(spec.layer_shapes)
    for li, (Wl, bl, off) in enumerate(_unpack(spec, W)):
        Z = np.einsum('mnd,mdo->mno', A, Wl).astype(dtype) + bl
        cache.append((A, Z, Wl, off))
        A = _act(spec, Z) if li < nl - 1 else Z
    return (Z, cache) if keep else Z


def _vjp(spec, cache, dF, G):
    """Accumulate dL/dW into G (M,D) given dL/df (M,P,sK)."""
    nl = len(cache)
    dZ = dF
    for li in range(nl - 1, -1, -1):
        A, Z, Wl, off = cache[li]
        m, n = Wl.shape[1], Wl.shape[2]
        G[:, off:off + m * n] += np.einsum('mnd,mno->mdo', A, dZ).reshape(G.shape[0], m * n)
        G[:, off + m * n:off + m * n + n] += dZ.sum(axis=1)
        if li > 0:
            dA = np.einsum('mno,mdo->mnd', dZ, Wl)
            Ap, Zp, _, _ = cache[li - 1]
            dZ = dA * _dact(spec, Zp, A)


def _points_term(spec, W, X, head, G, dtype):
    """value (M,) of head(f(X;W)) and its gradient accumulated into G."""
    F, cache = forward(spec, W, X, dtype=dtype, keep=True)
    val, dF = head(F)
    if G is not None:
        _vjp(spec, cache, dF.astype(dtype), G)
    return val


# --------------------------------------------------------------------------
# priors
# --------------------------------------------------------------------------

def gaussian_prior(W, mean, std, dtype=np.float32):
    """Normal(mean,std).log_prob(w).sum()   (base.py:127-133)."""
    W = np.asarray(W, dtype=dtype)
    mean = np.asarray(mean, dtype=dtype)
    std = np.asarray(std, dtype=dtype)
    var = std * std
    lp = -((W - mean) ** 2) / (2 * var) - np.log(std) - dtype(LOG_SQRT_2PI)
    return lp.sum(axis=1), -(W - mean) / var


def _phi(tau, z):
    """Sigmoidal constraint classifier (base.py:360-362) and its derivative."""
    t0 = np.tanh(-tau[0] * z)
    t1 = np.tanh(-tau[1] * z)
    val = 0.25 * (t0 + 1) * (t1 + 1)
    d = 0.25 * (-tau[0] * (1 - t0 * t0) * (t1 + 1) - tau[1] * (1 - t1 * t1) * (t0 + 1))
    return val, d


def neg_exponential_head(spec, dtype):
    """negative_exponential_cocp (base.py:364-388) as value/df of f (M,C,1)."""
    neg = spec.neg
    S, gamma = int(neg['S']), dtype(neg['gamma'])
    tau = (dtype(neg['tau'][0]), dtype(neg['tau'][1]))
    nc = len(neg['intervals'])

    def head(F):
        f = F[:, :, 0]
        M = f.shape[0]
        penalty = np.zeros((M, S), dtype=dtype)          # ONE vector shared by all constraints
        dpen = np.zeros_like(f)
        for c in range(nc):
            iv = np.asarray(neg['intervals'][c], dtype=dtype)   # (S,K,2)
            fc = f[:, c * S:(c + 1) * S]
            marker = np.full((S,), -np.inf, dtype=dtype)
            gaps = []
            for k in range(iv.shape[1]):
                lbs, ubs = iv[:, k, 0], iv[:, k, 1]
                if np.all(marker < lbs):                         # joint decision (:378)
                    gaps.append((marker, lbs))
                marker = ubs
            infs = np.full((S,), np.inf, dtype=dtype)
            if np.all(marker < infs):                            # (:383)
                gaps.append((marker, infs))
            for a, b in gaps:
                with np.errstate(invalid='ignore', over='ignore'):
                    p1, d1 = _phi(tau, a[None, :] - fc)
                    p2, d2 = _phi(tau, fc - b[None, :])
                penalty += p1 * p2
                dpen[:, c * S:(c + 1) * S] += -d1 * p2 + p1 * d2
        scale = -gamma / dtype(S * nc)
        return scale * penalty.sum(axis=1), (scale * dpen)[:, :, None]
    return head


def pos_gaussian_head(spec, dtype):
    """positive_gaussian_cocp (base.py:344-358): 1-D MVN with COVARIANCE sigma_c."""
    pos = spec.pos
    S, sc = int(pos['S']), dtype(pos['sigma_c'])
    ys = np.asarray(pos['ys'], dtype=dtype)[:, 0]
    ylens = [int(k) for k in pos['ylens']]

    def head(F):
        f = F[:, :, 0]
        M = f.shape[0]
        val = np.zeros((M,), dtype=dtype)
        dF = np.zeros_like(f)
        idx = 0
        for K in ylens:
            n = K * S
            diff = f[:, idx:idx + n] - ys[None, idx:idx + n]
            ll = (-0.5 * diff * diff / sc - dtype(0.5 * math.log(2 * math.pi)) - dtype(0.5) * np.log(sc)
                  + np.log(dtype(1.0 / K))).astype(dtype)
            ll = ll.reshape(M, K, S)
            mx = ll.max(axis=1, keepdims=True)
            ex = np.exp(ll - mx)
            se = ex.sum(axis=1, keepdims=True)
            val += (mx[:, 0] + np.log(se[:, 0])).sum(axis=1)
            wts = (ex / se).reshape(M, n)
            dF[:, idx:idx + n] = wts * (-diff / sc)
            idx += n
        return val, dF[:, :, None]
    return head


def dirichlet_conc(spec, c):
    """Concentration vector of constraint c (base.py:507-508)."""
    d = spec.dir
    sK = spec.layer_sizes[-1]
    cnt = np.bincount(np.asarray(d['forbidden'][c], dtype=np.int64), minlength=sK).astype(np.float64)
    return d['gamma_d'] - (d['gamma_d'] * cnt) * (1 - d['alpha'])


def dirichlet_head(spec, dtype):
    """positive_dirichlet_cocp (base.py:498-511): Dirichlet(conc).log_prob(softmax(f))."""
    d = spec.dir
    S = int(d['S'])
    nc = len(d['forbidden'])

    def head(F):
        M = F.shape[0]
        val = np.zeros((M,), dtype=dtype)
        dF = np.zeros_like(F)
        for c in range(nc):
            conc = dirichlet_conc(spec, c).astype(dtype)
            f = F[:, c * S:(c + 1) * S, :]
            mx = f.max(axis=2, keepdims=True)
            ex = np.exp(f - mx)
            se = ex.sum(axis=2, keepdims=True)
            p = ex / se
            with np.errstate(divide='ignore'):
                logp = np.log(p)                                   # torch: xlogy(conc-1, softmax)
            norm = dtype(math.lgamma(float(conc.sum())) - sum(math.lgamma(float(a)) for a in conc))
            cm1 = conc - 1
            val += (np.where(cm1 == 0, 0, cm1 * logp).sum(axis=2) + norm).sum(axis=1)
            dF[:, c * S:(c + 1) * S, :] = cm1[None, None, :] - cm1.sum() * p
        return val, dF
    return head


def log_prior(spec, W, dtype=np.float32, grad=True):
    """log_prior dispatcher (base.py:106-121)."""
    W = np.asarray(W, dtype=dtype)
    if W.ndim == 1:
        W = W[None]
    if spec.use_ocbnn and spec.aocp_mean is not None:
        lp, g = gaussian_prior(W, spec.aocp_mean, spec.aocp_std, dtype)
        return lp, (g if grad else None)
    lp, g = gaussian_prior(W, 0.0, spec.sigma_w, dtype)
    G = g.copy() if grad else None
    if spec.use_ocbnn:
        if spec.pos is not None:
            lp = lp + _points_term(spec, W, np.asarray(spec.pos['xs']), pos_gaussian_head(spec, dtype), G, dtype)
        if spec.neg is not None:
            lp = lp + _points_term(spec, W, np.asarray(spec.neg['xs']), neg_exponential_head(spec, dtype), G, dtype)
        if spec.dir is not None:
            lp = lp + _points_term(spec, W, np.asarray(spec.dir['xs']), dirichlet_head(spec, dtype), G, dtype)
    return lp, G


# --------------------------------------------------------------------------
# likelihoods
# --------------------------------------------------------------------------

def batch_indices(N, offset, stride):
    """torch.arange(t % nbatches, N_train, nbatches)  (inference.py:180,276,334)."""
    return np.arange(offset, N, stride)


def log_likelihood(spec, W, batch=None, dtype=np.float32, grad=True):
    W = np.asarray(W, dtype=dtype)
    if W.ndim == 1:
        W = W[None]
    X, Y = spec.X_train, spec.Y_train
    mult = dtype(1)
    if batch is not None:
        X, Y = X[batch], Y[batch]
        mult = dtype(spec.N_train / len(batch))
    G = np.zeros_like(W) if grad else None

    if spec.task == 'regression':
        Yt = np.asarray(Y, dtype=dtype)
        if spec.layer_sizes[-1] == 1:                              # base.py:264
            s = dtype(spec.sigma_noise)

            def head(F):
                r = F - Yt[None]
                v = (-(r * r) / (2 * s * s) - np.log(s) - dtype(LOG_SQRT_2PI)).sum(axis=(1, 2))
                return mult * v, mult * (-r / (s * s))
        else:                                                       # base.py:265 (variance = sigma_noise)
            s = dtype(spec.sigma_noise)
            k = spec.layer_sizes[-1]

            def head(F):
                r = F - Yt[None]
                v = (-0.5 * (r * r).sum(axis=2) / s - dtype(0.5 * k * math.log(2 * math.pi))
                     - dtype(0.5 * k) * np.log(s)).sum(axis=1)
                return mult * v, mult * (-r / s)
    elif spec.task == 'classifier':                                # base.py:432-444
        Yi = np.asarray(Y, dtype=np.int64)

        def head(F):
            mx = F.max(axis=2, keepdims=True)
            ex = np.exp(F - mx)
            se = ex.sum(axis=2, keepdims=True)
            logp = F - mx - np.log(se)
            idx = np.arange(len(Yi))
            v = logp[:, idx, Yi].sum(axis=1)
            dF = -(ex / se)
            dF[:, idx, Yi] += 1
            return mult * v, mult * dF
    else:                                                           # base.py:524-542
        t = np.asarray(Y, dtype=dtype)

        def head(F):
            z = F[:, :, 0]
            if spec.naive_sigmoid:
                with np.errstate(over='ignore', invalid='ignore', divide='ignore'):
                    # literal restatement of base.py:524-526,541 AND of what autograd differentiates:
                    # p = e/(e+1), d/dp [t log p + (1-t) log(1-p)], dp/de = 1/(e+1) - e/(e+1)^2, de/dz = e.
                    # In fp32 the (1-p) cancellation makes this noisy for |z| >~ 8; in float64 it equals t - sigmoid(z).
                    e = np.exp(z)
                    p = e / (e + dtype(1.0))
                    v = (t * np.log(p) + (dtype(1.0) - t) * np.log(1 - p)).sum(axis=1)
                    dp = np.where(t[None] != 0, t[None] / p, 0) - np.where(t[None] != 1, (1 - t[None]) / (1 - p), 0)
                    dz = dp * (e * (1 / (e + 1) - e / ((e + 1) * (e + 1))))
            else:
                v = (t * z - np.logaddexp(0, z)).sum(axis=1)
                dz = t[None] - 1 / (1 + np.exp(-z))
            return mult * v, (mult * dz)[:, :, None]

    val = _points_term(spec, W, X, head, G, dtype)
    return val, G


def log_posterior(spec, W, batch=None, with_data=True, dtype=np.float32):
    """(lp (M,), grad (M,D)) of log_prior + log_likelihood   (base.py:123-125)."""
    lp, G = log_prior(spec, W, dtype)
    if with_data:
        ll, Gl = log_likelihood(spec, W, batch, dtype)
        lp, G = lp + ll, G + Gl
    return lp, G


# --------------------------------------------------------------------------
# HMC                                                  (inference.py:29-78)
# --------------------------------------------------------------------------

def hmc_iterate(spec, q, p0, u, eps, L, with_data=True, restore_on_reject=False, dtype=np.float32):
    """n_it HMC iterations for M chains with injected momenta p0 (n_it,M,D) and uniforms u (n_it,M).

    Returns (q_final, accept (n_it,M) bool, H (n_it,M,4) = U0,K0,U1,K1).
    The reference never restores the state on reject (original_q aliases the
    live storage, inference.py:46,78): restore_on_reject=False reproduces that.
    """
    q = np.array(q, dtype=dtype, copy=True)
    if q.ndim == 1:
        q = q[None]
    n_it = p0.shape[0]
    acc = np.zeros((n_it, q.shape[0]), dtype=bool)
    H = np.zeros((n_it, q.shape[0], 4), dtype=dtype)
    e, eh = dtype(eps), dtype(eps / 2)

    def gradU(x):
        return -log_posterior(spec, x, None, with_data, dtype)[1]

    def U(x):
        lp = log_posterior(spec, x, None, with_data, dtype)[0] if with_data else log_prior(spec, x, dtype, grad=False)[0]
        return -lp

    for it in range(n_it):
        p = np.array(p0[it], dtype=dtype, copy=True)
        q0 = q.copy()
        U0 = U(q)
        K0 = (0.5 * (p * p).sum(axis=1)).astype(dtype)
        p -= eh * gradU(q)
        for l in range(1, L + 1):
            q += e * p
            if l < L:
                p -= e * gradU(q)
        p -= eh * gradU(q)
        U1 = U(q)
        K1 = (0.5 * (p * p).sum(axis=1)).astype(dtype)
        with np.errstate(over='ignore'):
            alpha = np.exp((U0 - U1 + K0 - K1).astype(dtype)).astype(np.float64)
        a = np.asarray(u[it], dtype=np.float64) < alpha          # strict <, float64 compare (:74)
        acc[it] = a
        H[it] = np.stack([U0, K0, U1, K1], axis=1)
        if restore_on_reject:
            q[~a] = q0[~a]
    return q, acc, H


# --------------------------------------------------------------------------
# SGLD                                                 (inference.py:300-340)
# --------------------------------------------------------------------------

def sgld_stepsize(a, b, gamma, t):
    return a * ((b + t) ** (-1 * gamma))


def sgld_step(spec, W, noise, eps_t, batch=None, with_data=True, dtype=np.float32):
    """w += (eps/2)(grad log prior + grad log lik) + noise ; noise already ~N(0, eps)."""
    W = np.array(W, dtype=dtype, copy=True)
    if W.ndim == 1:
        W = W[None]
    upd = log_prior(spec, W, dtype)[1].copy()
    if with_data:
        upd += log_likelihood(spec, W, batch, dtype)[1]
    upd *= dtype(eps_t / 2)
    upd += np.asarray(noise, dtype=dtype)
    return W + upd


# --------------------------------------------------------------------------
# Adagrad (torch.optim.Adagrad defaults: lr_decay 0, eps 1e-10, acc0 0)
# --------------------------------------------------------------------------

def adagrad_step(x, acc, g, lr, dtype=np.float32):
    acc = (acc + g * g).astype(dtype)
    x = (x - dtype(lr) * g / (np.sqrt(acc) + dtype(1e-10))).astype(dtype)
    return x, acc


# --------------------------------------------------------------------------
# SVGD                                                 (inference.py:208-259)
# --------------------------------------------------------------------------

def svgd_rbf_h(X, dtype=np.float32):
    """h = med^2 / ln n ; med = LOWER median of ||x_i - x_j + 1e-6|| over i<j, zeros dropped."""
    X = np.asarray(X, dtype=dtype)
    n = X.shape[0]
    iu = np.triu_indices(n, k=1)
    diff = X[iu[0]] - X[iu[1]] + dtype(1e-6)
    d = np.sqrt((diff * diff).sum(axis=1)).astype(dtype)
    d = d[d != 0]
    d.sort()
    med = d[(len(d) - 1) // 2]
    return dtype(med * med) / dtype(math.log(n)) if dtype == np.float32 else med * med / math.log(n)


def svgd_phi(X, G, h, dtype=np.float32):
    """phi_i = (1/n) sum_j [K_ji g_j + grad_{x_j} K_ji], K_ji = exp(-||x_j-x_i||^2/h)."""
    X = np.asarray(X, dtype=dtype)
    G = np.asarray(G, dtype=dtype)
    n = X.shape[0]
    diff = X[:, None, :] - X[None, :, :]                 # [j,i,:] = x_j - x_i
    K = np.exp(-(diff * diff).sum(axis=2) / dtype(h))   # K[j,i]
    rep = (-(2 / dtype(h)) * K)[:, :, None] * diff       # grad_{x_j} K_ji
    return ((K.T @ G) + rep.sum(axis=0)) / dtype(n)


def svgd_step(spec, X, acc, lr, batch=None, with_data=True, dtype=np.float32):
    """One SVGD epoch + Adagrad update (inference.py:225-259, 272-281)."""
    if with_data:
        G = log_posterior(spec, X, batch, True, dtype)[1]
    else:
        G = log_prior(spec, X, dtype)[1]
    h = svgd_rbf_h(X, dtype)
    phi = svgd_phi(X, G, h, dtype)
    Xn, accn = adagrad_step(np.asarray(X, dtype=dtype), np.asarray(acc, dtype=dtype), -phi, lr, dtype)
    return Xn, accn, phi, h


# --------------------------------------------------------------------------
# BBB                                                  (inference.py:132-153)
# --------------------------------------------------------------------------

def softplus(r):
    return np.logaddexp(r, 0)


def bbb_elbo_grad(spec, q_params, eps, batch=None, with_data=True, dtype=np.float32):
    """ELBO and its closed-form gradient wrt (mu, rho); eps (k,D) injected."""
    qp = np.asarray(q_params, dtype=dtype)
    eps = np.asarray(eps, dtype=dtype)
    mu, rho = qp[:, 0], qp[:, 1]
    sd = softplus(rho).astype(dtype)
    W = mu[None] + eps * sd[None]
    if with_data:
        lp, G = log_posterior(spec, W, batch, True, dtype)
    else:
        lp, G = log_prior(spec, W, dtype)
    logq = (-(eps * eps) / 2 - np.log(sd)[None] - dtype(LOG_SQRT_2PI)).sum(axis=1)
    elbo = (lp - logq).mean()
    sig = 1 / (1 + np.exp(-rho))
    g_mu = G.mean(axis=0)
    g_rho = sig * ((G * eps).mean(axis=0) + 1 / sd)
    return dtype(elbo), np.stack([g_mu, g_rho], axis=1).astype(dtype)


def bbb_step(spec, q_params, acc, eps, lr, batch=None, with_data=True, dtype=np.float32):
    elbo, g = bbb_elbo_grad(spec, q_params, eps, batch, with_data, dtype)
    qn, accn = adagrad_step(np.asarray(q_params, dtype=dtype), np.asarray(acc, dtype=dtype), -g, lr, dtype)
    return qn, accn, elbo


# --------------------------------------------------------------------------
# AOCP objective                            (base.py:390-419, 624-672)
# --------------------------------------------------------------------------

def _jac_at(spec, mu, xs, dtype):
    """f(x; mu) (C,) and b = df/dmu (C,D) for scalar-output nets."""
    D = mu.shape[0]
    C = xs.shape[0]
    B = np.zeros((C, D), dtype=dtype)
    fvals = np.zeros((C,), dtype=dtype)
    for c in range(C):
        F, cache = forward(spec, mu[None], xs[c:c + 1], dtype=dtype, keep=True)
        Gc = np.zeros((1, D), dtype=dtype)
        _vjp(spec, cache, np.ones_like(F), Gc)
        B[c] = Gc[0]
        fvals[c] = F[0, 0, 0]
    return fvals, B


def _ncdf(x):
    return 0.5 * (1 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


def aocp_objective_binary(spec, params, det=None, prob=None, dtype=np.float32):
    """Binary AOCP objective (base.py:624-672) and its gradient wrt params (D,2).

    det  : dict(xs (C,s0), classes (C,))   deterministic constraints, |class - p|
    prob : dict(xs (C,s0), p (C,))         probabilistic constraints, symmetric Bernoulli KL
    b = df/dmu is evaluated at a DETACHED copy of mu (no second-order terms).
    """
    P = np.asarray(params, dtype=dtype)
    mu, rho = P[:, 0], P[:, 1]
    s = softplus(rho).astype(dtype)
    sig = 1 / (1 + np.exp(-rho))
    loss = dtype(0)
    g = np.zeros_like(P)
    denom = 0

    def accumulate(xs, fn):
        nonlocal loss, g
        _, B = _jac_at(spec, mu, np.asarray(xs, dtype=dtype), dtype)
        s2 = (B * B * (s * s)[None]).sum(axis=1)
        kappa = (1 + dtype(math.pi) * s2 / 8) ** dtype(-0.5)
        a = B @ mu
        z = kappa * a
        e = np.exp(z)
        p = e / (e + 1)
        val, dp = fn(p)
        loss += val.sum()
        dz = dp * p * (1 - p)
        g[:, 0] += (dz * kappa) @ B
        dk = -0.5 * (1 + dtype(math.pi) * s2 / 8) ** dtype(-1.5) * dtype(math.pi) / 8
        ds2 = dz * a * dk
        g[:, 1] += (ds2[:, None] * B * B).sum(axis=0) * 2 * s * sig

    if det is not None:
        cls = np.asarray(det['classes'], dtype=dtype)
        accumulate(det['xs'], lambda p: (np.abs(cls - p), -np.sign(cls - p)))
        denom += len(det['xs'])
    if prob is not None:
        p1 = np.clip(np.asarray(prob['p'], dtype=dtype), 0.001, 0.999)

        def kl(p):
            inside = (p > 0.001) & (p < 0.999)
            ph = np.clip(p, 0.001, 0.999)
            v = (p1 * (np.log(p1) - np.log(ph)) + (1 - p1) * (np.log(1 - p1) - np.log(1 - ph))
                 + ph * (np.log(ph) - np.log(p1)) + (1 - ph) * (np.log(1 - ph) - np.log(1 - p1)))
            d = (-p1 / ph + (1 - p1) / (1 - ph)
                 + (np.log(ph) - np.log(p1)) - (np.log(1 - ph) - np.log(1 - p1)))
            return v, np.where(inside, d, 0)
        accumulate(prob['xs'], kl)
        denom += len(prob['xs'])
    return dtype(loss / denom), (g / denom).astype(dtype)


def aocp_objective_regression(spec, params, xs, intervals, dtype=np.float32):
    """Regression AOCP objective (base.py:390-419) and gradient; Xdim = 1 only.

    intervals: list (C) of lists of (lb, ub).  sd = sqrt(sigma_noise + sigma^2)
    (sigma_noise enters UNSQUARED, :405); the mean f(x;mu) is differentiable.
    """
    P = np.asarray(params, dtype=dtype)
    mu, rho = P[:, 0], P[:, 1]
    s = softplus(rho).astype(dtype)
    sig = 1 / (1 + np.exp(-rho))
    xs = np.asarray(xs, dtype=dtype)
    fvals, B = _jac_at(spec, mu, xs, dtype)
    s2 = (B * B * (s * s)[None]).sum(axis=1)
    sd = np.sqrt(dtype(spec.sigma_noise) + s2)
    C = xs.shape[0]
    loss = 0.0
    dm = np.zeros((C,), dtype=dtype)
    dsd = np.zeros((C,), dtype=dtype)
    pdf = lambda t: np.exp(-0.5 * t * t) / math.sqrt(2 * math.pi)
    for c in range(C):
        permitted = 0.0
        for lb, ub in intervals[c]:
            # cdf(+inf) = 1, cdf(-inf) = 0 (the three branches of base.py:411-416)
            for bound, sign in ((ub, +1.0), (lb, -1.0)):
                if np.isinf(bound):
                    permitted += sign * (1.0 if bound > 0 else 0.0)
                    continue
                t = (bound - fvals[c]) / sd[c]
                permitted += sign * float(_ncdf(t))
                dm[c] += sign * pdf(t) / sd[c]              # d(1-permitted)/dm
                dsd[c] += sign * pdf(t) * t / sd[c]         # d(1-permitted)/dsd
        loss += 1 - permitted
    g = np.zeros_like(P)
    g[:, 0] = (dm[:, None] * B).sum(axis=0)
    ds2 = dsd / (2 * sd)
    g[:, 1] = (ds2[:, None] * B * B).sum(axis=0) * 2 * s * sig
    return dtype(loss / C), (g / C).astype(dtype)


# --------------------------------------------------------------------------
# predict                                   (base.py:267-277, 446-461, 544-559)
# --------------------------------------------------------------------------

def predict(spec, samples, domain, return_probs=False, dtype=np.float32):
    F = forward(spec, np.asarray(samples, dtype=dtype), np.asarray(domain, dtype=dtype), dtype=dtype)
    if spec.task == 'regression':
        return F
    if spec.task == 'classifier':
        mx = F.max(axis=2, keepdims=True)
        ex = np.exp(F - mx)
        p = ex / ex.sum(axis=2, keepdims=True)
        return p if return_probs else p.argmax(axis=2)
    e = np.exp(F[:, :, 0])
    p = e / (e + 1)
    return p if return_probs else p >= 0.5
