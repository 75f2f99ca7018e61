"""ctypes wrapper of oracle/_ref/libocbnn_oracle.so (C restatement of the oracle; chains fanned out over host threads here).
TEST INFRASTRUCTURE ONLY — the multi-core CPU baseline of bench.py and a cross-check of the numpy oracle."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

import ocbnn_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libocbnn_oracle.so")
_fp, _ip = C.POINTER(C.c_float), C.POINTER(C.c_int)


class oc_spec(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("sizes", C.c_int * 9), ("act", C.c_int), ("task", C.c_int), ("sigma_w", C.c_float), ("sigma_noise", C.c_float),
                ("prior_mean", _fp), ("prior_std", _fp), ("n_train", C.c_int), ("x", _fp), ("y", _fp),
                ("n_neg", C.c_int), ("maxg", C.c_int), ("neg_x", _fp), ("neg_ngaps", _ip), ("neg_gaps", _fp), ("neg_scale", C.c_float),
                ("tau0", C.c_float), ("tau1", C.c_float), ("n_pos", C.c_int), ("pos_x", _fp), ("pos_y", _fp), ("sigma_c", C.c_float),
                ("n_dir", C.c_int), ("dir_x", _fp), ("dir_cm1", _fp), ("dir_norm", _fp)]


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.oc_logpost_grad.argtypes = [C.POINTER(oc_spec), _fp, C.c_int, C.c_int, _fp, _fp]
        _lib.oc_hmc_iterate.argtypes = [C.POINTER(oc_spec), _fp, _fp, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                        C.POINTER(C.c_ubyte)]
        _lib.oc_svgd_pair_dists.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
        _lib.oc_svgd_pair_dists.restype = C.c_long
        _lib.oc_svgd_phi_rows.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _fp]
    return _lib


def _fan_out(n_rows, threads, fn, balance_triangle=False):
    """fn(lo, hi) over `threads` contiguous row ranges (equal pair counts of the i<j triangle when balance_triangle)."""
    threads = max(1, min(int(threads), n_rows))
    if balance_triangle:            # rows [0, r) of the upper triangle hold r n - r(r+1)/2 pairs: cut at equal areas
        n = n_rows
        tot = n * (n - 1) / 2
        cuts = [0]
        for t in range(1, threads):
            target = tot * t / threads
            r = int(n - 0.5 - math.sqrt(max((n - 0.5) ** 2 - 2 * target, 0.0)))
            cuts.append(min(max(r, cuts[-1]), n))
        cuts.append(n)
        bounds = list(zip(cuts[:-1], cuts[1:]))
    else:
        bounds = [(n_rows * t // threads, n_rows * (t + 1) // threads) for t in range(threads)]
    if threads == 1:
        return [fn(*bounds[0])]
    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda b: fn(*b), bounds))


def svgd_rbf_h(X, threads=1):
    """h = med^2 / ln n, med = LOWER median of the nonzero ||x_i - x_j + 1e-6||, i < j (inference.py:208-217)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, d = X.shape

    def run(lo, hi):
        cnt = sum(n - 1 - i for i in range(lo, hi))
        out = np.empty(cnt, dtype=np.float32)
        k = lib().oc_svgd_pair_dists(X.ctypes.data_as(_fp), n, d, lo, hi, out.ctypes.data_as(_fp))
        assert k == cnt
        return out
    dist = np.concatenate(_fan_out(n, threads, run, balance_triangle=True))
    dist = dist[dist != 0]
    k = (len(dist) - 1) // 2
    med = np.partition(dist, k)[k]
    return np.float32(med * med) / np.float32(math.log(n))


def svgd_phi(X, G, h, threads=1):
    X, G = np.ascontiguousarray(X, dtype=np.float32), np.ascontiguousarray(G, dtype=np.float32)
    n, d = X.shape

    def run(lo, hi):
        out = np.empty((hi - lo, d), dtype=np.float32)
        lib().oc_svgd_phi_rows(X.ctypes.data_as(_fp), G.ctypes.data_as(_fp), n, d, float(h), lo, hi, out.ctypes.data_as(_fp))
        return out
    return np.concatenate(_fan_out(n, threads, run), 0)


def svgd_step(cspec, X, acc, lr, with_data=True, threads=1):
    """One SVGD epoch + Adagrad (full batch): the C counterpart of ocbnn_oracle.svgd_step."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    G = np.concatenate(_fan_out(X.shape[0], threads, lambda lo, hi: cspec.logpost_grad(X[lo:hi], with_data)[1]), 0)
    h = svgd_rbf_h(X, threads)
    phi = svgd_phi(X, G, h, threads)
    Xn, accn = orc.adagrad_step(X, np.asarray(acc, dtype=np.float32), -phi, lr, np.float32)
    return Xn, accn, phi, h


class CSpec:
    """oc_spec built from an ocbnn_oracle.Spec (the joint gap decision of the negative COCP is made here, base.py:374-384)."""

    def __init__(self, spec):
        self.keep = []
        s = oc_spec()
        s.n_layers = len(spec.layer_sizes) - 1
        for i, v in enumerate(spec.layer_sizes):
            s.sizes[i] = v
        s.act = {"rbf": 1, "relu": 2}.get(spec.activation, 0)
        s.task = {"regression": 0, "classifier": 1, "binary": 2}[spec.task]
        s.sigma_w, s.sigma_noise = spec.sigma_w, spec.sigma_noise
        f = self._f
        s.n_train = spec.N_train
        s.x, s.y = f(spec.X_train), f(np.asarray(spec.Y_train, dtype=np.float32))
        aocp = spec.use_ocbnn and spec.aocp_mean is not None
        if aocp:
            s.prior_mean, s.prior_std = f(spec.aocp_mean), f(spec.aocp_std)
        if spec.use_ocbnn and not aocp:
            if spec.neg is not None:
                S, nc = int(spec.neg["S"]), len(spec.neg["intervals"])
                gaps_all = []
                for c in range(nc):
                    iv = np.asarray(spec.neg["intervals"][c], dtype=np.float32)
                    marker = np.full((S,), -np.inf, dtype=np.float32)
                    gaps = []
                    for k in range(iv.shape[1]):
                        if np.all(marker < iv[:, k, 0]):
                            gaps.append(np.stack([marker, iv[:, k, 0]], 1))
                        marker = iv[:, k, 1]
                    if np.all(marker < np.inf):
                        gaps.append(np.stack([marker, np.full((S,), np.inf, dtype=np.float32)], 1))
                    gaps_all.append(gaps)
                maxg = max(1, max(len(g) for g in gaps_all))
                G = np.zeros((nc * S, maxg, 2), dtype=np.float32)
                ng = np.zeros((nc * S,), dtype=np.int32)
                for c, gaps in enumerate(gaps_all):
                    ng[c * S:(c + 1) * S] = len(gaps)
                    for k, g in enumerate(gaps):
                        G[c * S:(c + 1) * S, k] = g
                s.n_neg, s.maxg = nc * S, maxg
                s.neg_x, s.neg_gaps = f(spec.neg["xs"]), f(G)
                ng = np.ascontiguousarray(ng)
                self.keep.append(ng)
                s.neg_ngaps = ng.ctypes.data_as(_ip)
                s.neg_scale = -float(spec.neg["gamma"]) / (S * nc)
                s.tau0, s.tau1 = float(spec.neg["tau"][0]), float(spec.neg["tau"][1])
            if spec.pos is not None:
                assert all(int(k) == 1 for k in spec.pos["ylens"]), "the C port restates the reachable K = 1 case"
                s.n_pos = len(spec.pos["xs"])
                s.pos_x, s.pos_y, s.sigma_c = f(spec.pos["xs"]), f(np.asarray(spec.pos["ys"])[:, 0]), float(spec.pos["sigma_c"])
            if spec.dir is not None:
                S, nc, sk = int(spec.dir["S"]), len(spec.dir["forbidden"]), spec.layer_sizes[-1]
                cm1 = np.zeros((nc * S, sk), dtype=np.float32)
                nrm = np.zeros((nc * S,), dtype=np.float32)
                for c in range(nc):
                    conc = orc.dirichlet_conc(spec, c)
                    cm1[c * S:(c + 1) * S] = conc - 1
                    nrm[c * S:(c + 1) * S] = math.lgamma(conc.sum()) - sum(math.lgamma(a) for a in conc)
                s.n_dir = nc * S
                s.dir_x, s.dir_cm1, s.dir_norm = f(spec.dir["xs"]), f(cm1), f(nrm)
        self.s = s
        self.D = spec.nweights

    def _f(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self.keep.append(a)
        return a.ctypes.data_as(_fp)

    def logpost_grad(self, W, with_data=True):
        W = np.ascontiguousarray(W, dtype=np.float32)
        lp = np.empty(W.shape[0], dtype=np.float32)
        g = np.empty_like(W)
        lib().oc_logpost_grad(C.byref(self.s), W.ctypes.data_as(_fp), W.shape[0], int(with_data), lp.ctypes.data_as(_fp), g.ctypes.data_as(_fp))
        return lp, g

    def hmc_iterate(self, q, p0, u, eps, L, with_data=True, threads=1):
        """n_it HMC iterations of M chains; `threads` host threads each take a contiguous share of the chains."""
        q = np.array(q, dtype=np.float32, order="C", copy=True)
        M, n_it = q.shape[0], p0.shape[0]
        acc = np.empty((n_it, M), dtype=np.uint8)

        def run(lo, hi):
            qs = np.ascontiguousarray(q[lo:hi])
            ps = np.ascontiguousarray(p0[:, lo:hi], dtype=np.float32)
            us = np.ascontiguousarray(u[:, lo:hi], dtype=np.float64)
            ac = np.empty((n_it, hi - lo), dtype=np.uint8)
            lib().oc_hmc_iterate(C.byref(self.s), qs.ctypes.data_as(_fp), ps.ctypes.data_as(_fp), us.ctypes.data_as(C.POINTER(C.c_double)),
                                 hi - lo, n_it, float(eps), int(L), int(with_data), ac.ctypes.data_as(C.POINTER(C.c_ubyte)))
            q[lo:hi] = qs
            acc[:, lo:hi] = ac
        threads = max(1, min(int(threads), M))
        bounds = [(M * t // threads, M * (t + 1) // threads) for t in range(threads)]
        if threads == 1:
            run(0, M)
        else:
            import concurrent.futures as cf
            with cf.ThreadPoolExecutor(threads) as ex:
                list(ex.map(lambda b: run(*b), bounds))
        return q, acc.astype(bool)
