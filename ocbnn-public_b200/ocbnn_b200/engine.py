"""
ctypes binding of libocbnn_b200.so (the C-ABI declared in include/ocbnn_b200.h).

PyTorch is used for device memory, streams and torch.distributed only; every
number on the hot path is produced by the hand-written sm_100a kernels behind
this boundary.  There is NO CPU fallback: if the library is missing, or there is
no B200, every compute call raises EngineError.
"""

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libocbnn_b200.so")

OCBNN_ABI_VERSION = 3
OCBNN_MAX_LAYERS = 8
ACT_IDENTITY, ACT_RBF, ACT_RELU = 0, 1, 2
HEAD_LIK_GAUSS, HEAD_NEG_EXP, HEAD_POS_GAUSS, HEAD_DIRICHLET, HEAD_LIK_SOFTMAX, HEAD_LIK_BERNOULLI = range(6)
PREDICT_RAW, PREDICT_PROBS, PREDICT_CLASS = range(3)

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class EngineError(RuntimeError):
    pass


class ocbnn_desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_layers", C.c_int32),
        ("layer_sizes", C.c_int32 * (OCBNN_MAX_LAYERS + 1)),
        ("activation", C.c_int32),
        ("lik_head", C.c_int32),
        ("lik_scale", C.c_float),
        ("lik_const", C.c_double),
        ("prior_mean", C.c_float),
        ("prior_std", C.c_float),
        ("prior_mean_vec", _fp),
        ("prior_std_vec", _fp),
        ("prior_const", C.c_double),
        ("n_train", C.c_int64),
        ("x_train", _fp),
        ("y_train", _fp),
        ("y_width", C.c_int32),
        ("n_cpoints", C.c_int32),
        ("cx", _fp),
        ("ckind", _ip),
        ("cweight", _fp),
        ("cparam_stride", C.c_int32),
        ("cparam", _fp),
        ("neg_tau0", C.c_float),
        ("neg_tau1", C.c_float),
        ("pos_inv_var", C.c_float),
    ]


_lib = None

_SIGS = {
    "ocbnn_last_error": (C.c_char_p, []),
    "ocbnn_abi_version": (C.c_int, []),
    "ocbnn_launch_count": (C.c_int64, []),
    "ocbnn_problem_create": (C.c_int, [C.POINTER(ocbnn_desc), C.POINTER(C.c_void_p)]),
    "ocbnn_problem_destroy": (C.c_int, [C.c_void_p]),
    "ocbnn_problem_nweights": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ocbnn_logpost_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_void_p]),
    "ocbnn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ocbnn_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "ocbnn_violation_flags": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "ocbnn_hmc_iterate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_void_p]),
    "ocbnn_sgld_iterate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _fp, _ip, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "ocbnn_svgd_bandwidth": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "ocbnn_svgd_select_mode": (C.c_int, [C.c_int32]),
    "ocbnn_svgd_phi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "ocbnn_svgd_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64, C.c_int32]),
    "ocbnn_svgd_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ocbnn_comm_unique_id": (C.c_int, [C.c_void_p]),
    "ocbnn_comm_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "ocbnn_comm_wrap": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "ocbnn_comm_destroy": (C.c_int, [C.c_void_p]),
    "ocbnn_comm_shard": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ocbnn_bbb_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "ocbnn_bbb_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ocbnn_gemm_tf32x3_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64]),
    "ocbnn_gemm_tf32x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "ocbnn_adagrad_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p]),
    "ocbnn_bbb_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ocbnn_bbb_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ocbnn_aocp_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "ocbnn_aocp_objective": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                       C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ocbnn_probe_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "ocbnn_hmc_iterate_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
}


def lib():
    """Load the CUDA extension; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: build it with `python ocbnn-public_b200/build.py` "
                              "(the engine has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.ocbnn_abi_version() != OCBNN_ABI_VERSION:
            raise EngineError("libocbnn_b200.so ABI version mismatch")
        _lib = L
    return _lib


def exported_symbols():
    return list(_SIGS)


def check(rc):
    if rc != 0:
        raise EngineError(f"ocbnn engine error {rc}: {lib().ocbnn_last_error().decode()}")


def launch_count():
    return int(lib().ocbnn_launch_count())


def probe_peak(mode):
    """Measured instruction peak of this GPU: 0 scalar FFMA, 1 packed FFMA2 (FLOP/s), 2 MUFU.EX2 (op/s)."""
    require_cuda()
    r = C.c_double()
    check(lib().ocbnn_probe_peak(int(mode), C.byref(r)))
    return r.value


def require_cuda():
    if not torch.cuda.is_available():
        raise EngineError("no CUDA device visible: the OC-BNN B200 engine has no CPU fallback")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise EngineError(f"{name} must be a contiguous float32 CUDA tensor")
    return t


class DeviceProblem:
    """Device-resident compiled problem (opaque ocbnn_problem handle)."""

    def __init__(self, cp):
        require_cuda()
        L = lib()
        self.cp = cp
        d = ocbnn_desc()
        d.abi_version = OCBNN_ABI_VERSION
        d.n_layers = len(cp.layer_sizes) - 1
        for i, s in enumerate(cp.layer_sizes):
            d.layer_sizes[i] = int(s)
        d.activation, d.lik_head = cp.activation, cp.lik_head
        d.lik_scale, d.lik_const = cp.lik_scale, cp.lik_const
        d.prior_mean, d.prior_std, d.prior_const = cp.prior_mean, cp.prior_std, cp.prior_const
        keep = []

        def fptr(a):
            if a is None or a.size == 0:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            return a.ctypes.data_as(_fp)
        d.prior_mean_vec, d.prior_std_vec = fptr(cp.prior_mean_vec), fptr(cp.prior_std_vec)
        d.n_train = cp.X.shape[0]
        d.x_train, d.y_train, d.y_width = fptr(cp.X), fptr(cp.Y), cp.Y.shape[1] if cp.Y.ndim == 2 else 1
        d.n_cpoints = cp.cx.shape[0]
        d.cx, d.cweight, d.cparam = fptr(cp.cx), fptr(cp.cweight), fptr(cp.cparam)
        ck = np.ascontiguousarray(cp.ckind, dtype=np.int32)
        keep.append(ck)
        d.ckind = ck.ctypes.data_as(_ip) if ck.size else None
        d.cparam_stride = cp.cparam.shape[1] if cp.cparam.ndim == 2 and cp.cparam.size else 0
        d.neg_tau0, d.neg_tau1, d.pos_inv_var = cp.tau[0], cp.tau[1], cp.pos_inv_var
        h = C.c_void_p()
        check(L.ocbnn_problem_create(C.byref(d), C.byref(h)))
        self._h = h
        self.D = cp.nweights
        self.sK = cp.layer_sizes[-1]
        self.s0 = cp.layer_sizes[0]
        self.device = torch.device("cuda", torch.cuda.current_device())

    def close(self):
        if getattr(self, "_h", None):
            lib().ocbnn_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- K1 -------------------------------------------------------------------------------------------
    def logpost_grad(self, W, batch=(0, 1), with_data=True, want_lp=True, want_grad=True, lanes=0, out_lp=None, out_grad=None):
        W = _f32(W, "W")
        M = W.shape[0]
        lp = (out_lp if out_lp is not None else torch.empty(M, device=W.device)) if want_lp else None
        g = (out_grad if out_grad is not None else torch.empty_like(W)) if want_grad else None
        mode = 2 if with_data == "likelihood" else int(bool(with_data))
        check(lib().ocbnn_logpost_grad(self._h, _ptr(W), M, int(batch[0]), int(batch[1]), mode, _ptr(lp), _ptr(g),
                                       int(lanes), _stream()))
        return lp, g

    def _domain(self, W, X):
        W, X = _f32(W, "W"), _f32(X, "X")
        if W.ndim != 2 or W.shape[1] != self.D:
            raise EngineError(f"weights must be (M, {self.D}), got {tuple(W.shape)}")
        if X.ndim != 2 or X.shape[1] != self.s0:
            raise EngineError(f"domain must be (K, {self.s0}), got {tuple(X.shape)}")
        return W, X

    def forward(self, W, X):
        W, X = self._domain(W, X)
        out = torch.empty(W.shape[0], X.shape[0], self.sK, device=W.device)
        check(lib().ocbnn_forward(self._h, _ptr(W), W.shape[0], _ptr(X), X.shape[0], _ptr(out), _stream()))
        return out

    def predict(self, W, X, head=PREDICT_RAW, want_class=False, want_mean=False):
        """ocbnn_predict: (out (M, K, sK) raw outputs or probabilities, classes (M, K) int32 or None, mean over M (K, sK) or None)."""
        W, X = self._domain(W, X)
        M, K = W.shape[0], X.shape[0]
        out = torch.empty(M, K, self.sK, device=W.device)
        cls = torch.empty(M, K, dtype=torch.int32, device=W.device) if (want_class or head == PREDICT_CLASS) else None
        mean = torch.empty(K, self.sK, device=W.device) if want_mean else None
        check(lib().ocbnn_predict(self._h, _ptr(W), M, _ptr(X), K, int(head), _ptr(out), _ptr(cls), _ptr(mean), _stream()))
        return out, cls, mean

    def violation_flags(self, W, X, lo, hi):
        """ocbnn_violation_flags: (M,) uint8, 1 where sample m predicts inside (lo, hi) somewhere on the domain X."""
        W, X = self._domain(W, X)
        M, K = W.shape[0], X.shape[0]
        scratch = torch.empty(max(1, M * K * self.sK), device=W.device)
        flags = torch.empty(M, dtype=torch.uint8, device=W.device)
        check(lib().ocbnn_violation_flags(self._h, _ptr(W), M, _ptr(X), K, float(lo), float(hi), _ptr(scratch), _ptr(flags), _stream()))
        return flags

    # ---- K2 -------------------------------------------------------------------------------------------
    def hmc_iterate(self, q, p0, u, eps, L, with_data=True, restore_on_reject=False, want_h=False, want_flags=True, samples=None,
                    thin=0, lanes=0):
        q, p0 = _f32(q, "q"), _f32(p0, "p0")
        if u.dtype != torch.float64 or not u.is_cuda or not u.is_contiguous():
            raise EngineError("u must be a contiguous float64 CUDA tensor")
        n_it, M = p0.shape[0], q.shape[0]
        acc = torch.empty(n_it, M, dtype=torch.uint8, device=q.device)
        H = torch.empty(n_it, M, 4, device=q.device) if want_h else None
        flags = torch.zeros(M, dtype=torch.int32, device=q.device) if want_flags else None
        check(lib().ocbnn_hmc_iterate(self._h, _ptr(q), _ptr(p0), _ptr(u), M, n_it, float(eps), int(L), int(bool(with_data)),
                                      int(bool(restore_on_reject)), _ptr(acc), _ptr(H), _ptr(flags), _ptr(samples), int(thin),
                                      int(lanes), _stream()))
        return acc, H, flags

    # ---- K4 -------------------------------------------------------------------------------------------
    def sgld_iterate(self, w, noise, half_eps, batch_offsets, batch_stride, with_data=True, samples=None, thin=0, lanes=0):
        w, noise = _f32(w, "w"), _f32(noise, "noise")
        T = noise.shape[0]
        he = np.ascontiguousarray(half_eps, dtype=np.float32)
        bo = np.ascontiguousarray(batch_offsets if batch_offsets is not None else np.zeros(T), dtype=np.int32)
        check(lib().ocbnn_sgld_iterate(self._h, _ptr(w), _ptr(noise), he.ctypes.data_as(_fp), bo.ctypes.data_as(_ip),
                                       int(batch_stride), w.shape[0], T, int(bool(with_data)), _ptr(samples), int(thin), int(lanes),
                                       _stream()))


    # ---- K6 -------------------------------------------------------------------------------------------
    def aocp_objective(self, params, cx, targets, mode, max_iv, sigma_noise, inv_denom, out_loss, out_grad, accumulate=False):
        params, cx, targets = _f32(params, "params"), _f32(cx, "cx"), _f32(targets, "targets")
        Cn = cx.shape[0]
        nbytes = int(lib().ocbnn_aocp_workspace_bytes(Cn, self.D))
        ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=params.device)
        check(lib().ocbnn_aocp_objective(self._h, _ptr(params), _ptr(cx), _ptr(targets), Cn, int(mode), int(max_iv), float(sigma_noise),
                                         float(inv_denom), int(bool(accumulate)), _ptr(out_loss), _ptr(out_grad), _ptr(ws), ws.numel() * 4,
                                         _stream()))


# ---- K3 (problem independent) ------------------------------------------------------------------------------

SVGD_HDR_BYTES = 3 * 2048 * 4 + 256      # select histogram (3 x 2048 u32) + select state at the start of every SVGD workspace


def svgd_use_tensor(use_tensor, n):
    """Resolve the use_tensor knob (0 SIMT, 1 tcgen05, -1 choose) to a bool, the same way the library does."""
    return use_tensor > 0 or (use_tensor < 0 and n >= 1024)


def svgd_workspace(n, d, device, n_rows=None, use_tensor=-1):
    nbytes = int(lib().ocbnn_svgd_workspace_bytes(n, d, n if n_rows is None else n_rows, int(use_tensor)))
    return torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)


def svgd_ws_views(workspace):
    """(hist int32[3*2048], state int64[4]) views of the workspace header (what a multi-GPU caller all-reduces)."""
    hist = workspace[:3 * 2048 * 4 // 8].view(torch.int32)
    state = workspace[3 * 2048 * 4 // 8:3 * 2048 * 4 // 8 + 4]
    return hist, state


def svgd_bandwidth(X, workspace, out_h=None, use_tensor=0):
    X = _f32(X, "X")
    h = out_h if out_h is not None else torch.empty(1, device=X.device)
    check(lib().ocbnn_svgd_bandwidth(_ptr(X), X.shape[0], X.shape[1], _ptr(h), int(use_tensor), _ptr(workspace), workspace.numel() * 8, _stream()))
    return h


def svgd_select_mode(mode=-1):
    """Median-select algorithm of svgd_bandwidth's tensor path: 1 bracket select (default), 0 radix select, 2 sabotaged bracket
    (tests the in-kernel fallback); returns the previous mode."""
    return int(lib().ocbnn_svgd_select_mode(int(mode)))


def svgd_phi(X, G, h, row_begin, n_rows, workspace, use_tensor=0, out=None):
    X, G = _f32(X, "X"), _f32(G, "G")
    phi = out if out is not None else torch.empty(n_rows, X.shape[1], device=X.device)
    check(lib().ocbnn_svgd_phi(_ptr(X), _ptr(G), X.shape[0], X.shape[1], int(row_begin), int(n_rows), _ptr(h), _ptr(phi), int(use_tensor),
                               _ptr(workspace), workspace.numel() * 8 if workspace is not None else 0, _stream()))
    return phi


def svgd_step(problem, X_all, G_all, acc, phi, h, row_begin, n_rows, lr, batch, with_data, use_tensor, workspace, comm=None):
    """One SVGD iteration in one library call (K1 -> all-gather G -> K3a -> K3b -> K3c -> all-gather X), on the current stream.
    X_all / G_all: (n, D) device buffers replicated on every rank; acc / phi: (n_rows, D) of the owned rows; comm: Comm or None."""
    off, stride = batch
    check(lib().ocbnn_svgd_step(problem._h, _ptr(X_all), _ptr(G_all), _ptr(acc), _ptr(phi), _ptr(h), X_all.shape[0], int(row_begin), int(n_rows),
                                float(lr), int(off), int(stride), int(bool(with_data)), int(use_tensor), _ptr(workspace), workspace.numel() * 8,
                                comm.handle if comm is not None else None, _stream()))


def bbb_workspace(k_local, d, device):
    nbytes = int(lib().ocbnn_bbb_workspace_bytes(int(k_local), int(d)))
    return torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)


def bbb_step(problem, q_params, acc, eps, k_total, lr, batch, with_data, out_elbo, workspace, comm=None):
    """One BBB epoch in one library call (sample -> K1 -> closed-form ELBO gradient -> one all-reduce -> Adagrad)."""
    off, stride = batch
    check(lib().ocbnn_bbb_step(problem._h, _ptr(q_params), _ptr(acc), _ptr(_f32(eps, "eps")), eps.shape[0], int(k_total), float(lr), int(off),
                               int(stride), int(bool(with_data)), _ptr(out_elbo), _ptr(workspace), workspace.numel() * 8,
                               comm.handle if comm is not None else None, _stream()))


class Comm:
    """ocbnn_comm: the library's own NCCL communicator (one per process).  The id is created on rank 0 and handed to the other
    ranks by `exchange` (a callable: bytes-or-None -> bytes, e.g. a torch.distributed broadcast) - host plumbing only."""

    def __init__(self, rank, world, exchange):
        require_cuda()
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            check(lib().ocbnn_comm_unique_id(buf))
        ident = exchange(bytes(buf) if rank == 0 else None)
        buf = (C.c_uint8 * 128).from_buffer_copy(ident)
        h = C.c_void_p()
        check(lib().ocbnn_comm_create(buf, int(rank), int(world), C.byref(h)))
        self.handle, self.rank, self.world = h, int(rank), int(world)

    def close(self):
        if self.handle:
            lib().ocbnn_comm_destroy(self.handle)
            self.handle = None


def comm_shard(n, rank, world):
    lo, cnt = C.c_int64(), C.c_int64()
    check(lib().ocbnn_comm_shard(int(n), int(rank), int(world), C.byref(lo), C.byref(cnt)))
    return lo.value, lo.value + cnt.value


def gemm_tf32x3(A, B):
    """C = A @ B.T on tcgen05 in 3xTF32 precision (building block of the SVGD tensor path)."""
    A, B = _f32(A, "A"), _f32(B, "B")
    m, k = A.shape
    n = B.shape[0]
    Cm = torch.empty(m, n, device=A.device)
    nbytes = int(lib().ocbnn_gemm_tf32x3_workspace_bytes(m, n, k))
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=A.device)
    check(lib().ocbnn_gemm_tf32x3(_ptr(A), _ptr(B), _ptr(Cm), m, n, k, _ptr(ws), ws.numel() * 4, _stream()))
    return Cm


def adagrad_step(x, acc, direction, lr, grad_sign):
    check(lib().ocbnn_adagrad_step(_ptr(_f32(x, "x")), _ptr(_f32(acc, "acc")), _ptr(_f32(direction, "dir")), x.numel(), float(lr),
                                   float(grad_sign), _stream()))


def bbb_sample(q_params, eps):
    w = torch.empty_like(eps)
    check(lib().ocbnn_bbb_sample(_ptr(_f32(q_params, "q_params")), _ptr(_f32(eps, "eps")), eps.shape[0], eps.shape[1], _ptr(w), _stream()))
    return w


def bbb_reduce(q_params, eps, lp, grad, k_total, add_entropy=True):
    D = eps.shape[1]
    gq = torch.empty(D, 2, device=eps.device)
    elbo = torch.empty(1, device=eps.device)
    colstat = torch.empty(D, device=eps.device)
    check(lib().ocbnn_bbb_reduce(_ptr(q_params), _ptr(eps), _ptr(_f32(lp, "lp")), _ptr(_f32(grad, "grad")), eps.shape[0], D, int(k_total),
                                 int(bool(add_entropy)), _ptr(gq), _ptr(elbo), _ptr(colstat), _stream()))
    return gq, elbo
