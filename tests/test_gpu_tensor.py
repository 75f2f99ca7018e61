"""tcgen05 / TMA path: the 3xTF32 GEMM building block and the SVGD tensor path against float64 references."""
import numpy as np
import pytest
import torch

import ocbnn_oracle as orc
from ocbnn_b200 import engine as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (128, 64, 64), (300, 200, 100), (512, 64, 256), (257, 31, 40), (1000, 1000, 31), (130, 22464 // 16, 513)])
def test_gemm_tf32x3(m, n, k):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    A = torch.randn(m, k, generator=g)
    B = torch.randn(n, k, generator=g) * 3 + 0.5
    C = E.gemm_tf32x3(A.cuda(), B.cuda()).cpu().double()
    ref = A.double() @ B.double().T
    scale = (A.double().abs() @ B.double().abs().T)                 # error relative to sum |a||b| (no cancellation credit)
    err = ((C - ref).abs() / scale).max().item()
    assert err < 2e-6, err
    plain_tf32 = 5e-4
    assert err < plain_tf32 / 50


@pytest.mark.parametrize("n,D", [(1500, 31), (1111, 63), (1100, 200), (2050, 41)])
def test_svgd_tensor_path_matches_float64_oracle(n, D):
    """K3a/K3b on tcgen05 (3xTF32, TMA): bandwidth and phi against the float64 closed form and the exact SIMT arm."""
    gen = torch.Generator().manual_seed(n + D)
    X = torch.randn(n, D, generator=gen) * 0.8 + 0.3          # non-zero mean: the path centres X
    G = torch.randn(n, D, generator=gen) * 5
    Xd, Gd = X.cuda(), G.cuda()
    ws = E.svgd_workspace(n, D, Xd.device, use_tensor=1)
    h_t = E.svgd_bandwidth(Xd, ws, use_tensor=1)
    h_s = E.svgd_bandwidth(Xd, E.svgd_workspace(n, D, Xd.device, use_tensor=0), use_tensor=0)
    h_ref = float(orc.svgd_rbf_h(X.numpy()))
    assert abs(float(h_s) - h_ref) <= 2e-6 * h_ref
    assert abs(float(h_t) - h_ref) <= (3e-6 if D <= 64 else 1e-5) * h_ref, (float(h_t), h_ref)
    phi_t = E.svgd_phi(Xd, Gd, h_t, 0, n, ws, use_tensor=2)                     # workspace prepared by svgd_bandwidth
    ref = orc.svgd_phi(X.numpy().astype(np.float64), G.numpy().astype(np.float64), float(h_t), dtype=np.float64)
    err = np.linalg.norm(phi_t.cpu().numpy() - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert err.max() < 1e-5, err.max()
    phi_1 = E.svgd_phi(Xd, Gd, h_t, 0, n, ws, use_tensor=1)                     # prepares the workspace itself
    assert torch.equal(phi_1, phi_t)
    ws_part = E.svgd_workspace(n, D, Xd.device, n_rows=157, use_tensor=1)
    part = E.svgd_phi(Xd, Gd, h_t, 300, 157, ws_part, use_tensor=1)             # a rank's row share
    assert (part - phi_t[300:457]).norm() / phi_t[300:457].norm() < 5e-6      # different column splits: fp32 reassociation only


def test_svgd_epoch_tensor_vs_simt():
    import os, tempfile
    from ocbnn_b200 import workloads
    os.chdir(tempfile.mkdtemp())
    n = 2048
    m = workloads.build("W4", seed=0, infer_nsamples=n)
    X0 = torch.randn(n, m.nweights, generator=torch.Generator().manual_seed(3))
    st_t, st_s = m.svgd_setup(X0, use_tensor=1), m.svgd_setup(X0, use_tensor=0)
    assert st_t["tensor"] and not st_s["tensor"]
    for ep in range(1, 3):
        m.svgd_epoch(st_t, ep, True)
        m.svgd_epoch(st_s, ep, True)
        assert abs(float(st_t["h"]) - float(st_s["h"])) <= 1e-5 * float(st_s["h"])
        rel = (st_t["phi"] - st_s["phi"]).norm(dim=1) / st_s["phi"].norm(dim=1)
        assert float(rel.max()) < 2e-5, float(rel.max())


@pytest.mark.parametrize("n,D,dup", [(1500, 31, 0), (2050, 41, 0), (1100, 200, 0), (1200, 31, 300), (64, 10, 0)])
def test_svgd_bandwidth_select_modes_agree(n, D, dup):
    """The bracket select (default), its in-kernel fallback (sabotaged bracket) and the 3-pass radix select return the same
    exact lower median, bit for bit; `dup` duplicated particles give exact-zero distances (dropped) and heavy ties."""
    gen = torch.Generator().manual_seed(n * 3 + D)
    X = torch.randn(n, D, generator=gen) * 0.8 + 0.3
    if dup:
        X[-dup:] = X[:dup]
    Xd = X.cuda()
    ws = E.svgd_workspace(n, D, Xd.device, use_tensor=1)
    _, state = E.svgd_ws_views(ws)
    prev = E.svgd_select_mode(-1)
    try:
        out = {}
        for mode in (1, 0, 2):
            E.svgd_select_mode(mode)
            out[mode] = float(E.svgd_bandwidth(Xd, ws, use_tensor=1))
            if mode == 1:
                hit, ncand, below, zeros = [int(v) for v in state[:4].cpu()]
                pairs = n * (n - 1) // 2
                assert hit == 1, "the sampled bracket must contain the median"
                assert zeros == 0          # the +1e-6 inside the norm keeps duplicated particles at distance sqrt(D) 1e-6 (reference too)
                assert 0 < ncand <= max(4096, 0.08 * pairs), (ncand, pairs)
            if mode == 2:
                assert int(state[0]) == 0, "sabotaged bracket must take the fallback"
        assert out[1] == out[0] == out[2], out
        h_ref = float(orc.svgd_rbf_h(X.numpy()))
        assert abs(out[1] - h_ref) <= (3e-6 if D <= 64 else 1e-5) * h_ref
    finally:
        E.svgd_select_mode(prev)
