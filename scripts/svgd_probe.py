"""Per-kernel timing of one SVGD epoch (W4 shapes) with CUDA events: K1 gradients, K3a bandwidth, K3b phi, K3c Adagrad."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ocbnn-public_b200"))
from ocbnn_b200 import engine as E, workloads

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

os.chdir(tempfile.mkdtemp())
wl = sys.argv[1] if len(sys.argv) > 1 else "W4"
for n in [int(a) for a in sys.argv[2:]] or [4096, 16384]:
    m = workloads.build(wl, seed=0, infer_nsamples=n)
    prob = m.engine_problem()
    dev = torch.device("cuda")
    if m.aocp_mean is not None:
        X = (m.aocp_mean[None] + 3 * m.aocp_std[None] * torch.randn(n, m.nweights)).to(dev).contiguous()
    else:
        X = torch.randn(n, m.nweights, device=dev)
    ws = E.svgd_workspace(n, m.nweights, dev)
    h = torch.empty(1, device=dev); acc = torch.zeros_like(X)
    G = torch.empty_like(X); phi = torch.empty_like(X)
    use_t = int(os.environ.get("SVGD_TENSOR", "0"))
    r = {}
    r["K1 grad"] = t(lambda: prob.logpost_grad(X, (1, m.nbatches) if m.nbatches else (0, 1), True, want_lp=False, out_grad=G))
    r["K3a bandwidth"] = t(lambda: E.svgd_bandwidth(X, ws, h))
    r["K3b phi"] = t(lambda: E.svgd_phi(X, G, h, 0, n, ws, use_t, out=phi))
    r["K3c adagrad"] = t(lambda: E.adagrad_step(X.clone(), acc, phi, 0.5, -1.0))
    st = m.svgd_setup(X.cpu())
    r["epoch (host API)"] = t(lambda: m.svgd_epoch(st, 1, True, use_tensor=use_t))
    print(wl, "n", n, "D", m.nweights, {k: f"{v:.3f} ms" for k, v in r.items()}, "particle-iters/s %.3e" % (n / (r["epoch (host API)"] * 1e-3)))
