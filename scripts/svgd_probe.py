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
    use_t = int(os.environ.get("SVGD_TENSOR", "0"))
    ws = E.svgd_workspace(n, m.nweights, dev, use_tensor=use_t)
    h = torch.empty(1, device=dev); acc = torch.zeros_like(X)
    G = torch.empty_like(X); phi = torch.empty_like(X)
    r = {}
    r["K1 grad"] = t(lambda: prob.logpost_grad(X, (1, m.nbatches) if m.nbatches else (0, 1), True, want_lp=False, out_grad=G))
    r["K3a bandwidth"] = t(lambda: E.svgd_bandwidth(X, ws, h, use_tensor=use_t))
    r["K3b phi"] = t(lambda: E.svgd_phi(X, G, h, 0, n, ws, 2 if use_t else 0, out=phi))
    r["K3c adagrad"] = t(lambda: E.adagrad_step(X.clone(), acc, phi, 0.5, -1.0))
    def raw_epoch():
        prob.logpost_grad(X, (1, m.nbatches) if m.nbatches else (0, 1), True, want_lp=False, out_grad=G)
        E.svgd_bandwidth(X, ws, h, use_tensor=use_t)
        E.svgd_phi(X, G, h, 0, n, ws, 2 if use_t else 0, out=phi)
        E.adagrad_step(X, acc, phi, 1e-6, -1.0)
    r["epoch (raw engine calls)"] = t(raw_epoch)
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter(); raw_epoch(); t1 = time.perf_counter(); torch.cuda.synchronize()
    r["host time of raw epoch"] = (t1 - t0) * 1e3
    st = m.svgd_setup(X.cpu(), use_tensor=use_t)
    m.svgd_epoch(st, 1, True); m.svgd_epoch(st, 1, True)      # eager, then graph capture
    r["epoch (host API)"] = t(lambda: m.svgd_epoch(st, 1, True))
    print(wl, "n", n, "D", m.nweights, {k: f"{v:.3f} ms" for k, v in r.items()}, "particle-iters/s %.3e" % (n / (r["epoch (host API)"] * 1e-3)))
