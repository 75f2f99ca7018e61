"""GPU probe: SVGD epoch time (W4 = F3b shape, D = 31) and its two big phases, per particle count."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ocbnn-public_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from ocbnn_b200 import engine as E, workloads

def t_cuda(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

os.chdir(tempfile.mkdtemp())
wl = sys.argv[1] if len(sys.argv) > 1 else "W4"
for n in [int(v) for v in (sys.argv[2:] or ["4096", "16384", "32768"])]:
    m = workloads.build(wl, seed=0, infer_nsamples=n)
    X0 = torch.randn(n, m.nweights, generator=torch.Generator().manual_seed(3))
    st = m.svgd_setup(X0)
    ep = [0]
    def epoch():
        ep[0] += 1
        m.svgd_epoch(st, ep[0], True)
    t_ep = t_cuda(epoch, reps=20, warm=4)
    X, G = st["X"], st["G"]
    t_bw = t_cuda(lambda: E.svgd_bandwidth(X, st["ws"], st["h"], use_tensor=1))
    t_phi = t_cuda(lambda: E.svgd_phi(X, G, st["h"], 0, n, st["ws"], 2, out=st["phi"]))
    t_g = t_cuda(lambda: st["prob"].logpost_grad(X, (0, 1), True, want_lp=False, out_grad=G))
    print(f"{wl} n={n}: epoch {t_ep*1e3:.1f} us ({n/t_ep/1e3:.2f} M particle-iters/s); bandwidth {t_bw*1e3:.1f} us, phi {t_phi*1e3:.1f} us, grad {t_g*1e3:.1f} us, "
          f"ws {st['ws'].numel()*8/2**20:.0f} MiB, finite {bool(torch.isfinite(st['X']).all())}", flush=True)
