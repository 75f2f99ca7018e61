"""Timing of one BBB epoch at the credit.yaml shape (W6): K5 sample, K1 (generic net), K5 reduce, Adagrad."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ocbnn-public_b200"))
from ocbnn_b200 import engine as E, workloads

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

os.chdir(tempfile.mkdtemp())
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = workloads.build("W6", seed=0, bbb_esamples=k)
prob = m.engine_problem()
dev = torch.device("cuda")
D = m.nweights
qp = torch.cat([0.3 * torch.randn(D, 1), torch.ones(D, 1) * -3.0], 1).to(dev).contiguous()
acc = torch.zeros_like(qp)
eps = torch.randn(k, D, device=dev)
W = E.bbb_sample(qp, eps)
lp = torch.empty(k, device=dev); G = torch.empty_like(W)
r = {}
r["K5 sample"] = t(lambda: E.bbb_sample(qp, eps))
r["K1 logpost+grad"] = t(lambda: prob.logpost_grad(W, (1, m.nbatches), True, out_lp=lp, out_grad=G))
r["K5 reduce"] = t(lambda: E.bbb_reduce(qp, eps, lp, G, k))
r["epoch"] = t(lambda: m.bbb_epoch(qp, acc, eps, 1, True, k_total=k))
npts = len(range(1, m.N_train, m.nbatches)) + 500
flop = k * npts * workloads.flops_per_point(m.layer_sizes)
print("W6 BBB k", k, "D", D, "points/eval", npts, {a: f"{b:.3f} ms" for a, b in r.items()},
      "K1 %.2f TFLOP/s" % (flop / (r["K1 logpost+grad"] * 1e-3) / 1e12), "MC-sample-evals/s %.3e" % (k / (r["epoch"] * 1e-3)))
