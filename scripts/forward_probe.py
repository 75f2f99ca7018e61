"""Timing of the batched forward / predict path on the application nets (posterior-sample evaluation: M weight vectors x P points)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ocbnn-public_b200"))
import torch
from ocbnn_b200 import engine as E, workloads

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

os.chdir(tempfile.mkdtemp())
for wl, M, P in (("W6", 1000, 30000), ("W7", 50, 6172), ("W7", 1000, 6172)):
    m = workloads.build(wl, seed=0)
    prob = m.engine_problem()
    W = 0.3 * torch.randn(M, m.nweights, device="cuda")
    X = torch.randn(P, m.layer_sizes[0], device="cuda")
    ms = t(lambda: prob.forward(W, X))
    fl = 2.0 * sum(a * b for a, b in zip(m.layer_sizes[:-1], m.layer_sizes[1:])) * M * P
    print(f"{wl} forward {M} weight vectors x {P} points: {ms:.3f} ms = {fl / ms / 1e9:.2f} TFLOP/s ({M * P / ms / 1e3:.1f} M forwards/s), OCBNN_GEN_TC={os.environ.get('OCBNN_GEN_TC', '1')}", flush=True)
