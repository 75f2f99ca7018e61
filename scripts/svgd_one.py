"""A few SVGD epochs (W4 shape) at one particle count, eager: the command profiled under ncu."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ocbnn-public_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from ocbnn_b200 import workloads
os.chdir(tempfile.mkdtemp())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = sys.argv[2] if len(sys.argv) > 2 else "W4"
m = workloads.build(wl, seed=0, infer_nsamples=n)
m.update_config(svgd_graph=False)
st = m.svgd_setup(torch.randn(n, m.nweights, generator=torch.Generator().manual_seed(3)))
for ep in range(1, 4):
    m.svgd_epoch(st, ep, True)
torch.cuda.synchronize()
print("ok", float(st["h"]))
