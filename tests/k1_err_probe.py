"""Relative error of the generic-net K1 (value, gradient) against the float64 oracle on the credit / recid parity cases."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # test infrastructure: the only place besides tests that may import oracle/
for p in (ROOT, os.path.join(ROOT, "ocbnn-public_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import tempfile
os.chdir(tempfile.mkdtemp())
import ocbnn_oracle as orc
from helpers import load_golden, make_model, rel_err, spec_from_golden
for name in ("credit_small", "recid_small"):
    g = load_golden(name); m = make_model(name, g); spec = spec_from_golden(name, g); W = g["W"]
    prob = m.engine_problem()
    lp64, g64 = orc.log_posterior(spec, W, None, True, dtype=np.float64)
    lp32, g32 = orc.log_posterior(spec, W, None, True, dtype=np.float32)
    lp, gr = prob.logpost_grad(torch.as_tensor(W).cuda().float().contiguous(), (0, 1), True)
    print(name, "engine vs f64: lp %.2e grad %.2e | fp32 oracle vs f64: lp %.2e grad %.2e" % (
        rel_err(lp.cpu().numpy(), lp64).max(), rel_err(gr.cpu().numpy(), g64).max(), rel_err(lp32, lp64).max(), rel_err(g32, g64).max()))
