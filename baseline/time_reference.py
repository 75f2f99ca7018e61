#!/usr/bin/env python
"""Times the UNMODIFIED reference (dtak/ocbnn-public) on the host cores: `HMCMixin.single` (bnn/inference.py:39-78) and
`SVGDMixin.single` + its Adagrad step (bnn/inference.py:225-259, 272-283) exactly as its own run scripts set them up
(run_toys.py:46-74 for repro/F1a.yaml, :199-213 for repro/F3b.yaml).

The reference is read from baseline/_ref (a copy of /root/reference/{bnn,data,config.yaml,repro/*.yaml} made by
`__graft_entry__.build()` in the build container; git-ignored, shipped to the GPU box by gpurun).  Nothing of this repo is
imported here: it is the reference's public API and stock code path, in its own process.  Prints one JSON object.

    python baseline/time_reference.py [--hmc-iters 20] [--svgd-epochs 2] [--particles 50]
"""
import argparse
import json
import logging
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hmc-iters", type=int, default=20)
    ap.add_argument("--svgd-epochs", type=int, default=2)
    ap.add_argument("--particles", type=int, default=50)
    ap.add_argument("--skip-hmc", action="store_true")
    ap.add_argument("--skip-svgd", action="store_true")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "bnn")):
        print(json.dumps({"unavailable": "baseline/_ref is missing (run __graft_entry__.build() where /root/reference exists)"}))
        return
    sys.path.insert(0, REF)
    logging.disable(logging.CRITICAL)
    import numpy as np
    import torch
    from torch.distributions.normal import Normal
    import bnn                                   # the reference package
    from data.dataloader import toy1, toy4
    os.chdir(tempfile.mkdtemp(prefix="ocbnn_ref_"))          # the constructor writes history/ into the CWD (base.py:36-40)
    out = {"torch": torch.__version__, "torch_threads": torch.get_num_threads(), "cpu_count": os.cpu_count(),
           "source": "baseline/_ref (unmodified dtak/ocbnn-public)"}

    if not args.skip_hmc:
        torch.manual_seed(0)
        np.random.seed(0)
        m = bnn.BNNHMCRegressor(uid="F1a", configfile=os.path.join(REF, "repro", "F1a.yaml"))
        m.load(**toy1())
        m.add_deterministic_constraint(constrained_domain=(-0.3, 0.3), interval_func=lambda X: [[(2.5, 3.0)] for _ in range(len(X))],
                                       prior_type="negative_exponential_cocp")
        m.update_config(use_ocbnn=True)
        m.accepts = m.rejects = 0
        m.single(True)                           # warm-up iteration (excluded)
        t0 = time.perf_counter()
        for _ in range(args.hmc_iters):
            m.single(True)
        dt = time.perf_counter() - t0
        out["hmc"] = {"workload": "W2 repro/F1a.yaml HMC + negative-exponential COCP (100 constraint points), toy1, posterior; ONE chain (the reference has no batching)",
                      "iterations": args.hmc_iters, "leapfrog_l": int(m.hmc_l), "seconds": dt,
                      "value": args.hmc_iters * int(m.hmc_l) / dt, "unit": "chain-leapfrog-steps/s",
                      "accept_rate": m.accepts / max(1, m.accepts + m.rejects)}

    if not args.skip_svgd:
        torch.manual_seed(0)
        m = bnn.BNNSVGDRegressor(uid="F3b", configfile=os.path.join(REF, "repro", "F3b.yaml"))
        m.load(**toy4())
        inf = float("inf")
        m.add_deterministic_constraint(constrained_domain=(-1.0, 1.0), interval_func=lambda X: [[(-inf, 1.0), (2.5, inf)] for _ in range(len(X))],
                                       prior_type="negative_exponential_cocp")
        m.update_config(use_ocbnn=True, infer_nsamples=args.particles)
        m.particles = Normal(0, m.sigma_w).sample(torch.Size([m.infer_nsamples, m.nweights]))
        opt = torch.optim.Adagrad([m.particles], lr=m.svgd_init_lr)

        def epoch():                             # the body of SVGDMixin.infer's loop (inference.py:273-281)
            opt.zero_grad()
            m.particles.grad = -m.single(True)
            opt.step()
        epoch()                                  # warm-up epoch (excluded)
        t0 = time.perf_counter()
        for _ in range(args.svgd_epochs):
            epoch()
        dt = time.perf_counter() - t0
        out["svgd"] = {"workload": f"W4 repro/F3b.yaml SVGD + negative-exponential COCP, toy4, {args.particles} particles (the reference's own scale)",
                       "epochs": args.svgd_epochs, "particles": args.particles, "seconds": dt,
                       "value": args.particles * args.svgd_epochs / dt, "unit": "particle-iters/s"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
