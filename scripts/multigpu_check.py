"""Multi-GPU agreement check (run under torchrun, one rank per GPU): the sharded HMC / SVGD / BBB paths against the same
work done by one rank.  HMC chains must agree bit-for-bit; SVGD and BBB within 1e-5.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py
"""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ocbnn-public_b200"))

from ocbnn_b200 import dist as D, engine as E, workloads  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    os.chdir(tempfile.mkdtemp(prefix=f"mg_r{rank}_"))
    ok = True

    # ---- HMC: independent chains, no data-path collective ---------------------------------------------------------
    m = workloads.build("W2", seed=0)
    prob = m.engine_problem()
    M, n_it, L = 1000, 3, 10
    g = torch.Generator().manual_seed(5)
    q0 = 0.3 * torch.randn(M, m.nweights, generator=g)
    p0 = torch.randn(n_it, M, m.nweights, generator=g)
    u = torch.rand(n_it, M, generator=g, dtype=torch.float64)
    lo, hi = D.shard(M)
    q = q0[lo:hi].to(dev).contiguous()
    acc, _, _ = prob.hmc_iterate(q, p0[:, lo:hi].to(dev).contiguous(), u[:, lo:hi].to(dev).contiguous(), m.hmc_epsilon, L, True, lanes=4)
    q_all, acc_all = D.gather_cat(q, 0), D.gather_cat(acc, 1)
    q1 = q0.to(dev).contiguous()
    acc1, _, _ = prob.hmc_iterate(q1, p0.to(dev), u.to(dev), m.hmc_epsilon, L, True, lanes=4)
    same = torch.equal(q_all, q1) and torch.equal(acc_all, acc1)
    ok &= same
    if rank == 0:
        print(f"HMC {world} ranks vs 1 rank: bit-identical = {same} (M={M}, {n_it} iterations, L={L})")

    # ---- SVGD: all-gather of particles and gradients + all-reduced select histograms ------------------------------------
    ms = workloads.build("W4", seed=0, infer_nsamples=257)
    X0 = torch.randn(257, ms.nweights, generator=g)
    st = ms.svgd_setup(X0)
    for ep in range(1, 4):
        ms.svgd_epoch(st, ep, True, use_tensor=0)
    Xs, hs = st["Xall"].clone(), st["h"].clone()
    # single-rank arm: same model code with the process group hidden
    saved = D.active
    D.active = lambda: False
    st1 = ms.svgd_setup(X0)
    for ep in range(1, 4):
        ms.svgd_epoch(st1, ep, True, use_tensor=0)
    D.active = saved
    rel = float((Xs - st1["X"]).norm() / st1["X"].norm())
    hrel = float((hs - st1["h"]).abs() / st1["h"])
    ok &= rel < 1e-5 and hrel < 1e-6
    if rank == 0:
        print(f"SVGD {world} ranks vs 1 rank: particles rel diff {rel:.2e}, bandwidth rel diff {hrel:.2e} (n=257, 3 epochs)")

    # ---- SVGD tensor path: bracket select across ranks (and the legacy radix select) vs one rank ---------------------------
    from ocbnn_b200 import engine as E
    for n_t, wl in ((2500, "W4"), (1100, "W7")):
        mt = workloads.build(wl, seed=0, infer_nsamples=n_t)
        if mt.aocp_mean is not None:
            Xt = mt.aocp_mean[None] + 3 * mt.aocp_std[None] * torch.randn(n_t, mt.nweights, generator=g)
        else:
            Xt = torch.randn(n_t, mt.nweights, generator=g)
        res = {}
        for mode in (1, 0):
            E.svgd_select_mode(mode)
            stt = mt.svgd_setup(Xt, use_tensor=1)
            for ep in range(1, 3):
                mt.svgd_epoch(stt, ep, True)
            res[mode] = (stt["Xall"].clone(), stt["h"].clone())
        E.svgd_select_mode(1)
        D.active = lambda: False
        st1 = mt.svgd_setup(Xt, use_tensor=1)
        mt.update_config(svgd_graph=False)
        for ep in range(1, 3):
            mt.svgd_epoch(st1, ep, True)
        D.active = saved
        h_same = bool(res[1][1] == res[0][1]) and bool(res[1][1] == st1["h"])
        rel = float((res[1][0] - st1["X"]).norm() / st1["X"].norm())
        ok &= h_same and rel < 1e-5
        if rank == 0:
            print(f"SVGD tensor path {wl} n={n_t}: bandwidth bit-identical (bracket select x{world}, radix select x{world}, 1 rank) = {h_same}, "
                  f"particles rel diff {rel:.2e}")

    # ---- SVGD: CUDA-graph replay of the multi-rank epoch (the NCCL collectives ocbnn_svgd_step issues are captured with the
    # kernels) against the same ranks running eagerly: same work split, so bit-identical -----------------------------------------
    mg = workloads.build("W4", seed=0, infer_nsamples=1500)
    Xg = torch.randn(1500, mg.nweights, generator=g)
    mg.update_config(svgd_graph=True)
    stg = mg.svgd_setup(Xg, use_tensor=1)
    for ep in range(1, 7):
        mg.svgd_epoch(stg, ep, True)
    captured = any(not isinstance(v, str) for v in stg.get("graphs", {}).values())
    mg.update_config(svgd_graph=False)
    ste = mg.svgd_setup(Xg, use_tensor=1)
    for ep in range(1, 7):
        mg.svgd_epoch(ste, ep, True)
    same = torch.equal(stg["Xall"], ste["Xall"]) and torch.equal(stg["h"], ste["h"])
    ok &= captured and same
    if rank == 0:
        print(f"SVGD graph replay x{world} vs eager x{world} (n=1500, 6 epochs): captured = {captured}, bit-identical = {same}")

    # ---- infer() end to end: HMC / SGLD with host RNG (streams defined over all chains -> N ranks = 1 rank bit for bit, ragged
    # shares and nchains < world included), SVGD and BBB replicated state independent of per-rank seeding ----------------------
    import copy
    for nch in (5, 1):
        outs = {}
        for arm in ("sharded", "single"):
            if arm == "single":
                D.active = lambda: False
            mi = workloads.build("W2", seed=0, nchains=nch, hmc_nburnin=2, infer_nsamples=6, hmc_ninterval=2, hmc_l=5, rng="host")
            torch.manual_seed(123)
            import numpy as _np
            _np.random.seed(123)
            mi.infer(verbose=False)
            outs[arm] = mi.all_bayes_samples[-1][0].clone()
            D.active = saved
        same = torch.equal(outs["sharded"], outs["single"])
        ok &= same
        if rank == 0:
            print(f"HMC infer() nchains={nch} x{world} ranks vs 1 rank (host RNG, same seed): bit-identical = {same}")
    outs = {}
    for arm in ("sharded", "single"):
        if arm == "single":
            D.active = lambda: False
        mi = workloads.build("W1", sampler="SGLD", seed=0, nchains=3, sgld_nburnin=3, infer_nsamples=4, sgld_ninterval=2, rng="host")
        torch.manual_seed(321)
        mi.infer(verbose=False)
        outs[arm] = mi.all_bayes_samples[-1][0].clone()
        D.active = saved
    same = torch.equal(outs["sharded"], outs["single"])
    ok &= same
    if rank == 0:
        print(f"SGLD infer() nchains=3 x{world} ranks vs 1 rank: bit-identical = {same}")
    # ranks seeded DIFFERENTLY: replicated state (SVGD particles, BBB q_params) must still agree across ranks
    mi = workloads.build("W4", seed=0, infer_nsamples=64, svgd_epochs=3)
    torch.manual_seed(1000 + rank)
    mi.infer(verbose=False)
    pr = mi.particles.to(dev)
    p0_ = pr.clone()
    dist.broadcast(p0_, src=0)
    same = torch.equal(pr, p0_)
    ok &= same
    mi = workloads.build("W1", sampler="BBB", seed=0, bbb_epochs=4, bbb_esamples=8, infer_nsamples=3)
    torch.manual_seed(2000 + rank)
    mi.infer(verbose=False)
    qr = mi.q_params.to(dev)
    q0_ = qr.clone()
    dist.broadcast(q0_, src=0)
    same_q = torch.equal(qr, q0_)
    ok &= same_q
    if rank == 0:
        print(f"ranks seeded differently: SVGD particles identical on all ranks = {same}, BBB q_params identical = {same_q}")

    # ---- BBB: all-reduce of the (D,2) gradient ----------------------------------------------------------------------------
    mb = workloads.build("W1", sampler="BBB", seed=0)
    k = 64
    eps = torch.randn(k, mb.nweights, generator=g)
    qp0 = torch.cat([torch.randn(mb.nweights, 1, generator=g), -0.5 * torch.ones(mb.nweights, 1)], 1)
    lo, hi = D.shard(k)
    qp, ac = qp0.to(dev).contiguous(), torch.zeros(mb.nweights, 2, device=dev)
    e_sh = mb.bbb_epoch(qp, ac, eps[lo:hi].to(dev).contiguous(), 1, True, k_total=k)
    D.active = lambda: False
    qp1, ac1 = qp0.to(dev).contiguous(), torch.zeros(mb.nweights, 2, device=dev)
    e_1 = mb.bbb_epoch(qp1, ac1, eps.to(dev).contiguous(), 1, True, k_total=k)
    D.active = saved
    rel = float((qp - qp1).norm() / qp1.norm())
    erel = float((e_sh - e_1).abs() / e_1.abs())
    ok &= rel < 1e-5 and erel < 1e-5
    if rank == 0:
        print(f"BBB {world} ranks vs 1 rank: q_params rel diff {rel:.2e}, ELBO rel diff {erel:.2e} (k={k})")
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL")
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if float(t) == 1.0 else 1)


if __name__ == "__main__":
    main()
