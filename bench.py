#!/usr/bin/env python
"""
bench.py — headline benchmark of the OC-BNN B200 engine.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     (CPU arm: the oracle port on the host cores)

Metric (BASELINE.json): HMC chain-leapfrog-steps/s.  Workload: W2 = repro/F1a.yaml ([1,10,1] rbf regressor on toy1,
negative-exponential COCP with 100 constraint points, posterior) with the chain count scaled up; one "step" is one HMC
iteration (L = 50 leapfrog steps + Metropolis-Hastings) of all chains.  Chains are independent: ranks hold disjoint
chains, there is no data-path collective ("scaling": "weak", per-GPU chain count fixed).

`value`  : chain-leapfrog-steps/s with q, momenta and uniforms resident in HBM (CUDA events, L2 flushed between steps,
           max over ranks).
`e2e`    : the same through host buffers: every step copies q, momenta and uniforms from pinned host memory to the device,
           runs the kernel and copies q and the accept flags back.
`roofline`: the HMC kernel against the fp32 instruction roofline measured in the same run (packed FFMA2 probe); the toy
           nets keep all state in registers, so HBM traffic is ~7 B per chain-step and neither the HBM nor the tensor
           roofline binds (fractions of those are reported for context).
`cpu_baseline`: the oracle port (the faster of its C and numpy restatements) on all host cores (bounded sample).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "ocbnn-public_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "hmc_chain_leapfrog_steps_per_s"
UNIT = "chain-leapfrog-steps/s"
WORKLOAD_NAME = "W2 repro/F1a.yaml HMC + negative-exponential COCP (100 constraint points), toy1 (6 points), [1,10,1] rbf, posterior"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="W2")
    ap.add_argument("--chains", type=int, default=262144, help="chains per GPU")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--leapfrog", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-svgd", action="store_true", help="skip the secondary SVGD particle-iters/s measurement")
    ap.add_argument("--extras", action="store_true", help="also time SVGD/SGLD/BBB and the chain-count sweep (extra JSON keys)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=0.0, help="total CPU budget of --impl reference (default: max(20, 8 s per step))")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement oracle/ocbnn_oracle.c, pinned to the numpy oracle and the reference's golden vectors)
# ---------------------------------------------------------------------------------------------------------------------

def _oracle_spec_arrays(m):
    from ocbnn_b200 import workloads
    iv = np.array([[(float(a), float(b)) for a, b in row] for row in workloads.ifunc_f1a(m._cr_neg_xsamples)])
    return dict(sizes=list(m.layer_sizes), act=m.activation, X=m.X_train.numpy(), Y=m.Y_train.numpy(), sigma_w=float(m.sigma_w),
                sigma_noise=float(m.sigma_noise), xs=m._cr_neg_xsamples.numpy(), iv=iv, S=int(m.ocp_nsamples),
                gamma=float(m.cocp_expo_gamma), tau=[float(t) for t in m.cocp_expo_tau], eps=float(m.hmc_epsilon))


def _oracle_spec(arr):
    import ocbnn_oracle as orc
    return orc.Spec(arr["sizes"], arr["act"], "regression", arr["X"], arr["Y"], sigma_w=arr["sigma_w"], sigma_noise=arr["sigma_noise"],
                    use_ocbnn=True, neg=dict(xs=arr["xs"], intervals=[arr["iv"]], S=arr["S"], gamma=arr["gamma"], tau=arr["tau"]))


def _np_worker(args):
    arr, M, L, seed = args
    import ocbnn_oracle as orc
    spec = _oracle_spec(arr)
    rng = np.random.default_rng(seed)
    q = (0.5 * rng.standard_normal((M, spec.nweights))).astype(np.float32)
    p0 = rng.standard_normal((1, M, spec.nweights)).astype(np.float32)
    u = rng.random((1, M))
    t0 = time.perf_counter()
    orc.hmc_iterate(spec, q, p0, u, arr["eps"], L)
    return time.perf_counter() - t0


def cpu_hmc_rate(arr, L, seconds, steps=1, warmup=0):
    """chain-leapfrog-steps/s of the oracle port on all host cores.  Two restatements exist: the C one (oracle/ocbnn_oracle.c ->
    oracle/_ref/libocbnn_oracle.so, scalar libm, one host thread per core via ctypes, GIL released) and the numpy one
    (oracle/ocbnn_oracle.py, vectorised over chains, one process per core).  Both are calibrated on a small sample and the
    FASTER one is timed and reported (the stronger baseline).  Each step is one HMC iteration (L leapfrog steps + MH) of a
    bounded number of chains sized for ~`seconds` of wall time in total.  Returns (rate, cores, chains per step, s per step, how)."""
    import concurrent.futures as cf
    import c_port
    cs = c_port.CSpec(_oracle_spec(arr))
    cores = os.cpu_count() or 1
    D = cs.D

    def run_c(M, seed):
        r = np.random.default_rng(seed)
        q = (0.5 * r.standard_normal((M, D))).astype(np.float32)
        p0 = r.standard_normal((1, M, D)).astype(np.float32)
        u = r.random((1, M))
        t0 = time.perf_counter()
        cs.hmc_iterate(q, p0, u, arr["eps"], L, threads=cores)
        return time.perf_counter() - t0

    n = max(1, steps + warmup)
    with cf.ProcessPoolExecutor(max_workers=cores) as ex:
        def run_np(M, seed):
            t0 = time.perf_counter()
            list(ex.map(_np_worker, [(arr, M // cores, L, seed * 1000 + c) for c in range(cores)]))
            return time.perf_counter() - t0
        run_np(cores, 0)                                   # start the workers, import numpy there
        cal = {"c": run_c(4 * cores, 1) / (4 * cores), "numpy": run_np(256 * cores, 1) / (256 * cores)}
        how = min(cal, key=cal.get)
        run = run_c if how == "c" else run_np
        M = int(max(cores, min(1 << 20, seconds / n / max(cal[how], 1e-9))))
        M -= M % cores
        times = [run(M, 100 + s) for s in range(n)][warmup:]
    tot = sum(times)
    desc = ("C restatement of the oracle (gcc -O3 AVX2, scalar libm), one thread per core" if how == "c" else
            "numpy restatement of the oracle (vectorised over chains), one process per core")
    desc += f"; faster of the two ports here (calibration s/chain-iteration: C {cal['c']:.2e}, numpy {cal['numpy']:.2e})"
    return M * L * len(times) / tot, cores, M, tot / len(times), desc


# ---------------------------------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        from ocbnn_b200 import workloads
        m = workloads.build(args.workload, seed=0)
        arr = _oracle_spec_arrays(m)
        os.chdir(cwd)
    rate, cores, M, t_step, how = cpu_hmc_rate(arr, args.leapfrog, seconds=args.ref_seconds or max(20.0, 8.0 * (args.steps + args.warmup)), steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "chains": M,
                       "leapfrog_l": args.leapfrog},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{M} chains x 1 HMC iteration (L={args.leapfrog}) per step on {cores} host cores; {how}"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # NCCL's own messages (e.g. its version banner) stay off stdout: one JSON line
        dist.init_process_group("nccl", device_id=dev)
    from ocbnn_b200 import engine, workloads

    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix=f"ocbnn_bench_r{rank}_")
    os.chdir(scratch)
    m = workloads.build(args.workload, seed=0)
    prob = m.engine_problem()
    os.chdir(cwd)
    D, L, M = m.nweights, args.leapfrog, args.chains
    K, W = args.steps, max(args.warmup, 3)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    q = 0.5 * torch.randn(M, D, device=dev, generator=gen)
    # short burn-in so that the timed steps run on equilibrated chains (untimed)
    for _ in range(2):
        prob.hmc_iterate(q, torch.randn(1, M, D, device=dev, generator=gen), torch.rand(1, M, device=dev, dtype=torch.float64, generator=gen),
                         m.hmc_epsilon, L, True, lanes=args.lanes)
    p0 = torch.randn(W + K, M, D, device=dev, generator=gen)
    u = torch.rand(W + K, M, device=dev, dtype=torch.float64, generator=gen)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step(i):
        return prob.hmc_iterate(q, p0[i:i + 1], u[i:i + 1], m.hmc_epsilon, L, True, want_flags=False, lanes=args.lanes)

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = engine.launch_count()
    evs = []
    t_wall0 = time.perf_counter()
    acc_total = 0
    for i in range(K):
        flush.fill_(i & 0xFF)                                           # L2 flush between timed steps (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        acc, _, _ = step(W + i)
        e1.record()
        evs.append((e0, e1))
        acc_total = acc_total + acc.sum()
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = engine.launch_count() - launches0
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * M * L * K / (ms_total * 1e-3)
    accept_rate = float(acc_total) / (K * M)

    # ---- end to end through the public sampler driver with host buffers ---------------------------------------------------
    # HMCMixin.hmc_run (the body of infer()): every step uploads that step's momenta and uniforms from pinned host memory
    # (side stream, overlapped with the previous step's kernel), runs the fused kernel and returns the step's sample;
    # samples and accept counts are read back to the host inside the timed region.  q stays on the device between steps,
    # exactly as in infer().  Inputs of the K steps total more than the 126 MB L2, so no explicit flush is needed here.
    e2e = None
    if not args.no_e2e:
        p0h = p0[W:W + K].cpu().pin_memory()
        uh = u[W:W + K].cpu().pin_memory()
        samples_h = torch.empty(K, M, D).pin_memory()
        qd = q.clone()
        pos = [0]

        def draw_from_host(n, m_, device, lo=0, hi=None):
            i = pos[0]
            pos[0] += n
            return p0h[i:i + n].to(device, non_blocking=True), uh[i:i + n].to(device, non_blocking=True)
        m._hmc_draw = draw_from_host
        m.update_config(lanes=args.lanes, hmc_l=L)
        os.chdir(scratch)
        m.hmc_run(qd, min(K, 3), True, collect_every=1, chunk=1, samples_host=samples_h[:min(K, 3)])   # warm the streams / allocator / pinned pages (untimed)
        best, reps_ms = None, []
        for rep in range(2):                      # the K-step run is repeated once and the faster repetition reported: a one-off host
            pos[0] = 0                            # stall on a rank (seen twice at N = 2: +170 ms in a 37 ms region) is not the path's speed
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            acc_cnt, smp = m.hmc_run(qd, K, True, collect_every=1, chunk=1, samples_host=samples_h)
            acc_host = acc_cnt.cpu()
            e1.record()
            barrier()
            ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1:
                torch.distributed.all_reduce(ms2, op=torch.distributed.ReduceOp.MAX)
            reps_ms.append(round(float(ms2.item()), 3))
            best = ms2 if best is None or float(ms2.item()) < float(best.item()) else best
        ms2 = best
        os.chdir(cwd)
        # context for the number above: this box's pinned-memory copy bandwidth (64 MiB each way, untimed region)
        pin = torch.empty(16 << 20, dtype=torch.float32).pin_memory()
        dbuf = torch.empty_like(pin, device=dev)
        link = {}
        for name, (src_, dst_) in (("h2d", (pin, dbuf)), ("d2h", (dbuf, pin))):
            dst_.copy_(src_, non_blocking=True)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(4):
                dst_.copy_(src_, non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            link[name] = round(4 * pin.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9, 1)
        del pin, dbuf
        e2e = {"value": world * M * L * K / (float(ms2.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(M * D * 4 + M * 8), "d2h_bytes_per_step": int(M * D * 4 + M * 8 // K),
               "reps_ms": reps_ms, "pinned_copy_gbs": link,
               "path": "HMCMixin.hmc_run(collect_every=1, chunk=1, samples_host=pinned): momenta+uniforms H2D per step on a side stream, that step's samples D2H on a copy stream, accept counts D2H at the end; faster of two repetitions of the K-step run (max over ranks each)"}

    # ---- second headline metric of BASELINE.json: SVGD particle-iters/s (W4 = repro/F3b.yaml, 4096 particles per GPU) --------
    svgd = None
    if not args.no_svgd:
        os.chdir(scratch)
        try:
            svgd = svgd_secondary(dev, world, barrier)
        finally:
            os.chdir(cwd)

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_hmc_small) -------------------------------------------------------------------
    flop_step, mufu_step, bytes_step = workloads.hmc_step_work(m, True)
    hbm_gbs, bf16_tf, src = load_peaks()
    peak_ffma2 = engine.probe_peak(1)
    peak_ffma = engine.probe_peak(0)
    peak_mufu = engine.probe_peak(2)
    ker_s = ms_total * 1e-3 / K                      # one launch per step: the event brackets exactly that kernel
    ach_flops = M * L * flop_step / ker_s
    ach_mufu = M * L * mufu_step / ker_s
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_hmc_ncu.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        if t.get("chains") == M and t.get("leapfrog_l") == L:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]         # per launch, from the committed ncu --set full capture
    # The toy nets keep all state on chip: ~7 B of HBM traffic per chain-step, so neither the HBM nor the tensor roofline binds.
    # The binding pipe is the one with the larger fraction of its measured peak: MUFU (exp2/rcp, 16 op/clk/SM) vs fp32 FMA.
    mufu_frac, fp32_frac = ach_mufu / peak_mufu, ach_flops / peak_ffma2
    mufu_bound = mufu_frac >= fp32_frac
    roofline = {"bound": "mufu" if mufu_bound else "fp32-simt", "kernel": "k_hmc_small",
                "achieved": (ach_mufu if mufu_bound else ach_flops) / 1e12, "peak": (peak_mufu if mufu_bound else peak_ffma2) / 1e12,
                "unit": "Top/s" if mufu_bound else "TFLOP/s", "frac": mufu_frac if mufu_bound else fp32_frac, "traffic": traffic,
                "peak_source": "instruction-pipe probes measured in this run (ocbnn_probe_peak: MUFU.EX2, packed FFMA2); MEASURED_PEAKS.json has HBM and bf16 tensor peaks only",
                "flop_per_chain_step": flop_step, "mufu_per_chain_step": mufu_step, "hbm_bytes_per_chain_step": bytes_step,
                "algorithmic_hbm_bytes_per_launch": M * L * bytes_step,
                "fp32_frac": fp32_frac, "fp32_achieved_tflops": ach_flops / 1e12, "fp32_peak_tflops": peak_ffma2 / 1e12,
                "scalar_ffma_peak_tflops": peak_ffma / 1e12, "mufu_frac": mufu_frac, "mufu_peak_tops": peak_mufu / 1e12,
                "hbm_frac_for_context": (M * L * bytes_step / ker_s) / (hbm_gbs * 1e9), "hbm_peak_gbs": hbm_gbs, "hbm_peak_source": src}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        os.chdir(scratch)
        arr = _oracle_spec_arrays(m)
        os.chdir(cwd)
        rate, cores, Mc, _, how = cpu_hmc_rate(arr, L, seconds=args.cpu_seconds)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{Mc} chains x 1 HMC iteration (L={L}) of the same workload on {cores} host cores; {how}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME if args.workload.upper() == "W2" else f"{args.workload.upper()} (see ocbnn_b200/workloads.py)",
                       "chains_per_gpu": M, "chains_total": world * M, "leapfrog_l": L, "hmc_epsilon": m.hmc_epsilon, "lanes_per_chain": args.lanes or "auto",
                       "l2": "flushed between timed steps (256 MiB write)", "accept_rate": accept_rate},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall, "svgd": svgd}
    if args.extras:
        line["extras"] = extras(m, prob, dev, args)
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def svgd_secondary(dev, world, barrier, per_gpu=4096, epochs=20, warm=3):
    """SVGD particle-iters/s through the host driver (SVGDMixin.svgd_setup / svgd_epoch, the body of infer()): W4 = repro/F3b.yaml
    ([1,10,1] rbf regressor on toy4 + negative COCP), `per_gpu` particles per GPU.  Every epoch = K1 gradients of this rank's
    particles, all-gather of particles and gradients (NCCL, N > 1), exact median bandwidth, phi on tcgen05 (3xTF32), Adagrad.
    Weak scaling in the particle count; note that the pair work grows with n^2."""
    from ocbnn_b200 import workloads
    n = per_gpu * world
    ms = workloads.build("W4", seed=0, infer_nsamples=n)
    X0 = torch.randn(n, ms.nweights, generator=torch.Generator().manual_seed(5))
    st = ms.svgd_setup(X0)
    for ep in range(1, warm + 1):
        ms.svgd_epoch(st, ep, True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for ep in range(warm + 1, warm + epochs + 1):
        ms.svgd_epoch(st, ep, True)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(t.item())
    return {"metric": "svgd_particle_iters_per_s", "value": n * epochs / (ms_total * 1e-3), "unit": "particle-iters/s", "particles_total": n,
            "particles_per_gpu": per_gpu, "epochs": epochs, "ms_per_epoch": ms_total / epochs, "tensor_path": bool(st["tensor"]),
            "workload": "W4 repro/F3b.yaml SVGD + negative COCP, toy4, [1,10,1] rbf; particles ~ N(0,1)",
            "finite": bool(torch.isfinite(st["X"]).all())}


def _time_cuda(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def extras(m, prob, dev, args):
    """Secondary measurements (single GPU): HMC chain-count sweep, SVGD particle-iters/s (W4), SGLD chain-steps/s."""
    from ocbnn_b200 import engine as E, workloads
    out = {}
    D, L = m.nweights, args.leapfrog
    sweep = {}
    for M in (4096, 32768, 262144, 1048576):
        for lanes in ((0,) if M > 4096 else (0, 8, 16, 32)):
            q = 0.5 * torch.randn(M, D, device=dev)
            p0 = torch.randn(1, M, D, device=dev)
            u = torch.rand(1, M, device=dev, dtype=torch.float64)
            t = _time_cuda(lambda: prob.hmc_iterate(q, p0, u, m.hmc_epsilon, L, True, want_flags=False, lanes=lanes), reps=2)
            sweep[f"M{M}_lanes{lanes or 'auto'}"] = M * L / t
    out["hmc_sweep_chain_leapfrog_steps_per_s"] = sweep
    cwd = os.getcwd()
    d = tempfile.mkdtemp(prefix="ocbnn_extras_")
    os.chdir(d)
    try:
        for n in (4096, 16384):
            ms = workloads.build("W4", seed=0, infer_nsamples=n)
            for ut, tag in ((1, "tcgen05"), (0, "simt")):
                st = ms.svgd_setup(torch.randn(n, ms.nweights), use_tensor=ut)
                t = _time_cuda(lambda: ms.svgd_epoch(st, 1, True), reps=5, warm=3)
                out[f"svgd_W4_n{n}_{tag}_particle_iters_per_s"] = n / t
        mg = workloads.build("W1", sampler="SGLD", seed=0)
        pg = mg.engine_problem()
        M, T = 262144, 20
        w = 0.5 * torch.randn(M, D, device=dev)
        noise = 0.01 * torch.randn(T, M, D, device=dev)
        he = np.full(T, 0.003, dtype=np.float32)
        t = _time_cuda(lambda: pg.sgld_iterate(w, noise, he, None, 1, True), reps=2)
        out["sgld_W1_chain_steps_per_s"] = M * T / t
        # HMC on the classifier workload (F2a: [2,10,3] + Dirichlet COCP) and BBB on the credit shape (W6: K1 generic net + K5)
        m3 = workloads.build("W3", seed=0)
        p3 = m3.engine_problem()
        M3 = 131072
        q3 = 0.5 * torch.randn(M3, m3.nweights, device=dev)
        p03 = torch.randn(1, M3, m3.nweights, device=dev)
        u3 = torch.rand(1, M3, device=dev, dtype=torch.float64)
        t = _time_cuda(lambda: p3.hmc_iterate(q3, p03, u3, m3.hmc_epsilon, L, True, want_flags=False), reps=2)
        out["hmc_W3_classifier_chain_leapfrog_steps_per_s"] = M3 * L / t
        m4 = workloads.build("W4P", seed=0)         # F4 shape: [2,10,1] binary classifier, toy5 (20 points), AOCP prior
        p4 = m4.engine_problem()
        M4 = 262144
        q4 = (m4.aocp_mean[None] + m4.aocp_std[None] * torch.randn(M4, m4.nweights)).to(dev).contiguous()
        p04 = torch.randn(1, M4, m4.nweights, device=dev)
        u4 = torch.rand(1, M4, device=dev, dtype=torch.float64)
        t = _time_cuda(lambda: p4.hmc_iterate(q4, p04, u4, m4.hmc_epsilon, L, True, want_flags=False), reps=2)
        out["hmc_W4P_binary_aocp_chain_leapfrog_steps_per_s"] = M4 * L / t
        k = 4096
        mb = workloads.build("W6", seed=0, bbb_esamples=k)
        qp = torch.cat([0.3 * torch.randn(mb.nweights, 1), torch.ones(mb.nweights, 1) * -3.0], 1).to(dev).contiguous()
        acc = torch.zeros_like(qp)
        eps = torch.randn(k, mb.nweights, device=dev)
        t = _time_cuda(lambda: mb.bbb_epoch(qp, acc, eps, 1, True, k_total=k), reps=2)
        out["bbb_W6_mc_sample_evals_per_s"] = k / t
    finally:
        os.chdir(cwd)
    return out


if __name__ == "__main__":
    main()
