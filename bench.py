#!/usr/bin/env python
"""
bench.py — headline benchmarks of the OC-BNN B200 engine.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W     (CPU arm: the oracle port on all host cores)
    python bench.py --workload W4|W6|W7 ...                    (the other workloads of BASELINE.json as the headline line)

BASELINE.json's metric is two numbers: HMC chain-leapfrog-steps/s and SVGD particle-iters/s.

Default line (W2): HMC chain-leapfrog-steps/s on repro/F1a.yaml ([1,10,1] rbf regressor on toy1, negative-exponential COCP with
100 constraint points, posterior), 262144 chains per GPU; one "step" is one HMC iteration (L = 50 leapfrog steps +
Metropolis-Hastings) of all chains.  Chains are independent: no data-path collective ("scaling": "weak").  The line also
carries `svgd`: the SVGD half of the metric (W4 = repro/F3b.yaml, 4096 particles per GPU) as a complete measurement of its
own — value, e2e, roofline, cpu_baseline.

`--workload W4`: the same SVGD measurement as the headline line (one step = one SVGD epoch = `ocbnn_svgd_step`).
`--workload W7`: SVGD on the recid.yaml shape ([9,100,100,1], D = 11201, AOCP prior, 62-point minibatches), 4096 particles in
                 total sharded over the GPUs ("scaling": "strong"; 183.5 MB all-gather per epoch at N > 1).
`--workload W6`: BBB on the credit.yaml shape ([10,50,50,2], D = 3202, Dirichlet COCP with 500 points, 2185-point minibatches),
                 4096 Monte-Carlo samples in total sharded over the GPUs ("scaling": "strong"); one step = one `ocbnn_bbb_step`.

`value`  : throughput with all inputs resident in HBM (CUDA events, L2 flushed between timed steps, max over ranks).
`e2e`    : the same through host buffers: inputs copied from pinned host memory and results read back inside the timed region.
`roofline`: the dominant kernel against the pipe that bounds it (MUFU / fp32 FMA probes measured in the same run; tensor = TF32
           rate derived from MEASURED_PEAKS.json's bf16 figure; HBM from MEASURED_PEAKS.json).
`cpu_baseline`: the oracle port (C / numpy restatements under oracle/) on all host cores on a bounded sample, plus
           `reference_unmodified`: the reference's own `single()` timed as it is (baseline/_ref, one chain / 50 particles).
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

HMC_METRIC, HMC_UNIT = "hmc_chain_leapfrog_steps_per_s", "chain-leapfrog-steps/s"
SVGD_METRIC, SVGD_UNIT = "svgd_particle_iters_per_s", "particle-iters/s"
BBB_METRIC, BBB_UNIT = "bbb_mc_sample_evals_per_s", "MC-sample-evals/s"
NAMES = {
    "W2": "W2 repro/F1a.yaml HMC + negative-exponential COCP (100 constraint points), toy1 (6 points), [1,10,1] rbf, posterior",
    "W4": "W4 repro/F3b.yaml SVGD + negative-exponential COCP (100 constraint points), toy4 (4 points), [1,10,1] rbf, D = 31; particles ~ N(0,1)",
    "W6": "W6 repro/credit.yaml shape: BBB, classifier [10,50,50,2] D = 3202, Dirichlet COCP (500 points), synthetic 131072 x 10 data, nbatches 60 (2185 points per step)",
    "W7": "W7 repro/recid.yaml shape: SVGD, binary classifier [9,100,100,1] D = 11201, AOCP diagonal-Gaussian prior, synthetic 6172 x 9 data, nbatches 100 (62 points per step)",
}
WORKLOAD_NAME = NAMES["W2"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="W2")
    ap.add_argument("--chains", type=int, default=262144, help="HMC chains per GPU")
    ap.add_argument("--particles", type=int, default=4096, help="SVGD particles per GPU (W4) / in total (W7); BBB samples in total (W6)")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--leapfrog", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-svgd", action="store_true", help="W2 line: skip the SVGD half of the metric")
    ap.add_argument("--extras", action="store_true", help="also time SVGD/SGLD/BBB and the chain-count sweep (extra JSON keys)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=0.0, help="total CPU budget of --impl reference (default: max(20, 8 s per step))")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement oracle/ocbnn_oracle.c + numpy oracle, pinned to the reference's golden vectors)
# ---------------------------------------------------------------------------------------------------------------------

def oracle_spec(m, wl):
    """ocbnn_oracle.Spec of a workloads.build() model (the oracle never sees the engine)."""
    import ocbnn_oracle as orc
    from ocbnn_b200 import workloads
    wl = wl.upper()
    kw = dict(sigma_w=float(m.sigma_w), sigma_noise=float(m.sigma_noise), use_ocbnn=True)
    task = {"W2": "regression", "W4": "regression", "W6": "classifier", "W7": "binary"}[wl]
    if wl in ("W2", "W4"):
        fn = workloads.ifunc_f1a if wl == "W2" else workloads.ifunc_f3b
        iv = np.array([[(float(a), float(b)) for a, b in row] for row in fn(m._cr_neg_xsamples)])
        kw["neg"] = dict(xs=m._cr_neg_xsamples.numpy(), intervals=[iv], S=int(m.ocp_nsamples), gamma=float(m.cocp_expo_gamma),
                         tau=[float(t) for t in m.cocp_expo_tau])
    elif wl == "W6":
        kw["dir"] = dict(xs=m._cr_xsamples.numpy(), forbidden=[[1]], S=int(m.ocp_nsamples), gamma_d=float(m.cocp_dirichlet_gamma),
                         alpha=float(m.cocp_dirichlet_alpha))
    else:
        kw.update(aocp_mean=m.aocp_mean.numpy(), aocp_std=m.aocp_std.numpy(), naive_sigmoid=False)
    return orc.Spec(list(m.layer_sizes), m.activation, task, m.X_train.numpy(), m.Y_train.numpy(), **kw)


def _spec_pickle(m, wl):
    """Everything a worker process needs to rebuild the Spec (plain numpy)."""
    s = oracle_spec(m, wl)
    return dict(layer_sizes=s.layer_sizes, activation=s.activation, task=s.task, X=s.X_train, Y=s.Y_train, sigma_w=s.sigma_w,
                sigma_noise=s.sigma_noise, aocp_mean=s.aocp_mean, aocp_std=s.aocp_std, neg=s.neg, pos=s.pos, dir=s.dir,
                naive_sigmoid=s.naive_sigmoid, eps=float(getattr(m, "hmc_epsilon", 0.003)))


def _spec_from_pickle(a):
    import ocbnn_oracle as orc
    return orc.Spec(a["layer_sizes"], a["activation"], a["task"], a["X"], a["Y"], sigma_w=a["sigma_w"], sigma_noise=a["sigma_noise"],
                    use_ocbnn=True, aocp_mean=a["aocp_mean"], aocp_std=a["aocp_std"], neg=a["neg"], pos=a["pos"], dir=a["dir"],
                    naive_sigmoid=a["naive_sigmoid"])


def _np_worker(args):
    arr, M, L, seed = args
    import ocbnn_oracle as orc
    spec = _spec_from_pickle(arr)
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
    cs = c_port.CSpec(_spec_from_pickle(arr))
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


def cpu_svgd_rate(m, wl, n, seconds, steps=1, warmup=0):
    """particle-iters/s of the oracle's C port (oracle/ocbnn_oracle.c: K1, pair distances, phi rows; lower median by
    numpy.partition) on all host cores.  An epoch at n particles costs O(n^2 D); when a full epoch would exceed the budget the
    three phases are timed on a row SAMPLE (K1 of R particles, the distances of R rows spread over the triangle, phi of R rows
    against all n particles) and scaled to the full epoch - the sample is stated.  Returns (rate, cores, s per epoch, sample)."""
    import c_port
    import ocbnn_oracle as orc
    spec = oracle_spec(m, wl)
    cores = os.cpu_count() or 1
    D = spec.nweights
    rng = np.random.default_rng(3)
    if spec.aocp_mean is not None:
        X = (spec.aocp_mean[None] + 3 * spec.aocp_std[None] * rng.standard_normal((n, D))).astype(np.float32)
    else:
        X = rng.standard_normal((n, D)).astype(np.float32)
    nb = int(getattr(m, "nbatches", 0) or 0)
    mult = 1.0
    if nb > 1:                                             # the C port evaluates full batches: one strided minibatch becomes its data set
        bi = orc.batch_indices(spec.N_train, 1, nb)        # (the N/|B| likelihood multiplier only scales a gradient term: same work)
        spec.X_train, spec.Y_train = spec.X_train[bi], spec.Y_train[bi]
    cs = c_port.CSpec(spec)
    lib, fp = c_port.lib(), c_port._fp
    Xp = np.ascontiguousarray(X)

    def grads(R):
        return np.concatenate(c_port._fan_out(R, cores, lambda lo, hi: cs.logpost_grad(Xp[lo:hi], True)[1]), 0)

    def dists(rows):
        def run(lo, hi):
            out = []
            for i in rows[lo:hi]:
                buf = np.empty(n - 1 - int(i), dtype=np.float32)
                if len(buf):
                    lib.oc_svgd_pair_dists(Xp.ctypes.data_as(fp), n, D, int(i), int(i) + 1, buf.ctypes.data_as(fp))
                out.append(buf)
            return np.concatenate(out) if out else np.empty(0, np.float32)
        d = np.concatenate(c_port._fan_out(len(rows), cores, run))
        d = d[d != 0]
        k = (len(d) - 1) // 2
        return float(np.partition(d, k)[k]) ** 2 / math.log(n)

    def phi(R, G, h):
        def run(lo, hi):
            out = np.empty((hi - lo, D), dtype=np.float32)
            lib.oc_svgd_phi_rows(Xp.ctypes.data_as(fp), G.ctypes.data_as(fp), n, D, h, lo, hi, out.ctypes.data_as(fp))
            return out
        return np.concatenate(c_port._fan_out(R, cores, run), 0)

    def sampled_epoch(R):
        t0 = time.perf_counter()
        G = np.zeros_like(Xp)
        G[:R] = grads(R)
        h = dists(np.linspace(0, n - 1, R).astype(int))    # rows spread over the triangle: their mean pair count is (n - 1) / 2
        phi(R, G, h)
        return (time.perf_counter() - t0) * n / R

    nrep = max(1, steps + warmup)
    R0 = min(n, cores)
    est = sampled_epoch(R0)                                # calibration
    if est * nrep <= seconds or n <= 2 * cores:
        acc = np.zeros_like(X)
        times = []
        for _ in range(nrep):
            t0 = time.perf_counter()
            X, acc, _, _ = c_port.svgd_step(cs, X, acc, float(m.svgd_init_lr), True, threads=cores)
            times.append(time.perf_counter() - t0)
        times = times[warmup:]
        t_epoch = sum(times) / len(times)
        return n / t_epoch, cores, t_epoch, (f"{len(times)} full SVGD epoch(s) at n = {n} particles (K1 + exact lower median of all {n * (n - 1) // 2} pair "
                                             f"distances + phi + Adagrad), C port of the oracle, {cores} host threads")
    R = int(max(cores, min(n, seconds / nrep / max(est, 1e-9) * n)))
    R = max(cores, R - R % cores)
    times = [sampled_epoch(R) for _ in range(nrep)][warmup:]
    t_epoch = sum(times) / len(times)
    return n / t_epoch, cores, t_epoch, (f"row sample of one SVGD epoch at n = {n}: K1 of {R} particles, pair distances of {R} rows spread over the triangle "
                                         f"(+ their lower median), phi of {R} rows against all {n} particles, scaled by n/R = {n / R:.1f}; C port of the oracle, {cores} host threads")


def _bbb_worker(args):
    arr, k, seed, nb = args
    import ocbnn_oracle as orc
    spec = _spec_from_pickle(arr)
    rng = np.random.default_rng(seed)
    D = spec.nweights
    qp = np.stack([0.3 * rng.standard_normal(D), -3.0 * np.ones(D)], 1).astype(np.float32)
    eps = rng.standard_normal((k, D)).astype(np.float32)
    bi = orc.batch_indices(spec.N_train, 1, nb) if nb > 1 else None
    t0 = time.perf_counter()
    orc.bbb_elbo_grad(spec, qp, eps, bi)
    return time.perf_counter() - t0


def cpu_bbb_rate(m, wl, seconds, steps=1, warmup=0):
    """MC-sample-evals/s of the numpy oracle (ocbnn_oracle.bbb_elbo_grad: sample, log posterior + closed-form gradient on one
    strided minibatch, ELBO gradient), one process per core, a bounded number of Monte-Carlo samples per step."""
    import concurrent.futures as cf
    arr = _spec_pickle(m, wl)
    nb = int(getattr(m, "nbatches", 0) or 0)
    cores = os.cpu_count() or 1
    nrep = max(1, steps + warmup)
    with cf.ProcessPoolExecutor(max_workers=cores) as ex:
        list(ex.map(_bbb_worker, [(arr, 1, c, nb) for c in range(cores)]))
        t0 = time.perf_counter()
        list(ex.map(_bbb_worker, [(arr, 2, 10 + c, nb) for c in range(cores)]))
        per = (time.perf_counter() - t0) / 2
        k = int(max(1, min(64, seconds / nrep / max(per, 1e-9))))
        times = []
        for s in range(nrep):
            t0 = time.perf_counter()
            list(ex.map(_bbb_worker, [(arr, k, 100 * s + c, nb) for c in range(cores)]))
            times.append(time.perf_counter() - t0)
    times = times[warmup:]
    t_step = sum(times) / len(times)
    return k * cores / t_step, cores, t_step, (f"{k * cores} Monte-Carlo samples per step ({k} per process, {cores} processes), numpy oracle, "
                                               "one strided minibatch + all constraint points each")


def reference_unmodified(skip_hmc=False, skip_svgd=False, timeout=600):
    """The reference as it is (baseline/_ref): baseline/time_reference.py in its own process."""
    script = os.path.join(ROOT, "baseline", "time_reference.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "bnn")):
        return {"unavailable": "baseline/_ref is missing (made by __graft_entry__.build() where /root/reference exists)"}
    cmd = [sys.executable, script] + (["--skip-hmc"] if skip_hmc else []) + (["--skip-svgd"] if skip_svgd else [])
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=tempfile.mkdtemp(prefix="ocbnn_refarm_"))
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        return json.loads(lines[-1]) if lines else {"unavailable": (r.stderr or "no output")[-300:]}
    except Exception as exc:
        return {"unavailable": repr(exc)[:300]}


def _ref_part(ref, key):
    """sub-dict `key` of reference_unmodified()'s result, with the process-level facts (threads, torch) carried along"""
    if key not in ref:
        return ref
    d = dict(ref[key])
    d.update({k: ref[k] for k in ("torch", "torch_threads", "cpu_count", "source") if k in ref})
    return d


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
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name, **match):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/<name>), if it was
    taken at this size."""
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    if all(t.get(k) == v for k, v in match.items()):
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    return None


class Ctx:
    """Process / rank context shared by the workload functions."""

    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # NCCL's own messages stay off stdout: one JSON line
            dist.init_process_group("nccl", device_id=self.dev)
        self.cwd = os.getcwd()
        self.scratch = tempfile.mkdtemp(prefix=f"ocbnn_bench_r{self.rank}_")
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(self, step, K, first=0):
        """K steps, each bracketed by its own CUDA events on the current stream, L2 flushed before each (not timed); barrier +
        synchronize on both sides; returns (sum of the step times in ms as the max over ranks, wall seconds)."""
        self.barrier()
        evs = []
        t0 = time.perf_counter()
        for i in range(K):
            self.flush.fill_(i & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(first + i)
            e1.record()
            evs.append((e0, e1))
        self.barrier()
        wall = time.perf_counter() - t0
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)), wall

    def close(self):
        if self.world > 1:
            torch.distributed.destroy_process_group()


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


def _pinned_bw(dev):
    """this box's pinned-memory copy bandwidth (64 MiB each way), context for the e2e numbers"""
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
    return link


# ---------------------------------------------------------------------------------------------------------------------
# HMC (W2): the default headline
# ---------------------------------------------------------------------------------------------------------------------

def bench_hmc(cx):
    from ocbnn_b200 import engine, workloads
    args, world, rank, dev = cx.args, cx.world, cx.rank, cx.dev
    os.chdir(cx.scratch)
    m = workloads.build("W2", seed=0)
    prob = m.engine_problem()
    os.chdir(cx.cwd)
    D, L, M = m.nweights, args.leapfrog, args.chains
    K, W = args.steps, max(args.warmup, 3)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    q = 0.5 * torch.randn(M, D, device=dev, generator=gen)
    for _ in range(2):              # short burn-in so that the timed steps run on equilibrated chains (untimed)
        prob.hmc_iterate(q, torch.randn(1, M, D, device=dev, generator=gen), torch.rand(1, M, device=dev, dtype=torch.float64, generator=gen),
                         m.hmc_epsilon, L, True, lanes=args.lanes)
    p0 = torch.randn(W + K, M, D, device=dev, generator=gen)
    u = torch.rand(W + K, M, device=dev, dtype=torch.float64, generator=gen)
    acc_total = [0]

    def step(i):
        acc, _, _ = prob.hmc_iterate(q, p0[i:i + 1], u[i:i + 1], m.hmc_epsilon, L, True, want_flags=False, lanes=args.lanes)
        acc_total[0] = acc_total[0] + acc.sum()

    for i in range(W):
        step(i)
    acc_total[0] = 0
    sampler = ClockSampler(cx.local)
    cx.barrier()
    sampler.start()
    launches0 = engine.launch_count()
    ms_total, wall = cx.timed_steps(step, K, first=W)
    launches = engine.launch_count() - launches0
    clocks = sampler.stop()
    value = world * M * L * K / (ms_total * 1e-3)
    accept_rate = float(acc_total[0]) / (K * M)

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
        m._hmc_draw = draw_from_host        # the momenta / uniforms are injected (pre-drawn, pinned): the timed path is copy + kernel, not host RNG
        m.update_config(lanes=args.lanes, hmc_l=L)
        os.chdir(cx.scratch)
        m.hmc_run(qd, min(K, 3), True, collect_every=1, chunk=1, samples_host=samples_h[:min(K, 3)])   # warm the streams / allocator / pinned pages
        best, reps_ms = None, []
        for rep in range(2):                      # the K-step run is repeated once and the faster repetition reported: a one-off host
            pos[0] = 0                            # stall on a rank (seen at N = 2: +170 ms in a 37 ms region) is not the path's speed
            cx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            acc_cnt, smp = m.hmc_run(qd, K, True, collect_every=1, chunk=1, samples_host=samples_h)
            acc_cnt.cpu()
            e1.record()
            cx.barrier()
            t = cx.max_over_ranks(e0.elapsed_time(e1))
            reps_ms.append(round(t, 3))
            best = t if best is None or t < best else best
        os.chdir(cx.cwd)
        # what infer() does by default (rng = 'host', several chains): ONE torch.randn + numpy call per chunk for all chains
        # (round 1: a Python loop per chain); timed here on a bounded chain count so the line says what the default path costs
        m_h = 16384
        del m._hmc_draw
        m.update_config(rng="host")
        t0 = time.perf_counter()
        m._hmc_draw(1, m_h, torch.device("cpu"))
        host_rng_s = time.perf_counter() - t0
        e2e = {"value": world * M * L * K / (best * 1e-3), "unit": HMC_UNIT,
               "h2d_bytes_per_step": int(M * D * 4 + M * 8), "d2h_bytes_per_step": int(M * D * 4 + M * 8 // K),
               "reps_ms": reps_ms, "pinned_copy_gbs": _pinned_bw(dev),
               "host_rng_default_path": {"chains": m_h, "seconds_per_iteration": host_rng_s,
                                         "note": "rng='host' draw of one iteration's momenta + uniforms for these chains on one host thread (vectorised; rng='device' draws on the GPU)"},
               "path": "HMCMixin.hmc_run(collect_every=1, chunk=1, samples_host=pinned): injected (pre-drawn) momenta+uniforms H2D per step on a side stream, that step's samples D2H on a copy stream, accept counts D2H at the end; faster of two repetitions of the K-step run (max over ranks each)"}

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (k_hmc_small) -------------------------------------------------------------------
    flop_step, mufu_step, bytes_step = workloads.hmc_step_work(m, True)
    hbm_gbs, bf16_tf, src = load_peaks()
    peak_ffma2, peak_ffma, peak_mufu = engine.probe_peak(1), engine.probe_peak(0), engine.probe_peak(2)
    ker_s = ms_total * 1e-3 / K                      # one launch per step: the event brackets exactly that kernel
    ach_flops, ach_mufu = M * L * flop_step / ker_s, M * L * mufu_step / ker_s
    traffic = ncu_traffic("r02_hmc_ncu.json", chains=M, leapfrog_l=L)
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
        os.chdir(cx.scratch)
        arr = _spec_pickle(m, "W2")
        os.chdir(cx.cwd)
        rate, cores, Mc, _, how = cpu_hmc_rate(arr, L, seconds=args.cpu_seconds)
        cpu = {"value": rate, "unit": HMC_UNIT, "cores": cores, "kind": "port",
               "sample": f"{Mc} chains x 1 HMC iteration (L={L}) of the same workload on {cores} host cores; {how}"}

    return {"metric": HMC_METRIC, "value": value, "unit": HMC_UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": NAMES["W2"], "chains_per_gpu": M, "chains_total": world * M, "leapfrog_l": L, "hmc_epsilon": m.hmc_epsilon,
                       "lanes_per_chain": args.lanes or "auto", "l2": "flushed between timed steps (256 MiB write)", "accept_rate": accept_rate},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall}


# ---------------------------------------------------------------------------------------------------------------------
# SVGD (W4 / W7): the second half of BASELINE.json's metric
# ---------------------------------------------------------------------------------------------------------------------

def bench_svgd(cx, wl, K=None, W=None, with_clocks=False):
    """SVGD particle-iters/s through the sampler driver (SVGDMixin.svgd_setup / svgd_epoch, the body of infer()): one step = one
    epoch = ONE `ocbnn_svgd_step` call (K1 gradients of this rank's particles, NCCL all-gather inside the library, exact median
    bandwidth, phi on tcgen05 in 3xTF32, Adagrad, all-gather), replayed from a CUDA graph."""
    from ocbnn_b200 import engine as E, workloads
    args, world, rank = cx.args, cx.world, cx.rank
    wl = wl.upper()
    K = args.steps if K is None else K
    W = max(args.warmup, 3) if W is None else W
    weak = wl == "W4"
    n = args.particles * world if weak else args.particles
    os.chdir(cx.scratch)
    ms = workloads.build(wl, seed=0, infer_nsamples=n)
    g = torch.Generator().manual_seed(5)
    if ms.aocp_mean is not None:
        X0 = ms.aocp_mean[None] + 3 * ms.aocp_std[None] * torch.randn(n, ms.nweights, generator=g)
    else:
        X0 = torch.randn(n, ms.nweights, generator=g)
    D = ms.nweights
    nb = int(getattr(ms, "nbatches", 0) or 0)
    npts = (len(range(1, ms.N_train, nb)) if nb > 1 else ms.N_train) + ms.engine_problem().cp.cx.shape[0]
    st = ms.svgd_setup(X0)
    rows = st["hi"] - st["lo"]
    # launches of one epoch (eager), then the graph path: two minibatch offsets alternate so that both graphs exist after warm-up
    ms.update_config(svgd_graph=False)
    l0 = E.launch_count()
    ms.svgd_epoch(st, 1, True)
    launches_per_epoch = E.launch_count() - l0
    ms.update_config(svgd_graph=True)
    ep_of = (lambda i: 1 + (i % 2)) if nb > 1 else (lambda i: 1)
    for i in range(max(W, 3) * 2):
        ms.svgd_epoch(st, ep_of(i), True)
    sampler = ClockSampler(cx.local) if with_clocks else None
    if sampler:
        cx.barrier()
        sampler.start()
    ms_total, wall = cx.timed_steps(lambda i: ms.svgd_epoch(st, ep_of(i), True), K)
    clocks = sampler.stop() if sampler else None
    value = n * K / (ms_total * 1e-3)
    finite = bool(torch.isfinite(st["Xall"]).all())

    # ---- e2e: what infer() does around the epochs - particles in from pinned host memory, K epochs, particles back ----------
    e2e = None
    if not args.no_e2e:
        xh = X0.clone().pin_memory()
        outh = torch.empty_like(xh).pin_memory()
        best, reps_ms = None, []
        for rep in range(2):
            st["acc"].zero_()
            cx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st["Xall"].copy_(xh, non_blocking=True)
            for i in range(K):
                ms.svgd_epoch(st, ep_of(i), True)
            outh.copy_(st["Xall"], non_blocking=True)
            e1.record()
            cx.barrier()
            t = cx.max_over_ranks(e0.elapsed_time(e1))
            reps_ms.append(round(t, 3))
            best = t if best is None or t < best else best
        e2e = {"value": n * K / (best * 1e-3), "unit": SVGD_UNIT, "h2d_bytes_per_step": int(n * D * 4 // K), "d2h_bytes_per_step": int(n * D * 4 // K),
               "reps_ms": reps_ms,
               "path": f"as SVGDMixin.infer(): all {n} particles H2D from pinned host memory ({n * D * 4} B), {K} epochs (svgd_epoch -> ocbnn_svgd_step, graph replay), particles D2H; bytes per step = those copies amortised over the {K} epochs; faster of two repetitions (max over ranks each)"}

    # ---- stage times on this rank (CUDA events, eager calls of the same kernels; the workspace was prepared by the last step) ----
    Xall, Gall, h = st["Xall"], st["Gall"], st["h"]
    tensor = bool(st["tensor"])
    batch = (1, nb) if nb > 1 else (0, 1)
    t_grad = _time_cuda(lambda: st["prob"].logpost_grad(st["X"], batch, True, want_lp=False, out_grad=st["G"]), reps=5, warm=2)
    t_phi = _time_cuda(lambda: E.svgd_phi(Xall, Gall, h, st["lo"], rows, st["ws"], 2 if tensor else 0, out=st["phi"]), reps=5, warm=2)
    t_bw = None
    if world == 1:
        t_bw = _time_cuda(lambda: E.svgd_bandwidth(Xall, st["ws"], h, use_tensor=1 if tensor else 0), reps=5, warm=2)
    os.chdir(cx.cwd)
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel: the phi stage -------------------------------------------------------------------
    hbm_gbs, bf16_tf, src = load_peaks()
    dp = (D + 31) // 32 * 32
    tf32_peak = bf16_tf / 2.0                         # dense TF32 rate = half the bf16 rate on tcgen05 (kind::tf32 vs kind::f16)
    flash = dp <= 64
    # algorithmic fp32-equivalent FLOP of phi for this rank's rows: S = Xc Xc^T (2 rows n D, recomputed only on the flash path) and
    # O = K [G | Xc] (2 rows n 2D); each runs as THREE tf32 MMAs per product (hi.hi + hi.lo + lo.hi) on padded D
    fp32_flop = (2.0 * rows * n * D if flash else 0.0) + 4.0 * rows * n * D
    mma_flop = 3.0 * ((2.0 * rows * n * dp if flash else 0.0) + 4.0 * rows * n * dp)
    kernel = "k_phi_flash (+ k_vt_split, k_phi_finalize in the same event bracket)" if flash else \
        "k_kmat_split + k_gemm_tf32x3<EpiStore> + k_vt_split + k_phi_finalize (phi stage, D > 64)"
    roofline = {"bound": "tensor", "kernel": kernel, "achieved": mma_flop / t_phi / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": mma_flop / t_phi / 1e12 / tf32_peak,
                "traffic": ncu_traffic("r02_svgd_phi_ncu.json", workload=wl, particles=n, rows=rows),
                "peak_source": f"TF32 dense = bf16 / 2, bf16 = {bf16_tf} TFLOP/s {src}",
                "stage_seconds": t_phi, "rows": rows, "columns": n, "d": D, "d_padded": dp,
                "tf32_mma_flop_per_launch": mma_flop, "fp32_equivalent_flop_per_launch": fp32_flop,
                "fp32_equivalent_tflops": fp32_flop / t_phi / 1e12,
                "algorithmic_hbm_bytes_per_launch": int(4 * (3 * n * D + rows * D)),
                "mufu_floor_seconds": rows * n / E.probe_peak(2), "tensor_floor_seconds": mma_flop / (tf32_peak * 1e12),
                "stages_ms": {"k1_gradients": t_grad * 1e3, "k3a_bandwidth": None if t_bw is None else t_bw * 1e3, "k3b_phi": t_phi * 1e3,
                              "epoch": ms_total / K},
                "note": "3xTF32: an fp32-accurate product costs three tf32 MMAs, so `achieved` counts issued tensor work; fp32_equivalent_tflops is the useful rate"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        os.chdir(cx.scratch)
        rate, cores, t_ep, sample = cpu_svgd_rate(ms, wl, n, seconds=args.cpu_seconds)
        os.chdir(cx.cwd)
        cpu = {"value": rate, "unit": SVGD_UNIT, "cores": cores, "kind": "port", "seconds_per_epoch": t_ep, "sample": sample}

    return {"metric": SVGD_METRIC, "value": value, "unit": SVGD_UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32 (3xTF32 on tcgen05 for the contractions)",
            "data": "synthetic",
            "config": {"workload": NAMES[wl], "particles_total": n, "particles_per_gpu": rows, "weights_per_particle": D, "points_per_gradient": int(npts),
                       "tensor_path": tensor, "cuda_graph": True, "l2": "flushed between timed steps (256 MiB write)", "finite": finite,
                       "collectives_per_step": 0 if world == 1 else "2 all-gathers (G, X: %d B each) + 4 small all-reduces of the median select, NCCL inside ocbnn_svgd_step" % (n * D * 4)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_epoch * K),
            "gpu_launches_note": f"{launches_per_epoch} kernels per epoch (counted on an eager epoch), replayed from a CUDA graph",
            "roofline": roofline, "cpu_baseline": cpu, "wall_s_timed_region": wall}


# ---------------------------------------------------------------------------------------------------------------------
# BBB (W6)
# ---------------------------------------------------------------------------------------------------------------------

def bench_bbb(cx, wl="W6"):
    """BBB MC-sample-evals/s through BBBMixin.bbb_epoch: one step = ONE `ocbnn_bbb_step` call (sample, K1 on this rank's Monte-Carlo
    samples, closed-form (D,2) ELBO gradient, one all-reduce, Adagrad)."""
    from ocbnn_b200 import dist as Dm, engine as E, workloads
    from ocbnn_b200.inference import H2DPipe
    args, world, rank, dev = cx.args, cx.world, cx.rank, cx.dev
    K, W = args.steps, max(args.warmup, 3)
    k = args.particles
    os.chdir(cx.scratch)
    mb = workloads.build(wl, seed=0, bbb_esamples=k)
    D = mb.nweights
    nb = int(getattr(mb, "nbatches", 0) or 0)
    npts = (len(range(1, mb.N_train, nb)) if nb > 1 else mb.N_train) + mb.engine_problem().cp.cx.shape[0]
    lo, hi = Dm.shard(k)
    g = torch.Generator().manual_seed(9)
    qp = torch.cat([0.3 * torch.randn(D, 1, generator=g), -3.0 * torch.ones(D, 1)], 1).to(dev).contiguous()
    acc = torch.zeros_like(qp)
    eps_all = torch.randn(W + K, hi - lo, D, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
    state = {}
    elbo = [None]

    def step(i):
        elbo[0] = mb.bbb_epoch(qp, acc, eps_all[i], 1 + (i % 2), True, k_total=k, state=state)
    l0 = E.launch_count()
    step(0)
    launches_per_step = E.launch_count() - l0
    for i in range(1, W):
        step(i)
    sampler = ClockSampler(cx.local)
    cx.barrier()
    sampler.start()
    ms_total, wall = cx.timed_steps(step, K, first=W)
    clocks = sampler.stop()
    value = k * K / (ms_total * 1e-3)
    e2e = None
    if not args.no_e2e:
        eps_h = eps_all[W:W + K].cpu().pin_memory()
        best, reps_ms = None, []
        for rep in range(2):
            cx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            elbos = []
            pipe = H2DPipe(tuple(eps_h[0].shape), dev)            # what BBBMixin.infer() uses: the next epoch's draws copy under this epoch's kernels
            slot = pipe.put(eps_h[0])
            for i in range(K):
                cur, slot = slot, (pipe.put(eps_h[i + 1]) if i + 1 < K else None)
                elbos.append(mb.bbb_epoch(qp, acc, pipe.take(cur), 1 + (i % 2), True, k_total=k, state=state))
                pipe.done(cur)
            torch.cat(elbos).cpu()
            e1.record()
            cx.barrier()
            t = cx.max_over_ranks(e0.elapsed_time(e1))
            reps_ms.append(round(t, 3))
            best = t if best is None or t < best else best
        e2e = {"value": k * K / (best * 1e-3), "unit": BBB_UNIT, "h2d_bytes_per_step": int((hi - lo) * D * 4), "d2h_bytes_per_step": 4, "reps_ms": reps_ms,
               "path": "as BBBMixin.infer() with rng='host': this rank's (k_local, D) standard-normal draws H2D from pinned memory every epoch on a copy stream, two deep (inference.H2DPipe: the copy of epoch e+1 under the kernels of epoch e), bbb_epoch -> ocbnn_bbb_step, the ELBO scalars D2H at the end; faster of two repetitions"}
    # K1 alone (the dominant kernel) on this rank's samples
    Wt = E.bbb_sample(qp, eps_all[0])
    batch = (1, nb) if nb > 1 else (0, 1)
    t_k1 = _time_cuda(lambda: mb.engine_problem().logpost_grad(Wt, batch, True), reps=3, warm=1)
    os.chdir(cx.cwd)
    if rank != 0:
        return None
    peak_ffma2 = E.probe_peak(1)
    flop_pt = workloads.flops_per_point(mb.layer_sizes)
    flop = float(hi - lo) * npts * flop_pt
    tc_k1 = os.environ.get("OCBNN_GEN_TC", "1") != "0"
    ls = mb.layer_sizes
    up16 = lambda v: (v + 15) // 16 * 16
    # bf16 MMA work the tcgen05 kernel issues per point (padded operand extents, 6 part products, 2 FLOP per MAC): GEMM1 + GEMM2 + GEMM3
    macs = sum(up16(ls[l - 1] + 1) * up16(ls[l]) for l in range(1, len(ls))) + sum(up16(ls[l]) * up16(ls[l - 1]) for l in range(2, len(ls))) \
        + sum(min(64 * -(-(ls[l - 1] + 1) // 64) * up16(ls[l]), 64 * -(-ls[l] // 64) * up16(ls[l - 1] + 1)) for l in range(1, len(ls)))
    bf16_peak = load_peaks()[1]
    roofline = {"bound": "fp32-simt",
                "kernel": ("k_logpost_gen_tc (K1 on the credit net: tcgen05 kind::f16, every fp32 operand split into three bf16 parts, six MMAs per product; "
                           "the result is fp32-accurate, so the fp32 FMA peak is the yardstick and the issued tensor work is reported beside it)") if tc_k1 else
                          "k_logpost_generic (K1 on the credit net; mma.sync 3xTF32 variant, fp32-accurate result: the fp32 FMA peak is the yardstick)",
                "achieved": flop / t_k1 / 1e12, "peak": peak_ffma2 / 1e12, "unit": "TFLOP/s", "frac": flop / t_k1 / peak_ffma2,
                "tensor_issued_tflops": (float(hi - lo) * npts * macs * 12 / t_k1 / 1e12) if tc_k1 else None,
                "tensor_issued_frac_of_bf16_peak": (float(hi - lo) * npts * macs * 12 / t_k1 / 1e12 / bf16_peak) if (tc_k1 and bf16_peak) else None,
                "traffic": ncu_traffic("r02_generic_k1_ncu.json", workload=wl, samples=hi - lo),
                "peak_source": "packed FFMA2 probe measured in this run (ocbnn_probe_peak)", "flop_per_point": flop_pt, "points_per_sample": int(npts),
                "samples_this_rank": hi - lo, "kernel_seconds": t_k1, "algorithmic_hbm_bytes_per_launch": int(8 * D * (hi - lo)),
                "stages_ms": {"k1": t_k1 * 1e3, "step": ms_total / K}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        os.chdir(cx.scratch)
        rate, cores, t_step, sample = cpu_bbb_rate(mb, wl, seconds=args.cpu_seconds)
        os.chdir(cx.cwd)
        cpu = {"value": rate, "unit": BBB_UNIT, "cores": cores, "kind": "port", "sample": sample}
    return {"metric": BBB_METRIC, "value": value, "unit": BBB_UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (bf16x3 on tcgen05 for the layer GEMMs)", "data": "synthetic",
            "config": {"workload": NAMES[wl], "mc_samples_total": k, "mc_samples_per_gpu": hi - lo, "weights": D, "points_per_sample": int(npts),
                       "l2": "flushed between timed steps (256 MiB write)", "elbo": float(elbo[0]),
                       "collectives_per_step": 0 if world == 1 else f"1 all-reduce of {2 * D + 1} floats (NCCL inside ocbnn_bbb_step)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * K), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall}


# ---------------------------------------------------------------------------------------------------------------------

def run_reference(args):
    """CPU arm: the oracle port on all host cores on this arm's config / metric (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload.upper()
    cwd = os.getcwd()
    from ocbnn_b200 import workloads
    budget = args.ref_seconds or max(20.0, 8.0 * (args.steps + args.warmup))
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            if wl == "W2":
                m = workloads.build("W2", seed=0)
                arr = _spec_pickle(m, "W2")
                rate, cores, M, t_step, how = cpu_hmc_rate(arr, args.leapfrog, seconds=budget, steps=args.steps, warmup=args.warmup)
                metric, unit, cfg = HMC_METRIC, HMC_UNIT, {"workload": NAMES["W2"], "chains": M, "leapfrog_l": args.leapfrog}
                sample = f"{M} chains x 1 HMC iteration (L={args.leapfrog}) per step on {cores} host cores; {how}"
            elif wl in ("W4", "W7"):
                n = args.particles * (args.gpus if wl == "W4" else 1)
                m = workloads.build(wl, seed=0, infer_nsamples=n)
                rate, cores, t_step, sample = cpu_svgd_rate(m, wl, n, seconds=budget, steps=args.steps, warmup=args.warmup)
                metric, unit, cfg = SVGD_METRIC, SVGD_UNIT, {"workload": NAMES[wl], "particles_total": n}
            elif wl == "W6":
                m = workloads.build(wl, seed=0, bbb_esamples=args.particles)
                rate, cores, t_step, sample = cpu_bbb_rate(m, wl, seconds=budget, steps=args.steps, warmup=args.warmup)
                metric, unit, cfg = BBB_METRIC, BBB_UNIT, {"workload": NAMES[wl], "mc_samples_total": args.particles}
            else:
                raise SystemExit(f"unknown workload {wl}")
        finally:
            os.chdir(cwd)
    line = {"impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak" if wl in ("W2", "W4") else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    # libraries write to fd 1 behind Python's back (NCCL prints its version banner there when the engine's own communicator
    # comes up): everything but the JSON line goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    wl = args.workload.upper()
    cx = Ctx(args)
    try:
        if wl == "W2":
            line = bench_hmc(cx)
            svgd = None
            if not args.no_svgd:
                # second half of BASELINE.json's metric, a complete measurement of its own (20 epochs)
                svgd = bench_svgd(cx, "W4", K=20, W=3)
            if cx.rank == 0:
                if cx.world == 1 and not args.no_cpu_baseline:
                    ref = reference_unmodified()
                    if line.get("cpu_baseline") is not None:
                        line["cpu_baseline"]["reference_unmodified"] = _ref_part(ref, "hmc")
                    if svgd is not None and svgd.get("cpu_baseline") is not None:
                        svgd["cpu_baseline"]["reference_unmodified"] = _ref_part(ref, "svgd")
                line["svgd"] = svgd
                if args.extras:
                    line["extras"] = extras(cx)
        elif wl in ("W4", "W7"):
            line = bench_svgd(cx, wl, with_clocks=True)
            if cx.rank == 0 and wl == "W4" and cx.world == 1 and not args.no_cpu_baseline and line.get("cpu_baseline") is not None:
                line["cpu_baseline"]["reference_unmodified"] = _ref_part(reference_unmodified(skip_hmc=True), "svgd")
        elif wl == "W6":
            line = bench_bbb(cx, wl)
        else:
            raise SystemExit(f"unknown workload {wl}")
        if cx.rank == 0:
            emit(line)
    finally:
        cx.close()


def extras(cx):
    """Secondary measurements (single GPU): HMC chain-count sweep, larger SVGD sizes, SGLD, classifier HMC."""
    from ocbnn_b200 import workloads
    args, dev = cx.args, cx.dev
    out = {}
    L = args.leapfrog
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="ocbnn_extras_"))
    try:
        m = workloads.build("W2", seed=0)
        prob = m.engine_problem()
        D = m.nweights
        sweep = {}
        for M in (4096, 8192, 16384, 32768, 262144, 1048576):
            q = 0.5 * torch.randn(M, D, device=dev)
            p0 = torch.randn(1, M, D, device=dev)
            u = torch.rand(1, M, device=dev, dtype=torch.float64)
            t = _time_cuda(lambda: prob.hmc_iterate(q, p0, u, m.hmc_epsilon, L, True, want_flags=False, lanes=0), reps=3)
            sweep[f"M{M}"] = M * L / t
        out["hmc_sweep_chain_leapfrog_steps_per_s"] = sweep
        for n in (16384, 32768):
            ms = workloads.build("W4", seed=0, infer_nsamples=n)
            st = ms.svgd_setup(torch.randn(n, ms.nweights))
            t = _time_cuda(lambda: ms.svgd_epoch(st, 1, True), reps=5, warm=3)
            out[f"svgd_W4_n{n}_particle_iters_per_s"] = n / t
        mg = workloads.build("W1", sampler="SGLD", seed=0)
        pg = mg.engine_problem()
        M, T = 262144, 20
        w = 0.5 * torch.randn(M, D, device=dev)
        noise = 0.01 * torch.randn(T, M, D, device=dev)
        he = np.full(T, 0.003, dtype=np.float32)
        t = _time_cuda(lambda: pg.sgld_iterate(w, noise, he, None, 1, True), reps=2)
        out["sgld_W1_chain_steps_per_s"] = M * T / t
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
    finally:
        os.chdir(cwd)
    return out


if __name__ == "__main__":
    main()
