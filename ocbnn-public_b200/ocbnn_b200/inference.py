"""
Inference mixins with the interface of the reference (bnn/inference.py:5-10): each provides
`infer(verbose=True, with_data=True)`, which appends `(samples, name)` to `self.all_bayes_samples`
(`samples`: float32 (infer_nsamples, nweights) on the host) and saves `history/{uid}_{sampler}{id}.pt`.

The bodies are drivers only: random numbers are drawn by PyTorch / numpy (host generators in the reference's
draw order by default, so that a seeded run consumes the same streams as the reference), everything else runs
in the CUDA engine:
  HMC   K2  fused trajectory + Metropolis-Hastings        (replaces inference.py:39-78)
  SGLD  K4  fused Langevin steps                           (replaces inference.py:304-340)
  SVGD  K1 + K3a/b/c  gradients, bandwidth, phi, Adagrad  (replaces inference.py:208-281)
  BBB   K5 + K1  reparameterised ELBO gradient + Adagrad   (replaces inference.py:132-186)
Chains / particles / Monte-Carlo samples shard over the ranks of torch.distributed when it is initialised
(one process per GPU); SVGD all-gathers particles and gradients each epoch, BBB all-reduces its (D,2) gradient.
"""

import logging
import math
import time

import numpy as np
import torch
from torch.distributions.normal import Normal

from . import dist as D
from . import engine as E
from .compiler import batch_spec


def _name(self, sampler):
    return f"{sampler}_{'ocbnn' if self.use_ocbnn else 'baseline'}"


class H2DPipe:
    """Host draws -> device, two deep: the copy of the next step's host tensor runs on a copy stream under the kernels of the current
    step.  put() stages a host tensor (through a pinned buffer unless it is pinned already) and returns a slot; take() makes the
    compute stream wait for that slot's copy and hands out the device tensor; done() marks it reusable."""

    def __init__(self, shape, dev, dtype=torch.float32):
        self.dev = dev
        self.stream = torch.cuda.Stream(device=dev)
        self.dbuf = [torch.empty(shape, dtype=dtype, device=dev) for _ in range(2)]
        self.hbuf = [None, None]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [None, None]
        self.i = 0

    def put(self, host):
        i = self.i
        self.i ^= 1
        if not host.is_pinned():
            if self.hbuf[i] is None:
                self.hbuf[i] = torch.empty(host.shape, dtype=host.dtype).pin_memory()
            self.ready[i].synchronize()                   # the previous copy out of this pinned buffer has finished
            self.hbuf[i].copy_(host)
            host = self.hbuf[i]
        if self.consumed[i] is not None:
            self.stream.wait_event(self.consumed[i])      # the step that read this device buffer two steps ago
        with torch.cuda.stream(self.stream):
            self.dbuf[i].copy_(host, non_blocking=True)
            self.ready[i].record(self.stream)
        return i

    def take(self, i):
        torch.cuda.current_stream().wait_event(self.ready[i])
        return self.dbuf[i]

    def done(self, i):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.consumed[i] = ev


def _batch_for(self, t):
    """(offset, stride) of torch.arange(t % nbatches, N_train, nbatches) (inference.py:180,276,334)."""
    nb = int(getattr(self, "nbatches", 0) or 0)
    if nb <= 1:
        return (0, 1)
    return (t % nb, nb)


def _device_generator(self, dev):
    """rng='device': one CUDA generator per model and rank, seeded from rank 0's torch seed + the rank."""
    gen = getattr(self, "_dev_gen", None)
    if gen is None or gen.device != dev:
        seed = torch.tensor([int(torch.initial_seed()) & 0x7FFFFFFFFFFF], dtype=torch.int64)
        D.broadcast_(seed)
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed) + 7919 * D.rank())
        self._dev_gen = gen
    return gen


def _raise_nan(self):
    logging.error(f"[{self.uid}] NaNs encountered in current set of weights.")
    raise Exception


class HMCMixin:
    """Hamiltonian Monte Carlo (inference.py:24-124), M chains at once."""

    def _hmc_chunk(self):
        # iterations per kernel launch: bounded by the momentum buffer (<= ~256 MB)
        M = max(1, int(getattr(self, "nchains", 1) or 1))
        return max(1, min(512, (64 << 20) // max(1, M * self.nweights)))

    def _hmc_draw(self, n_it, M, dev, lo=0, hi=None):
        """Momenta (n_it, hi-lo, D) ~ N(0,1) and uniforms (n_it, hi-lo) float64 for chains [lo, hi) of M.

        rng='host', one chain: the reference's streams exactly - one torch Normal(0,1).sample([D]) and one numpy uniform per
        iteration (inference.py:45,74).  rng='host', M > 1 (no reference stream exists): ONE torch.randn / numpy call for all
        chains of the chunk; every rank draws the full (n_it, M, D) block and keeps its slice, so N ranks seeded alike reproduce
        one rank bit for bit and ranks seeded differently still run independent chains.  rng='device': a generator per rank
        (rank 0's seed + rank), so equal user seeds on all ranks do not couple the chains."""
        Dw = self.nweights
        hi = M if hi is None else hi
        if getattr(self, "rng", "host") == "device":
            gen = _device_generator(self, dev)
            p0 = torch.randn(n_it, hi - lo, Dw, device=dev, generator=gen)
            u = torch.rand(n_it, hi - lo, device=dev, dtype=torch.float64, generator=gen)
            return p0, u
        if M == 1:
            p0 = torch.empty(n_it, 1, Dw)
            u = np.empty((n_it, 1))
            for it in range(n_it):
                p0[it, 0] = Normal(0, 1).sample(torch.Size([Dw]))
                u[it, 0] = np.random.uniform()
            return p0[:, lo:hi].to(dev), torch.from_numpy(u[:, lo:hi]).to(dev)
        p0 = torch.randn(n_it, M, Dw)
        u = np.random.uniform(size=(n_it, M))
        return p0[:, lo:hi].contiguous().to(dev), torch.from_numpy(np.ascontiguousarray(u[:, lo:hi])).to(dev)

    def _hmc_state(self, M, dev):
        """Chain 0 continues from self.weights (like the reference); further chains continue from the previous
        infer() when available, else start from N(0, sigma_w)."""
        prev = getattr(self, "chain_states", None)
        q = torch.empty(M, self.nweights)
        q[0] = self.weights.data
        for c in range(1, M):
            q[c] = prev[c] if prev is not None and prev.shape[0] > c else Normal(0, self.sigma_w).sample(torch.Size([self.nweights]))
        return q.to(dev).contiguous()

    def hmc_run(self, q, n_it, with_data, collect_every=0, chunk=None, samples_host=None, chains=None):
        """n_it iterations from state q (M,D, device; updated in place).  Returns (accepts (M,), samples or None).
        chains = (lo, M_total): q holds chains [lo, lo + M) of M_total (a rank's share; the random streams are defined over
        all M_total chains, see _hmc_draw).

        The iterations run in chunks (one kernel launch each).  The momenta / uniforms of chunk k+1 are drawn and uploaded
        on a side stream while the kernel of chunk k runs, so host RNG and H2D copies overlap the compute.  With
        `samples_host` (pinned CPU tensor (n_it // collect_every, M, D)) the samples of every chunk are copied to the host on
        a third stream while the next chunk runs; the copies are complete when this method returns."""
        prob = self.engine_problem()
        M = q.shape[0]
        c_lo, M_total = (0, M) if chains is None else chains
        acc_total = torch.zeros(M, dtype=torch.int64, device=q.device)
        nsamp = n_it // collect_every if collect_every else 0
        samples = torch.empty(nsamp, M, self.nweights, device=q.device) if nsamp else None
        if M == 0:                                       # a rank without chains (nchains < world size) only joins the collectives
            return acc_total, samples
        chunk = self._hmc_chunk() if chunk is None else int(chunk)
        if collect_every:
            chunk = max(collect_every, chunk - chunk % collect_every)
        main = torch.cuda.current_stream()
        side = getattr(self, "_hmc_side_stream", None)
        if side is None or side.device != q.device:
            side = self._hmc_side_stream = torch.cuda.Stream(device=q.device)

        side.wait_stream(main)                           # everything queued so far (q, the problem upload) is visible to the side stream
        d2h = None
        if samples_host is not None and nsamp:
            d2h = getattr(self, "_hmc_d2h_stream", None)
            if d2h is None or d2h.device != q.device:
                d2h = self._hmc_d2h_stream = torch.cuda.Stream(device=q.device)

        def draw(n):
            with torch.cuda.stream(side):
                p0, u = self._hmc_draw(n, M_total, q.device, c_lo, c_lo + M)
                ev = torch.cuda.Event()
                ev.record(side)
            return p0, u, ev

        nan_seen = torch.zeros((), dtype=torch.int32, device=q.device)
        done = 0
        nxt = draw(min(chunk, n_it)) if n_it > 0 else None
        while done < n_it:
            p0, u, ev = nxt
            n = p0.shape[0]
            main.wait_event(ev)
            p0.record_stream(main)
            u.record_stream(main)
            sl = samples[done // collect_every:(done + n) // collect_every] if collect_every else None
            acc, _, flags = prob.hmc_iterate(q, p0, u, self.hmc_epsilon, self.hmc_l, with_data,
                                             bool(getattr(self, "restore_on_reject", False)), samples=sl, thin=collect_every,
                                             lanes=int(getattr(self, "lanes", 0) or 0))
            if d2h is not None and sl is not None and sl.shape[0]:
                d2h.wait_stream(main)
                with torch.cuda.stream(d2h):             # overlaps the next chunk's kernel
                    samples_host[done // collect_every:(done + n) // collect_every].copy_(sl, non_blocking=True)
            done += n
            if done < n_it:
                nxt = draw(min(chunk, n_it - done))      # overlaps the kernel just launched
            acc_total += acc.sum(dim=0)
            nan_seen = torch.maximum(nan_seen, flags.max())
        if d2h is not None:
            main.wait_stream(d2h)
        if int(nan_seen) != 0:                           # one host sync per run (the reference checks every iteration)
            _raise_nan(self)
        if d2h is not None:
            d2h.synchronize()
        return acc_total, samples

    def infer(self, verbose=True, with_data=True):
        """Burn-in, then collect every hmc_ninterval-th state (inference.py:80-124).  With nchains = M > 1 every
        collection round yields M samples (round-major order); the first infer_nsamples are kept."""
        infer_id = len(self.all_bayes_samples) + 1
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: Beginning HMC inference...")
        if not getattr(self, "restore_on_reject", False):
            # the reference keeps a rejected proposal (inference.py:46,78: original_q aliases the live weights); reproduced by default
            logging.warning(f"[{self.uid}] restore_on_reject is false: rejected HMC proposals are kept, as in the reference. "
                            "The chain is not Metropolis-corrected and the logged acceptance rate is diagnostic only; "
                            "set restore_on_reject: true for textbook HMC.")
        start_time = time.time()
        dev = self._device()
        M_total = max(1, int(getattr(self, "nchains", 1) or 1))
        lo, hi = D.shard(M_total)
        q = self._hmc_state(M_total, dev)[lo:hi].contiguous()
        M = hi - lo

        acc, _ = self.hmc_run(q, self.hmc_nburnin, with_data, chains=(lo, M_total))
        self.accepts = int(D.all_sum(acc.sum()))
        self.rejects = self.hmc_nburnin * M_total - self.accepts
        if self.hmc_nburnin:
            logging.info(f"  All {self.hmc_nburnin} burn-in steps completed. Acceptance rate is "
                         f"{100 * (self.accepts / (self.hmc_nburnin * M_total)):.2f}%.")
        logging.info("  Collecting samples now...")
        rounds = -(-self.infer_nsamples // M_total)
        n_it = rounds * self.hmc_ninterval
        acc, samples = self.hmc_run(q, n_it, with_data, collect_every=self.hmc_ninterval, chains=(lo, M_total))
        self.accepts = int(D.all_sum(acc.sum()))
        self.rejects = n_it * M_total - self.accepts
        end_time = time.time()
        if self.infer_nsamples:
            logging.info(f"  All {self.infer_nsamples} samples collected. Acceptance rate is "
                         f"{100 * (self.accepts / max(1, n_it * M_total)):.2f}%. Time took: {(end_time - start_time):.0f} seconds.")
        samples = D.gather_cat(samples, dim=1) if samples is not None else torch.empty(0, M_total, self.nweights, device=dev)
        samples = samples.reshape(-1, self.nweights)[:self.infer_nsamples].cpu()
        states = D.gather_cat(q, dim=0).cpu()
        self.chain_states = states
        self.weights = states[0].clone().requires_grad_(True)
        if D.rank() == 0:
            torch.save(samples, f"history/{self.uid}_hmc{infer_id}.pt")
        self.all_bayes_samples.append((samples, _name(self, "hmc")))
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: HMC inference completed. Samples saved as `history/{self.uid}_hmc{infer_id}.pt`.")


class SGLDMixin:
    """Stochastic Gradient Langevin Dynamics (inference.py:295-375), M chains at once."""

    def get_stepsize(self, t):
        """epsilon_t = a (b + t)^(-gamma) in Python float64 (inference.py:300-302)."""
        return self.sgld_epa * ((self.sgld_epb + t) ** (-1 * self.sgld_epgamma))

    def _sgld_noise(self, ts, M, dev, lo=0, hi=None):
        """Injected noise (len(ts), hi-lo, D) ~ N(0, eps_t) for chains [lo, hi) of M; same conventions as HMCMixin._hmc_draw
        (one chain: the reference's stream, inference.py:339; several chains: one draw per chunk over all M chains, sliced)."""
        Dw = self.nweights
        hi = M if hi is None else hi
        sd = torch.tensor([self.get_stepsize(t) ** 0.5 for t in ts], dtype=torch.float32)
        if getattr(self, "rng", "host") == "device":
            return torch.randn(len(ts), hi - lo, Dw, device=dev, generator=_device_generator(self, dev)) * sd.to(dev)[:, None, None]
        if M == 1:
            z = torch.empty(len(ts), 1, Dw)
            for i, t in enumerate(ts):
                eps = self.get_stepsize(t)
                z[i, 0] = Normal(torch.zeros(Dw), (eps ** 0.5) * torch.ones(Dw)).sample()     # inference.py:339
            return z[:, lo:hi].to(dev)
        z = torch.randn(len(ts), M, Dw) * sd[:, None, None]
        return z[:, lo:hi].contiguous().to(dev)

    def sgld_run(self, w, t_first, n_steps, with_data, collect_every=0, chains=None):
        prob = self.engine_problem()
        M = w.shape[0]
        c_lo, M_total = (0, M) if chains is None else chains
        nsamp = n_steps // collect_every if collect_every else 0
        samples = torch.empty(nsamp, M, self.nweights, device=w.device) if nsamp else None
        if M == 0:                                       # a rank without chains only joins the collectives of infer()
            return samples
        chunk = max(1, min(1024, (64 << 20) // max(1, M * self.nweights)))
        if collect_every:
            chunk = max(collect_every, chunk - chunk % collect_every)
        nb = int(getattr(self, "nbatches", 0) or 0)
        done = 0
        while done < n_steps:
            n = min(chunk, n_steps - done)
            ts = list(range(t_first + done, t_first + done + n))
            noise = self._sgld_noise(ts, M_total, w.device, c_lo, c_lo + M)
            half = np.array([np.float32(self.get_stepsize(t) / 2) for t in ts], dtype=np.float32)
            offs = np.array([t % nb if nb > 1 else 0 for t in ts], dtype=np.int32)
            sl = samples[done // collect_every:(done + n) // collect_every] if collect_every else None
            prob.sgld_iterate(w, noise, half, offs, nb if nb > 1 else 1, with_data, samples=sl, thin=collect_every,
                              lanes=int(getattr(self, "lanes", 0) or 0))
            done += n
            if bool(torch.isnan(w).any()):
                _raise_nan(self)
        return samples

    def infer(self, verbose=True, with_data=True):
        infer_id = len(self.all_bayes_samples) + 1
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: Beginning SGLD inference...")
        start_time = time.time()
        dev = self._device()
        M_total = max(1, int(getattr(self, "nchains", 1) or 1))
        lo, hi = D.shard(M_total)
        w = HMCMixin._hmc_state(self, M_total, dev)[lo:hi].contiguous()
        self.sgld_run(w, 1, self.sgld_nburnin, with_data, chains=(lo, M_total))
        if self.sgld_nburnin:
            logging.info(f"  All {self.sgld_nburnin} burn-in steps completed.")
        logging.info("  Collecting samples now...")
        rounds = -(-self.infer_nsamples // M_total)
        samples = self.sgld_run(w, self.sgld_nburnin + 1, rounds * self.sgld_ninterval, with_data, collect_every=self.sgld_ninterval,
                                chains=(lo, M_total))
        end_time = time.time()
        if self.infer_nsamples:
            logging.info(f"  All {self.infer_nsamples} samples collected. Time took: {(end_time - start_time):.0f} seconds.")
        samples = D.gather_cat(samples, dim=1).reshape(-1, self.nweights)[:self.infer_nsamples].cpu()
        states = D.gather_cat(w, dim=0).cpu()
        self.chain_states = states
        self.weights = states[0].clone().requires_grad_(True)
        if D.rank() == 0:
            torch.save(samples, f"history/{self.uid}_sgld{infer_id}.pt")
        self.all_bayes_samples.append((samples, _name(self, "sgld")))
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: SGLD inference completed. Samples saved as `history/{self.uid}_sgld{infer_id}.pt`.")


class SVGDMixin:
    """Stein Variational Gradient Descent (inference.py:203-292); particles shard over ranks."""

    def svgd_setup(self, particles, use_tensor=-1):
        """particles: (n, D) host or device tensor, the same on every rank (rank 0's copy is broadcast); every rank keeps all n
        particles on its device and owns rows [lo, hi).  use_tensor: 0 exact SIMT kernels, 1 tcgen05 path, -1 choose (tensor path
        from 1024 particles up)."""
        dev = self._device()
        n = particles.shape[0]
        lo, hi = D.shard(n)
        tensor = E.svgd_use_tensor(int(use_tensor), n)
        ws = E.svgd_workspace(n, self.nweights, dev, n_rows=hi - lo, use_tensor=1 if tensor else 0)
        Xall = particles.to(dev, torch.float32).contiguous().clone()
        D.broadcast_(Xall)
        Gall = torch.empty_like(Xall)
        return {"n": n, "lo": lo, "hi": hi, "tensor": tensor, "ws": ws, "prob": self.engine_problem(), "comm": D.comm(),
                "Xall": Xall, "Gall": Gall, "X": Xall[lo:hi], "G": Gall[lo:hi], "phi": torch.empty(hi - lo, self.nweights, device=dev),
                "acc": torch.zeros(hi - lo, self.nweights, device=dev), "h": torch.empty(1, device=dev)}

    def svgd_epoch(self, st, epoch, with_data, use_tensor=None):
        """One SVGD epoch = ONE library call (`ocbnn_svgd_step`: gradients K1, all-gather, bandwidth K3a, phi K3b, Adagrad K3c,
        all-gather), on any number of ranks.

        The ~16 launches (and, across ranks, the NCCL collectives the library issues on the same stream) are captured once per
        (minibatch offset, with_data, path) into a CUDA graph and replayed - the epoch is launch-gap bound at a few thousand
        particles.  `svgd_graph: false` in the config or a failed capture use the eager path."""
        if not st.get("graph_ok", True) or not bool(getattr(self, "svgd_graph", True)):
            return self._svgd_epoch_eager(st, epoch, with_data, use_tensor)
        key = (_batch_for(self, epoch), bool(with_data), use_tensor)
        graphs = st.setdefault("graphs", {})
        ent = graphs.get(key)
        if ent is None:                         # first use of this key: eager (also warms per-kernel attributes and the side stream)
            graphs[key] = "warm"
            return self._svgd_epoch_eager(st, epoch, with_data, use_tensor)
        if ent == "warm":
            try:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._svgd_epoch_eager(st, epoch, with_data, use_tensor)
                graphs[key] = ent = g
            except Exception as exc:            # capture not possible here: stay eager from now on
                if D.world() > 1:               # the ranks must agree on the path; a one-sided fallback would deadlock the collectives
                    raise
                logging.warning(f"[{self.uid}] SVGD epoch graph capture failed ({exc}); running eagerly.")
                st["graph_ok"] = False
                graphs.clear()
                torch.cuda.synchronize()
                return self._svgd_epoch_eager(st, epoch, with_data, use_tensor)
        ent.replay()
        return st

    def _svgd_epoch_eager(self, st, epoch, with_data, use_tensor=None):
        tensor = st["tensor"] if use_tensor is None else E.svgd_use_tensor(int(use_tensor), st["n"]) and st["tensor"]
        E.svgd_step(st["prob"], st["Xall"], st["Gall"], st["acc"], st["phi"], st["h"], st["lo"], st["hi"] - st["lo"], self.svgd_init_lr,
                    _batch_for(self, epoch), with_data, 1 if tensor else 0, st["ws"], st["comm"])
        return st

    def infer(self, verbose=True, with_data=True):
        infer_id = len(self.all_bayes_samples) + 1
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: Beginning SVGD inference...")
        start_time = time.time()
        self.particles = Normal(0, self.sigma_w).sample(torch.Size([self.infer_nsamples, self.nweights]))
        st = self.svgd_setup(self.particles, int(getattr(self, "svgd_tensor", -1)))
        for epoch in range(1, self.svgd_epochs + 1):
            self.svgd_epoch(st, epoch, with_data)
            if verbose and epoch % 10 == 0:
                logging.info(f"  Epoch {epoch} reached.")
        end_time = time.time()
        logging.info(f"  {self.svgd_epochs} epochs completed. Time took: {(end_time - start_time):.0f} seconds.")
        self.particles = st["Xall"].cpu()             # every rank holds all particles after the step's closing all-gather
        if bool(torch.isnan(self.particles).any()):
            _raise_nan(self)
        if D.rank() == 0:
            torch.save(self.particles, f"history/{self.uid}_svgd{infer_id}.pt")
        self.all_bayes_samples.append((self.particles.clone(), _name(self, "svgd")))
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: SVGD inference completed. Samples saved as `history/{self.uid}_svgd{infer_id}.pt`.")


class BBBMixin:
    """Bayes by Backprop (inference.py:127-200); Monte-Carlo samples shard over ranks."""

    def sample_q(self, k, q_params=None, eps=None):
        """k weight samples from the variational Gaussian and their log q (inference.py:132-140), on the host."""
        qp = self.q_params if q_params is None else q_params
        mean, std = qp[:, 0], torch.logsumexp(torch.stack((qp[:, 1], torch.zeros(self.nweights))), 0)
        eps = torch.randn(k, self.nweights) if eps is None else eps
        weights = mean + eps * std
        return weights, Normal(mean, std).log_prob(weights)

    def bbb_epoch(self, qp, acc, eps, epoch, with_data, k_total=None, state=None):
        """One ELBO-gradient step = ONE library call (`ocbnn_bbb_step`: sample, K1, closed-form (D,2) gradient, one all-reduce of
        gradient + ELBO, Adagrad).  eps: this rank's (k_local, D) standard normal draws.  state: dict reused across epochs
        (workspace, communicator, ELBO scalar); returns the ELBO as a 1-element device tensor."""
        state = {} if state is None else state
        k_total = eps.shape[0] * D.world() if k_total is None else k_total
        if state.get("k") != eps.shape[0]:
            state.update(k=eps.shape[0], ws=E.bbb_workspace(eps.shape[0], self.nweights, qp.device), comm=D.comm(), prob=self.engine_problem())
        elbo = torch.empty(1, device=qp.device)
        E.bbb_step(state["prob"], qp, acc, eps, k_total, self.bbb_init_lr, _batch_for(self, epoch), with_data, elbo, state["ws"], state["comm"])
        return elbo

    def infer(self, verbose=True, with_data=True):
        infer_id = len(self.all_bayes_samples) + 1
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: Beginning BBB inference...")
        start_time = time.time()
        dev = self._device()
        init_means = self.bbb_init_mean + torch.randn(self.nweights, 1)
        init_log_stds = self.bbb_init_std * torch.ones(self.nweights, 1)
        qp = torch.cat([init_means, init_log_stds], dim=1).to(dev).contiguous()
        D.broadcast_(qp)                        # replicated state: rank 0's draw, whatever the other ranks' seeds are
        acc = torch.zeros_like(qp)
        k = int(self.bbb_esamples)
        lo, hi = D.shard(k)
        device_rng = getattr(self, "rng", "host") == "device"
        gen = None
        if device_rng:                          # distinct Monte-Carlo draws on every rank even when the user seeded all ranks alike
            gen = torch.Generator(device=dev)
            gen.manual_seed(int(torch.initial_seed()) + 7919 * (D.rank() + 1))
        elbos, state = [], {}

        def host_draw():                        # inference.py:138; rank 0's stream, so that N ranks reproduce one rank
            eps_all = torch.randn(k, self.nweights)
            D.broadcast_(eps_all)
            return eps_all[lo:hi]

        # host draws go to the device two deep (H2DPipe): the draw and the copy of epoch e + 1 run under the kernels of epoch e
        pipe = None if device_rng or self.bbb_epochs < 1 else H2DPipe((hi - lo, self.nweights), dev)
        slot = pipe.put(host_draw()) if pipe else None
        for epoch in range(1, self.bbb_epochs + 1):
            if device_rng:
                eps = torch.randn(hi - lo, self.nweights, device=dev, generator=gen)
            else:
                cur, slot = slot, (pipe.put(host_draw()) if epoch < self.bbb_epochs else None)
                eps = pipe.take(cur)
            elbos.append(self.bbb_epoch(qp, acc, eps, epoch, with_data, k_total=k, state=state))
            if pipe:
                pipe.done(cur)
            if verbose and epoch % 100 == 0:
                logging.info(f"  Epoch {epoch}: {float(elbos[-1]):.2f}")
        self.elbos = [float(e) for e in torch.cat(elbos).cpu()] if elbos else []
        end_time = time.time()
        if self.elbos:
            logging.info(f"  {self.bbb_epochs} epochs completed. Final ELBO is {self.elbos[-1]:.2f}. Time took: {(end_time - start_time):.0f} seconds.")
        self.q_params = qp.cpu()
        samples, _ = self.sample_q(self.infer_nsamples)
        if D.rank() == 0:
            torch.save(self.q_params, f"history/{self.uid}_bbb{infer_id}_qparams.pt")
            torch.save(samples, f"history/{self.uid}_bbb{infer_id}.pt")
        self.all_bayes_samples.append((samples, _name(self, "bbb")))
        logging.info(f"[{self.uid}] Posterior Sample #{infer_id}: BBB inference completed.")
        logging.info(f"[{self.uid}] Variational parameters saved as `history/{self.uid}_bbb{infer_id}_qparams.pt`. Samples saved as `history/{self.uid}_bbb{infer_id}.pt`.")
