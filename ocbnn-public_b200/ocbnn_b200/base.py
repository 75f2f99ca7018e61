"""
Host-side mirror of the reference's model core (bnn/base.py): same class and method names, same YAML-driven
flat config, same constraint-registration API and RNG draw order, so that user code written against the
reference runs unchanged.  Everything numeric on the inference path — forward, log_prior, log_likelihood,
log_posterior, predict — is evaluated by the CUDA engine (engine.py -> libocbnn_b200.so); there is no PyTorch
or CPU implementation of those here, and they raise EngineError when no B200 is present.

Engine-only config keys (ignored by the reference, whose config is a flat dict too, base.py:33-35):
  nchains            number of independent HMC/SGLD chains run at once          (default 1)
  restore_on_reject  textbook Metropolis-Hastings instead of the reference's
                     keep-the-proposal behaviour (inference.py:46,78)            (default false)
  rng                "host": draw momenta/noise/eps from the CPU torch/numpy generators in the reference's
                     order (bit-identical streams) ; "device": draw on the GPU   (default "host")
  lanes              threads cooperating on one chain, 0 = choose                 (default 0)
"""

import logging
import math
import os

import numpy as np
import torch
import yaml
from torch.distributions.normal import Normal
from torch.distributions.uniform import Uniform

from . import engine as E
from .compiler import batch_spec, compile_problem


# knobs debug_mode() shrinks (value while debugging); the first five are put back by switch_off_debug_mode (base.py:207-242)
_DEBUG_KNOBS = {"infer_nsamples": 3, "hmc_nburnin": 5, "bbb_epochs": 100, "svgd_epochs": 10, "sgld_nburnin": 5}


def _tag(bnn, text):
    logging.info(f"[{bnn.uid}] {text}")


class BayesianNeuralNetwork:
    """Bayesian inference over an MLP.  API of bnn/base.py:27-242."""

    def __init__(self, uid="bnn-0", configfile="config.yaml"):
        """Flat YAML config -> attributes; a copy of it goes to history/{uid}.yaml (base.py:28-48)."""
        with open(configfile) as fh:
            settings = yaml.load(fh, Loader=yaml.FullLoader)
        self.uid = uid
        vars(self).update(settings)
        os.makedirs("history", exist_ok=True)
        with open(os.path.join("history", f"{uid}.yaml"), "w") as fh:
            yaml.dump(settings, fh)
        self.all_bayes_samples, self.dconstraints, self.pconstraints = [], {}, {}
        self.aocp_mean = self.aocp_std = None
        self._problem = self._problem_key = None
        _tag(self, f"BNN instantiated. Configs saved as `history/{self.uid}.yaml`.")

    # ---- data ------------------------------------------------------------------------------------------------
    def load(self, **dataset):
        """Attach a dataset dict (X_train, Y_train, [X_test, Y_test], X_train_min/max, dataset_name ...), derive the layer
        shapes and draw the initial weights from N(0, sigma_w) — one draw of nweights values, as base.py:50-70."""
        vars(self).update(dataset)
        self.Xdim, self.Ydim, self.N_train = self.X_train.shape[1], self._Ydim, self.Y_train.shape[0]
        counted = f"{self.N_train} training points"
        if hasattr(self, "Y_test"):
            self.N_test = self.Y_test.shape[0]
            counted += f" and {self.N_test} test points"
        widths = [self.Xdim, *self.architecture, self.Ydim]
        self.layer_sizes = widths
        self.layer_shapes = list(zip(widths, widths[1:]))
        self.nweights = sum(fan_out * (fan_in + 1) for fan_in, fan_out in self.layer_shapes)
        self.weights = Normal(0, self.sigma_w).sample(torch.Size([self.nweights])).requires_grad_(True)
        _tag(self, f"Dataset <{self.dataset_name}> loaded: {counted}.")

    # ---- engine plumbing ----------------------------------------------------------------------------------------
    def _device(self):
        E.require_cuda()
        return torch.device("cuda", torch.cuda.current_device())

    def _fingerprint(self):
        def tid(t):
            return None if t is None else (t.data_ptr(), tuple(t.shape), t._version)
        cr = tuple((k, tid(v)) for k, v in sorted(self.__dict__.items()) if k.startswith("_cr_") and isinstance(v, torch.Tensor))
        cfg = tuple((k, repr(getattr(self, k, None))) for k in (
            "use_ocbnn", "sigma_w", "sigma_noise", "activation", "ocp_nsamples", "cocp_expo_gamma", "cocp_expo_tau",
            "cocp_dirichlet_gamma", "cocp_dirichlet_alpha", "cocp_gaussian_sigma_c", "architecture"))
        cons = tuple((k, len(v)) for k, v in sorted(self.dconstraints.items()))
        return (tid(self.X_train), tid(self.Y_train), tid(self.aocp_mean), tid(self.aocp_std), cr, cfg, cons)

    def engine_problem(self):
        """Device-resident compiled problem for the current config/data/constraints (rebuilt when they change)."""
        key = self._fingerprint()
        if self._problem is None or key != self._problem_key:
            if self._problem is not None:
                self._problem.close()
            self._problem = E.DeviceProblem(compile_problem(self))
            self._problem_key = key
        return self._problem

    def _weights_matrix(self, weights=None):
        w = self.weights if weights is None else weights
        w = torch.as_tensor(w).detach()
        if w.ndimension() == 1:
            w = w.unsqueeze(0)
        return w.to(self._device(), torch.float32).contiguous()

    def _predict_operands(self, bayes_samples, domain):
        W = self._weights_matrix(torch.as_tensor(bayes_samples))
        return W, torch.as_tensor(domain).to(W.device, torch.float32).contiguous()

    # ---- model evaluation (CUDA) -----------------------------------------------------------------------------------
    def forward(self, X, weights=None):
        """Forward pass (base.py:89-104): (P,s0) x (D,)|(M,D) -> (P,sK) | (M,P,sK)."""
        W = self._weights_matrix(weights)
        out = self.engine_problem().forward(W, torch.as_tensor(X).to(W.device, torch.float32).contiguous()).cpu()
        return out.squeeze(dim=0)

    def log_posterior_and_grad(self, weights=None, batch_indices=None, with_data=True):
        """(lp (M,), grad (M,D)) of the log posterior (or log prior) for M weight vectors, on the host."""
        W = self._weights_matrix(weights)
        lp, g = self.engine_problem().logpost_grad(W, batch_spec(batch_indices, self.N_train), with_data)
        return lp.cpu(), g.cpu()

    def log_prior(self):
        """log_prior dispatcher (base.py:106-121); value only (no autograd graph)."""
        lp, _ = self.engine_problem().logpost_grad(self._weights_matrix(), (0, 1), False, want_grad=False)
        return lp.cpu()[0]

    def log_posterior(self, batch_indices=None):
        """log_prior + log_likelihood (base.py:123-125)."""
        lp, _ = self.engine_problem().logpost_grad(self._weights_matrix(), batch_spec(batch_indices, self.N_train), True, want_grad=False)
        return lp.cpu()[0]

    def log_likelihood(self, batch_indices=None):
        """log_likelihood of the task mixin (base.py:252-265, 432-444, 524-542), evaluated on its own by K1 (no prior table, no
        constraint rows) rather than as posterior - prior."""
        lp, _ = self.engine_problem().logpost_grad(self._weights_matrix(), batch_spec(batch_indices, self.N_train), "likelihood", want_grad=False)
        return lp.cpu()[0]

    def isotropic_gaussian_prior(self, weights=None, mean=0, std=None):
        """Isotropic Gaussian prior (base.py:127-133) — host arithmetic, not on the sampler path."""
        if weights is None:
            weights = self.weights
        if std is None:
            std = self.sigma_w
        return Normal(mean, std).log_prob(weights.detach()).sum()

    # ---- constraint point sampling (base.py:135-157) --------------------------------------------------------------------
    def _region_bounds(self, region):
        """[(lower, upper)] per input dimension: the region's bound where it is finite, the training range otherwise."""
        lows, highs = region[0::2], region[1::2]
        return [(float(lows[d] if lows[d] > -np.inf else self.X_train_min[d]),
                 float(highs[d] if highs[d] < np.inf else self.X_train_max[d])) for d in range(self.Xdim)]

    def _sample_from_hypercube(self, region, nsamples, mode="uniform"):
        """(nsamples, Xdim) points of the hypercube `region` = (lo_0, hi_0, lo_1, hi_1, ...).  'uniform' draws one
        Uniform(lo, hi) vector per dimension, in dimension order (the reference's RNG order); 'even' is a linspace."""
        draw = {"uniform": lambda lo, hi: Uniform(lo, hi).sample(torch.Size([nsamples])),
                "even": lambda lo, hi: torch.linspace(lo, hi, nsamples)}.get(mode)
        if draw is None:
            return torch.empty(0)
        return torch.stack([draw(lo, hi) for lo, hi in self._region_bounds(region)], dim=1)

    def _sample_region_blocks(self, constraints):
        """One (ocp_nsamples, Xdim) block per registered constraint, stacked (re-sampled on every add_* call)."""
        total = self.ocp_nsamples * len(constraints)
        xs = torch.zeros(total, self.Xdim)
        index = 0
        for dom, _ in constraints:
            xs[index:index + self.ocp_nsamples, :] = self._sample_from_hypercube(dom, self.ocp_nsamples)
            index += self.ocp_nsamples
        return xs

    # ---- samples / AOCP parameters / config -------------------------------------------------------------------------------
    def load_bayes_samples(self, filename, descriptor="sample"):
        """Append a saved (M, nweights) sample tensor to all_bayes_samples (base.py:159-162)."""
        entry = (torch.load(filename), descriptor)
        self.all_bayes_samples.append(entry)
        _tag(self, f"Posterior Sample #{len(self.all_bayes_samples)}: Loaded from `{filename}`.")

    def _set_aocp(self, params):
        self.aocp_mean = params[:, 0]
        self.aocp_std = torch.logsumexp(torch.stack((params[:, 1], torch.zeros(self.nweights))), 0) / self.aocp_std_multiplier

    def load_gaussian_aocp_parameters(self, filename):
        """Load AOCP parameters (base.py:190-195)."""
        self._set_aocp(torch.load(filename))
        logging.info(f"[{self.uid}] Loaded AOCP parameters from `{filename}`. This replaces any previously loaded or learnt AOCP parameters.")

    def learn_gaussian_aocp(self, verbose=True):
        """Amortized output-constrained prior (base.py:164-188): Adagrad on the AOCP objective."""
        from .aocp import learn_gaussian_aocp
        learn_gaussian_aocp(self, verbose)

    def gaussian_aocp_objective(self):
        """AOCP objective at self._aocp_params (base.py:390-419, 624-672); value only."""
        from .aocp import gaussian_aocp_objective
        return gaussian_aocp_objective(self)[0].cpu()[0]

    def update_config(self, **overrides):
        vars(self).update(overrides)
        _tag(self, "Configs updated.")

    def clear_all_samples(self):
        self.all_bayes_samples = []
        _tag(self, "All posterior samples cleared from memory.")

    def debug_mode(self):
        """A few samples / epochs only, to check a pipeline end to end (base.py:207-222); the previous values are kept
        as old_<name> attributes."""
        for knob, small in _DEBUG_KNOBS.items():
            setattr(self, "old_" + knob, getattr(self, knob))
            setattr(self, knob, small)
        self.aocp_nepochs = 1
        _tag(self, "Debug mode activated.")

    def switch_off_debug_mode(self):
        if not hasattr(self, "old_infer_nsamples"):
            logging.error("You cannot switch off debug mode without having turned it on first!")
            raise Exception
        for knob in _DEBUG_KNOBS:
            setattr(self, knob, vars(self).pop("old_" + knob))
        _tag(self, "Debug mode turned off.")


class RegressorMixin:
    """Regression-specific functionality (base.py:245-419)."""

    _task = "regression"

    @property
    def _Ydim(self):
        return self.Y_train.shape[1]

    def predict(self, bayes_samples, domain):
        """(M,D),(K,Xdim) -> (M,K,Ydim) network outputs (base.py:267-277), one batched kernel instead of a per-sample loop."""
        W, X = self._predict_operands(bayes_samples, domain)
        return self.engine_problem().predict(W, X, E.PREDICT_RAW)[0].cpu()

    def add_deterministic_constraint(self, constrained_domain, interval_func, prior_type):
        """Positive or negative deterministic constraint, 1-D output (base.py:279-342)."""
        if prior_type not in self.dconstraints:
            self.dconstraints[prior_type] = []
        self.dconstraints[prior_type].append((constrained_domain, interval_func))
        S = self.ocp_nsamples
        if prior_type == "positive_gaussian_cocp":
            cons = self.dconstraints["positive_gaussian_cocp"]
            # the reference sizes its buffers from one probe point per constraint (base.py:311); the probes consume
            # RNG draws, so they are made here too.  (The reference's size is wrong for K>1 targets and raises there;
            # this implementation sizes the buffers from the real K.)
            for dom, ifunc in cons:
                ifunc(self._sample_from_hypercube(dom, 1))
            blocks, ylens = [], []
            for dom, ifunc in cons:
                xsamp = self._sample_from_hypercube(dom, S)
                intervals = ifunc(xsamp)
                K = len(intervals[0])
                ylens.append(K)
                ys = torch.stack([torch.tensor([float(sub[i][0]) for sub in intervals]) for i in range(K)], 0)   # (K,S)
                blocks.append((xsamp.repeat(K, 1), ys.reshape(K * S, 1)))
            self._cr_pos_xsamples = torch.cat([b[0] for b in blocks], 0)
            self._cr_pos_ysamples = torch.cat([b[1] for b in blocks], 0).expand(-1, self.Ydim).contiguous()
            self._cr_ylens = ylens
        elif prior_type == "negative_exponential_cocp":
            self._cr_neg_xsamples = self._sample_region_blocks(self.dconstraints["negative_exponential_cocp"])
        elif prior_type == "gaussian_aocp":
            self._cr_aocp_xsamples = self._sample_region_blocks(self.dconstraints["gaussian_aocp"])
        logging.info(f"[{self.uid}] Defined constrained region for `{prior_type}`.")


class ClassifierMixin:
    """Classification-specific functionality (base.py:422-511).  Classes are 0-indexed."""

    _task = "classifier"

    @property
    def _Ydim(self):
        return torch.max(self.Y_train).int().item() + 1

    def predict(self, bayes_samples, domain, return_probs=False):
        """(M,K) predicted classes, or (M,K,Ydim) softmax probabilities (base.py:446-461)."""
        W, X = self._predict_operands(bayes_samples, domain)
        probs, classes, _ = self.engine_problem().predict(W, X, E.PREDICT_PROBS if return_probs else E.PREDICT_CLASS)
        return probs.cpu() if return_probs else classes.cpu().long()

    def add_deterministic_constraint(self, constrained_domain, forbidden_classes, prior_type="positive_dirichlet_cocp"):
        """Forbidden classes over a region (base.py:463-496)."""
        if not forbidden_classes or len(forbidden_classes) >= self.Ydim:
            logging.error("Number of forbidden classes must be between 0 and total number of classes (exclusive).")
            raise Exception
        if prior_type not in self.dconstraints:
            self.dconstraints[prior_type] = []
        self.dconstraints[prior_type].append((constrained_domain, forbidden_classes))
        self._cr_xsamples = self._sample_region_blocks(self.dconstraints[prior_type])
        logging.info(f"[{self.uid}] Defined constrained region for `{prior_type}`.")


class BinaryClassifierMixin:
    """Binary classification with a single sigmoid node (base.py:514-672); needed for the AOCP."""

    _task = "binary"

    @property
    def _Ydim(self):
        return 1

    def _sigmoid(self, batch):
        return torch.sigmoid(batch)

    def predict(self, bayes_samples, domain, return_probs=False):
        """(M,K) p(Y=1) or the decision p >= 0.5 (base.py:544-559)."""
        W, X = self._predict_operands(bayes_samples, domain)
        probs, classes, _ = self.engine_problem().predict(W, X, E.PREDICT_PROBS if return_probs else E.PREDICT_CLASS)
        return probs[:, :, 0].cpu() if return_probs else classes.cpu().bool()

    def add_deterministic_constraint(self, constrained_domain, desired_class, prior_type="gaussian_aocp"):
        """Desired class over a region (base.py:561-590)."""
        if prior_type not in self.dconstraints:
            self.dconstraints[prior_type] = []
        self.dconstraints[prior_type].append((constrained_domain, desired_class))
        self._cr_det_xsamples = self._sample_region_blocks(self.dconstraints[prior_type])
        logging.info(f"[{self.uid}] Defined constrained region for `{prior_type}`.")

    def add_probabilistic_constraint(self, constrained_domain, prob_func, prior_type="gaussian_aocp"):
        """Desired p(Y=1|x) over a region (base.py:592-622)."""
        if prior_type not in self.pconstraints:
            self.pconstraints[prior_type] = []
        self.pconstraints[prior_type].append((constrained_domain, prob_func))
        self._cr_prob_xsamples = self._sample_region_blocks(self.pconstraints[prior_type])
        logging.info(f"[{self.uid}] Defined constrained region for `{prior_type}`.")
