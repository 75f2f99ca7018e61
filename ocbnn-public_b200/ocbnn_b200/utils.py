"""
Evaluation on top of the posterior samples: the reference's `bnn/utils.py:217-232`
(`eval_accuracy_and_f1_score`), `run_apps.py:31-41` (`recid_evaluate_high_risk`), `run_apps.py:130-143`
(`credit_evaluate_recourse`), the constraint-violation fraction of `run_toys.py:229-250` (Figure 3c) and the feature standardisation of `data/dataloader.py:134-142, 167-176`.

Same names, arguments and return values as the reference.  The posterior-mean predictive probability
`predict(samples, X, return_probs=True).mean(dim=0)` — the only heavy part — is formed on the device: one batched
forward kernel (`ocbnn_forward`) over all samples and points, the softmax / sigmoid and the mean over samples stay on
the GPU and only the (K, classes) means come back.  No CPU fallback: without the CUDA engine these raise EngineError.
The plotting helpers of `bnn/utils.py` are out of scope.
"""

import numpy as np
import torch

from . import engine as E


def posterior_mean_probs(bnn, samples, X, chunk=1 << 26):
    """(K, Ydim) for classifiers, (K,) for binary classifiers: mean over the samples of the predictive probabilities
    (what every evaluation of the reference starts from).  Points are processed in chunks that keep the (M, chunk, Ydim)
    logits under `chunk` floats."""
    E.require_cuda()
    W = bnn._weights_matrix(torch.as_tensor(samples))
    X = torch.as_tensor(X).to(W.device, torch.float32).contiguous()
    prob = bnn.engine_problem()
    binary = getattr(bnn, "_task", "") == "binary"
    per = max(1, chunk // max(1, W.shape[0] * prob.sK))
    outs = []
    for k0 in range(0, X.shape[0], per):
        # forward + softmax / sigmoid + mean over the samples in one library call (ocbnn_predict)
        mean = prob.predict(W, X[k0:k0 + per].contiguous(), E.PREDICT_PROBS, want_mean=True)[2]
        outs.append(mean[:, 0] if binary else mean)
    return torch.cat(outs, 0).cpu()


def constraint_violation_fraction(bnn, samples, domain, forbidden):
    """Fraction of posterior samples whose prediction enters the open interval `forbidden` = (lo, hi) at any point of
    `domain` (K, Xdim): the "fraction of rejected posterior samples" of Figure 3c (run_toys.py:229-233, 245-249, where the
    interval is (1.0, 2.5) and the domain arange(-1, 1, 0.05)).  Forward pass and the per-sample any() run in the library
    (ocbnn_violation_flags); only the M flags come back."""
    E.require_cuda()
    W = bnn._weights_matrix(torch.as_tensor(samples))
    X = torch.as_tensor(domain).to(W.device, torch.float32).contiguous()
    flags = bnn.engine_problem().violation_flags(W, X, forbidden[0], forbidden[1])
    return int(flags.sum()) / max(1, W.shape[0])


def _binary_f1(y_true, y_pred):
    """F1 of the positive class, zero when undefined (sklearn precision_recall_fscore_support(average='binary'))."""
    y_true, y_pred = torch.as_tensor(y_true).bool(), torch.as_tensor(y_pred).bool()
    tp = int((y_true & y_pred).sum())
    fp = int((~y_true & y_pred).sum())
    fn = int((y_true & ~y_pred).sum())
    prec = tp / (tp + fp) if tp + fp else 0.0
    rec = tp / (tp + fn) if tp + fn else 0.0
    return 2 * prec * rec / (prec + rec) if prec + rec else 0.0


def eval_accuracy_and_f1_score(bnn, is_binary=False, X_eval=None, Y_eval=None):
    """Accuracy and F1 on the test set from the most recent posterior sample (bnn/utils.py:217-232)."""
    if X_eval is None:
        X_eval, Y_eval = bnn.X_test, bnn.Y_test
    samples, _ = bnn.all_bayes_samples[-1]
    probs = posterior_mean_probs(bnn, samples, X_eval)
    preds = (probs >= 0.5) if is_binary else probs.argmax(dim=1)
    Y = torch.as_tensor(Y_eval)
    acc = float((preds.long() == Y.long()).float().mean())
    return acc, _binary_f1(Y == 1, preds == 1)


def recid_evaluate_high_risk(bnn):
    """High-risk counts of both groups on the training set (run_apps.py:31-41): (aa_highrisk, aa_total, nonaa_highrisk,
    nonaa_total); bnn.X_train_race holds the group indicator."""
    samples, _ = bnn.all_bayes_samples[-1]
    preds = posterior_mean_probs(bnn, samples, bnn.X_train) >= 0.5
    race = torch.as_tensor(bnn.X_train_race)
    aa, non = race == 1.0, race == 0.0
    return int((aa & preds).sum()), int(aa.sum()), int((non & preds).sum()), int(non.sum())


def credit_evaluate_recourse(bnn):
    """Mean revolving utilisation (feature 0) of young adults (age < 35, feature 1 de-standardised) by ground-truth and by
    predicted class (run_apps.py:130-143): (gt_neg, gt_pos, pred_neg, pred_pos)."""
    samples, _ = bnn.all_bayes_samples[-1]
    preds = posterior_mean_probs(bnn, samples, bnn.X_test).argmax(dim=1)
    X, Y = torch.as_tensor(bnn.X_test), torch.as_tensor(bnn.Y_test)
    young = (X[:, 1] * bnn.X_train_std[1] + bnn.X_train_mean[1] < 35).nonzero().squeeze(dim=1)
    Xy, Yy, Py = X[young], Y[young], preds[young]
    return (Xy[(Yy == 0).nonzero(), 0].mean(), Xy[(Yy == 1).nonzero(), 0].mean(),
            Xy[(Py == 0).nonzero(), 0].mean(), Xy[(Py == 1).nonzero(), 0].mean())


def standardize_compas(X_train):
    """Per-feature mean / (unbiased) std of the training matrix and in-place standardisation of columns 0, 2, 3, then
    the per-feature min / max (data/dataloader.py:134-148).  Returns the dict entries the reference adds."""
    X_train = torch.as_tensor(X_train).float()
    mean = [X_train[:, i].mean().item() for i in range(X_train.shape[1])]
    std = [X_train[:, i].std().item() for i in range(X_train.shape[1])]
    for j in (0, 2, 3):
        X_train[:, j] = (X_train[:, j] - mean[j]) / std[j]
    return {"X_train": X_train, "X_train_mean": mean, "X_train_std": std,
            "X_train_min": [X_train[:, i].min().item() for i in range(X_train.shape[1])],
            "X_train_max": [X_train[:, i].max().item() for i in range(X_train.shape[1])]}


def standardize_credit(X_train, X_test):
    """Population mean / std (numpy) of the training features, standardisation of columns 1, 4, 5 of both splits
    (data/dataloader.py:167-176; the reference computes the statistics before upsampling)."""
    X_train, X_test = np.array(X_train, dtype=np.float64), np.array(X_test, dtype=np.float64)
    mean = [X_train[:, i].mean() for i in range(X_train.shape[1])]
    std = [X_train[:, i].std() for i in range(X_train.shape[1])]
    for j in (1, 4, 5):
        X_test[:, j] = (X_test[:, j] - mean[j]) / std[j]
        X_train[:, j] = (X_train[:, j] - mean[j]) / std[j]
    return {"X_train": torch.tensor(X_train).float(), "X_test": torch.tensor(X_test).float(), "X_train_mean": mean, "X_train_std": std}
