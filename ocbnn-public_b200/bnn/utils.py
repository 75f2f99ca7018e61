"""Drop-in alias of the reference's `bnn.utils` evaluation helpers (the plotting functions are out of scope)."""
from ocbnn_b200.utils import *  # noqa: F401,F403
from ocbnn_b200.utils import (constraint_violation_fraction, credit_evaluate_recourse, eval_accuracy_and_f1_score, posterior_mean_probs,  # noqa: F401
                              recid_evaluate_high_risk, standardize_compas, standardize_credit)
