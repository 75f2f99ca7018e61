"""The 12 model classes of the reference (bnn/models.py:10-79): {HMC,BBB,SVGD,SGLD} x {Regressor,Classifier,BinaryClassifier}."""

from .base import BayesianNeuralNetwork, BinaryClassifierMixin, ClassifierMixin, RegressorMixin
from .inference import BBBMixin, HMCMixin, SGLDMixin, SVGDMixin

_TASKS = {"Regressor": RegressorMixin, "Classifier": ClassifierMixin, "BinaryClassifier": BinaryClassifierMixin}
_SAMPLERS = {"HMC": HMCMixin, "BBB": BBBMixin, "SVGD": SVGDMixin, "SGLD": SGLDMixin}

__all__ = []
for _s, _smix in _SAMPLERS.items():
    for _t, _tmix in _TASKS.items():
        _cls = type(f"BNN{_s}{_t}", (BayesianNeuralNetwork, _tmix, _smix), {"__doc__": f"{_s} inference for the {_t} task."})
        _cls.__module__ = __name__
        globals()[_cls.__name__] = _cls
        __all__.append(_cls.__name__)
