"""Drop-in alias: `import bnn` with ocbnn-public_b200/ on sys.path gives the reference's package name and classes."""
from ocbnn_b200 import *  # noqa: F401,F403
from ocbnn_b200 import __all__  # noqa: F401
