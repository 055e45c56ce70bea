"""ursabench_b200 -- B200-native engine for URSABench's SG-MCMC / SWAG / BMA hot path.

Drop-in surface (same names as the reference package):
    from ursabench_b200 import inference, tasks, models, util
    inference.SGLD / SGHMC / cSGLD / cSGHMC / SWA / SWAG, inference.optimSGHMC
    tasks.Prediction
The CUDA kernels live in ``libursa_b200.so`` (C ABI: include/ursa_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import inference, models, tasks, util          # noqa: E402,F401
from .util import set_random_seed                      # noqa: E402,F401
