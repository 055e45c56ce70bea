"""Inference classes with the reference's names and signatures (reference inference/__init__.py)."""
from .csghmc import cSGHMC
from .csgld import cSGLD
from .hmc import HMC
from .inference_base import _Inference
from .optim_sghmc import optimSGHMC
from .pca_subspace import PCASubspaceSampler
from .sghmc import SGHMC
from .sgld import SGLD
from .projection_model import SubspaceModel
from .subspaces import CovarianceSpace, PCASpace, Subspace
from .swa import SWA
from .swag import SWAG

__all__ = ["_Inference", "optimSGHMC", "SGHMC", "SGLD", "cSGHMC", "cSGLD", "SWA", "SWAG", "HMC", "Subspace",
           "CovarianceSpace", "PCASpace", "SubspaceModel", "PCASubspaceSampler"]
