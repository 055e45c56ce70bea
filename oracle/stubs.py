"""Stub modules that let the unmodified reference package import in the build
container (test harness only; see oracle/__init__.py).

The reference's ``__init__`` chain pulls in packages that are not installed
here and are not on the hot path: hamiltorch (util.py:11, inference/hmc.py:10),
botorch / gpytorch (hyperopt/hyper_optimization.py:5-17) and a private sklearn
symbol removed in modern sklearn (inference/subspaces.py:13).
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # oracle/make_ref.py (travels to the GPU box)


def _default_root():
    env = os.environ.get("URSA_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/URSABench"):
        return "/root/reference"
    return _STAGED


REF_ROOT = _default_root()


class _Anything:
    """Callable, attribute-able placeholder for never-executed symbols."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "URSABench"))


def install(root=None):
    """Register the stubs and put the reference root (``root`` or REF_ROOT) on sys.path."""
    global REF_ROOT
    if root is not None:
        REF_ROOT = root
    os.environ.setdefault("WANDB_MODE", "disabled")
    if "hamiltorch" not in sys.modules:
        ham = _module("hamiltorch")
        ham.util = _module("hamiltorch.util")
    wanted = {
        "botorch": {},
        "botorch.acquisition": {"UpperConfidenceBound": _Anything},
        "botorch.fit": {"fit_gpytorch_model": _Anything},
        "botorch.models": {"SingleTaskGP": _Anything},
        "botorch.optim": {"initializers": _Anything(), "optimize_acqf": _Anything},
        "botorch.utils": {"standardize": _Anything},
        "gpytorch": {},
        "gpytorch.constraints": {},
        "gpytorch.constraints.constraints": {"GreaterThan": _Anything},
        "gpytorch.likelihoods": {},
        "gpytorch.likelihoods.gaussian_likelihood": {"GaussianLikelihood": _Anything},
        "gpytorch.mlls": {"ExactMarginalLogLikelihood": _Anything},
        "gpytorch.priors": {},
        "gpytorch.priors.torch_priors": {"GammaPrior": _Anything},
    }
    for name, attrs in wanted.items():
        if name not in sys.modules:
            _module(name, **attrs)
    if "sklearn.decomposition.pca" not in sys.modules:
        # scikit-learn <= 0.22's private four-argument function, absent from the installed release: the restatement of its
        # published algorithm (oracle/restate.py::assess_dimension_sklearn022) stands in, so the reference's own 'mle'
        # branch (subspaces.py:133-153) runs around it.
        from . import restate as _restate
        _module("sklearn.decomposition.pca", _assess_dimension_=_restate.assess_dimension_sklearn022)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            _module("wandb", log=lambda *a, **k: None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
