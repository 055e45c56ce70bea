"""The drop-in boundary (include/ursa_b200.h <-> libursa_b200.so <-> ursabench_b200/_C.py), checked without a GPU:
every function the header declares is exported by the library and bound by the ctypes table with the declared number of
arguments; there is no CPU path to fall back to (constructors reject non-CUDA devices, a missing library is an error)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ursa_b200.h")


def _declared():
    """{name: n_args} for every `ursa_*(...)` prototype in the header (comments stripped)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(ursa_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_header_declares_the_path():
    d = _declared()
    for name in ("ursa_sgmcmc_step", "ursa_swag_collect", "ursa_swag_variance", "ursa_swag_draw", "ursa_swag_gram",
                 "ursa_bma_accumulate", "ursa_bma_metrics", "ursa_bma_mlp_forward", "ursa_bma_preresnet_forward",
                 "ursa_bma_wrn_forward", "ursa_wrn_bn_update", "ursa_hmc_leapfrog", "ursa_last_error", "ursa_abi_version"):
        assert name in d, name


def test_library_exports_every_declared_symbol():
    from ursabench_b200 import _C
    handle = ctypes.CDLL(_C.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(handle, n)]
    assert not missing, missing


def test_ctypes_table_matches_the_header():
    from ursabench_b200 import _C
    decl = _declared()
    bound = dict(_C.SIGNATURES)
    unbound = sorted(set(decl) - set(bound) - {"ursa_abi_version", "ursa_last_error", "ursa_device_info"})
    assert not unbound, "declared in the header but not bound in _C.SIGNATURES: %s" % unbound
    for name, (_, argtypes) in bound.items():
        assert name in decl, "%s is bound but not declared in include/ursa_b200.h" % name
        assert len(argtypes) == decl[name], (name, len(argtypes), decl[name])


def test_missing_library_is_an_error_not_a_fallback():
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['URSA_B200_LIB'] = '/nonexistent/libursa_b200.so';"
            "from ursabench_b200 import _C\n"
            "try:\n    _C.lib()\nexcept _C.UrsaError as e:\n    print('UrsaError')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "UrsaError" in out.stdout, out.stdout + out.stderr


def test_no_cpu_path():
    from ursabench_b200 import inference, tasks
    from ursabench_b200.flat import FlatParams
    from ursabench_b200.models import MLP
    m = MLP(8, 4, 3)
    with pytest.raises(ValueError):
        FlatParams.from_model(m, torch.device("cpu"))
    x, y = torch.randn(6, 4), torch.randint(0, 3, (6,))
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(x, y), batch_size=3)
    with pytest.raises((RuntimeError, ValueError)):
        tasks.Prediction({"in_distribution_test": loader}, 3, torch.device("cpu"), "ALL")
    with pytest.raises((RuntimeError, ValueError)):
        inference.SGHMC(None, model=m, train_loader=loader, device=torch.device("cpu"))


def test_argument_errors_are_reported_without_touching_the_gpu():
    """Error behaviour of the C ABI (host-side checks only: no kernel is launched, so this runs without a GPU): a negative return
    code plus a message from ursa_last_error(), never a crash or a silent success."""
    from ursabench_b200 import _C
    lib = _C.lib()

    def err():
        return lib.ursa_last_error().decode()

    assert lib.ursa_swag_gram(None, 0, 3, 10, None, None) < 0 and "ursa_swag_gram" in err()
    assert lib.ursa_bma_wrn_forward(None, 0, None, 0, 1, None, 1, 28, 10, 100, None, None, None, 1e-4, None, 0,
                                    _C.ALGO_TCGEN05, None) < 0 and "null pointer" in err()
    assert lib.ursa_wrn_bn_update(None, None, None, 10, 128, 28, 10, 100, None, 0, None) < 0 and "null pointer" in err()
    assert lib.ursa_gemm_nt_3xtf32(None, 0, 0, None, 0, 0, None, 0, 0, None, 0, 0, 1, 1, 1, 1, None, 0, None) < 0
    assert "ursa_gemm_nt_3xtf32" in err()
    # workspace queries answer 0 for shapes an engine does not cover (the Python layer then picks another engine)
    assert lib.ursa_bma_wrn_workspace(1, 8, 28, 10, 100, _C.ALGO_FFMA) == 0
    assert lib.ursa_wrn_bn_update_workspace(100, 127, 28, 10, 100) == 0          # odd batch
    assert lib.ursa_wrn_bn_update_workspace(100, 1024, 28, 10, 100) == 0         # batch larger than a chunk
    assert lib.ursa_gemm_nt_3xtf32_workspace(0, 1, 1, 1, 0) == 0
    assert lib.ursa_gemm_nt_3xtf32_workspace(2, 100, 10, 50, 1) > 0
