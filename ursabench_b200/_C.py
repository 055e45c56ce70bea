"""ctypes binding of ``libursa_b200.so`` (C ABI in ``include/ursa_b200.h``).

There is NO fallback: if the library is missing or a call fails, this module
raises.  Device pointers come from ``tensor.data_ptr()`` and the stream from
``torch.cuda.current_stream()``; torch is used here only as the owner of
device memory and streams.
"""
import ctypes
import functools
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("URSA_B200_LIB") or os.path.join(_HERE, "libursa_b200.so")   # env override: debug builds

STEP_FIRST, STEP_NOISE, STEP_ZERO_GRAD = 1, 2, 4
ALGO_FFMA, ALGO_TCGEN05, ALGO_TCGEN05_FUSED, ALGO_TCGEN05_FUSED_F16, ALGO_TCGEN05_F16 = 0, 1, 2, 3, 4
ALGO_FLAG_WS_KEPT = 0x100      # ursa_bma_preresnet_forward: workspace untouched since this caller's previous identical-shape call
DRAW_MAX_S, DRAW_MAX_K = 30, 24

_c = ctypes
_vp, _i64, _i32, _u32, _u64, _f32, _f64, _sz = (_c.c_void_p, _c.c_int64, _c.c_int, _c.c_uint32, _c.c_uint64,
                                                _c.c_float, _c.c_double, _c.c_size_t)

# name -> (restype, argtypes); mirrors include/ursa_b200.h one to one
SIGNATURES = {
    "ursa_abi_version": (_i32, []),
    "ursa_last_error": (_c.c_char_p, []),
    "ursa_launch_count": (_u64, []),
    "ursa_profile_begin": (_i32, [_i32]),
    "ursa_profile_end": (_i32, [_c.POINTER(_f64), _c.POINTER(_i64), _i32]),
    "ursa_device_info": (_i32, [_c.POINTER(_i32)] * 3),
    "ursa_sgmcmc_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _u32, _u64, _u64, _u64, _vp]),
    "ursa_sgmcmc_step_dyn": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _u32, _u64, _u64, _vp]),
    "ursa_sgmcmc_set_dyn": (_i32, [_vp, _f32, _f32, _f32, _f32, _u64, _vp]),
    "ursa_philox_normal": (_i32, [_vp, _i64, _u64, _u64, _u64, _vp]),
    "ursa_swag_collect": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp]),
    "ursa_swag_variance": (_i32, [_vp, _vp, _vp, _i64, _f32, _vp]),
    "ursa_swag_draw": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _i64, _f32, _u64, _u64, _vp]),
    "ursa_bma_accumulate": (_i32, [_vp, _i64, _i64, _i32, _i64, _vp, _vp, _f64, _vp]),
    "ursa_bma_metrics_workspace": (_sz, [_i64, _i32]),
    "ursa_bma_metrics": (_i32, [_vp, _i64, _i32, _f32, _vp, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ursa_bma_mlp_workspace": (_sz, [_i32, _i64, _i32, _i32, _i32, _i32]),
    "ursa_bma_mlp_forward": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _f64, _vp, _sz, _i32, _vp]),
    "ursa_bma_preresnet_workspace": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "ursa_bma_preresnet_forward": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _f64, _vp,
                                          _sz, _i32, _vp]),
    "ursa_gemm_nt_3xtf32_workspace": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "ursa_gemm_nt_3xtf32": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i32, _vp, _i64, _i64, _i32, _i64, _i32, _i32,
                                   _vp, _sz, _vp]),
    "ursa_hmc_mlp_grad_workspace": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "ursa_hmc_mlp_grad": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ursa_hmc_mlp_grad_f16_workspace": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "ursa_hmc_mlp_grad_f16": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ursa_bma_wrn_workspace": (_sz, [_i32, _i64, _i32, _i32, _i32, _i32]),
    "ursa_bma_wrn_forward": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _f64, _vp,
                                    _sz, _i32, _vp]),
    "ursa_swag_gram": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp]),
    "ursa_wrn_bn_update_workspace": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "ursa_wrn_bn_update": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    "ursa_wrn_bn_update_algo": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _sz, _i32, _vp]),
    "ursa_preresnet_bn_update_workspace": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "ursa_preresnet_bn_update": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _sz, _vp]),
    "ursa_hmc_momentum": (_i32, [_vp, _vp, _i64, _f32, _u64, _u64, _u64, _vp]),
    "ursa_hmc_leapfrog": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp]),
    "ursa_hmc_energy_workspace": (_sz, [_i64, _i64]),
    "ursa_hmc_energy": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "ursa_hmc_accept": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _u64, _u64, _u64, _vp]),
}

_lib = None


class UrsaError(RuntimeError):
    pass


def lib():
    """Load the shared library once; raise if it is absent (no CPU / eager fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UrsaError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C ursabench_b200/csrc`; there is no fallback path" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if handle.ursa_abi_version() != 1:
            raise UrsaError("libursa_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def _check(rc, what):
    if rc != 0:
        msg = lib().ursa_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError("%s: %s" % (what, msg))
        raise UrsaError("%s failed (%d): %s" % (what, rc, msg))


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _on_device(fn):
    """Run ``fn`` with the CUDA device of its first tensor argument current: the kernels launch on the *current*
    device (cudaGetDevice), the reference's classes accept any ``device=`` without a prior ``set_device``."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapped


def _dev_f32(t, name, allow_none=False):
    if t is None:
        if allow_none:
            return
        raise ValueError("%s is required" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor: this engine has no CPU path" % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError("%s must be contiguous float32" % name)


@_on_device
def sgmcmc_step(p, g, v=None, snapshot=None, noise=None, *, lr, momentum, wd_over_n, noise_mul=0.0, noise_div=1.0,
                first_step=False, add_noise=True, zero_grad=False, seed=0, step=0, elem_offset=0):
    """Fused optimSGHMC update over flat buffers (see ursa_sgmcmc_step in the header)."""
    for t, nm, opt in ((p, "p", False), (g, "g", False), (v, "v", True), (snapshot, "snapshot", True), (noise, "noise", True)):
        _dev_f32(t, nm, opt)
    n = p.numel()
    for t, nm in ((g, "g"), (v, "v"), (snapshot, "snapshot"), (noise, "noise")):
        if t is not None and t.numel() < n:
            raise ValueError("%s has fewer elements than p" % nm)
    flags = (STEP_FIRST if first_step else 0) | (STEP_NOISE if add_noise else 0) | (STEP_ZERO_GRAD if zero_grad else 0)
    rc = lib().ursa_sgmcmc_step(_ptr(p), _ptr(g), _ptr(v), _ptr(snapshot), _ptr(noise), n, float(lr), float(momentum),
                                float(wd_over_n), float(noise_mul), float(noise_div), flags, int(seed), int(step),
                                int(elem_offset), _stream(p))
    _check(rc, "ursa_sgmcmc_step")


@_on_device
def sgmcmc_step_dyn(p, g, v, snapshot, noise, dyn, *, first_step=False, add_noise=True, zero_grad=False, seed=0,
                    elem_offset=0):
    """K1 with [lr, momentum, wd_over_n, noise_scale, step_lo, step_hi] read from the device tensor ``dyn``
    (graph-capturable: replays pick up new scalars)."""
    for t, nm, opt in ((p, "p", False), (g, "g", False), (v, "v", True), (snapshot, "snapshot", True),
                       (noise, "noise", True), (dyn, "dyn", False)):
        _dev_f32(t, nm, opt)
    flags = (STEP_FIRST if first_step else 0) | (STEP_NOISE if add_noise else 0) | (STEP_ZERO_GRAD if zero_grad else 0)
    rc = lib().ursa_sgmcmc_step_dyn(_ptr(p), _ptr(g), _ptr(v), _ptr(snapshot), _ptr(noise), p.numel(), _ptr(dyn),
                                    flags, int(seed), int(elem_offset), _stream(p))
    _check(rc, "ursa_sgmcmc_step_dyn")


@_on_device
def sgmcmc_set_dyn(dyn, lr, momentum, wd_over_n, noise_scale, step):
    _dev_f32(dyn, "dyn")
    if dyn.numel() < 6:
        raise ValueError("dyn needs 6 words")
    _check(lib().ursa_sgmcmc_set_dyn(_ptr(dyn), float(lr), float(momentum), float(wd_over_n), float(noise_scale),
                                     int(step), _stream(dyn)), "ursa_sgmcmc_set_dyn")


@_on_device
def philox_normal(out, seed, step, elem_offset=0):
    _dev_f32(out, "out")
    _check(lib().ursa_philox_normal(_ptr(out), out.numel(), seed, step, elem_offset, _stream(out)), "ursa_philox_normal")
    return out


@_on_device
def swag_collect(w, mean, sq_mean, dev_row, n_collected):
    for t, nm in ((w, "w"), (mean, "mean"), (sq_mean, "sq_mean"), (dev_row, "dev_row")):
        _dev_f32(t, nm)
    n = w.numel()
    keep = n_collected / (n_collected + 1.0)
    rc = lib().ursa_swag_collect(_ptr(w), _ptr(mean), _ptr(sq_mean), _ptr(dev_row), n, keep, n_collected + 1.0, _stream(w))
    _check(rc, "ursa_swag_collect")


@_on_device
def swag_variance(mean, sq_mean, var, clamp=1e-30):
    for t, nm in ((mean, "mean"), (sq_mean, "sq_mean"), (var, "var")):
        _dev_f32(t, nm)
    _check(lib().ursa_swag_variance(_ptr(mean), _ptr(sq_mean), _ptr(var), mean.numel(), clamp, _stream(mean)),
           "ursa_swag_variance")
    return var


@_on_device
def swag_draw(out, mean, var, D, ring=None, z2=None, z1=None, rank_div=1.0, seed=0, step=0):
    """out: [S, ld]; ring: [K, ld] or None; z2: [S, K]; z1: [S, ld] or None (Philox)."""
    _dev_f32(out, "out"), _dev_f32(mean, "mean"), _dev_f32(var, "var")
    _dev_f32(ring, "ring", True), _dev_f32(z2, "z2", True), _dev_f32(z1, "z1", True)
    S = out.shape[0]
    K = 0 if ring is None else ring.shape[0]
    rc = lib().ursa_swag_draw(_ptr(out), out.stride(0), _ptr(mean), _ptr(var), _ptr(ring),
                              0 if ring is None else ring.stride(0), K, _ptr(z2), _ptr(z1),
                              0 if z1 is None else z1.stride(0), S, D, rank_div, seed, step, _stream(out))
    _check(rc, "ursa_swag_draw")
    return out


def swag_draw_noise(S, D, seed, step, device):
    """The [S, roundup4(D)] N(0, 1) stream ``swag_draw`` uses when z1 is None: a diagonal draw with mean 0 and variance 1
    returns it exactly (0 + 1 * z)."""
    ld = (D + 3) // 4 * 4
    out = torch.zeros(S, ld, device=device)
    return swag_draw(out, torch.zeros(ld, device=device), torch.ones(ld, device=device), D, seed=seed, step=step)


@_on_device
def swag_gram(ring, D):
    """ring: [K, ld] device fp32 -> [K, K] float64 device tensor R R^T over the first D columns."""
    _dev_f32(ring, "ring")
    K = ring.shape[0]
    gram = torch.empty(K, K, dtype=torch.float64, device=ring.device)
    _check(lib().ursa_swag_gram(_ptr(ring), ring.stride(0), K, D, _ptr(gram), _stream(ring)), "ursa_swag_gram")
    return gram


@_on_device
def bma_accumulate(logits, proba_sum, entropy_sum, gamma=1e-4):
    """logits: [S, N, C] contiguous."""
    _dev_f32(logits, "logits"), _dev_f32(proba_sum, "proba_sum"), _dev_f32(entropy_sum, "entropy_sum")
    S, N, C = logits.shape
    rc = lib().ursa_bma_accumulate(_ptr(logits), S, N, C, N * C, _ptr(proba_sum), _ptr(entropy_sum), gamma,
                                   _stream(logits))
    _check(rc, "ursa_bma_accumulate")


@_on_device
def bma_metrics(proba_sum, num_samples, targets, gamma=1e-4, n_bins=15, want_rows=False):
    """Returns (i64[1+2*nb], f64[2+nb], pred|None, conf|None) as device tensors."""
    _dev_f32(proba_sum, "proba_sum")
    if targets.dtype != torch.int64 or not targets.is_cuda:
        raise ValueError("targets must be a CUDA int64 tensor")
    N, C = proba_sum.shape
    dev = proba_sum.device
    out_i = torch.empty(1 + 2 * n_bins, dtype=torch.int64, device=dev)
    out_f = torch.empty(2 + n_bins, dtype=torch.float64, device=dev)
    pred = torch.empty(N, dtype=torch.int32, device=dev) if want_rows else None
    conf = torch.empty(N, dtype=torch.float32, device=dev) if want_rows else None
    wsb = lib().ursa_bma_metrics_workspace(N, n_bins)
    ws = torch.empty((wsb + 7) // 8, dtype=torch.int64, device=dev)
    rc = lib().ursa_bma_metrics(_ptr(proba_sum), N, C, float(num_samples), _ptr(targets), gamma, n_bins, _ptr(out_i),
                                _ptr(out_f), _ptr(pred), _ptr(conf), _ptr(ws), ws.numel() * 8, _stream(proba_sum))
    _check(rc, "ursa_bma_metrics")
    return out_i, out_f, pred, conf


@_on_device
def bma_mlp_forward(bank, S, x, in_dim, hidden, C, proba_sum, entropy_sum, logits_out=None, gamma=1e-4,
                    algo=ALGO_FFMA, workspace=None):
    _dev_f32(bank, "bank"), _dev_f32(x, "x"), _dev_f32(proba_sum, "proba_sum"), _dev_f32(entropy_sum, "entropy_sum")
    _dev_f32(logits_out, "logits_out", True)
    N = x.shape[0]
    need = lib().ursa_bma_mlp_workspace(S, N, in_dim, hidden, C, algo)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    rc = lib().ursa_bma_mlp_forward(_ptr(bank), bank.stride(0), S, _ptr(x), N, in_dim, hidden, C, _ptr(proba_sum),
                                    _ptr(entropy_sum), _ptr(logits_out), gamma, _ptr(workspace),
                                    workspace.numel() * workspace.element_size(), algo, _stream(x))
    _check(rc, "ursa_bma_mlp_forward")
    return workspace


@_on_device
def bma_preresnet_forward(bank, bufbank, S, x, depth, C, proba_sum, entropy_sum, logits_out=None, gamma=1e-4,
                          algo=ALGO_FFMA, workspace=None):
    _dev_f32(bank, "bank"), _dev_f32(bufbank, "bufbank"), _dev_f32(x, "x")
    _dev_f32(proba_sum, "proba_sum"), _dev_f32(entropy_sum, "entropy_sum"), _dev_f32(logits_out, "logits_out", True)
    N = x.shape[0]
    need = lib().ursa_bma_preresnet_workspace(S, N, depth, C, algo)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    rc = lib().ursa_bma_preresnet_forward(_ptr(bank), bank.stride(0), _ptr(bufbank), bufbank.stride(0), S, _ptr(x), N,
                                          depth, C, _ptr(proba_sum), _ptr(entropy_sum), _ptr(logits_out), gamma,
                                          _ptr(workspace), workspace.numel() * workspace.element_size(), algo,
                                          _stream(x))
    _check(rc, "ursa_bma_preresnet_forward")
    return workspace


@_on_device
def gemm_nt(A, B, out, bias=None, relu=False, workspace=None):
    """out[b] = A[b or shared] @ B[b]^T (+ bias[b]) (ReLU) on the 3xTF32 tcgen05 GEMM.  A: [M, K] (shared) or [batch, M, K];
    B: [batch, N, K]; out: [batch, M, N]; bias: [batch, N] or None.  Innermost dims contiguous.  Returns the workspace."""
    for t, nm in ((A, "A"), (B, "B"), (out, "out"), (bias, "bias")):        # strided views are fine: strides are passed down
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or t.stride(-1) != 1):
            raise ValueError("%s must be a CUDA float32 tensor with a contiguous last dimension" % nm)
    batch, N, K = B.shape
    shared = A.dim() == 2
    M = A.shape[-2]
    if A.shape[-1] != K or tuple(out.shape) != (batch, M, N) or A.stride(-1) != 1 or B.stride(-1) != 1 or out.stride(-1) != 1:
        raise ValueError("gemm_nt: shape / stride mismatch")
    need = lib().ursa_gemm_nt_3xtf32_workspace(batch, M, N, K, 0 if shared else 1)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=B.device)
    rc = lib().ursa_gemm_nt_3xtf32(_ptr(A), A.stride(-2), 0 if shared else A.stride(0), _ptr(B), B.stride(1), B.stride(0),
                                   _ptr(bias), 0 if bias is None else bias.stride(0), 1 if relu else 0, _ptr(out), out.stride(1),
                                   out.stride(0), batch, M, N, K, _ptr(workspace), workspace.numel() * workspace.element_size(),
                                   _stream(B))
    _check(rc, "ursa_gemm_nt_3xtf32")
    return workspace


@_on_device
def bma_wrn_forward(bank, bufbank, S, x, depth, widen, C, proba_sum, entropy_sum, logits_out=None, gamma=1e-4,
                    algo=ALGO_TCGEN05, workspace=None):
    """WideResNet (WRN-depth-widen) BMA forward: every conv is a persistent 3xTF32 tcgen05 implicit GEMM."""
    _dev_f32(bank, "bank"), _dev_f32(bufbank, "bufbank"), _dev_f32(x, "x")
    _dev_f32(proba_sum, "proba_sum"), _dev_f32(entropy_sum, "entropy_sum"), _dev_f32(logits_out, "logits_out", True)
    N = x.shape[0]
    need = lib().ursa_bma_wrn_workspace(S, N, depth, widen, C, algo)
    if need == 0:
        raise UrsaError("ursa_bma_wrn_workspace: unsupported WRN-%d-%d / algo %d" % (depth, widen, algo))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    rc = lib().ursa_bma_wrn_forward(_ptr(bank), bank.stride(0), _ptr(bufbank), bufbank.stride(0), S, _ptr(x), N,
                                    depth, widen, C, _ptr(proba_sum), _ptr(entropy_sum), _ptr(logits_out), gamma,
                                    _ptr(workspace), workspace.numel() * workspace.element_size(), algo, _stream(x))
    _check(rc, "ursa_bma_wrn_forward")
    return workspace


@_on_device
def wrn_bn_update(bank_row, buf_row, x, batch, depth, widen, C, workspace=None, algo=ALGO_TCGEN05):
    """Re-estimate the BatchNorm running statistics of ONE WideResNet sample with a train-mode pass over ``x`` (batches of
    ``batch`` images); ``buf_row`` [nb] is overwritten.  ``algo``: ALGO_TCGEN05 (3xTF32) or ALGO_TCGEN05_F16.  Returns the
    workspace (None if the shape is not covered)."""
    _dev_f32(bank_row, "bank_row"), _dev_f32(buf_row, "buf_row"), _dev_f32(x, "x")
    N = x.shape[0]
    need = lib().ursa_wrn_bn_update_workspace(N, batch, depth, widen, C)
    if need == 0:
        return None
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    rc = lib().ursa_wrn_bn_update_algo(_ptr(bank_row), _ptr(buf_row), _ptr(x), N, batch, depth, widen, C, _ptr(workspace),
                                       workspace.numel() * workspace.element_size(), algo, _stream(x))
    _check(rc, "ursa_wrn_bn_update")
    return workspace


@_on_device
def preresnet_bn_update(bank_rows, buf_rows, x, batch, depth, C, workspace=None):
    """Re-estimate the BatchNorm running statistics of S PreResNet samples (rows of ``bank_rows`` [S, ld]) with ONE
    sample-batched train-mode pass over ``x``; ``buf_rows`` [S, ldb] is overwritten.  Returns the workspace (None if the
    shape is not covered)."""
    _dev_f32(bank_rows, "bank_rows"), _dev_f32(buf_rows, "buf_rows"), _dev_f32(x, "x")
    S, N = bank_rows.shape[0], x.shape[0]
    need = lib().ursa_preresnet_bn_update_workspace(S, N, batch, depth, C)
    if need == 0:
        return None
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    rc = lib().ursa_preresnet_bn_update(_ptr(bank_rows), bank_rows.stride(0), _ptr(buf_rows), buf_rows.stride(0), S, _ptr(x), N,
                                        batch, depth, C, _ptr(workspace), workspace.numel() * workspace.element_size(),
                                        _stream(x))
    _check(rc, "ursa_preresnet_bn_update")
    return workspace


@_on_device
def hmc_mlp_grad(theta, x, y, in_dim, hidden, n_classes, grad, ce, workspace=None, engine="tf32"):
    """Chain-batched MLP likelihood gradient on tcgen05 (see ursa_hmc_mlp_grad / ursa_hmc_mlp_grad_f16).  theta, grad: [C, ld];
    x: [N, in]; y: [N] int64; ce: [C].  ``engine``: "tf32" (3xTF32 kernel) or "f16" (persistent 2xFP16-split kernel).  Returns
    the workspace (None if the shape is not covered)."""
    _dev_f32(theta, "theta"), _dev_f32(x, "x"), _dev_f32(grad, "grad"), _dev_f32(ce, "ce")
    if y.dtype != torch.int64 or not y.is_cuda or not y.is_contiguous():
        raise ValueError("y must be a contiguous CUDA int64 tensor")
    if engine not in ("tf32", "f16"):
        raise ValueError("hmc_mlp_grad: engine must be 'tf32' or 'f16'")
    C, ld = theta.shape
    N = x.shape[0]
    if tuple(grad.shape) != (C, ld) or ce.numel() < C or x.shape[1] != in_dim or y.numel() != N:
        raise ValueError("hmc_mlp_grad: shape mismatch")
    ws_fn, fn = ((lib().ursa_hmc_mlp_grad_f16_workspace, lib().ursa_hmc_mlp_grad_f16) if engine == "f16"
                 else (lib().ursa_hmc_mlp_grad_workspace, lib().ursa_hmc_mlp_grad))
    need = ws_fn(C, N, in_dim, hidden, n_classes)
    if need == 0:
        return None
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    rc = fn(_ptr(theta), ld, C, _ptr(x), y.data_ptr(), N, in_dim, hidden, n_classes, _ptr(grad), _ptr(ce),
            _ptr(workspace), workspace.numel() * workspace.element_size(), _stream(theta))
    _check(rc, "ursa_hmc_mlp_grad" + ("_f16" if engine == "f16" else ""))
    return workspace


@_on_device
def hmc_momentum(r, sqrt_mass, noise=None, seed=0, step=0, elem_offset=0):
    """r[C, ld] = sqrt_mass * z (z from ``noise`` or Philox)."""
    _dev_f32(r, "r"), _dev_f32(noise, "noise", True)
    if noise is not None and noise.numel() < r.numel():
        raise ValueError("noise has fewer elements than r")
    _check(lib().ursa_hmc_momentum(_ptr(r), _ptr(noise), r.numel(), float(sqrt_mass), int(seed), int(step),
                                   int(elem_offset), _stream(r)), "ursa_hmc_momentum")
    return r


@_on_device
def hmc_leapfrog(theta, r, g_nll, *, kick, drift, tau, tau_out=1.0, snapshot=None):
    """r += kick * grad_logp ; theta += drift * r  with grad_logp = -(tau_out * g_nll + tau * theta)."""
    for t, nm, opt in ((theta, "theta", False), (r, "r", False), (g_nll, "g_nll", False), (snapshot, "snapshot", True)):
        _dev_f32(t, nm, opt)
    n = theta.numel()
    for t, nm in ((r, "r"), (g_nll, "g_nll"), (snapshot, "snapshot")):
        if t is not None and t.numel() < n:
            raise ValueError("%s has fewer elements than theta" % nm)
    _check(lib().ursa_hmc_leapfrog(_ptr(theta), _ptr(r), _ptr(g_nll), _ptr(snapshot), n, float(kick), float(drift),
                                   float(tau), float(tau_out), _stream(theta)), "ursa_hmc_leapfrog")


@_on_device
def hmc_energy(theta, r, D, out=None, workspace=None):
    """theta, r: [C, ld].  Returns (sums[2, C] float64 = [sum theta^2, sum r^2], workspace)."""
    _dev_f32(theta, "theta"), _dev_f32(r, "r")
    C, ld = theta.shape
    if out is None:
        out = torch.empty(2, C, dtype=torch.float64, device=theta.device)
    need = lib().ursa_hmc_energy_workspace(C, D)
    if workspace is None or workspace.numel() * 8 < need:
        workspace = torch.empty(max(1, (need + 7) // 8), dtype=torch.float64, device=theta.device)
    _check(lib().ursa_hmc_energy(_ptr(theta), _ptr(r), C, D, ld, out[0].data_ptr(), out[1].data_ptr(), _ptr(workspace),
                                 workspace.numel() * 8, _stream(theta)), "ursa_hmc_energy")
    return out, workspace


@_on_device
def hmc_accept(theta, saved, h_old, h_new, accept, *, logu=None, keep_dst=None, keep_src=None, out=None, seed=0, step=0,
               chain_offset=0):
    """Metropolis accept / restore per chain (see ursa_hmc_accept).  h_old, h_new: float64 [C]; accept: int32 [C]."""
    for t, nm, opt in ((theta, "theta", False), (saved, "saved", False), (logu, "logu", True), (keep_dst, "keep_dst", True),
                       (keep_src, "keep_src", True)):
        _dev_f32(t, nm, opt)
    C, ld = theta.shape
    if out is not None and not (out.is_cuda and out.dtype == torch.float32 and out.dim() == 2 and out.stride(1) == 1
                                and out.shape[0] >= C and out.shape[1] >= ld):
        raise ValueError("out must be a CUDA float32 [C, >= ld] tensor with unit column stride")
    for t, nm in ((h_old, "h_old"), (h_new, "h_new")):
        if t.dtype != torch.float64 or not t.is_cuda or t.numel() < C or not t.is_contiguous():
            raise ValueError("%s must be a contiguous CUDA float64 tensor of C elements" % nm)
    if accept.dtype != torch.int32 or not accept.is_cuda or accept.numel() < C:
        raise ValueError("accept must be a CUDA int32 tensor of C elements")
    _check(lib().ursa_hmc_accept(_ptr(theta), _ptr(saved), _ptr(keep_dst), _ptr(keep_src), _ptr(out),
                                 0 if out is None else out.stride(0), C, ld, h_old.data_ptr(), h_new.data_ptr(),
                                 _ptr(logu), accept.data_ptr(), int(seed), int(step), int(chain_offset), _stream(theta)),
           "ursa_hmc_accept")
    return accept


def launch_count():
    """Kernel launches issued by the library in this process so far."""
    return int(lib().ursa_launch_count())


PROF_KINDS = ("stage_c16", "stage_c32", "stage_c64", "stem", "shortcut", "conv_s2", "head", "prep")


def profile_begin(capacity=1 << 16):
    _check(lib().ursa_profile_begin(int(capacity)), "ursa_profile_begin")


def profile_end():
    """{kind: (summed ms, launches)} of the instrumented PreResNet-forward kernels since ``profile_begin``."""
    n = len(PROF_KINDS)
    ms, cnt = (_f64 * n)(), (_i64 * n)()
    _check(lib().ursa_profile_end(ms, cnt, n), "ursa_profile_end")
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(PROF_KINDS)}


def device_info():
    sm, major, minor = _i32(), _i32(), _i32()
    _check(lib().ursa_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "ursa_device_info")
    return sm.value, major.value, minor.value
