"""ctypes binding of the C-ABI library ``libdsmnet_b200.so`` (include/dsmnet_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is
raised.  Building is ``make`` at the repo root or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libdsmnet_b200.so")

DSM_F32, DSM_BF16 = 0, 1
DSM_NCDHW, DSM_NDHWC_PADDED = 0, 1
DSM_VOL_PSM, DSM_VOL_GC, DSM_VOL_GC_RIGHT = 0, 1, 2
VOLUME_MODES = {"psm": DSM_VOL_PSM, "gc": DSM_VOL_GC, "gc_right": DSM_VOL_GC_RIGHT}


class DsmError(RuntimeError):
    pass


_P, _I, _F = c_void_p, c_int, c_float

# name -> argtypes; every function returns int except the two noted below
SIGNATURES = {
    "dsm_abi_version": [],
    "dsm_strerror": [_I],
    "dsm_corr1d_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dsm_corr1d_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dsm_concat_volume_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_concat_volume_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_conv3d_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, c_size_t, _P],
    "dsm_conv3d_fwd_ex": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_conv3d_wgrad_workspace_bytes": [_I, _I],
    "dsm_conv3d_wgrad_workspace_bytes_ex": [_I, _I, _I, _I, _I, _I, _I],
    "dsm_conv3d_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, c_size_t, _P],
    "dsm_zero_rim": [_P, _I, _I, _I, _I, _I, _P],
    "dsm_bn_stats": [_P, _I, _I, _I, _I, _I, _P, _P],
    "dsm_bn_finalize_fwd": [_P, _P, _P, _P, _I, ctypes.c_longlong, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "dsm_bn_act_fwd": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "dsm_bn_act_bwd_reduce": [_P, _P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "dsm_bn_finalize_bwd": [_P, _P, _P, _P, _I, ctypes.c_longlong, _P, _P, _P, _P],
    "dsm_bn_act_bwd": [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "dsm_conv3d_c1_bwd_workspace_bytes": [],
    "dsm_conv3d_c1_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, c_size_t, _P],
    "dsm_debug_conv_timeouts": [],
    "dsm_debug_conv_set_trap": [_I],
    "dsm_crop_add_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_crop_add_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_debug_wgrad_mode": [_I],
    "dsm_debug_wgrad_timeouts": [],
    "dsm_pack_weight": [_P, _P, _I, _I, _I, _P],
    "dsm_pack_ndhwc": [_P, _P, _I, _I, _I, _I, _I, _P],
    "dsm_unpack_ndhwc": [_P, _P, _I, _I, _I, _I, _I, _P],
    "dsm_conv3d_volume_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_pack_nhwc_bf16": [_P, _P, _I, _I, _I, _I, _I, _P],
    "dsm_pack_nhwc_bf16_pair": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "dsm_conv2d_fwd": [_P, _P, _P, _P, _P, _P] + [_I] * 16 + [_P],
    "dsm_conv2d_rs_fwd": [_P, _P, _P, _P, _P, _P] + [_I] * 13 + [_P],
    "dsm_conv2d_first_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "dsm_spp_workspace_bytes": [_I, _I, _I],
    "dsm_spp_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, c_size_t, _P],
    "dsm_softargmin_fwd": [_P, _P, _I, _I, _I, _I, _F, _P],
    "dsm_softargmin_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "dsm_disparity_regression_fwd": [_P, _P, _I, _I, _I, _I, _P],
    "dsm_disparity_regression_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "dsm_upsample_softargmin_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_upsample_softargmin_fwd_lse": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_upsample_softargmin_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "dsm_warp_fwd": [_P, _P, _P, _P, _F, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "dsm_warp_indices": [_P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "dsm_debug_conv_set_progress": [_P],
    "dsm_warp_fwd_batched": [_P, _I, _P],
    "dsm_warp_bwd_batched": [_P, _I, _P],
    "dsm_ssim_fwd": [_P, _P, _P, _I, _I, _I, _I, _P],
    "dsm_ssim_bwd_workspace_bytes": [_I, _I, _I],
    "dsm_ssim_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P, c_size_t, _P],
    "dsm_warp_bwd": [_P, _P, _P, _P, _P, _F, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
}



class DsmWarpJob(ctypes.Structure):
    """include/dsmnet_b200.h: one warp of a batched launch"""
    _fields_ = [("src", _P), ("disp", _P), ("row", _P), ("col", _P), ("out", _P), ("gout", _P), ("gsrc", _P), ("gdisp", _P),
                ("delt", _F), ("fliplr", _I), ("B", _I), ("C", _I), ("H0", _I), ("W0", _I), ("H", _I), ("W", _I)]


WARP_MAX_JOBS = 32

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (see Makefile)."""
    r = subprocess.run(["make", "-C", _ROOT, "-j8", "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:]); print(r.stderr[-4000:])
    if r.returncode != 0:
        raise DsmError("building libdsmnet_b200.so failed (nvcc / make)")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DsmError("%s not found: run `make` (or __graft_entry__.build()); there is no fallback path" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError here == ABI mismatch; let it surface
        fn.argtypes = argtypes
        fn.restype = c_char_p if name == "dsm_strerror" else (c_size_t if (name.endswith("_workspace_bytes") or name.endswith("_workspace_bytes_ex")) else c_int)
    if L.dsm_abi_version() != 1:
        raise DsmError("ABI version mismatch")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().dsm_strerror(rc)
        raise DsmError("%s failed: %s (code %d)" % (what, msg.decode() if msg else "?", rc))


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DsmError("dsmnet_b200 ops run on CUDA (sm_100a) tensors only; got a %s tensor — there is no CPU path" % t.device)
