"""CPU: the C-ABI library loads and exports every symbol include/dsmnet_b200.h declares
(no compute calls without a GPU), error strings work, and argument validation returns codes."""
import ctypes
import os
import re

import pytest

from dsmnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dsmnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsm_[a-z0-9_]+)\s*\(", text)))


def test_library_present():
    assert os.path.isfile(_lib.LIB_PATH), "run `make` or __graft_entry__.build() first"


def test_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "library does not export %s" % s
    # and the Python binding table covers the header too
    assert set(syms) <= set(_lib.SIGNATURES), set(syms) - set(_lib.SIGNATURES)


def test_version_and_strerror():
    L = _lib.lib()
    assert L.dsm_abi_version() == 1
    assert b"ok" == L.dsm_strerror(0)
    assert b"DSM_EINVAL" in L.dsm_strerror(-1)
    assert b"DSM_EUNSUPPORTED" in L.dsm_strerror(-2)


def test_argument_validation_needs_no_gpu():
    L = _lib.lib()
    # null pointers / bad shapes are rejected before any CUDA call
    assert L.dsm_corr1d_fwd(0, 0, 0, 1, 1, 1, 1, 1, 1, 0) == -1
    assert L.dsm_concat_volume_fwd(0, 0, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0) == -1
    assert L.dsm_softargmin_fwd(0, 0, 1, 1, 1, 1, 1.0, 0) == -1
    assert L.dsm_warp_fwd(0, 0, 0, 0, 0.0, 0, 0, 1, 1, 2, 2, 2, 2, 0) == -1
    assert L.dsm_conv3d_fwd(0, 0, 0, 0, 0, 0, 1, 32, 32, 4, 4, 4, 1, 0, 0, 1, 0, 0, 0) == -1


def test_no_cpu_fallback():
    import torch
    from dsmnet_b200.corr1d import corr1d
    with pytest.raises(_lib.DsmError):
        corr1d(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 8), 3)


def test_training_entry_points_validate_arguments():
    """the training-path entry points reject null pointers / unsupported channel counts before any CUDA call"""
    L = _lib.lib()
    assert L.dsm_bn_stats(0, 1, 32, 2, 2, 2, 0, 0) == -1                       # null pointers
    assert L.dsm_bn_stats(64, 1, 48, 2, 2, 2, 64, 0) == -2                     # C not in {32, 64, 128}
    assert L.dsm_bn_act_fwd(0, 0, 0, 0, 1, 0, 1, 32, 2, 2, 2, 0) == -1
    assert L.dsm_bn_act_fwd(64, 64, 64, 0, 3, 64, 1, 32, 2, 2, 2, 0) == -1     # relu mode out of range
    assert L.dsm_bn_act_bwd_reduce(64, 64, 0, 64, 64, 1, 64, 1, 32, 2, 2, 2, 0) == -1    # relu 1 needs z
    assert L.dsm_bn_act_bwd(64, 64, 0, 64, 64, 64, 1, 64, 0, 1, 32, 2, 2, 2, 0) == -1
    assert L.dsm_bn_finalize_fwd(0, 0, 0, 0, 32, 10, 1e-5, 0.1, 0, 0, 0, 0, 0, 0, 0) == -1
    assert L.dsm_zero_rim(0, 1, 32, 2, 2, 2, 0) == -1
    assert L.dsm_zero_rim(64, 1, 12, 2, 2, 2, 0) == -1                         # C % 8 != 0
    assert L.dsm_conv3d_c1_bwd(0, 0, 0, 0, 0, 1, 2, 2, 2, 2, 2, 2, 0, 0, 0, 0) == -1
    assert L.dsm_upsample_softargmin_bwd(0, 0, 0, 0, 0, 1, 2, 2, 2, 8, 8, 8, 1, 0) == -1
    assert L.dsm_pack_weight(0, 0, 32, 32, 0, 0) == -1
    assert L.dsm_pack_weight(64, 64, 32, 32, 3, 0) == -1                       # unknown mode
    assert L.dsm_conv3d_wgrad(0, 0, 0, 1, 32, 32, 2, 2, 2, 2, 2, 2, 1, 32, 32, 0, 0, 0, 0, 0, 0) == -1
    # workspace queries are pure arithmetic
    assert L.dsm_conv3d_wgrad_workspace_bytes(32, 32) > 0
    assert L.dsm_conv3d_wgrad_workspace_bytes_ex(1, 64, 32, 24, 48, 156, 2) > L.dsm_conv3d_wgrad_workspace_bytes(64, 32)
    assert L.dsm_conv3d_wgrad_workspace_bytes_ex(1, 32, 32, 48, 96, 312, 1) == L.dsm_conv3d_wgrad_workspace_bytes(32, 32)
    assert L.dsm_conv3d_c1_bwd_workspace_bytes() > 0


def test_training_wrappers_have_no_cpu_fallback():
    import torch
    import torch.nn as nn
    from dsmnet_b200 import train3d as T
    from dsmnet_b200.cost_volume import concat_volume_padded
    from dsmnet_b200.softargmin import upsample_softargmin
    from dsmnet_b200.volume_layout import PaddedVolume
    y = torch.zeros(1 * 4 * 4 * 4 * 32, dtype=torch.bfloat16)
    with pytest.raises(_lib.DsmError):
        T.bn_act(PaddedVolume(y, 1, 32, 2, 2, 2), nn.BatchNorm3d(32).train())
    with pytest.raises(_lib.DsmError):
        concat_volume_padded(torch.zeros(1, 8, 4, 8), torch.zeros(1, 8, 4, 8), 3, "psm")
    with pytest.raises(_lib.DsmError):
        upsample_softargmin(torch.zeros(1, 2, 3, 4, requires_grad=True), (8, 12, 16), True)
    with pytest.raises(_lib.DsmError):
        T.conv_c1(PaddedVolume(y, 1, 32, 2, 2, 2), nn.Conv3d(32, 1, 3, padding=1))


def test_trunk_entry_points_validate_arguments():
    """the 2-D trunk entry points reject bad pointers / shapes / alignment before any CUDA call
    (x, w, scale, shift, residual, y, B, Cin, Cout, H, W, dilation, relu, rim_in, rim_out, ldx, ldy, ldr, variant, stream)"""
    L = _lib.lib()
    ok = (64, 64, 0, 0, 0, 64)
    assert L.dsm_conv2d_rs_fwd(0, 0, 0, 0, 0, 0, 1, 64, 64, 8, 8, 1, 1, 1, 1, 64, 64, 64, 0, 0) == -1      # null pointers
    assert L.dsm_conv2d_rs_fwd(*ok, 1, 48, 64, 8, 8, 1, 1, 1, 1, 64, 64, 64, 0, 0) == -2                  # Cin not in {32, 64, 128}
    assert L.dsm_conv2d_rs_fwd(*ok, 1, 128, 32, 8, 8, 1, 1, 2, 2, 128, 32, 32, 0, 0) == -2                # 128 -> 32 stays per-tile
    assert L.dsm_conv2d_rs_fwd(*ok, 1, 64, 64, 8, 8, 2, 1, 1, 1, 64, 64, 64, 0, 0) == -1                  # rim narrower than the dilation
    assert L.dsm_conv2d_rs_fwd(64, 64, 0, 0, 0, 80, 1, 64, 64, 8, 8, 1, 1, 1, 1, 64, 64, 64, 0, 0) == -3  # y not 32-byte aligned
    assert L.dsm_conv2d_rs_fwd(*ok, 1, 64, 64, 8, 8, 1, 1, 1, 1, 64, 72, 64, 0, 0) == -1                  # ldy not a multiple of 16
    assert L.dsm_conv2d_rs_fwd(*ok, 1, 64, 64, 8, 8, 1, 2, 1, 1, 64, 64, 64, 0, 0) == -1                  # relu mode 2 is 3-D only
    assert L.dsm_spp_workspace_bytes(1, 63, 200) == 0 and L.dsm_spp_workspace_bytes(1, 96, 312) > 0       # AvgPool2d(64) needs 64 x 64
