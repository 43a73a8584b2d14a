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
