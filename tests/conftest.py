import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200 / sm_100a); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        a = z[k]
        if a.dtype.kind in "US":
            out[k] = str(a)
        elif a.ndim == 0:
            out[k] = a.item()
        elif a.dtype.kind == "f":
            out[k] = torch.from_numpy(a.astype(np.float32))
        else:
            out[k] = a.tolist() if a.ndim == 1 and a.size <= 4 else torch.from_numpy(a)
    return out


@pytest.fixture
def golden():
    return load_golden


def rel_err(a, b):
    """NORM-WISE relative error: max |a-b| / max |b| over the whole tensor (an error bound relative to the tensor's
    scale, not per element — see rel_err_elem for the element-wise reading of north_star's "1e-4 relative")."""
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def rel_err_elem(a, b, floor_frac=1e-2):
    """ELEMENT-WISE relative error with an absolute floor: max_i |a_i - b_i| / max(|b_i|, floor_frac * max |b|).
    Elements smaller than 1 % of the tensor's scale (results of cancellation, which no fp32 summation order can hold to a
    relative 1e-4) are judged against that floor; every other element against its own magnitude."""
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    floor = max(float(b.abs().max()) * floor_frac, 1e-30)
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())
