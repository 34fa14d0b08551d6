import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')

# parity tolerances (BASELINE.json north_star): max-norm relative per tensor
RTOL_FLUX = 1e-5
RTOL_GRAD = 1e-4


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> nested dict of torch tensors / numpy scalars from tests/golden/<name>.npz."""
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    out = {}
    for k in z.files:
        v = z[k]
        val = torch.from_numpy(v) if v.dtype.kind == 'f' and v.ndim > 0 else v
        if '/' in k:
            a, b = k.split('/', 1)
            out.setdefault(a, {})[b] = val
        else:
            out[k] = val
    return out


def maxnorm_err(a, b):
    """||a - b||_inf / ||b||_inf (b = reference)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    num = (a - b).abs().max().item()
    return num / denom if denom > 0 else num


def assert_close(a, b, rtol, what):
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    assert torch.isfinite(a).all(), f'{what}: non-finite values'
    e = maxnorm_err(a, b)
    assert e <= rtol, f'{what}: max-norm relative error {e:.3e} > {rtol:.1e}'
