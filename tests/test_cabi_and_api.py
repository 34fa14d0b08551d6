"""CPU: the C-ABI library builds/loads and exports every symbol include/hbv_b200.h declares;
the host-side API mirrors the reference's (tests/test_methods.py of the reference)."""

import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'hbv_b200.h')).read()
    return sorted(set(re.findall(r'HBV_API\s+[\w\s\*]+?\b(hbv_b200_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from hydrodl2_b200 import _build, _cabi
    _build.build()
    lib = ctypes.CDLL(_cabi.lib_path())
    syms = _declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f'missing export {s}'
    assert set(syms) == set(_cabi.EXPORTS)
    assert _cabi.load().hbv_b200_abi_version() == _cabi.ABI_VERSION


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C layout (checked against a tiny C program's sizeof)."""
    import subprocess
    import tempfile
    from hydrodl2_b200 import _cabi
    code = '#include <stdio.h>\n#include "hbv_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",' \
           'sizeof(hbv_desc_t),sizeof(hbv_fwd_io_t),sizeof(hbv_bwd_io_t),sizeof(hbv_route_desc_t),' \
           'sizeof(hbv_pair_desc_t));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, 't.c')
        open(c, 'w').write(code)
        exe = os.path.join(td, 't')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_cabi.HbvDesc), ctypes.sizeof(_cabi.HbvFwdIO),
                     ctypes.sizeof(_cabi.HbvBwdIO), ctypes.sizeof(_cabi.HbvRouteDesc),
                     ctypes.sizeof(_cabi.HbvPairDesc)]


def test_argument_validation_without_gpu():
    """Argument errors are reported before any CUDA call (safe on a CPU box)."""
    from hydrodl2_b200 import _cabi as A
    lib = A.load()
    d = A.HbvDesc()
    io = A.HbvFwdIO()
    assert lib.hbv_b200_fwd(None, None, None) == -1
    d.abi_version = 999
    io.forcing = 8
    io.state_in = 8
    assert lib.hbv_b200_fwd(ctypes.byref(d), ctypes.byref(io), None) < 0
    assert b'ABI' in lib.hbv_b200_last_error() or b'parameter' in lib.hbv_b200_last_error()
    assert lib.hbv_b200_route_chunks(730, 531) >= 1


def test_available_models():
    import hydrodl2_b200 as hydrodl2
    models = hydrodl2.available_models()
    assert isinstance(models, dict) and len(models) > 0
    for k, v in models.items():
        assert isinstance(k, str) and isinstance(v, list) and all(isinstance(i, str) for i in v)


def test_load_each_model():
    import hydrodl2_b200 as hydrodl2
    for _, names in hydrodl2.available_models().items():
        for name in names:
            cls = hydrodl2.load_model(name)
            assert isinstance(cls, type) and issubclass(cls, torch.nn.Module)


def test_load_model_by_version_and_missing():
    import hydrodl2_b200 as hydrodl2
    assert hydrodl2.load_model('hbv', ver_name='Hbv').__name__ == 'Hbv'
    with pytest.raises(ImportError):
        hydrodl2.load_model('nope')
    with pytest.raises(NotImplementedError):
        hydrodl2.load_module()


def test_model_attributes_match_reference_contract():
    import hydrodl2_b200 as hydrodl2
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 365, 'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': 16},
          device=torch.device('cpu'))
    assert m.learnable_param_count == 13 * 16 + 2
    assert list(m.parameter_bounds)[-1] == 'parBETAET'
    assert m.state_names == ['SNOWPACK', 'MELTWATER', 'SM', 'SUZ', 'SLZ']
    assert m.flux_names[0] == 'streamflow' and m.flux_names[-1] == 'BFI' and len(m.flux_names) == 17
    m0 = M({'dynamic_params': {'Hbv': []}, 'nmul': 16}, device=torch.device('cpu'))
    assert m0.learnable_param_count == 12 * 16 + 2
    with pytest.raises(KeyError):
        M({'nmul': 4})                                   # reference: config['dynamic_params'] required
    with pytest.raises(ValueError):
        m.load_states([torch.zeros(1)] * 5)              # list, not tuple (hbv.py:163-164)
    M2 = hydrodl2.load_model('hbv_1_1p', ver_name='Hbv_1_1p')
    m2 = M2({'dynamic_params': {'Hbv_1_1p': []}, 'nmul': 16}, device=torch.device('cpu'))
    assert m2.learnable_param_count == 226 and 'capillary' in m2.flux_names


def test_host_side_policies():
    """Choices made on the host (no GPU needed): checkpoint interval, gradient-plane mode."""
    from hydrodl2_b200 import _cabi as A, ops
    import hydrodl2_b200 as hydrodl2
    lib = A.load()
    # every state while it fits 16 GiB (20 B per lane-step), else every 16th
    assert lib.hbv_b200_auto_ckpt(730, 531, 16) == 1           # C2
    assert lib.hbv_b200_auto_ckpt(730, 22500, 16) == 1         # north-star shard: 5.3 GB
    assert lib.hbv_b200_auto_ckpt(17520, 2500, 16) == 1        # C4 per GPU: 14 GB
    assert lib.hbv_b200_auto_ckpt(17520, 20000, 16) == 16      # C4 on one GPU: 112 GB
    dsc = A.HbvDesc()
    dsc.T, dsc.B, dsc.nmul, dsc.ckpt_interval = 730, 531, 16, 0
    assert lib.hbv_b200_workspace_bytes(ctypes.byref(dsc)) == 730 * 5 * 531 * 16 * 4
    dsc.ckpt_interval = 16
    assert lib.hbv_b200_workspace_bytes(ctypes.byref(dsc)) == 46 * 5 * 531 * 16 * 4
    dsc.T = 0
    assert lib.hbv_b200_workspace_bytes(ctypes.byref(dsc)) < 0
    dev = torch.device('cpu')
    hbv = hydrodl2.load_model('hbv', ver_name='Hbv')({'dynamic_params': {'Hbv': ['parBETA', 'parBETAET']}, 'nmul': 16}, device=dev)
    p11 = hydrodl2.load_model('hbv_1_1p', ver_name='Hbv_1_1p')
    d14 = list(p11({'dynamic_params': {'Hbv_1_1p': []}, 'nmul': 16}, device=dev).parameter_bounds)
    m14 = p11({'dynamic_params': {'Hbv_1_1p': d14}, 'nmul': 16}, device=dev)
    prev, ops.FUSED_ZERO_FILL = ops.FUSED_ZERO_FILL, None
    try:
        s2 = hbv._spec(hbv.dynamic_params, True)
        assert ops._fused_zero_fill(s2, 210, 531) is False      # small grid: memset in stream order
        assert ops._fused_zero_fill(s2, 210, 22500) is True     # large grid: the adjoint writes its rows
        s14 = m14._spec(m14.dynamic_params, True)
        assert ops._fused_zero_fill(s14, 226, 531) is True      # dense-dynamic: always
        hbv4 = hydrodl2.load_model('hbv', ver_name='Hbv')({'dynamic_params': {'Hbv': ['parBETA']}, 'nmul': 4}, device=dev)
        assert ops._fused_zero_fill(hbv4._spec(hbv4.dynamic_params, True), 50, 100000) is False
    finally:
        ops.FUSED_ZERO_FILL = prev


def test_cpu_tensors_are_rejected():
    """There is no CPU path: a forward on CPU tensors fails loudly instead of falling back."""
    import hydrodl2_b200 as hydrodl2
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'dynamic_params': {'Hbv': []}, 'nmul': 4}, device=torch.device('cpu'))
    with pytest.raises(RuntimeError, match='no CPU path|CUDA'):
        m({'x_phy': torch.zeros(5, 3, 3)}, torch.zeros(5, 3, 12 * 4 + 2))
