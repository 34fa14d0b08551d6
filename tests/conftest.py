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


# per-parameter-block gradient check: floor (fraction of the tensor's max-norm) below which a
# block's error is fp32 accumulation noise — the reference's own fp32 gradient differs from its
# fp64 evaluation by up to 4e-5 of a block that is 1e-3 of the largest (measured on the golden
# cases), i.e. ~1e-7 of the tensor's max-norm
GRAD_BLOCK_FLOOR = 2e-7


def assert_grad_close(a, b, what, nmul, rtol=RTOL_GRAD):
    """Parameter-gradient parity PER PARAMETER BLOCK: the last dimension is cut into runs of `nmul`
    columns (one physical parameter each; a shorter tail = the routing columns) and every block
    must satisfy  ||a_blk - b_blk||_inf <= rtol * ||b_blk||_inf + GRAD_BLOCK_FLOOR * ||b||_inf.
    A single max-norm over the whole tensor would let a block whose gradient is 1000x smaller
    than the largest (parCFR, parCWH next to parFC) be entirely wrong and still pass."""
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    assert torch.isfinite(a).all(), f'{what}: non-finite values'
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    gmax = b.abs().max().item()
    ncol = a.shape[-1]
    worst = (0.0, -1)
    for c0 in range(0, ncol, nmul):
        ab, bb = a[..., c0:c0 + nmul], b[..., c0:c0 + nmul]
        nb = bb.abs().max().item()
        e = (ab - bb).abs().max().item()
        allowed = rtol * nb + GRAD_BLOCK_FLOOR * gmax
        assert e <= allowed, (f'{what}: parameter block {c0 // nmul} (columns {c0}..{min(ncol, c0 + nmul) - 1}): '
                              f'error {e:.3e} > {allowed:.3e} (block max-norm {nb:.3e}, tensor max-norm {gmax:.3e})')
        # (reported: the worst block among those whose tolerance is the relative term, i.e. whose
        # max-norm is above GRAD_BLOCK_FLOOR / rtol = 2e-3 of the tensor's)
        if nb * rtol >= GRAD_BLOCK_FLOOR * gmax and e / nb > worst[0]:
            worst = (e / nb, c0 // nmul)
    return worst


# `excs` = max(SM1 - FC, 0) (hbv.py:469-471) is a difference of two O(100) storages: one float32
# ulp of SM is up to 1e-3 of the excess series' max-norm where little excess occurs, and the
# reference's own fp32 evaluation differs from its fp64 evaluation by up to 9e-6 of that norm on
# the golden cases (hbv_2_d3_nowarm; scripts/parity_report.py).  That one series is held to 1e-4
# against the goldens; the fp64-arbiter gate of SURVEY §8 d6 (err(new, fp64) <= 2 err(ref32, fp64),
# tests/test_arbiter_gpu.py) is what bounds it tightly.
RTOL_BY_KEY = {'excs': 1e-4}


def flux_rtol(key: str) -> float:
    return RTOL_BY_KEY.get(key, RTOL_FLUX)


# Storages start at 0.001 (hbv.py:133) and the hourly model lifts anything below `nearzero` back
# at the next step (hbv_2_hourly.py:529-533); an upper-zone storage that empties every step
# (SUZ = SUZ1 - (SUZ1 / dt) * dt) is pure rounding residue, ~1e-11, whose value depends on whether
# `/ dt` is a division (ATen on the CPU) or a multiplication by 1/dt (ATen on CUDA, and here).
# State tensors are therefore compared relative to max(||ref||_inf, STATE_FLOOR).
STATE_FLOOR = 1e-3


def assert_close(a, b, rtol, what, floor=0.0):
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    assert torch.isfinite(a).all(), f'{what}: non-finite values'
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = max(b.abs().max().item(), floor)
    num = (a - b).abs().max().item()
    e = num / denom if denom > 0 else num
    assert e <= rtol, f'{what}: max-norm relative error {e:.3e} > {rtol:.1e}'


@pytest.fixture(autouse=True)
def _reset_library_options():
    """Tests switch kernel families through `_cabi.set_option`; put every switch back to "unset"
    (the library's own policy) afterwards."""
    yield
    if torch.cuda.is_available():
        from hydrodl2_b200 import _cabi
        for name in ('lean', 'pipe', 'pipe_max', 'ring', 'lean_small', 'lean_bwd_ring', 'lean_deep', 'dense',
                     'dense_ns', 'dense_ns_bwd', 'dense_minb', 'ckpt', 'adj_bpb', 'copy_blocks', 'ckpt_layout'):
            _cabi.set_option(name, -1)


def arbiter_gate(new, ref32, ref64, what, slack=1e-6, floor=0.0):
    """the CUDA result may be no further from the float64 evaluation than twice the
    reference's own float32 evaluation is:  err(new, fp64) <= 2 err(ref32, fp64) + slack
    (max-norm, relative to max(||fp64||_inf, floor)).  Returns (err_new, err_ref)."""
    r64 = ref64.detach().double().cpu()
    den = max(r64.abs().max().item(), floor)
    if den == 0:
        den = 1.0
    e_new = (new.detach().double().cpu() - r64).abs().max().item() / den
    e_ref = (ref32.detach().double().cpu() - r64).abs().max().item() / den
    assert e_new <= 2 * e_ref + slack, (f'{what}: err(new, fp64) = {e_new:.3e} > 2 x err(ref fp32, fp64) = '
                                       f'{e_ref:.3e} (+ {slack:.0e})')
    return e_new, e_ref


def check_split_slice_vs_oracle(model, x_dict, p0, p1, dyn_names, out, grads, loss_key, what, n=24, nmul=16):
    """Oracle leg of the kernel-family tests on the hbv_2 family: basins are independent, so the
    first `n` units of a large-grid CUDA run (loss = out[loss_key].sum(), no routing) must match
    the CPU oracle run on those units alone — outputs to RTOL_FLUX, the dynamic / static parameter
    gradients per parameter block — and pass the float64 arbiter gate.  Series listed in
    RTOL_BY_KEY (cancellation-limited) are held to the gate alone: on these short runs their
    max-norm can be ~1e-4 mm, where one ulp of the storages they are a difference of is 1e-3 of it
    and the reference's own fp32 result is that far from fp64."""
    from oracle import hbv_oracle as O
    xs = {'x_phy': x_dict['x_phy'][:, :n].detach().cpu(), 'ac_all': x_dict['ac_all'][:n].detach().cpu(),
          'elev_all': x_dict['elev_all'][:n].detach().cpu()}
    q0 = p0[:, :n].detach().cpu().clone().requires_grad_(True)
    q1 = p1[:n].detach().cpu().clone().requires_grad_(True)
    kw = dict(nmul=nmul, dynamic_params=list(dyn_names), routing=False, use_distr_routing=False)
    ref, _ = O.forward_split(model, xs, [q0, q1], **kw)
    ref[loss_key].sum().backward()
    with torch.no_grad():
        ref64, _ = O.forward_split(model, xs, [q0.detach(), q1.detach()], dtype=torch.float64, **kw)
    for k, v in ref.items():
        if k in out and out[k] is not None and out[k].dim() == 3:
            if k not in RTOL_BY_KEY:
                assert_close(out[k][:, :n], v, RTOL_FLUX, f'{what} vs oracle:{k}')
            arbiter_gate(out[k][:, :n], v, ref64[k], f'{what} arbiter:{k}')
    assert_grad_close(grads[0][:, :n], q0.grad, f'{what} vs oracle: grad dyn', nmul)
    assert_grad_close(grads[1][:n], q1.grad, f'{what} vs oracle: grad static', nmul)
