"""hostio.PipelinedSteps: overlapped upload / step / download gives the same results, step by
step, as the plain serial loop (different inputs every step, so a stale buffer would show)."""

import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu


def test_pipelined_steps_match_serial():
    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200.hostio import PipelinedSteps
    from oracle import hbv_oracle as O
    dev = torch.device('cuda:0')
    T, B, nmul = 40, 48, 16
    dyn = ['parBETA', 'parBETAET']
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': 8, 'dynamic_params': {'Hbv': dyn}, 'nmul': nmul}, device=dev)
    batches = []
    for s in range(5):
        x = O.synthetic_forcing(T, B, seed=100 + s).pin_memory()
        p = torch.randn(T, B, 13 * nmul + 2, generator=torch.Generator().manual_seed(200 + s)).pin_memory()
        batches.append({'x_phy': x, 'parameters': p})

    def step(inp):
        inp['parameters'].grad = None
        out = m({'x_phy': inp['x_phy']}, inp['parameters'])
        loss = out['streamflow'].sum()
        loss.backward()
        return {'streamflow': out['streamflow'], 'loss': loss, 'grad': inp['parameters'].grad}

    serial = []
    for bt in batches:
        inp = {'x_phy': bt['x_phy'].to(dev), 'parameters': bt['parameters'].to(dev).requires_grad_(True)}
        serial.append({k: v.detach().cpu() for k, v in step(inp).items()})

    pipe = PipelinedSteps(step, batches[0], dev, leaf_names=('parameters',))
    got = []
    for bt in batches:
        hb = pipe.step(bt)
        # a host buffer set is reused every `depth` steps: read it once its download is over
        pipe.ev_out[(pipe.i - 1) % pipe.depth].synchronize()
        got.append({k: v.clone() for k, v in hb.items()})
    pipe.drain()
    for i, (a, b) in enumerate(zip(got, serial)):
        for k in b:
            assert_close(a[k], b[k], 1e-7, f'step {i}: {k}')
    # free-running (no host wait between steps): the last `depth` steps' buffers are still intact
    pipe2 = PipelinedSteps(step, batches[0], dev, leaf_names=('parameters',))
    hbs = [pipe2.step(bt) for bt in batches]
    pipe2.drain()
    for i in (len(batches) - 2, len(batches) - 1):
        for k in serial[i]:
            assert_close(hbs[i][k], serial[i][k], 1e-7, f'free-running step {i}: {k}')


@pytest.mark.parametrize('block_copy', ['auto', 'dma', 'kernel'])
def test_sparse_staging_gives_the_dense_gradient(block_copy, monkeypatch):
    """Column-sparse staging (hostio.sparse_copy + Model.io_footprint): uploading only the entries
    of `parameters` the kernels read and downloading only the non-zero part of the gradient gives,
    on the host, bit for bit the dense gradient of the whole-tensor loop — with different
    parameters every step and garbage in the entries that are never uploaded."""
    import hydrodl2_b200 as hydrodl2
    from hydrodl2_b200 import hostio
    from hydrodl2_b200.hostio import PipelinedSteps
    from oracle import hbv_oracle as O
    monkeypatch.setattr(hostio, 'BLOCK_COPY', block_copy)     # copy engine / SM-driven / timed pick
    dev = torch.device('cuda:0')
    T, B, nmul, warm = 40, 50, 16, 8
    dyn = ['parBETA', 'parBETAET']
    M = hydrodl2.load_model('hbv', ver_name='Hbv')
    m = M({'warm_up': warm, 'dynamic_params': {'Hbv': dyn}, 'nmul': nmul}, device=dev)
    batches = []
    for s in range(4):
        x = O.synthetic_forcing(T, B, seed=300 + s).pin_memory()
        p = torch.randn(T, B, 13 * nmul + 2, generator=torch.Generator().manual_seed(400 + s)).pin_memory()
        batches.append({'x_phy': x, 'parameters': p})

    def step(inp):
        inp['parameters'].grad = None
        out = m({'x_phy': inp['x_phy']}, inp['parameters'])
        loss = out['streamflow'].sum()
        loss.backward()
        return {'streamflow': out['streamflow'], 'loss': loss, 'grad': inp['parameters'].grad}

    dense = []
    for bt in batches:
        inp = {'x_phy': bt['x_phy'].to(dev), 'parameters': bt['parameters'].to(dev).requires_grad_(True)}
        dense.append({k: v.detach().cpu() for k, v in step(inp).items()})

    fp = m.io_footprint(T)
    pipe = PipelinedSteps(step, batches[0], dev, leaf_names=('parameters',),
                          in_footprints={'parameters': fp['read']}, out_footprints={'grad': fp['grad']})
    for d in pipe.dev_in:                       # entries outside the footprint must never matter
        with torch.no_grad():
            d['parameters'].fill_(float('nan'))
    for i, bt in enumerate(batches):
        hb = pipe.step(bt)
        pipe.ev_out[(pipe.i - 1) % pipe.depth].synchronize()
        for k in dense[i]:
            assert torch.equal(hb[k], dense[i][k]), f'sparse staging step {i}: {k} differs from the dense loop'
    pipe.drain()
    assert set(pipe.block_copy) == {('in', 'parameters'), ('out', 'grad')}
    assert all(m in ('dma', 'kernel') and (block_copy == 'auto' or m == block_copy) for m in pipe.block_copy.values())
    full = 2 * T * B * (13 * nmul + 2) * 4
    assert pipe.h2d_bytes + pipe.d2h_bytes < 0.35 * full
    assert pipe.h2d_bytes == (T * B * 3 + (2 * B * (13 * nmul + 2)) + (T - warm - 1) * B * 2 * nmul) * 4


def test_fill_zero_and_copy_cols():
    """C-ABI helpers of csrc/fill.cu: thin TMA zero fill (unaligned head / tail included) and the
    column-block copy between pinned host memory and the device, both directions."""
    from hydrodl2_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device('cuda:0')
    st = torch.cuda.current_stream(dev).cuda_stream
    buf = torch.full((1_000_003,), 7.0, device=dev)
    view = buf[1:-2]                                    # 4 B-aligned start, odd length
    _cabi.check(lib.hbv_b200_fill_zero(view.data_ptr(), view.numel() * 4, 0, st), 'fill_zero')
    torch.cuda.synchronize()
    assert float(buf[0]) == 7.0 and float(buf[-1]) == 7.0 and float(buf[-2]) == 7.0
    assert int((view != 0).sum()) == 0
    rows, ncol = 777, 210
    host = torch.randn(rows, ncol).pin_memory()
    d = torch.zeros(rows, ncol, device=dev)
    for c0, n in ((0, 16), (192, 16), (205, 3)):
        _cabi.check(lib.hbv_b200_copy_cols(d.data_ptr(), host.data_ptr(), rows, ncol, c0, n, st), 'copy_cols h2d')
    torch.cuda.synchronize()
    ref = torch.zeros(rows, ncol)
    for c0, n in ((0, 16), (192, 16), (205, 3)):
        ref[:, c0:c0 + n] = host[:, c0:c0 + n]
    assert torch.equal(d.cpu(), ref)
    back = torch.zeros(rows, ncol).pin_memory()
    _cabi.check(lib.hbv_b200_copy_cols(back.data_ptr(), d.data_ptr(), rows, ncol, 192, 16, st), 'copy_cols d2h')
    torch.cuda.synchronize()
    assert torch.equal(back[:, 192:208], host[:, 192:208]) and int((back[:, :192] != 0).sum()) == 0
