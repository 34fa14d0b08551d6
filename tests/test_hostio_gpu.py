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
