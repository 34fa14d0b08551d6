"""Print the measured parity margins (max-norm relative error) of the CUDA path against the
golden vectors of the unmodified reference.  GPU box:  python scripts/parity_report.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden, maxnorm_err  # noqa: E402
import test_parity_gpu as T  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    worst = {'flux': 0.0, 'state': 0.0, 'grad': 0.0}
    for case in T.PACKED:
        g = load_golden(case)
        m, out, p = T._run_packed(g, dev)
        ef = max(maxnorm_err(out[k], ref) for k, ref in g['out'].items())
        kf = max(g['out'], key=lambda k: maxnorm_err(out[k], g['out'][k]))
        es = max(maxnorm_err(s, g['states'][n]) for n, s in zip(m.state_names, m.get_states()))
        loss = sum((out[k] * c.to(dev)).sum() for k, c in g['cot'].items())
        loss.backward()
        eg = maxnorm_err(p.grad, g['grad_parameters'])
        worst['flux'] = max(worst['flux'], ef)
        worst['state'] = max(worst['state'], es)
        worst['grad'] = max(worst['grad'], eg)
        print(f'{case:22s} flux {ef:.2e} ({kf})  state {es:.2e}  grad {eg:.2e}')
    for case in T.SPLIT:
        g = load_golden(case)
        m, out, params = T._run_split(g, dev)
        ef = max(maxnorm_err(out[k], ref) for k, ref in g['out'].items())
        kf = max(g['out'], key=lambda k: maxnorm_err(out[k], g['out'][k]))
        es = max(maxnorm_err(s, g['series'][n]) for n, s in zip(m.state_names, m._state_cache))
        loss = sum((out[k] * c.to(dev)).sum() for k, c in g['cot'].items())
        loss.backward()
        eg = max(maxnorm_err(p.grad, g['grad'][f'p{i}']) for i, p in enumerate(params))
        worst['flux'] = max(worst['flux'], ef)
        worst['state'] = max(worst['state'], es)
        worst['grad'] = max(worst['grad'], eg)
        print(f'{case:22s} flux {ef:.2e} ({kf})  state {es:.2e}  grad {eg:.2e}')
    print('worst', {k: f'{v:.2e}' for k, v in worst.items()}, ' tolerances: flux/state 1e-5, grad 1e-4')


if __name__ == '__main__':
    main()
