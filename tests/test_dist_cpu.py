"""CPU (gloo, world_size 2): the N>1 path — basin sharding with no data-path collective and the
single all-reduce of a shared-parameter gradient (hydrodl2_b200/dist.py).
Each rank runs the oracle on its basin shard; sharded fluxes concatenated == unsharded fluxes
and the all-reduced shared gradient == the unsharded one."""

import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hydrodl2_b200.dist import allreduce_shared_grad, max_over_ranks, shard_bounds


def test_shard_bounds_partition():
    for n in (1, 2, 7, 531, 180000):
        for world in (1, 2, 3, 8):
            parts = [shard_bounds(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from oracle import hbv_oracle as O
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    T, B, nmul, warm = 40, 10, 4, 8
    dyn = ['parBETA', 'parBETAET']
    x = O.synthetic_forcing(T, B, seed=3)
    p = torch.randn(T, B, 13 * nmul + 2, generator=torch.Generator().manual_seed(4))
    lo, hi = shard_bounds(B, rank, world)
    ps = p[:, lo:hi].clone().requires_grad_(True)
    out, _ = O.forward_packed('hbv', x[:, lo:hi], ps, nmul=nmul, warm_up=warm, dynamic_params=dyn)
    out['streamflow'].sum().backward()
    gshared = ps.grad[-1].sum(dim=0)
    allreduce_shared_grad(gshared)
    slow = max_over_ranks(float(rank + 1), torch.device('cpu'))
    torch.save({'q': out['streamflow'].detach(), 'g': gshared, 'lo': lo, 'hi': hi, 'slow': slow},
               os.path.join(outdir, f'r{rank}.pt'))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from oracle import hbv_oracle as O
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(world, _free_port(), td), nprocs=world, join=True)
        parts = [torch.load(os.path.join(td, f'r{r}.pt')) for r in range(world)]
    T, B, nmul, warm = 40, 10, 4, 8
    x = O.synthetic_forcing(T, B, seed=3)
    p = torch.randn(T, B, 13 * nmul + 2, generator=torch.Generator().manual_seed(4)).requires_grad_(True)
    out, _ = O.forward_packed('hbv', x, p, nmul=nmul, warm_up=warm,
                              dynamic_params=['parBETA', 'parBETAET'])
    out['streamflow'].sum().backward()
    q = torch.cat([pt['q'] for pt in parts], dim=1)
    # no cross-basin coupling at all (differences are vectorisation-order rounding only)
    assert torch.allclose(q, out['streamflow'].detach(), rtol=1e-6, atol=1e-8)
    g_ref = p.grad[-1].sum(dim=0)
    for pt in parts:
        assert torch.allclose(pt['g'], g_ref, rtol=1e-5, atol=1e-7)
        assert pt['slow'] == float(world)
    assert parts[0]['lo'] == 0 and parts[-1]['hi'] == B


def test_shard_gages_keeps_nested_gages_together():
    """Units that drain to several (nested) gages stay on the rank of all of those gages; every
    gage and unit is placed exactly once; the unit counts are balanced."""
    import torch
    from hydrodl2_b200.dist import shard_gages
    g = torch.Generator().manual_seed(3)
    n_g, n_u = 12, 96
    topo = torch.zeros(n_g, n_u)
    for k in range(6):                       # six independent river systems of 16 units ...
        topo[2 * k, 16 * k:16 * k + 16] = 1  # ... each with an outlet gage
        topo[2 * k + 1, 16 * k:16 * k + int(torch.randint(4, 12, (1,), generator=g))] = 1  # and a nested one
    parts = shard_gages(topo, 4)
    all_g = torch.cat([p[0] for p in parts])
    all_u = torch.cat([p[1] for p in parts])
    assert sorted(all_g.tolist()) == list(range(n_g)) and sorted(all_u.tolist()) == list(range(n_u))
    for gs, us in parts:
        sub = topo[gs]
        inside = torch.zeros(n_u, dtype=torch.bool)
        inside[us] = True
        assert not sub[:, ~inside].any()     # no gage of this rank needs a unit of another rank
        assert (topo[:, us].sum(0) == sub[:, us].sum(0)).all()   # and none of its units feeds a foreign gage
    sizes = sorted(len(p[1]) for p in parts)
    assert sizes == [16, 16, 32, 32]


def _allreduce_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from hydrodl2_b200.dpl import allreduce_gradients
    lin = torch.nn.Linear(3, 2)
    lin.weight.grad = torch.full((2, 3), float(rank + 1))
    lin.bias.grad = torch.arange(2, dtype=torch.float32) * (rank + 1)
    n = allreduce_gradients(lin.parameters(), average=True)
    q.put((rank, n, lin.weight.grad.clone(), lin.bias.grad.clone()))
    dist.destroy_process_group()


def test_allreduce_gradients_gloo_world2():
    """dpl.allreduce_gradients: one flat collective, averaged, written back in place."""
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, 29541, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, n, w, b in res:
        assert n == 8
        assert torch.allclose(w, torch.full((2, 3), 1.5))
        assert torch.allclose(b, torch.tensor([0.0, 1.5]))
