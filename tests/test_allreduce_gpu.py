"""The library's one-shot all-reduce (csrc/allreduce.cu) on ONE GPU: two "ranks" are two comm
buffers on the same device and two kernels on two streams — the protocol (peer stores, release /
acquire flags, parity double-buffering, the in-buffer step counter) is the same as across NVLink;
the multi-process form (symmetric memory, torchrun) is exercised by bench.py --gpus N."""

import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_oneshot_allreduce_two_ranks_on_one_gpu():
    from hydrodl2_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device('cuda:0')
    world, n = 2, 210
    nfl = int(lib.hbv_b200_allreduce_buffer_floats(world, n))
    bufs = [torch.zeros(nfl, device=dev) for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    torch.cuda.synchronize()
    g = torch.Generator(device=dev).manual_seed(5)
    for step in range(7):                      # odd and even parities, the slots are reused
        xs = [torch.randn(n, generator=g, device=dev) for _ in range(world)]
        want = xs[0] + xs[1]
        outs = [x.clone() for x in xs]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                _cabi.check(lib.hbv_b200_oneshot_allreduce(ptrs.data_ptr(), r, world, outs[r].data_ptr(),
                                                           outs[r].data_ptr(), n, streams[r].cuda_stream), 'allreduce')
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(outs[r], want), f'step {step} rank {r}'
        assert torch.equal(outs[0], outs[1])
    for b in bufs:                              # step counter advanced, no timeout recorded
        tail = b[-2:].view(torch.int32).tolist()
        assert tail == [7, 0], tail


def test_oneshot_allreduce_single_rank_is_identity():
    from hydrodl2_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device('cuda:0')
    n = 33
    buf = torch.zeros(int(lib.hbv_b200_allreduce_buffer_floats(1, n)), device=dev)
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=dev)
    x = torch.arange(n, dtype=torch.float32, device=dev)
    y = torch.empty_like(x)
    st = torch.cuda.current_stream(dev).cuda_stream
    _cabi.check(lib.hbv_b200_oneshot_allreduce(ptrs.data_ptr(), 0, 1, x.data_ptr(), y.data_ptr(), n, st), 'allreduce')
    torch.cuda.synchronize()
    assert torch.equal(x, y)
    assert lib.hbv_b200_oneshot_allreduce(None, 0, 1, x.data_ptr(), y.data_ptr(), n, st) < 0
