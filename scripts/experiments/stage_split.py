#!/usr/bin/env python
"""Do several copy streams raise the column-block staging rate?  BASELINE config 2's shape
(two 64-byte blocks per 840-byte row, 387,630 rows, 49.6 MB): cudaMemcpy2DAsync (copy engine) and
hbv_b200_copy_cols (SM-driven) with the work split over 1 / 2 / 4 / 8 streams (by block, then by
row range), each direction alone and both at once."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from hydrodl2_b200 import _cabi as A  # noqa: E402


def main():
    lib = A.load()
    dev = torch.device('cuda:0')
    T, B, ncol, warm = 1095, 531, 210, 365
    host = torch.randn(T, B, ncol).pin_memory()
    hback = torch.zeros(T, B, ncol).pin_memory()
    d = torch.zeros(T, B, ncol, device=dev)
    blocks = [(0, 16), (192, 16)]
    rows = (T - warm) * B
    off0 = warm * B * ncol * 4
    nbytes = rows * 32 * 4
    streams = [torch.cuda.Stream(dev) for _ in range(8)]
    main_s = torch.cuda.current_stream(dev)
    res = {'async_engines': torch.cuda.get_device_properties(dev).multi_processor_count}

    def issue(mode, h2d, ns):
        dst, src = (d, host) if h2d else (hback, d)
        pieces = []
        nr = max(1, ns // len(blocks))
        step = (rows + nr - 1) // nr
        for c0, n in blocks:
            for r0 in range(0, rows, step):
                pieces.append((c0, n, r0, min(step, rows - r0)))
        for i, (c0, n, r0, nrw) in enumerate(pieces):
            s = streams[i % ns]
            off = off0 + r0 * ncol * 4
            if mode == 'dma':
                A.check(lib.hbv_b200_memcpy2d(dst.data_ptr() + off, src.data_ptr() + off, nrw, ncol, c0, n,
                                              1 if h2d else 2, s.cuda_stream), 'm2d')
            else:
                A.check(lib.hbv_b200_copy_cols(dst.data_ptr() + off, src.data_ptr() + off, nrw, ncol, c0, n,
                                               s.cuda_stream), 'cc')

    def timed(fn, ns, n=5):
        def run():
            for s in streams[:ns]:
                s.wait_stream(main_s)
            fn()
            for s in streams[:ns]:
                main_s.wait_stream(s)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for mode in ('dma', 'kernel'):
        for ns in (1, 2, 4, 8):
            up = timed(lambda: issue(mode, True, ns), ns)
            dn = timed(lambda: issue(mode, False, ns), ns)

            def both():
                issue(mode, True, max(1, ns // 2) if ns > 1 else 1)
                # the download on the other half of the streams
                dst, src = hback, d
                half = streams[ns // 2:ns] if ns > 1 else streams[1:2]
                for i, (c0, n) in enumerate(blocks):
                    s = half[i % len(half)]
                    fn = lib.hbv_b200_memcpy2d if mode == 'dma' else lib.hbv_b200_copy_cols
                    args = (dst.data_ptr() + off0, src.data_ptr() + off0, rows, ncol, c0, n)
                    A.check(fn(*args, 2, s.cuda_stream) if mode == 'dma' else fn(*args, s.cuda_stream), 'x')
            bt = timed(both, max(ns, 2))
            res[f'{mode}_{ns}'] = {'h2d_ms': round(up, 3), 'd2h_ms': round(dn, 3), 'both_ms': round(bt, 3),
                                   'h2d_GBps': round(nbytes / up / 1e6, 1), 'd2h_GBps': round(nbytes / dn / 1e6, 1)}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
