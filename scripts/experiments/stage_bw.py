#!/usr/bin/env python
"""PCIe bandwidth of the column-block staging strategies at BASELINE config 2's shape
(parameters [1095, 531, 210]: two 64-byte blocks per 840-byte row, rows 365..1094):
  whole   plain async copy of the whole tensor (round 1)
  kernel  hbv_b200_copy_cols — the GPU reads / writes the pinned host tensor directly
  dma     hbv_b200_memcpy2d — cudaMemcpy2DAsync through the copy engine
  packed  gather on the host into a contiguous pinned buffer (CPU time counted) + 1-D copy
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from hydrodl2_b200 import _cabi as A  # noqa: E402


def main():
    lib = A.load()
    dev = torch.device('cuda:0')
    T, B, ncol, warm = 1095, 531, 210, 365
    host = torch.randn(T, B, ncol).pin_memory()
    d = torch.zeros(T, B, ncol, device=dev)
    blocks = [(0, 16), (192, 16)]
    rows = (T - warm) * B
    off = warm * B * ncol * 4
    st = torch.cuda.current_stream(dev)
    res = {}

    def timed(fn, n=10):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    nbytes = rows * 32 * 4
    for name, h2d in (('h2d', True), ('d2h', False)):
        def whole():
            (d.copy_(host, non_blocking=True) if h2d else host.copy_(d, non_blocking=True))

        def kernel():
            for c0, n in blocks:
                dst, src = (d, host) if h2d else (host, d)
                A.check(lib.hbv_b200_copy_cols(dst.data_ptr() + off, src.data_ptr() + off, rows, ncol, c0, n, st.cuda_stream), 'cc')

        def dma():
            for c0, n in blocks:
                dst, src = (d, host) if h2d else (host, d)
                A.check(lib.hbv_b200_memcpy2d(dst.data_ptr() + off, src.data_ptr() + off, rows, ncol, c0, n, 1 if h2d else 2, st.cuda_stream), 'm2d')

        ms = timed(whole, 5)
        res[f'{name}_whole'] = {'ms': ms, 'GBps': host.numel() * 4 / ms / 1e6}
        ms = timed(kernel)
        res[f'{name}_kernel'] = {'ms': ms, 'GBps_useful': nbytes / ms / 1e6}
        ms = timed(dma, 3)
        res[f'{name}_dma2d'] = {'ms': ms, 'GBps_useful': nbytes / ms / 1e6}
    # host-side pack (CPU gather) + contiguous copy
    packed = torch.empty(T - warm, B, 32).pin_memory()
    t0 = time.perf_counter()
    for _ in range(3):
        packed[..., :16].copy_(host[warm:, :, 0:16])
        packed[..., 16:].copy_(host[warm:, :, 192:208])
    res['host_pack_ms'] = (time.perf_counter() - t0) / 3 * 1e3
    dp = torch.empty_like(packed, device=dev)
    ms = timed(lambda: dp.copy_(packed, non_blocking=True))
    res['h2d_packed_copy'] = {'ms': ms, 'GBps': nbytes / ms / 1e6}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
