"""Host <-> device staging for training loops whose batches live in pinned host memory.

A training step of the hot path moves far more bytes over PCIe than the kernels take to run
(BASELINE config 2: 495 MB of `parameters` in, 490 MB of gradient out, ~1 ms of kernels), so a
loop that copies, computes and copies back in sequence is bound by the SUM of the two copy
directions.  `PipelinedSteps` keeps two device-side input sets and runs three streams — upload,
compute, download — so step i+1's upload and step i-1's download overlap step i's kernels; PCIe is
full duplex, so the loop becomes bound by the slower direction alone.  Every step still uploads
its own inputs and downloads its own results; nothing is cached between steps.

Column-sparse staging.  The kernels read only part of a packed `parameters` tensor — the column
blocks of the time-varying parameters plus the last row (static values + routing, hbv.py:201-214)
and the last warm-up row — and most of the dense gradient they return is structural zeros.  Given
the model's `io_footprint()` the loop uploads exactly those entries (whole rows with a plain async
copy, column blocks with copy-engine 2-D copies or with the library's `hbv_b200_copy_cols`, which
lets the GPU read the pinned host tensor directly over PCIe — whichever is faster on this host,
see BLOCK_COPY) and downloads only the gradient's non-zero entries into a pinned
host plane that was zeroed once — the host still ends up with the full dense gradient.  BASELINE
config 2: 985 MB -> ~135 MB over PCIe per step.

PyTorch is used for streams, events and pinned memory only.
"""

from __future__ import annotations

import os
from typing import Callable, Dict, Optional, Sequence

import torch


def _merge_blocks(blocks):
    out = []
    for c0, n in sorted(blocks):
        if out and out[-1][0] + out[-1][1] == c0:
            out[-1] = (out[-1][0], out[-1][1] + n)
        else:
            out.append((c0, n))
    return out


# How column blocks cross PCIe: 'dma' = cudaMemcpy2DAsync on a copy engine (hbv_b200_memcpy2d),
# 'kernel' = hbv_b200_copy_cols (the GPU reads / writes the pinned host tensor in place).
# Measured on B200 at BASELINE config 2's shape (two 64-byte blocks per 840-byte row, 49.6 MB;
# scripts/experiments/stage_bw.py): dma 1.88 ms up / 2.38 ms down, kernel 2.20 / 2.51 ms, the whole
# tensor 8.8 ms each way — and the copy engine leaves the SMs to the latency-bound kernels.
# The strided copy-engine rate is a property of the HOST, not of the GPU: the same code measured
# 15.9 / 14.3 GB/s (up / down) on one B200 box and 6.5 / 5.8 GB/s on another whose contiguous copies
# ran at the same 55 GB/s.  'auto' (default) therefore times both ways once per tensor and
# direction on the caller's own buffers (`PipelinedSteps`, first step) and keeps the faster one; the
# SM-driven copy has to win by 15 % to be taken, since it competes with the step's kernels.
# Upload and download do not run at full duplex for these 64-byte pieces whatever the mix: both
# directions at once take the SUM of their times (copy engine: 1.88 + 2.31 -> 3.77 ms; SM-driven:
# 2.49 + 2.49 -> 5.07 ms; more copy streams change nothing — scripts/experiments/stage_split.py), and
# the pipelined step measures 3.6 / 3.9 / 3.6 / 4.4 ms for (up, down) = (dma, dma) / (dma, kernel) /
# (kernel, dma) / (kernel, kernel).
BLOCK_COPY = os.environ.get('HBV_B200_BLOCK_COPY', 'auto')


def sparse_copy(dst: torch.Tensor, src: torch.Tensor, fp: dict, stream: torch.cuda.Stream,
                mode: Optional[str] = None) -> int:
    """Copy the entries of the [T, B, ncol] float32 tensor `src` named by footprint `fp` into
    `dst` (same shape, both contiguous; either may be pinned host memory) on `stream`.
    fp = {'rows_full': [t, ...], 'col_blocks': [(col0, ncols), ...], 'row_range': (t0, t1)}.
    mode: 'dma' / 'kernel' for the column blocks (None: BLOCK_COPY, 'auto' counting as 'dma').
    Returns the number of bytes moved."""
    if mode is None:
        mode = BLOCK_COPY if BLOCK_COPY in ('dma', 'kernel') else 'dma'
    from . import _cabi as A
    lib = A.load()
    T, B, ncol = src.shape
    assert dst.shape == src.shape and dst.is_contiguous() and src.is_contiguous()
    assert dst.dtype == torch.float32 and src.dtype == torch.float32
    moved = 0
    with torch.cuda.stream(stream):
        with torch.no_grad():
            for t in fp.get('rows_full', ()):
                dst[t].copy_(src[t], non_blocking=True)
                moved += B * ncol * 4
        t0, t1 = fp.get('row_range', (0, T))
        full = set(fp.get('rows_full', ()))
        while t1 > t0 and (t1 - 1) in full:      # a trailing full row already carries its blocks
            t1 -= 1
        if t1 > t0:
            off = t0 * B * ncol * 4
            kind = 1 if dst.is_cuda else 2
            for c0, n in _merge_blocks(fp.get('col_blocks', ())):
                if mode == 'dma' and dst.is_cuda != src.is_cuda:
                    A.check(lib.hbv_b200_memcpy2d(dst.data_ptr() + off, src.data_ptr() + off, (t1 - t0) * B, ncol,
                                                  c0, n, kind, stream.cuda_stream), 'memcpy2d')
                else:
                    A.check(lib.hbv_b200_copy_cols(dst.data_ptr() + off, src.data_ptr() + off, (t1 - t0) * B, ncol,
                                                   c0, n, stream.cuda_stream), 'copy_cols')
                moved += (t1 - t0) * B * n * 4
    return moved


def pick_block_copy(dst: torch.Tensor, src: torch.Tensor, fp: dict, stream: torch.cuda.Stream) -> str:
    """'dma' or 'kernel' for this (tensor, direction, footprint): BLOCK_COPY when it names one
    ('dma', 'kernel', or 'up,down' e.g. 'dma,kernel'), else both timed on `stream` (one warm copy +
    one timed copy each; the copies are idempotent)."""
    forced = BLOCK_COPY.split(',')
    if all(f in ('dma', 'kernel') for f in forced):
        return forced[0] if dst.is_cuda or len(forced) == 1 else forced[1]
    if not fp.get('col_blocks'):
        return 'dma'
    ms = {}
    for mode in ('dma', 'kernel'):
        sparse_copy(dst, src, fp, stream, mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sparse_copy(dst, src, fp, stream, mode)
        e1.record(stream)
        e1.synchronize()
        ms[mode] = e0.elapsed_time(e1)
    return 'kernel' if ms['kernel'] < 0.85 * ms['dma'] else 'dma'


class PipelinedSteps:
    """Double-buffered upload -> step -> download loop.

    step_fn(inputs: dict[str, Tensor]) -> dict[str, Tensor]
        runs on the compute stream with device inputs and returns the device tensors to download
        (e.g. {'streamflow': ..., 'loss': ..., 'grad': ...}).
    leaf_names: inputs that must be autograd leaves (`requires_grad_`), e.g. ('parameters',).
    in_footprints / out_footprints: optional {name: footprint} (see `sparse_copy`) — only those
        entries of the named input are uploaded / of the named output are downloaded; a sparse
        output lands in a pinned host plane that was zeroed when it was allocated.
    `h2d_bytes` / `d2h_bytes`: bytes moved by the last `step()`.
    """

    def __init__(self, step_fn: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
                 host_inputs: Dict[str, torch.Tensor], device: torch.device,
                 leaf_names: Sequence[str] = (), depth: int = 2,
                 in_footprints: Optional[Dict[str, dict]] = None,
                 out_footprints: Optional[Dict[str, dict]] = None):
        self.step_fn, self.dev, self.depth = step_fn, device, depth
        self.in_fp = dict(in_footprints or {})
        self.out_fp = dict(out_footprints or {})
        self.h2d_bytes = self.d2h_bytes = 0
        self.block_copy = {}                   # ('in' | 'out', name) -> 'dma' | 'kernel', picked at first use
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self.dev_in = []
        for _ in range(depth):
            d = {}
            for k, h in host_inputs.items():
                # (a footprinted input is only partly overwritten by the uploads: start it from zeros)
                t = (torch.zeros if k in (in_footprints or {}) else torch.empty)(h.shape, dtype=h.dtype, device=device)
                if k in leaf_names:
                    t.requires_grad_(True)
                d[k] = t
            self.dev_in.append(d)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free = [None] * depth          # compute that last read input set k has finished
        self.ev_out = [None] * depth           # download into host set k has finished
        self.host_out = [None] * depth
        self._live = [None] * depth            # device results of set k, kept until their download is over
        self.i = 0

    def _host_buffers(self, k, outs):
        if self.host_out[k] is None:
            self.host_out[k] = {n: (torch.zeros(t.shape, dtype=t.dtype) if n in self.out_fp
                                    else torch.empty(t.shape, dtype=t.dtype)).pin_memory()
                                for n, t in outs.items()}
        return self.host_out[k]

    def step(self, host_inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Enqueue one step; returns the pinned host buffers its results are being written to
        (valid after `wait(k)` / `drain()`)."""
        k = self.i % self.depth
        self.i += 1
        comp = torch.cuda.current_stream(self.dev)
        # upload on its own stream, once the compute that last used this input set is done
        with torch.cuda.stream(self.s_in):
            if self.ev_free[k] is not None:
                self.s_in.wait_event(self.ev_free[k])
            moved = 0
            with torch.no_grad():
                for n, h in host_inputs.items():
                    if n in self.in_fp:
                        if ('in', n) not in self.block_copy:
                            self.block_copy['in', n] = pick_block_copy(self.dev_in[k][n], h, self.in_fp[n], self.s_in)
                        moved += sparse_copy(self.dev_in[k][n], h, self.in_fp[n], self.s_in, self.block_copy['in', n])
                    else:
                        self.dev_in[k][n].copy_(h, non_blocking=True)
                        moved += h.numel() * h.element_size()
            self.h2d_bytes = moved
            self.ev_in[k].record(self.s_in)
        comp.wait_event(self.ev_in[k])
        # the host buffers of this set are reused, and so may be the device memory of its previous
        # results (the library hands a released, unmodified gradient plane out again without
        # re-zeroing it, ops._clean_plane): their previous download must be over before this
        # step's kernels are enqueued, and only then are the previous result tensors let go
        if self.ev_out[k] is not None:
            self.ev_out[k].synchronize()
        self._live[k] = None
        outs = self.step_fn(self.dev_in[k])
        self._live[k] = outs
        ev_done = torch.cuda.Event()
        ev_done.record(comp)
        self.ev_free[k] = ev_done
        hb = self._host_buffers(k, outs)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_done)
            moved = 0
            for n, t in outs.items():
                if n in self.out_fp:
                    if ('out', n) not in self.block_copy:
                        self.block_copy['out', n] = pick_block_copy(hb[n], t.detach(), self.out_fp[n], self.s_out)
                    moved += sparse_copy(hb[n], t.detach(), self.out_fp[n], self.s_out, self.block_copy['out', n])
                else:
                    hb[n].copy_(t.detach(), non_blocking=True)
                    moved += t.numel() * t.element_size()
                t.record_stream(self.s_out)
            self.d2h_bytes = moved
            ev = torch.cuda.Event()
            ev.record(self.s_out)
        self.ev_out[k] = ev
        return hb

    def drain(self):
        """Block until every enqueued upload, step and download has finished."""
        for ev in self.ev_out:
            if ev is not None:
                ev.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
