"""Host <-> device staging for training loops whose batches live in pinned host memory.

A training step of the hot path moves far more bytes over PCIe than the kernels take to run
(BASELINE config 2: 495 MB of `parameters` in, 490 MB of gradient out, ~1 ms of kernels), so a
loop that copies, computes and copies back in sequence is bound by the SUM of the two copy
directions.  `PipelinedSteps` keeps two device-side input sets and runs three streams — upload,
compute, download — so step i+1's upload and step i-1's download overlap step i's kernels; PCIe is
full duplex, so the loop becomes bound by the slower direction alone.  Every step still uploads
its own inputs and downloads its own results; nothing is cached between steps.

PyTorch is used for streams, events and pinned memory only.
"""

from __future__ import annotations

from typing import Callable, Dict, Sequence

import torch


class PipelinedSteps:
    """Double-buffered upload -> step -> download loop.

    step_fn(inputs: dict[str, Tensor]) -> dict[str, Tensor]
        runs on the compute stream with device inputs and returns the device tensors to download
        (e.g. {'streamflow': ..., 'loss': ..., 'grad': ...}).
    leaf_names: inputs that must be autograd leaves (`requires_grad_`), e.g. ('parameters',).
    """

    def __init__(self, step_fn: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
                 host_inputs: Dict[str, torch.Tensor], device: torch.device,
                 leaf_names: Sequence[str] = (), depth: int = 2):
        self.step_fn, self.dev, self.depth = step_fn, device, depth
        self.s_in = torch.cuda.Stream(device)
        self.s_out = torch.cuda.Stream(device)
        self.dev_in = []
        for _ in range(depth):
            d = {}
            for k, h in host_inputs.items():
                t = torch.empty(h.shape, dtype=h.dtype, device=device)
                if k in leaf_names:
                    t.requires_grad_(True)
                d[k] = t
            self.dev_in.append(d)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free = [None] * depth          # compute that last read input set k has finished
        self.ev_out = [None] * depth           # download into host set k has finished
        self.host_out = [None] * depth
        self.i = 0

    def _host_buffers(self, k, outs):
        if self.host_out[k] is None:
            self.host_out[k] = {n: torch.empty(t.shape, dtype=t.dtype).pin_memory() for n, t in outs.items()}
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
            with torch.no_grad():
                for n, h in host_inputs.items():
                    self.dev_in[k][n].copy_(h, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        comp.wait_event(self.ev_in[k])
        outs = self.step_fn(self.dev_in[k])
        ev_done = torch.cuda.Event()
        ev_done.record(comp)
        self.ev_free[k] = ev_done
        # the host buffers of this set are reused: their previous download must be over
        if self.ev_out[k] is not None:
            self.ev_out[k].synchronize()
        hb = self._host_buffers(k, outs)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_done)
            for n, t in outs.items():
                hb[n].copy_(t.detach(), non_blocking=True)
                t.record_stream(self.s_out)
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
