"""CUDA-graph capture of a whole training / inference step.

The 531-basin configurations are launch- and host-bound: one `Hbv.forward` + `backward` is ~20
kernel launches and ~1 ms of Python for ~1.1 ms of GPU work.  Every launch of this package goes
to `torch.cuda.current_stream()` through the C-ABI, allocates nothing itself and never
synchronises, so the whole step (warm-up run, recurrence, routing, adjoint, the side-stream
gradient memset, and an NCCL all-reduce if there is one) can be captured once and replayed.

    step = GraphedStep(lambda: train_step(model, x_static, p_static))
    out, loss = step.outputs          # static tensors, refreshed by every replay
    step.replay()

Measured on B200 (531 basins, 0.55 ms of kernels per training step): replay 0.55 ms; the eager
step 0.57-0.71 ms depending on the box's host CPU (one eager step costs the host about as much as
the kernels take).  `bench.py` times the replay; with several GPUs the NCCL all-reduce of the
shared gradient stays outside the graph and runs eagerly on the graph's static output after each
replay (capturing a step that contains the collective hung on 2 GPUs): 0.565 ms per step on 2 GPUs.

Inputs must live in static tensors (copy new data into them before `replay()`); the dynamic-
parameter dropout draw (`dy_drop > 0`, a CPU RNG draw per forward, hbv.py:240-246) would be
frozen into the graph, so capture only with `dy_drop == 0` or for inference.
"""

from __future__ import annotations

from typing import Any, Callable

import torch


class GraphedStep:
    def __init__(self, fn: Callable[[], Any], warmup: int = 3, device=None, capture_error_mode: str = 'global'):
        """capture_error_mode='thread_local' is what a step containing an NCCL collective needs:
        NCCL's watchdog thread issues CUDA calls of its own while this thread captures, which the
        default 'global' mode turns into a capture error."""
        if not torch.cuda.is_available():
            raise RuntimeError('hydrodl2_b200.graphs: CUDA is required (no CPU path)')
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):      # warm-up off the default stream, as capture requires
            for _ in range(max(1, warmup)):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.outputs = fn()

    def replay(self):
        self.graph.replay()
        return self.outputs
