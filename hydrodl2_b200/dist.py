"""Multi-GPU plumbing: basins shard, nothing else does.

Every (basin, component) lane of the recurrence and every basin of the UH convolution is
independent, so the basin axis is partitioned contiguously over the ranks with NO data-path
collective.  The only exchange in a training step is the all-reduce (sum) of gradients of
parameters that are *shared* across basins (the parameter network of δMG; in `bench.py` a
shared per-column bias on the raw parameters) — tiny and latency-bound.
One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(n_basins: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of the basin axis: ranks < n % world get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f'bad rank/world {rank}/{world}')
    base, rem = divmod(n_basins, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_gages(outlet_topo: torch.Tensor, world: int) -> list[tuple[torch.Tensor, torch.Tensor]]:
    """Partition the pair routing of `Hbv_2_hourly` (hbv_2_hourly.py:800-850) so that it needs no
    exchange: `outlet_topo[g, u] != 0` says unit u drains to gage g, and a unit may drain to
    several (nested) gages, so the atoms that can be placed independently are the connected
    components of that bipartite graph.  Components are assigned to ranks largest first, each to
    the rank with the fewest units so far.  Returns, per rank, (gage indices, unit indices), both
    sorted; rank r then runs the model on `x[:, units_r]`, `outlet_topo[gages_r][:, units_r]`."""
    if world <= 0:
        raise ValueError(f'bad world size {world}')
    topo = (outlet_topo != 0).cpu()
    n_g, n_u = topo.shape
    parent = list(range(n_g + n_u))          # union-find over gages [0, n_g) and units [n_g, n_g + n_u)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for g, u in topo.nonzero().tolist():
        ra, rb = find(g), find(n_g + u)
        if ra != rb:
            parent[rb] = ra
    comps: dict[int, tuple[list, list]] = {}
    for g in range(n_g):
        comps.setdefault(find(g), ([], []))[0].append(g)
    for u in range(n_u):
        comps.setdefault(find(n_g + u), ([], []))[1].append(u)
    order = sorted(comps.values(), key=lambda c: (-len(c[1]), c[0][:1], c[1][:1]))
    out = [([], []) for _ in range(world)]
    for gs, us in order:
        r = min(range(world), key=lambda k: (len(out[k][1]), k))
        out[r][0].extend(gs)
        out[r][1].extend(us)
    return [(torch.tensor(sorted(g), dtype=torch.long), torch.tensor(sorted(u), dtype=torch.long))
            for g, u in out]


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the process group
    when WORLD_SIZE > 1."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def bind_to_gpu_numa_node(device_index: int) -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (from sysfs), so pinned
    host buffers are first-touched in memory local to that GPU's PCIe root — matters for the
    host <-> device legs when several ranks share a two-socket host.  Best effort; always returns
    a one-line description: what was done, or 'unchanged (<why>)'."""
    try:
        import pynvml as nv
    except Exception as exc:
        return f'unchanged (pynvml not importable: {type(exc).__name__})'
    try:
        nv.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = device_index
        if vis:
            tok = vis.split(',')[device_index]
            if not tok.strip().isdigit():
                return f'unchanged (CUDA_VISIBLE_DEVICES entry {tok!r} is not an index)'
            phys = int(tok)
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(':')[0]) == 8:          # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        base = f'/sys/bus/pci/devices/{bus}'
        if not os.path.exists(f'{base}/numa_node'):
            return f'unchanged ({base}/numa_node not present: no sysfs PCI topology in this container)'
        node = int(open(f'{base}/numa_node').read())
        if node < 0:
            return 'unchanged (sysfs reports numa_node = -1: single-node host or virtualised topology)'
        cpus = set()
        for part in open(f'{base}/local_cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        both = cpus & allowed
        if not both:
            return (f'unchanged (numa node {node}: none of its {len(cpus)} cpus is in this process\'s '
                    f'affinity mask of {len(allowed)})')
        if both == allowed:
            return f'unchanged (already confined to numa node {node}: {len(allowed)} cpus)'
        os.sched_setaffinity(0, both)
        return f'numa node {node}, {len(both)} cpus'
    except Exception as exc:
        return f'unchanged ({type(exc).__name__}: {exc})'


class OneShotAllReduce:
    """Sum of a small float32 vector over the ranks of one NVLink / NVSwitch box with the library's
    own kernel (csrc/allreduce.cu) instead of NCCL: every rank stores its vector into every peer's
    symmetric buffer and sums what it received — one launch, a few microseconds, capturable in the
    CUDA graph of a training step.  The symmetric buffer comes from
    torch.distributed._symmetric_memory (plumbing: allocation + the exchange of peer addresses).
    `available()` is False where that is not possible; callers then keep NCCL."""

    def __init__(self, n: int, device: torch.device):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _cabi as A
        self.lib, self.C, self.n, self.dev = A.load(), C, int(n), device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        nfl = int(self.lib.hbv_b200_allreduce_buffer_floats(self.world, self.n))
        self.buf = symm.empty(nfl, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        self.ptrs = int(self.hdl.buffer_ptrs_dev)
        torch.cuda.synchronize(device)
        dist.barrier()                    # every rank's flags are zero before anyone writes

    def __call__(self, g: torch.Tensor) -> torch.Tensor:
        """In place, like dist.all_reduce; g: contiguous float32 of `n` elements on the device."""
        from . import _cabi as A
        assert g.is_contiguous() and g.dtype == torch.float32 and g.numel() == self.n
        A.check(self.lib.hbv_b200_oneshot_allreduce(self.ptrs, self.rank, self.world, g.data_ptr(), g.data_ptr(),
                                                    self.n, torch.cuda.current_stream(self.dev).cuda_stream),
                'oneshot_allreduce')
        return g

    def timed_out(self) -> bool:
        """True if any call gave up waiting for a peer (results are then invalid)."""
        err = self.buf[-1:].view(torch.int32)
        return bool(int(err.item()) != 0)


_ONESHOT: dict = {}


def enable_oneshot_allreduce(n: int, device: torch.device) -> bool:
    """Route `allreduce_shared_grad` of n-element vectors through OneShotAllReduce (all ranks must
    call this together).  Returns False — and NCCL stays — if symmetric memory cannot be set up."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return False
    ok = 1
    try:
        obj = OneShotAllReduce(n, device)
    except Exception as exc:     # pragma: no cover - depends on driver / fabric support
        print(f'hydrodl2_b200.dist: one-shot all-reduce unavailable ({type(exc).__name__}: {exc}); using NCCL',
              flush=True)
        obj, ok = None, 0
    flag = torch.tensor([ok], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # all ranks or none
    if int(flag.item()) == 1:
        _ONESHOT[int(n)] = obj
        return True
    return False


def allreduce_shared_grad(g: torch.Tensor) -> torch.Tensor:
    """Sum a shared-parameter gradient over ranks (in place); no-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        one = _ONESHOT.get(g.numel()) if g.is_cuda else None
        if one is not None:
            return one(g)
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
    return g


def max_over_ranks(value: float, device) -> float:
    """Max of a host scalar over ranks (timing: the slowest rank defines the step)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
