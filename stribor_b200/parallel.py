"""Data-parallel plumbing for the coupling-flow path: one process per GPU, `torch.distributed`.

The path shards by ROWS (every op is row-independent):
* inference (`log_prob`, `inverse`, `sample`): rank r evaluates rows [start_r, stop_r) of the batch;
  no collective is needed (an optional all-gather returns the whole vector);
* NLL training: each rank back-propagates  -sum(log_prob(shard)) / global_rows  over micro-batches,
  then ONE flat all-reduce (sum) of the gradients and one of the scalar loss over NCCL/NVLink.
  The result equals the single-process full-batch gradient up to summation order.
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rows owned by `rank`; the first (n_rows % world) ranks get one extra."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def init_from_env(backend: Optional[str] = None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT
    (torchrun).  Returns (rank, world, local_rank).  No-op for a single process."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this process (and therefore the pinned host buffers it allocates afterwards: first touch) to the CPU
    cores of the NUMA node the GPU hangs off.  With 8 ranks pulling rows over PCIe at once, host buffers on the far
    socket halve the per-GPU copy rate.  No-op when the platform does not expose the topology (returns why)."""
    info = {'bound': False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = None
        if all(hasattr(pr, a) for a in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
            bdf = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        if bdf is None:
            import subprocess
            out = subprocess.run(['nvidia-smi', '--query-gpu=pci.bus_id', '--format=csv,noheader', '-i', str(local_rank)],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout.strip()
            bdf = out.splitlines()[0].strip() if out else None
        if not bdf:
            info['why'] = 'no PCI bus id'
            return info
        bdf = bdf.lower()
        if len(bdf.split(':')[0]) == 8:                      # nvidia-smi prints an 8-digit domain, sysfs has 4
            bdf = bdf[4:]
        node_path = f'/sys/bus/pci/devices/{bdf}/numa_node'
        node = int(open(node_path).read().strip()) if os.path.exists(node_path) else -1
        info['pci'] = bdf
        info['numa_node'] = node
        if node < 0:
            info['why'] = 'numa_node not exposed'
            return info
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            info['why'] = 'no allowed cores on that node'
            return info
        os.sched_setaffinity(0, cpus)
        info.update({'bound': True, 'cores': len(cpus)})
    except Exception as e:                                   # topology probing must never break a run
        info['why'] = f'{type(e).__name__}: {e}'
    return info


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Sum the gradients of `params` over the ranks with ONE flat all-reduce.  Returns the number
    of bytes reduced (0 for a single process)."""
    grads = [p.grad for p in params if p.grad is not None]
    if _world() == 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()


def gather_rows(local: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """All-gather row shards produced with `shard_rows` back into the full [n_rows, ...] tensor."""
    world = _world()
    if world == 1:
        return local
    rank = dist.get_rank()
    sizes = [shard_rows(n_rows, r, world) for r in range(world)]
    m = max(b - a for a, b in sizes)
    pad = local.new_zeros((m,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], 0)


class DataParallelNLL:
    """NLL training step of a flow, batch sharded over the ranks.

        dp = DataParallelNLL(flow, micro_rows=1 << 18)
        loss = dp.step(y_shard, global_rows)     # grads are in p.grad, already all-reduced
        optimizer.step()

    `log_prob_fn(rows) -> [n, 1]` defaults to `flow.log_prob`; it is a parameter so the host logic
    can be exercised on CPU (gloo) with any differentiable stand-in.
    """

    def __init__(self, flow: torch.nn.Module, micro_rows: int = 1 << 18,
                 log_prob_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, group=None):
        self.flow = flow
        self.micro_rows = int(micro_rows)
        self.log_prob_fn = log_prob_fn if log_prob_fn is not None else flow.log_prob
        self.group = group
        self.last_allreduce_bytes = 0

    def step(self, y_shard: torch.Tensor, global_rows: int, zero_grad: bool = True) -> torch.Tensor:
        if zero_grad:
            for p in self.flow.parameters():
                p.grad = None
        total = y_shard.new_zeros(())
        for r0 in range(0, y_shard.shape[0], self.micro_rows):
            lp = self.log_prob_fn(y_shard[r0:r0 + self.micro_rows])
            loss = -(lp.sum() / global_rows)             # this micro-batch's share of the global mean
            loss.backward()
            total += loss.detach()
        self.last_allreduce_bytes = allreduce_gradients(self.flow.parameters(), self.group)
        if _world() > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return total
