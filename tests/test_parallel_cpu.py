"""CPU, world_size 2 over gloo: row sharding and the gradient / loss all-reduce of the
data-parallel NLL step equal the single-process full-batch result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stribor_b200.parallel import DataParallelNLL, allreduce_gradients, gather_rows, shard_rows


def test_shard_rows_partition():
    for n in (0, 1, 7, 8, 1 << 22, (1 << 22) + 5):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


class _ToyFlow(torch.nn.Module):
    """Differentiable stand-in with the same calling convention (rows -> log-density [n, 1])."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 6))

    def log_prob(self, y):
        s = self.net(y)
        z = y * torch.exp(s) + s.flip(-1)
        return (-(z ** 2) / 2 - 0.9189385332046727).sum(-1, keepdim=True) + s.sum(-1, keepdim=True)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(11)
    y = torch.randn(n_rows, 6)
    a, b = shard_rows(n_rows, rank, world)
    flow = _ToyFlow()
    dp = DataParallelNLL(flow, micro_rows=5)
    loss = dp.step(y[a:b], n_rows)
    with torch.no_grad():
        lp_full = gather_rows(flow.log_prob(y[a:b]), n_rows)
    q.put((rank, loss.item(), [p.grad.clone() for p in flow.parameters()], lp_full, dp.last_allreduce_bytes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_data_parallel_nll_equals_full_batch():
    world, n_rows = 2, 37                     # ragged: 19 + 18 rows, micro-batches of 5
    ctx = mp.get_context('spawn')

    def run_world():
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            return [q.get(timeout=50) for _ in range(world)]
        finally:
            for p in procs:
                p.join(timeout=20)
                if p.is_alive():              # a rank that never rendezvoused (port taken in between): do not leak it
                    p.terminate()
                    p.join(timeout=10)

    results = None
    for attempt in range(2):                  # the free port is released before the ranks bind it: one retry
        try:
            results = run_world()
            break
        except Exception:
            if attempt == 1:
                raise

    torch.manual_seed(11)
    y = torch.randn(n_rows, 6)
    ref = _ToyFlow()
    loss = -ref.log_prob(y).mean()
    loss.backward()
    n_param_bytes = sum(p.numel() for p in ref.parameters()) * 4
    for rank, l, grads, lp_full, nbytes in results:
        assert abs(l - loss.item()) < 1e-5
        assert nbytes == n_param_bytes
        for g, p in zip(grads, ref.parameters()):
            torch.testing.assert_close(g, p.grad, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(lp_full, ref.log_prob(y).detach(), rtol=1e-6, atol=1e-6)


def test_single_process_is_a_noop():
    f = _ToyFlow()
    y = torch.randn(9, 6)
    loss = DataParallelNLL(f, micro_rows=4).step(y, 9)
    ref = _ToyFlow()
    l2 = -ref.log_prob(y).mean()
    l2.backward()
    assert abs(loss.item() - l2.item()) < 1e-6
    assert allreduce_gradients(f.parameters()) == 0
    for a, b in zip(f.parameters(), ref.parameters()):
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-6)
