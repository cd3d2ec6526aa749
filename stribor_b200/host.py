"""Host-buffer entry point: log_prob of rows that live in (pinned) host memory.

The rows are cut into chunks; chunk i is copied H2D, evaluated and its log-probabilities copied
D2H on stream i % 2, so the PCIe copies of one chunk overlap the kernels of the other.  This is
what ``bench.py`` times as ``e2e``.
"""
from __future__ import annotations

import torch


class HostPipeline:
    def __init__(self, flow, dim: int, device, chunk_rows: int = 1 << 19, n_streams: int = 2):
        self.flow, self.dim, self.device = flow, dim, torch.device(device)
        self.chunk_rows = int(chunk_rows)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self.bufs = [torch.empty(self.chunk_rows, dim, device=self.device) for _ in range(n_streams)]

    @torch.no_grad()
    def log_prob(self, y_host: torch.Tensor, lp_host: torch.Tensor, **kwargs) -> torch.Tensor:
        """y_host [rows, dim] (pinned for real overlap) -> lp_host [rows, 1], both on the host.
        Returns after everything is enqueued AND the current stream waits on the work."""
        rows = y_host.shape[0]
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for i, r0 in enumerate(range(0, rows, self.chunk_rows)):
            r1 = min(rows, r0 + self.chunk_rows)
            s = self.streams[i % len(self.streams)]
            buf = self.bufs[i % len(self.bufs)][: r1 - r0]
            with torch.cuda.stream(s):
                buf.copy_(y_host[r0:r1], non_blocking=True)
                lp = self.flow.log_prob(buf, **kwargs)
                lp_host[r0:r1].copy_(lp, non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        return lp_host

    @torch.no_grad()
    def map(self, fn, ins_host, outs_host, chunk_rows: int = None):
        """Generic form of ``log_prob``: ``fn(*device_chunks) -> tensor or tuple`` applied chunk-wise to host
        tensors that share their leading (row) dimension; results are copied into ``outs_host``.  Device staging
        buffers are allocated per call shape and cached."""
        rows = ins_host[0].shape[0]
        chunk = int(chunk_rows or self.chunk_rows)
        cur = torch.cuda.current_stream(self.device)
        key = tuple((tuple(t.shape[1:]), t.dtype) for t in ins_host) + (chunk,)
        if getattr(self, '_map_key', None) != key:
            self._map_bufs = [[torch.empty((chunk,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
                               for t in ins_host] for _ in self.streams]
            self._map_key = key
        for s in self.streams:
            s.wait_stream(cur)
        for i, r0 in enumerate(range(0, rows, chunk)):
            r1 = min(rows, r0 + chunk)
            s = self.streams[i % len(self.streams)]
            bufs = [b[: r1 - r0] for b in self._map_bufs[i % len(self.streams)]]
            with torch.cuda.stream(s):
                for b, h in zip(bufs, ins_host):
                    b.copy_(h[r0:r1], non_blocking=True)
                res = fn(*bufs)
                res = res if isinstance(res, (tuple, list)) else (res,)
                for o, r in zip(outs_host, res):
                    o[r0:r1].copy_(r, non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        return outs_host
