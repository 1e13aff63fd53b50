"""Row-band sharding across ranks (one process per GPU) and the image gather.

The ray pass has no cross-pixel state (/root/reference/project/src/plane.cpp:161-169), so an image
shards by rows: the mesh is replicated, rank r renders the contiguous band rows[r] and ONE
exchange — a gather-v of the bands to rank 0 — assembles the x-fastest image
(object2d.cpp:17-21 makes a row band a contiguous span of the output). The exchange is a grouped
point-to-point send/recv (``ncclSend``/``ncclRecv`` over NVLink with the ``nccl`` backend; ``gloo``
in the CPU tests), because cost-balanced bands have unequal sizes.

Equal-height bands are badly unbalanced (the mesh sits in the middle rows), so bands are cut by
per-row tet-step counts from the previous view (``Context.last_row_cost``), all-reduced so every
rank derives the same cuts.

Sweeps can pipeline: with ``pipeline=True`` the gather of view k is left in flight while view k+1
renders into the other of two buffer sets, so the exchange costs no time on the critical path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import api


class BandRenderer:
    """Renders row bands of successive views on this rank's device and gathers them on rank 0."""

    def __init__(self, ctx: api.Context, *, device: torch.device, rank: int, world: int,
                 base_cost: float = 64.0):
        self.ctx, self.device, self.rank, self.world = ctx, device, rank, world
        self.base_cost = base_cost
        self.row_cost: np.ndarray | None = None
        self._band_buf = [None, None]      # two buffer sets: the gather of one view may still be
        self._image = [None, None]         # reading/writing set k while view k+1 fills the other
        self._pending = [[], []]
        self._count = 0
        self._bands = None                 # cached cut, valid until the row costs change

    def bands(self, res_y: int) -> list[tuple[int, int]]:
        if self._bands is not None and self._bands[0] == res_y:
            return self._bands[1]
        if self.row_cost is None or self.row_cost.shape[0] != res_y:
            cost = np.ones(res_y)          # first view: equal heights
            cut = api.balanced_bands(cost, self.world)
        else:
            cut = api.balanced_bands(self.row_cost, self.world, base_cost=self.base_cost)
        self._bands = (res_y, cut)
        return cut

    def _buffers(self, view: api.View, rows: int, par: int):
        n = rows * view.res_x * 2
        if self._band_buf[par] is None or self._band_buf[par].numel() < n:
            self._band_buf[par] = torch.empty(n, dtype=torch.float64, device=self.device)
        if self.rank == 0:
            full = view.res_y * view.res_x * 2
            if self._image[par] is None or self._image[par].numel() != full:
                self._image[par] = torch.empty(full, dtype=torch.float64, device=self.device)

    def _drain(self, par: int):
        for req in self._pending[par]:
            req.wait()          # orders the current stream after that exchange
        self._pending[par] = []

    def finish(self):
        """Waits (on the current stream) for every gather still in flight."""
        self._drain(0)
        self._drain(1)

    def render(self, view: api.View, *, gather: bool = True, rebalance: bool = True, stats: bool = True,
               pipeline: bool = False):
        """Renders this rank's band of `view`; returns (image on rank 0 or None, stats, bands).

        The image is a (res_y, res_x, 2) float64 tensor on rank 0's device. With stats=False (and
        rebalance=False) nothing is read back to the host: render and gather are only enqueued on the
        current stream. With pipeline=True the gather is additionally left in flight (call finish(),
        or render two more views, before reading the returned image)."""
        if not stats:
            rebalance = False
        par = self._count & 1
        self._count += 1
        self._drain(par)        # the buffers of this parity are about to be overwritten
        bands = self.bands(view.res_y)
        lo, hi = bands[self.rank]
        self._buffers(view, hi - lo, par)
        v = api.View.from_buffer_copy(view)
        v.row_begin, v.row_end = lo, hi
        if self.rank == 0 and gather:
            target = self._image[par][lo * view.res_x * 2: hi * view.res_x * 2]
        else:
            target = self._band_buf[par][: (hi - lo) * view.res_x * 2]
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0
        st = self.ctx.render_device(v, target.data_ptr(), stream, stats=stats)

        if self.world > 1 and gather:
            ops = []
            if self.rank == 0:
                for r in range(1, self.world):
                    rlo, rhi = bands[r]
                    ops.append(dist.P2POp(dist.irecv, self._image[par][rlo * view.res_x * 2: rhi * view.res_x * 2], r))
            else:
                ops.append(dist.P2POp(dist.isend, target, 0))
            self._pending[par] = list(dist.batch_isend_irecv(ops))
            if not pipeline:
                self._drain(par)

        if rebalance:
            cost = torch.from_numpy(self.ctx.last_row_cost(view.res_y).astype(np.int64))
            if self.world > 1:
                cost = cost.to(self.device)
                dist.all_reduce(cost, op=dist.ReduceOp.SUM)
                cost = cost.cpu()
            self.row_cost = cost.numpy().astype(np.float64)
            self._bands = None

        image = None
        if self.rank == 0 and gather:
            image = self._image[par].view(view.res_y, view.res_x, 2)
        return image, st, bands
