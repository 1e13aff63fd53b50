"""Row-band sharding across ranks (one process per GPU) and the assembly of the image.

The ray pass has no cross-pixel state (/root/reference/project/src/plane.cpp:161-169), so an image
shards by rows: the mesh is replicated, rank r renders the contiguous band rows[r], and the bands
must end up in ONE x-fastest image (object2d.cpp:17-21 makes a row band a contiguous span of it).
Three ways to get them there, all behind :class:`BandRenderer` / :class:`SharedHostImage`:

``gather="p2p"`` (default on CUDA)
    The image lives on rank 0's GPU; the other ranks map it through CUDA IPC (``c5_image_open``)
    and their walk kernels store their pixels straight into it over NVLink — the exchange is fused
    into the kernel that produces the data, 16 B posted stores spread over the kernel's run time.
    What remains per view is one tiny all-reduce used as a barrier ("every band of view k is in").
``gather="sendrecv"``
    Each rank renders into a local band buffer and ONE grouped point-to-point exchange
    (``ncclSend``/``ncclRecv`` with the ``nccl`` backend, ``gloo`` in the CPU tests) gathers the
    bands on rank 0; a gather-v, because cost-balanced bands have unequal sizes. This is the
    baseline the fused path is measured against.
:class:`SharedHostImage`
    For callers that want the image in HOST memory (the .vti writer): the image is a POSIX
    shared-memory segment that every rank pins (``c5_host_register``); ``c5_render`` with a row band
    then writes the band in place over that GPU's own PCIe link — N links in parallel instead of one
    device-to-host copy of the whole image on rank 0.

Equal-height bands are badly unbalanced (the mesh sits in the middle rows), so bands are cut by
per-row tet-step counts of the previous view (``Context.last_row_cost``), all-reduced so every
rank derives the same cuts. Tet-steps are not quite time (a band of short silhouette rays runs at
a lower rate than one of long central rays), so ``rebalance="time"`` additionally scales each
band's row costs by the time that band actually took.

View groups (``groups`` > 1): the ranks are split into groups of consecutive ranks, each group renders
whole views by row bands and the groups take alternate views (view k belongs to group k % groups);
every image is still assembled in rank 0's memory. ``groups=1`` is plain row bands, ``groups=world``
one whole view per rank.

Sweeps can pipeline: with ``pipeline=True`` the exchange (or barrier) of view k is left in flight
while the next views render into the other buffer sets. With L views in flight per GPU ("lanes":
the context and L-1 siblings sharing its mesh, each on its own stream) and G view groups there are
S = 2 L G + 1 sets (L + 1 would do; the second round lets a rank run a round of views ahead of the
slowest one), and a rank starts view k+S (which reuses view k's set) only after barrier k+1 has
completed. Rank 0 enqueues barrier k+1 inside ``render(k+1)``, so the contract for the consumer of
rank 0's image is: whatever reads image k must be finished, or enqueued on the current stream,
BEFORE ``render(k+1)`` is called; the peers cannot overwrite it earlier.
"""
from __future__ import annotations

import contextlib
from multiprocessing import shared_memory

import numpy as np
import torch
import torch.distributed as dist

from . import api


class BandRenderer:
    """Renders row bands of successive views on this rank's device and assembles them on rank 0."""

    def __init__(self, ctx: api.Context, *, device: torch.device, rank: int, world: int,
                 base_cost: float = 64.0, gather: str = "auto", lanes: int = 2, sets: int | None = None,
                 groups: int = 1):
        self.ctx, self.device, self.rank, self.world = ctx, device, rank, world
        # View groups: the ranks are split into `groups` groups of world / groups consecutive ranks; every
        # group renders WHOLE views by row bands, and the groups take alternate views (view k belongs to group
        # k % groups). groups = 1 is plain row bands (every rank a band of every view); groups = world is one
        # whole view per rank (what sweep.py does). In between it trades bands per view for views per node: a
        # 1/8 band is a single wave of blocks and sustains a lower rate than a 1/4 band.
        # Whatever the groups, every image is assembled in rank 0's memory.
        groups = max(1, int(groups))
        if world % groups:
            raise ValueError("groups must divide the number of ranks")
        self.groups = groups
        self.per_group = world // groups            # bands per view
        self.group = rank // self.per_group         # the group this rank renders for
        self.band_index = rank % self.per_group     # which band of the group's views
        self.base_cost = base_cost
        if gather == "auto":
            gather = "p2p" if (device.type == "cuda" and world > 1) else "sendrecv"
        if gather not in ("p2p", "sendrecv"):
            raise ValueError("gather must be 'auto', 'p2p' or 'sendrecv'")
        self.gather_mode = gather
        self.row_cost: np.ndarray | None = None
        self.n_lanes = max(1, int(lanes))
        # view k + n_sets reuses view k's buffers (module docstring). lanes + 1 sets are the minimum; with that a
        # lane's next view waits for EVERY rank to have finished the lane's previous one, so each view costs the
        # slowest rank's time for it. A second round of sets lets a rank run a whole round of views ahead, and
        # the view-to-view scatter of the ranks (+-10 % measured) averages out instead of adding up.
        self.n_sets = max(self.n_lanes + 1, int(sets)) if sets else (2 * self.n_lanes * self.groups + 1 if world > 1 else self.n_lanes + 1)
        self._band_buf = [None] * self.n_sets  # buffer sets: the exchange of one view may still be reading /
        self._image = [None] * self.n_sets     # writing its set while the next views fill the others
        self._peer = [None] * self.n_sets      # p2p: (device pointer of rank 0's image, bytes) per set
        self._pending = {}                 # view number -> requests of its exchange / barrier
        self._count = 0
        self._own_count = 0                # views this rank has rendered itself (lanes rotate over these)
        self._aux = None                   # stream for the barriers of views that belong to other groups
        self._bands = None                 # cached cut, valid until the row costs change
        self._cost_has_base = False        # row_cost is measured time (the per-row constant is in it)
        self._flag = None
        self._time_cost = None             # calibrate(): last per-row time estimate (milliseconds)
        self.calibration_log = []          # calibrate(): per round, the cut that was timed and what every rank sustained
        self._lanes = []                   # [(context, side stream)] on CUDA, created on first use

    # -- band cuts --------------------------------------------------------------------------------
    def bands(self, res_y: int) -> list[tuple[int, int]]:
        if self._bands is not None and self._bands[0] == res_y:
            return self._bands[1]
        if self.row_cost is None or self.row_cost.shape[0] != res_y:
            cost = np.ones(res_y)          # first view: equal heights
            cut = api.balanced_bands(cost, self.per_group)
        else:
            cut = api.balanced_bands(self.row_cost, self.per_group, base_cost=0.0 if self._cost_has_base else self.base_cost)
        self._bands = (res_y, cut)
        return cut

    # -- buffers ----------------------------------------------------------------------------------
    def _buffers(self, view: api.View, rows: int, par: int):
        n = rows * view.res_x * 2
        if self._band_buf[par] is None or self._band_buf[par].numel() < n:
            self._band_buf[par] = torch.empty(n, dtype=torch.float64, device=self.device)
        if self.rank == 0:
            full = view.res_y * view.res_x * 2
            if self._image[par] is None or self._image[par].numel() != full:
                self._image[par] = torch.empty(full, dtype=torch.float64, device=self.device)

    def _peer_image(self, view: api.View, par: int) -> int:
        """Device pointer of rank 0's image (set `par`) in THIS process, (re)created on a size change."""
        nbytes = view.res_y * view.res_x * 16
        if self._peer[par] is not None and self._peer[par][1] == nbytes:
            return self._peer[par][0]
        self.finish()
        if self._peer[par] is not None:
            dist.barrier()                 # nobody may still be writing the old mapping
            self.ctx.image_close(self._peer[par][0])
            self._peer[par] = None
        handle = torch.zeros(api.IPC_HANDLE_BYTES, dtype=torch.uint8)
        if self.rank == 0:
            ptr, raw = self.ctx.image_create(nbytes)
            handle = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
        handle = handle.to(self.device)
        dist.broadcast(handle, src=0)
        if self.rank != 0:
            ptr = self.ctx.image_open(bytes(handle.cpu().numpy().tobytes()))
        self._peer[par] = (ptr, nbytes)
        if self.rank == 0:
            self._image[par] = _tensor_from_pointer(ptr, view.res_y * view.res_x * 2, self.device)
        return ptr

    def _drain(self, upto: int):
        """Orders the current stream after the exchanges / barriers of all views numbered <= upto."""
        for k in sorted(k for k in self._pending if k <= upto):
            for req in self._pending.pop(k):
                req.wait()

    def finish(self):
        """Orders the current stream after every view and every exchange still in flight."""
        if self._lanes:
            self._join_lanes()
        self._drain(self._count)

    def close(self):
        self.finish()
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        for par in range(self.n_sets):
            if self._peer[par] is not None:
                if self.world > 1:
                    dist.barrier()
                self._image[par] = None
                self.ctx.image_close(self._peer[par][0])
                self._peer[par] = None
        for ctx, _ in self._lanes[1:]:
            ctx.close()
        self._lanes = []

    # -- lanes ------------------------------------------------------------------------------------
    def _lane(self, k: int):
        """(context, torch stream or None) that renders view k. On CUDA there are n_lanes lanes — this
        context and siblings sharing its mesh (c5_create_sibling), each with its own side stream —
        so that consecutive pipelined views overlap on the device: the last rays of view k no longer
        leave most SMs idle, because view k+1's blocks are already there to take them."""
        if not self._lanes:
            cuda = self.device.type == "cuda"   # (the CPU test build has no streams: its lanes only rotate contexts)
            self._lanes = [(self.ctx, torch.cuda.Stream(self.device) if cuda else None)]
            for _ in range(self.n_lanes - 1):
                self._lanes.append((self.ctx.sibling(), torch.cuda.Stream(self.device) if cuda else None))
        return self._lanes[k % len(self._lanes)]

    def enable_timeline(self, n_views: int):
        """Keeps the phase moments of the last n_views views of every lane (c5_debug_set "timeline")."""
        self._lane(0)
        self.ctx.debug_set("timeline", n_views)      # reaches the siblings (the other lanes) too

    def timeline(self, origin_event: int) -> list[np.ndarray]:
        """Per lane: (n, 6) ms since origin_event — start, rotated, refitted, mask, pixel kernel, grazing-ray kernel."""
        return [c.timeline(origin_event) for c, _ in self._lanes]

    def kernel_launches(self) -> int:
        """Kernels launched so far by every context this renderer drives."""
        if not self._lanes:
            return self.ctx.kernel_launches()
        return sum(c.kernel_launches() for c, _ in self._lanes)

    def _join_lanes(self):
        """Orders the caller's current stream after everything enqueued on the lane streams."""
        if self.device.type != "cuda":
            return
        cur = torch.cuda.current_stream(self.device)
        for _, s in self._lanes:
            cur.wait_stream(s)
        if self._aux is not None:
            cur.wait_stream(self._aux)

    def prepare(self, view: api.View):
        """Everything a pipelined run of `view`-sized images creates lazily, created now: the lanes and,
        for gather="p2p", rank 0's images and their peer mappings (cudaMalloc, CUDA IPC open and a
        broadcast each — milliseconds that do not belong in the first views of a sweep). Collective."""
        self._lane(0)
        if self.gather_mode == "p2p" and self.world > 1:
            for par in range(self.n_sets):
                self._peer_image(view, par)
        else:
            for par in range(self.n_sets):
                self._buffers(view, view.res_y, par)

    # -- band cuts from measured throughput -------------------------------------------------------
    def calibrate(self, view: api.View, *, rounds: int = 4, views: int = 8) -> list[tuple[int, int]]:
        """Cuts the bands so that every rank SUSTAINS the same time per view.

        What an N-rank sweep runs at is the slowest rank's pipelined rate — several views in flight,
        tails and grazing-ray kernels overlapped — not the time one view takes alone, and a band's
        rate per tet-step depends on what is in it (short silhouette rays, the solid mask, rows of
        grazing rays). So each round renders `views` pipelined views of this rank's band without
        any exchange, times them with CUDA events (the clock starts once the pipeline is full), corrects the cost of the band's rows by measured / predicted and re-cuts. Collective:
        every rank must call it with the same arguments."""
        cuda = self.device.type == "cuda"

        def sustained(n: int, skip: int) -> float:
            """Milliseconds per view over the last n of skip + n pipelined views of this rank's band, no
            exchange. The clock starts when view number `skip` completes (an event on its lane's stream):
            the pipeline is full by then. Timing a whole short run instead charges every band the latency
            of its first view, which differs between bands by more than their sustained times do (C3 at
            N = 8: first views 1.06 .. 1.73 ms, sustained 0.63 .. 0.76 ms)."""
            if cuda:
                torch.cuda.synchronize(self.device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            else:
                import time
            for i in range(skip + n):
                self.render(view, gather=False, stats=False, pipeline=True)
                if i == skip - 1:
                    if cuda:
                        self._lane(self._own_count - 1)[1].record_event(e0)      # the lane that view went to
                    else:
                        t0 = time.perf_counter()
            self.finish()
            if cuda:
                e1.record()
                torch.cuda.synchronize(self.device)
                return e0.elapsed_time(e1) / n
            return 1e3 * (time.perf_counter() - t0) / n

        steps = None
        if self._time_cost is None or self._time_cost.shape[0] != view.res_y:
            self._time_cost = None
            self.render(view, gather=False, rebalance="steps")      # per-row tet-steps, all-reduced; bands re-cut by them
            steps = self.row_cost.copy()
        for _ in range(rounds):
            # the cut that is timed is the cut the measurement is attributed to: nothing below changes it
            # before the update at the end of the round (render(rebalance=False) leaves the cuts alone)
            bands = self.bands(view.res_y)
            lo, hi = bands[self.band_index]
            ms = sustained(views, self.n_lanes + 2)
            every = torch.zeros(self.world, dtype=torch.float64)   # (with view groups, several ranks time the same band)
            every[self.rank] = ms
            if self.world > 1:
                every = every.to(self.device)
                dist.all_reduce(every, op=dist.ReduceOp.SUM)
                every = every.cpu()
            self.calibration_log.append({"bands": [list(b) for b in bands], "sustained_ms": [round(float(x), 4) for x in every]})
            if self._time_cost is None:
                # first estimate: the band's time spread over its rows in proportion to their tet-steps
                mine = torch.from_numpy(api.time_weighted_row_cost(steps, (lo, hi), ms, base_cost=self.base_cost))
            else:
                # afterwards: the rows keep the cost they have (it carries what earlier rounds learnt about how
                # the rate differs from band to band) and the band as a whole is corrected towards what it
                # measured now; rows that change bands take their cost with them
                mine = torch.zeros(view.res_y, dtype=torch.float64)
                have = self._time_cost[lo:hi]
                predicted = float(have.sum())
                if predicted > 0.0:
                    # decreasing gain: one measurement scatters by +-10 % (how the streams happen to interleave),
                    # so later rounds average rather than chase it
                    gain = max(0.25, 0.9 / (1 + 0.25 * max(0, len(self.calibration_log) - 2)))
                    mine[lo:hi] = torch.from_numpy(have * (ms / predicted) ** gain)
                else:
                    mine[lo:hi] = ms / max(hi - lo, 1)
            if self.world > 1:
                mine = mine.to(self.device)
                dist.all_reduce(mine, op=dist.ReduceOp.SUM)
                mine = mine.cpu()
            new_cost = mine.numpy().astype(np.float64) / self.groups   # mean over the ranks that hold the same band
            self._time_cost = new_cost
            self.row_cost, self._cost_has_base = new_cost, True
            self._bands = None
        return self.bands(view.res_y)

    # -- one view ---------------------------------------------------------------------------------
    def render(self, view: api.View, *, gather: bool = True, rebalance: bool | str = True, stats: bool = True,
               pipeline: bool = False):
        """Renders this rank's band of `view`; returns (image on rank 0 or None, stats, bands).

        The image is a (res_y, res_x, 2) float64 tensor on rank 0's device. With stats=False (and
        rebalance=False) nothing is read back to the host: render and exchange are only enqueued.
        With pipeline=True they are additionally left in flight on the lane's side stream (call
        finish() before reading the returned image); otherwise the caller's current stream is
        ordered after them on return.
        rebalance: True / "steps" = cut the next view's bands by this view's per-row tet-steps;
        "time" = additionally weight each band by the device time it took."""
        if not stats:
            rebalance = False
        k = self._count
        self._count += 1
        par = k % self.n_sets
        bands = self.bands(view.res_y)
        lo, hi = bands[self.band_index]
        # view groups: an assembled view belongs to ONE group; a local render (gather=False) is everybody's
        view_group = k % self.groups
        mine = (not gather) or view_group == self.group
        v = api.View.from_buffer_copy(view)
        v.row_begin, v.row_end = lo, hi
        p2p = self.gather_mode == "p2p" and self.world > 1 and gather
        if p2p:
            self._peer_image(view, par)     # (re)creates the mapping outside the lane stream if needed
        if mine:
            ctx, lane_stream = self._lane(self._own_count)
            self._own_count += 1
        else:   # only this view's barrier is enqueued here: on a stream of its own, the lanes are not held up
            ctx = self.ctx
            self._lane(0)
            if self._aux is None and self.device.type == "cuda":
                self._aux = torch.cuda.Stream(self.device)
            lane_stream = self._aux

        st = {}
        if lane_stream is not None:
            # what the caller enqueued so far (e.g. its reads of older images) comes first
            lane_stream.wait_stream(torch.cuda.current_stream(self.device))
            scope = torch.cuda.stream(lane_stream)
            stream = lane_stream.cuda_stream
        else:
            scope = contextlib.nullcontext()
            stream = 0
        with scope:
            # this view reuses the set of view k - n_sets: every rank must be past the barrier that
            # follows it, k - n_sets + 1 (module docstring)
            self._drain(k - self.n_sets + 1)
            if p2p:
                if mine:
                    base = self._peer[par][0]
                    st = ctx.render_device(v, base + lo * view.res_x * 16, stream, stats=stats)
                # barrier: when it completes on a rank's stream, every band of this view is in
                if self._flag is None:
                    self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
                self._pending[k] = [dist.all_reduce(self._flag, async_op=True)]
            else:
                self._buffers(view, hi - lo, par)
                if self.rank == 0 and gather and mine:
                    target = self._image[par][lo * view.res_x * 2: hi * view.res_x * 2]
                else:
                    target = self._band_buf[par][: (hi - lo) * view.res_x * 2]
                if mine:
                    st = ctx.render_device(v, target.data_ptr(), stream, stats=stats)
                if self.world > 1 and gather:
                    ops = []
                    if self.rank == 0:
                        for b in range(self.per_group):
                            r = view_group * self.per_group + b
                            if r == 0:
                                continue    # rank 0's own band is in place already
                            rlo, rhi = bands[b]
                            ops.append(dist.P2POp(dist.irecv, self._image[par][rlo * view.res_x * 2: rhi * view.res_x * 2], r))
                    elif mine:
                        ops.append(dist.P2POp(dist.isend, target, 0))
                    if ops:
                        self._pending[k] = list(dist.batch_isend_irecv(ops))
            if not pipeline:
                self._drain(k)
        if not pipeline and lane_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(lane_stream)

        if rebalance:
            # a local render is done by every rank: each band then arrives once per group
            share = 1.0 if gather else 1.0 / self.groups
            steps = torch.zeros(view.res_y, dtype=torch.float64)
            if mine:
                steps = torch.from_numpy(ctx.last_row_cost(view.res_y).astype(np.float64)) * share
            both = torch.zeros((2, view.res_y), dtype=torch.float64)
            both[0] = steps
            if rebalance == "time" and self.world > 1 and mine:
                # spread the time this band took over its rows in proportion to their tet-steps (plus
                # the per-row constant): bands whose rays run at a lower rate, or that carry the solid
                # mask, get proportionally fewer rows next time
                both[1] = torch.from_numpy(api.time_weighted_row_cost(
                    steps.numpy(), (lo, hi), float(st["ms_mask"] + st["ms_walk"]) * share, base_cost=self.base_cost))
            if self.world > 1:
                both = both.to(self.device)
                dist.all_reduce(both, op=dist.ReduceOp.SUM)
                both = both.cpu()
            if rebalance == "time" and float(both[1].sum()) > 0.0:
                new_cost = both[1].numpy().astype(np.float64)                                    # milliseconds per row
                if self._cost_has_base and self.row_cost is not None and self.row_cost.shape == new_cost.shape:
                    new_cost = 0.5 * (new_cost + self.row_cost)   # damped: a band's rate depends on its own cut
                self.row_cost, self._cost_has_base = new_cost, True
            else:
                self.row_cost, self._cost_has_base = both[0].numpy().astype(np.float64), False  # tet-steps per row
            self._bands = None

        image = None
        if self.rank == 0 and gather:
            image = self._image[par].view(view.res_y, view.res_x, 2)
        return image, st, bands


def bind_to_device_numa_node(device_index: int) -> dict:
    """Restricts this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that host
    memory it touches first — page-locked images, its rows of a shared image — is allocated on that
    node: N GPUs writing results into host memory then load every socket's memory controllers
    instead of the one the launcher happened to start on. Returns what it did; a box without NUMA
    information (one node, or sysfs hidden) is left alone."""
    import os
    info = {"bound": False}
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(f"{base}/numa_node") as f:
            node = int(f.read().strip())
        info.update(pci=bdf, numa_node=node)
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus=len(cpus))
    except Exception as e:       # no sysfs, no permission, not Linux: nothing to do
        info["note"] = f"{type(e).__name__}: {e}"
    return info


def _tensor_from_pointer(ptr: int, n_doubles: int, device: torch.device) -> torch.Tensor:
    """A float64 tensor view of library-owned device memory (no copy, no ownership)."""
    class _Cai:
        __cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Cai(), device=device)


class SharedHostImage:
    """`sets` (res_y, res_x, 2) float64 HOST images shared by all ranks of the node.

    Rank 0 creates a POSIX shared-memory segment, the others attach; every rank pins it with its
    own context (``c5_host_register``), after which a render of a row band into it makes the walk
    kernel store the band straight into the shared image over that GPU's own PCIe link.

    One view at a time: ``render_band`` (synchronous ``c5_render``) then ``barrier()``; the
    consumer on rank 0 calls ``barrier()`` once more when it is done with the image.

    Pipelined (``sets`` > 1): view k goes to image k % sets. ``submit_band(view, band, k)`` enqueues
    this rank's band (``c5_render_submit``; it first waits until rank 0 has released the image's
    previous view), ``complete_band(ticket, k)`` waits for it and publishes "rank r has finished view
    k" in the segment; rank 0's ``wait_image(k)`` returns the image once every rank has, and
    ``release(k)`` hands the set back. The flags are 8-byte words in the same segment (plain stores,
    polled): no collective on the path. In the CPU tests (hostsim build) the same calls degrade to
    plain memcpys into the shared segment."""

    _POLL_S = 2e-5

    def __init__(self, ctx: api.Context, res_x: int, res_y: int, *, rank: int, world: int, sets: int = 1):
        self.ctx, self.rank, self.world, self.sets = ctx, rank, world, max(1, int(sets))
        image_bytes = res_y * res_x * 16
        flag_bytes = 8 * (world + 1)
        nbytes = self.sets * image_bytes + flag_bytes
        names = [None]
        if rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            self._shm.buf[self.sets * image_bytes: nbytes] = bytes(flag_bytes)
            names = [self._shm.name]
        if world > 1:
            dist.broadcast_object_list(names, src=0)
        if rank != 0:
            self._shm = shared_memory.SharedMemory(name=names[0])
            try:  # the creator unlinks; keep Python's resource tracker from doing it twice
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.arrays = [np.ndarray((res_y, res_x, 2), dtype=np.float64, buffer=self._shm.buf, offset=s * image_bytes)
                       for s in range(self.sets)]
        self.array = self.arrays[0]
        # done[r] = views rank r has completed; released = views rank 0 has handed back
        self._flags = np.ndarray((world + 1,), dtype=np.int64, buffer=self._shm.buf, offset=self.sets * image_bytes)
        self._pinned = np.ndarray((self.sets * res_y, res_x, 2), dtype=np.float64, buffer=self._shm.buf)
        # first touch: every rank writes zeros into an equal share of the rows of every image before anyone pins
        # the segment, so those pages live on the NUMA node of the process that will write most of them
        # (bind_to_device_numa_node); row bands move with the calibration, the share stays close to them
        if world > 1:
            r0, r1 = res_y * rank // world, res_y * (rank + 1) // world
            for a in self.arrays:
                a[r0:r1] = 0.0
            dist.barrier()
        ctx.host_register(self._pinned)
        self._registered = True

    def render_band(self, view: api.View, band: tuple[int, int]) -> dict:
        v = api.View.from_buffer_copy(view)
        v.row_begin, v.row_end = band
        _, st = self.ctx.render(v, out=self.array)
        return st

    # -- pipelined ----------------------------------------------------------------------------------
    def _spin(self, cond, what: str, timeout_s: float = 120.0):
        import time
        t0 = time.monotonic()
        while not cond():
            time.sleep(self._POLL_S)
            if time.monotonic() - t0 > timeout_s:
                raise TimeoutError(f"SharedHostImage: rank {self.rank} waited {timeout_s:.0f} s for {what}; flags {self._flags.tolist()}")

    def submit_band(self, view: api.View, band: tuple[int, int], k: int) -> int:
        """Enqueues this rank's band of view number k (numbered from 0, in order) into image k % sets."""
        self._spin(lambda: int(self._flags[self.world]) >= k - self.sets + 1, f"the release of view {k - self.sets}")
        v = api.View.from_buffer_copy(view)
        v.row_begin, v.row_end = band
        return self.ctx.render_submit(v, self.arrays[k % self.sets])

    def complete_band(self, ticket: int, k: int) -> dict:
        st = self.ctx.render_wait(ticket)
        self._flags[self.rank] = k + 1
        return st

    def wait_image(self, k: int, ranks=None) -> np.ndarray:
        """Rank 0: image of view k once the band of every rank (or of `ranks`, the view group that rendered
        this view) is in."""
        who = list(range(self.world)) if ranks is None else list(ranks)
        self._spin(lambda: int(self._flags[who].min()) >= k + 1, f"the bands of view {k}")
        return self.arrays[k % self.sets]

    def release(self, k: int):
        """Rank 0: view k has been consumed; its image may be overwritten by view k + sets."""
        self._flags[self.world] = k + 1

    def barrier(self):
        if self.world > 1:
            dist.barrier()

    def close(self):
        if self._registered:
            self.ctx.host_unregister(self._pinned)
            self._registered = False
        self.array, self.arrays, self._flags, self._pinned = None, [], None, None
        if self.world > 1:
            dist.barrier()
        self._shm.close()
        if self.rank == 0:
            self._shm.unlink()
