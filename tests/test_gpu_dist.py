"""The N > 1 path on real GPUs: two processes, one per device, over NCCL/NVLink. Bands stored by
the walk kernel straight into rank 0's device image (CUDA IPC peer mapping), the grouped
ncclSend/ncclRecv gather, and the pinned shared-memory host image must all give the image a single
rank renders, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from course5_b200 import api, synth
    from course5_b200.dist import BandRenderer, SharedHostImage
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    mesh = synth.kuhn_cube(24, seed=72)
    ctx = api.Context(devices=(rank,))
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    results = {}
    views = [api.make_view(480, 360, X=0.4, Y=Y) for Y in (0.2, 0.9, 1.4)]
    if rank == 0:
        for k, v in enumerate(views):
            results[f"full{k}"] = ctx.render(v)[0]
    for mode in ("p2p", "sendrecv"):
        br = BandRenderer(ctx, device=device, rank=rank, world=world, gather=mode)
        for k, v in enumerate(views):          # synchronous, bands re-cut by time after every view
            image, st, bands = br.render(v, rebalance="time")
            torch.cuda.synchronize(device)
            if rank == 0:
                results[f"{mode}{k}"] = image.cpu().numpy().copy()
                results[f"{mode}_bands{k}"] = np.array(bands)
        imgs = [br.render(v, stats=False, pipeline=True)[0] for v in views]   # pipelined, fire and forget
        br.finish()
        torch.cuda.synchronize(device)
        if rank == 0:
            # three buffer sets: all three views are intact
            results[f"{mode}_pipe1"] = imgs[1].cpu().numpy().copy()
            results[f"{mode}_pipe2"] = imgs[2].cpu().numpy().copy()
        dist.barrier()
        br.close()
    shared = SharedHostImage(ctx, 480, 360, rank=rank, world=world)
    bands = api.balanced_bands(np.ones(360), world)
    for k, v in enumerate(views[:2]):
        if rank == 0:
            shared.array[:] = -1.0
        shared.barrier()
        shared.render_band(v, bands[rank])
        shared.barrier()
        if rank == 0:
            results[f"shared{k}"] = shared.array.copy()
    shared.close()
    if rank == 0:
        np.savez(out_path, **results)
    ctx.close()
    dist.destroy_process_group()


def test_two_gpus_one_image(gpu_lib, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    out = str(tmp_path / "bands.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)

    def differing_rows(name, k):
        rows = np.where(~np.all((r[name] == r[f"full{k}"]) | (np.isnan(r[name]) & np.isnan(r[f"full{k}"])), axis=(1, 2)))[0]
        return None if rows.size == 0 else (name, int(rows.size), int(rows[0]), int(rows[-1]))

    bad = []
    for k in range(3):
        bad += [differing_rows(f"p2p{k}", k), differing_rows(f"sendrecv{k}", k)]
    for mode in ("p2p", "sendrecv"):
        bad += [differing_rows(f"{mode}_pipe1", 1), differing_rows(f"{mode}_pipe2", 2)]
    for k in range(2):
        bad.append(differing_rows(f"shared{k}", k))
    bad = [b for b in bad if b]
    bands = {m: [r[f"{m}_bands{k}"].tolist() for k in range(3)] for m in ("p2p", "sendrecv")}
    assert not bad, f"(image, rows that differ, first, last): {bad}; bands: {bands}"
    for mode in ("p2p", "sendrecv"):
        assert r[f"{mode}_bands0"][0][1] == 180                    # first view: equal heights


def _group_worker(rank, world, port, out_path):
    """View groups on real GPUs: two groups of one rank each take alternate views; every image is assembled
    in rank 0's memory, by stores over NVLink (p2p) and by the grouped send/recv."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from course5_b200 import api, synth
    from course5_b200.dist import BandRenderer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    mesh = synth.kuhn_cube(20, seed=74)
    ctx = api.Context(devices=(rank,))
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    views = [api.make_view(400, 300, X=0.4, Y=Y) for Y in (0.2, 0.9, 1.4, 0.6, 1.1)]
    results = {}
    if rank == 0:
        for k, v in enumerate(views):
            results[f"full{k}"] = ctx.render(v)[0]
    for mode in ("p2p", "sendrecv"):
        br = BandRenderer(ctx, device=device, rank=rank, world=world, gather=mode, lanes=2, groups=world)
        br.prepare(views[0])
        imgs = [br.render(v, stats=False, pipeline=True)[0] for v in views]
        br.finish()
        torch.cuda.synchronize(device)
        dist.barrier()
        if rank == 0:
            for k in range(len(views)):
                results[f"{mode}{k}"] = imgs[k].cpu().numpy().copy()
        dist.barrier()
        br.close()
    if rank == 0:
        np.savez(out_path, **results)
    ctx.close()
    dist.destroy_process_group()


def test_view_groups_on_two_gpus(gpu_lib, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    out = str(tmp_path / "groups.npz")
    mp.spawn(_group_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    for mode in ("p2p", "sendrecv"):
        for k in range(5):
            assert np.array_equal(r[f"{mode}{k}"], r[f"full{k}"], equal_nan=True), f"{mode}: view {k} (rendered by rank {k % 2})"
