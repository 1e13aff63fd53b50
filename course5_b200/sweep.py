"""In-process view sweep: the mesh stays resident and only the rotations change per frame.

Replaces the reference's process-per-frame tooling (utility/rotate_traces.py:11-21: 1 500 launches
of `course ... -Y theta`, each re-reading the VTK file, regenerating the solids and re-allocating
~0.8 KB per pixel). Under torchrun the frames are dealt round-robin to the ranks (one whole view
per GPU): frames are independent, so there is NO collective on the data path.

    python -m course5_b200.sweep --config C4 --frames 360 [--out-dir frames/]
    python -m torch.distributed.run --nproc-per-node 8 -m course5_b200.sweep --config C4 --frames 360
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np

from . import api, hostlib, synth


def sweep_views(res_x, res_y, *, X, I, alpha_limit, y_from=0.0, y_to=2.0, frames=360, **extra):
    """The frame list of a full turn about y (angles in units of pi, like the CLI)."""
    return [api.make_view(res_x, res_y, X=X, Y=y_from + (y_to - y_from) * k / frames, I=I,
                          alpha_limit=alpha_limit, **extra) for k in range(frames)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--frames", type=int, default=360)
    ap.add_argument("--n", type=int, default=None, help="lattice size override")
    ap.add_argument("--out-dir", default=None, help="write frame_%%04d.vti here")
    args = ap.parse_args()

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    mesh, view = synth.make_config(args.config, n=args.n)
    roche, sphere = hostlib.make_solids(view["D"])
    ctx = api.Context(devices=(local,))
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    ctx.upload_solids(roche, True)
    ctx.upload_solids(sphere, False)
    views = sweep_views(view["res_x"], view["res_y"], X=view["X"], I=view["I"], alpha_limit=view["alpha_limit"],
                        frames=args.frames)
    mine = list(range(rank, args.frames, world))
    host = torch.empty((view["res_y"], view["res_x"], 2), dtype=torch.float64).pin_memory().numpy()
    if args.out_dir and rank == 0:
        os.makedirs(args.out_dir, exist_ok=True)

    ctx.render(views[mine[0]], out=host)  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    steps = 0
    for k in mine:
        _, st = ctx.render(views[k], out=host)
        steps += st["tet_steps"]
        if args.out_dir:
            hostlib.write_vti(os.path.join(args.out_dir, f"frame_{k:04d}.vti"), host)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tot = torch.tensor([dt, float(steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        t_max = tot.clone()
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dt, steps = float(t_max[0]), float(tot[1])
    if rank == 0:
        print(json.dumps({"config": args.config, "frames": args.frames, "n_gpus": world, "seconds": dt,
                          "views_per_sec": args.frames / dt, "tet_steps_per_sec": steps / dt,
                          "res": [view["res_x"], view["res_y"]], "n_tets": mesh.n_tets,
                          "includes": "rotate + BVH refit + solid mask + walk + D2H per view"
                                      + (" + .vti write" if args.out_dir else "")}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
