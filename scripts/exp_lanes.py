#!/usr/bin/env python
"""Steady-state time per view of ONE GPU's share of the work (a row band, as at N GPUs) with 1 and 2
views in flight: what a rank of an N-GPU run can sustain, without the other ranks.

    python scripts/exp_lanes.py C3 --rows "0,1800;430,555;800,925" --lanes 1,2 --views 24
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from course5_b200 import api, hostlib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--rows", default="0,0")
    ap.add_argument("--lanes", default="1,2")
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--view", default=None)
    ap.add_argument("--no-solids", action="store_true")
    ap.add_argument("--debug", default="", help="c5_debug_set knobs, key=value[,key=value]")
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    mesh, view = synth.make_config(args.config)
    if args.view:
        view["X"], view["Y"] = (float(x) for x in args.view.split(","))
    ctx = api.Context(devices=(0,))
    debug = dict(kv.split("=") for kv in args.debug.split(",") if kv)
    for key, value in debug.items():
        ctx.debug_set(key, int(value))
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    if not args.no_solids:
        solids = hostlib.make_solids(view["D"])
        ctx.upload_solids(solids[0], True)
        ctx.upload_solids(solids[1], False)
    max_lanes = max(int(x) for x in args.lanes.split(","))
    lanes = [(ctx, torch.cuda.Stream(dev))] + [(ctx.sibling(), torch.cuda.Stream(dev)) for _ in range(max_lanes - 1)]
    outs = [torch.empty((view["res_y"], view["res_x"], 2), dtype=torch.float64, device=dev) for _ in range(max_lanes)]
    for band in args.rows.split(";"):
        lo, hi = (int(x) for x in band.split(","))
        extra = {} if hi == 0 else dict(row_begin=lo, row_end=hi)
        v = api.make_view(view["res_x"], view["res_y"], X=view["X"], Y=view["Y"], I=view["I"],
                          alpha_limit=view["alpha_limit"], **extra)
        for c, s in lanes:
            st = c.render_device(v, outs[0].data_ptr(), s.cuda_stream)     # warm, and the band's step count
        for n_lanes in (int(x) for x in args.lanes.split(",")):
            times = []
            for rep in range(5):
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for s in (s for _, s in lanes[:n_lanes]):
                    s.wait_stream(torch.cuda.current_stream(dev))
                for k in range(args.views):
                    c, s = lanes[k % n_lanes]
                    c.render_device(v, outs[k % n_lanes].data_ptr(), s.cuda_stream, stats=False)
                for s in (s for _, s in lanes[:n_lanes]):
                    torch.cuda.current_stream(dev).wait_stream(s)
                e1.record()
                torch.cuda.synchronize(dev)
                times.append(e0.elapsed_time(e1) / args.views)
            ms = float(np.median(times))
            print(json.dumps({"config": args.config, "view": [view["X"], view["Y"]], "rows": [lo, hi], "lanes": n_lanes,
                              "ms_per_view": round(ms, 4), "tet_steps": st["tet_steps"],
                              "Gsteps_per_s": round(st["tet_steps"] / ms / 1e6, 2),
                              "reps_ms": [round(t, 4) for t in times], "one_view_ms_total": round(st["ms_total"], 4), "one_view_ms_graze": round(st["ms_graze"], 4),
                              "debug": debug}), flush=True)
    for c, _ in lanes[1:]:
        c.close()
    ctx.close()


if __name__ == "__main__":
    main()
