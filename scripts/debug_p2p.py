"""2-rank diagnosis of the peer-mapped image: which rows of rank 0's image differ from a single-rank
render after each view, and do they hold the buffer's previous content (stale) or something else."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from course5_b200 import api, synth
from course5_b200.dist import BandRenderer

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); device = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=device)
mesh = synth.kuhn_cube(24, seed=72)
ctx = api.Context(devices=(rank,))
ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
views = [api.make_view(480, 360, X=0.4, Y=Y) for Y in (0.2, 0.9, 1.4)]
full = [ctx.render(v)[0] for v in views] if rank == 0 else None

def rows_differ(a, b):
    same = np.all((a == b) | (np.isnan(a) & np.isnan(b)), axis=(1, 2))
    r = np.where(~same)[0]
    return (int(r.size), int(r[0]), int(r[-1])) if r.size else None

def run(label, extra_wait=False):
    br = BandRenderer(ctx, device=device, rank=rank, world=world, gather="p2p")
    prev = [None, None]
    for it in range(6):
        k = it % 3
        image, st, bands = br.render(views[k], rebalance="time")
        torch.cuda.synchronize(device)
        if extra_wait:
            dist.barrier(); time.sleep(0.01); torch.cuda.synchronize(device)
        if rank == 0:
            got = image.cpu().numpy().copy()
            bad = rows_differ(got, full[k])
            stale = None
            if bad and prev[it & 1] is not None:
                lo, hi = bad[1], bad[2] + 1
                stale = bool(np.array_equal(got[lo:hi], prev[it & 1][lo:hi], equal_nan=True))
            time.sleep(0.05)
            again = rows_differ(image.cpu().numpy(), full[k])
            print(f"{label} it={it} view={k} bands={bands} differ={bad} stale={stale} reread_after_50ms={again}", flush=True)
            prev[it & 1] = got
        dist.barrier()
    br.close()

mode = sys.argv[1] if len(sys.argv) > 1 else "base"
run(mode, extra_wait=(mode == "wait"))
ctx.close()
dist.destroy_process_group()
