"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libc5ref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
Each fixture stores the reference's pre-float-cast doubles (tau, I), per-pixel record counts and
the solid mask for one synthetic mesh + view; the mesh itself is regenerated from (n, seed,
generator kwargs) by course5_b200.synth, and its sha256 is stored to catch generator drift.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from course5_b200 import synth  # noqa: E402
from oracle import refbind  # noqa: E402

# NOTE: with the reference's solids res_y must be even: the (unrotated) sphere has vertices at
# y == 0 exactly, and an odd res_y puts a pixel row there (a ray through a vertex hits an odd number
# of faces and the reference aborts, plane.cpp:39-41).
CASES = {
    # name: (lattice n, seed, generator kwargs, res_x, res_y, flags, with reference solids)
    "cube_front": (6, 11, {}, 96, 72, dict(X=0.0, Y=0.0, D=0.0, I=0.0, alpha_limit=2.5), False),
    "cube_tilted_solids": (6, 12, {}, 120, 90, dict(X=0.4, Y=0.7, D=0.1, I=-0.03, alpha_limit=2.5), True),
    "cavity_readme_view": (10, 13, dict(scalars="sphere", carve_sphere=True), 120, 90,
                           dict(X=0.5, Y=0.0, D=0.0, I=0.0, alpha_limit=3.0), False),
    "graded_clamped": (8, 14, dict(grade_beta=1.5), 100, 76,
                       dict(X=0.45, Y=1.6, D=0.0, I=0.0, alpha_limit=0.7), True),
}


def mesh_digest(mesh):
    h = hashlib.sha256()
    for a in (mesh.points, mesh.tets, mesh.alpha, mesh.q):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    ref = refbind.Ref()
    index = {}
    for name, (n, seed, gen, rx, ry, flags, solids) in CASES.items():
        mesh = synth.kuhn_cube(n, seed, **gen)
        img = ref.render(mesh.tet_points(), mesh.alpha, mesh.q, res_x=rx, res_y=ry, threads=4,
                         solids=1 if solids else 0, raw=True, **flags)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), tau=img.tau, inten=img.inten,
                            steps=img.steps, solid=img.solid)
        index[name] = dict(n=n, seed=seed, gen=gen, res_x=rx, res_y=ry, flags=flags, solids=solids,
                           total_steps=img.total_steps, mesh_sha256=mesh_digest(mesh))
        print(name, img.total_steps, int(img.hit.sum()), int(img.solid.sum()))
    # the reference's solid objects (Roche lobe with its donor rotation, sphere) are ~650 K tets:
    # too big to commit, so only their counts and digests are pinned here; tests regenerate them
    # (oracle/_ref where present, else the host generator) and compare digests.
    index["_solids"] = {}
    for D in sorted({c[5]["D"] for c in CASES.values() if c[6]}):
        roche, sphere = ref.solids(D)
        index["_solids"][f"{D:g}"] = dict(
            n_roche=int(roche.shape[0]), n_sphere=int(sphere.shape[0]),
            roche_sha256=hashlib.sha256(roche.tobytes()).hexdigest(),
            sphere_sha256=hashlib.sha256(sphere.tobytes()).hexdigest())
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
