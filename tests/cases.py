"""Helpers shared by the CPU and GPU test modules."""
import hashlib
import json
import os

import numpy as np

from course5_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

with open(os.path.join(GOLDEN, "index.json")) as _f:
    GOLDEN_INDEX = json.load(_f)
GOLDEN_CASES = sorted(k for k in GOLDEN_INDEX if not k.startswith("_"))


def golden_case(name):
    """(mesh, meta, arrays) of a committed fixture; checks the regenerated mesh's digest."""
    meta = GOLDEN_INDEX[name]
    mesh = synth.kuhn_cube(meta["n"], meta["seed"], **meta["gen"])
    h = hashlib.sha256()
    for a in (mesh.points, mesh.tets, mesh.alpha, mesh.q):
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest() == meta["mesh_sha256"], "synthetic mesh generator drifted from the golden fixtures"
    arrays = np.load(os.path.join(GOLDEN, name + ".npz"))
    return mesh, meta, arrays


_SOLIDS_CACHE = {}


def reference_solids(D):
    """(roche, sphere) tet points in the pre-view frame: the reference's Roche lobe and sphere, made
    by the host-side generator (course5_b200/host/solids.cpp) and digest-checked against the pins
    taken from the reference's own generator (tests/golden/index.json)."""
    key = f"{D:g}"
    if key not in _SOLIDS_CACHE:
        from course5_b200 import hostlib
        roche, sphere = hostlib.make_solids(float(D))
        pin = GOLDEN_INDEX["_solids"].get(key)
        if pin:
            assert hashlib.sha256(roche.tobytes()).hexdigest() == pin["roche_sha256"]
            assert hashlib.sha256(sphere.tobytes()).hexdigest() == pin["sphere_sha256"]
        _SOLIDS_CACHE[key] = (roche, sphere)
    return _SOLIDS_CACHE[key]


def view_kwargs(meta):
    f = meta["flags"]
    return dict(X=f["X"], Y=f["Y"], I=f["I"], alpha_limit=f["alpha_limit"])
