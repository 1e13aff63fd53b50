"""Synthetic tetrahedral grids for tests and benchmarks.

Every mesh is a conforming Kuhn 6-tet split of an n^3 hexahedral lattice whose
bounding cube (side 1.0) is centred on the accretor position (1, 0, 0)
(``ACC_X0``, /root/reference/project/include/config.hpp:55), so that any view
rotation keeps it inside the reference's hard-coded window
``{2.2, -0.2, 0.9, -0.9}`` (/root/reference/project/src/main.cpp:83): the cube's
half diagonal is 0.866 < 0.9.  Geometry outside the window aborts the reference
(plane.cpp:39-41), so this is a precondition for every parity test.

Vertices are jittered with a counter-based hash RNG (splitmix64 of
``(seed, stream, index)``) so a C++ generator can reproduce the same mesh.  The
jitter amplitude is 0.15 of the local lattice spacing per axis, which keeps the
edge matrix of every Kuhn tet strictly diagonally dominant, i.e. no tet can
invert and the mesh stays a valid, conforming partition (a face-neighbour walk
and the reference's tet-soup scan conversion then see the same geometry).

Cell scalars are named like the reference's VTK arrays: ``AbsorpCoef`` (alpha) and
``radEnLooseRate`` (Q) (object3d_accretion_disk.cpp:4).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

ACC_X0 = 1.0
_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def hash_uniform(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """U[0,1) doubles from (seed, stream, index); 53 random mantissa bits."""
    with np.errstate(over="ignore"):
        key = splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        h = splitmix64(index.astype(np.uint64) ^ key)
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@dataclass
class TetMesh:
    points: np.ndarray   # (n_pts, 3) float64, file frame
    tets: np.ndarray     # (n_tets, 4) int32 vertex ids
    alpha: np.ndarray    # (n_tets,) float64  "AbsorpCoef"
    q: np.ndarray        # (n_tets,) float64  "radEnLooseRate"

    @property
    def n_tets(self) -> int:
        return int(self.tets.shape[0])

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])

    def tet_points(self) -> np.ndarray:
        """(n_tets, 4, 3) private per-tet point copies: the reference's `tetra` layout
        (tetra.hpp:42, filled at object3d_base.cpp:37-51)."""
        return np.ascontiguousarray(self.points[self.tets])

    def signed_volumes(self) -> np.ndarray:
        p = self.points[self.tets]
        e = p[:, 1:] - p[:, :1]
        return np.einsum("ni,ni->n", np.cross(e[:, 0], e[:, 1]), e[:, 2]) / 6.0


_KUHN_PERMS = list(itertools.permutations((0, 1, 2)))


def _lattice_axis(n: int, grade_beta: float) -> np.ndarray:
    u = np.arange(n + 1, dtype=np.float64) / n
    if grade_beta > 0.0:
        u = 0.5 * (1.0 + np.tanh(grade_beta * (2.0 * u - 1.0)) / np.tanh(grade_beta))
        u[0], u[-1] = 0.0, 1.0
    return u


def kuhn_cube(n: int, seed: int, *, jitter: float = 0.15, grade_beta: float = 0.0,
              scalars: str = "uniform", side: float = 1.0,
              centre=(ACC_X0, 0.0, 0.0), carve_sphere: bool = False) -> TetMesh:
    """n^3 lattice, 6 n^3 tets.

    scalars: "uniform"  alpha, Q ~ U(0.1, 4.0)                      (configs C1, C3, C5)
             "sphere"   centroid within r < 0.25 of the centre: alpha = 3.5 (1 +- 0.1 u), Q = 2;
                        else alpha = 0.3 (1 +- 0.1 u), Q = 0.1      (config C2)
             "const"    alpha = 1.5, Q = 0.75                       (known-answer slabs)
    carve_sphere: drop the tets of the inner sphere (config C2b: a cavity, so rays leave and
                  re-enter the mesh).
    """
    m = n + 1
    ax = _lattice_axis(n, grade_beta)
    # local spacing per axis index (min of the two neighbouring intervals)
    d = np.diff(ax)
    hmin = np.minimum(np.concatenate(([d[0]], d)), np.concatenate((d, [d[-1]])))

    ii, jj, kk = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    vid = np.arange(m * m * m, dtype=np.uint64)
    pts = np.empty((m * m * m, 3), dtype=np.float64)
    for c, (idx, stream) in enumerate(((ii, 1), (jj, 2), (kk, 3))):
        u = hash_uniform(seed, stream, vid)
        pts[:, c] = ax[idx] + jitter * hmin[idx] * (2.0 * u - 1.0)
    pts = (pts - 0.5) * side + np.asarray(centre, dtype=np.float64)

    # Kuhn split: for each cube and axis permutation (a,b,c): corner, +e_a, +e_a+e_b, +e_a+e_b+e_c
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = ((ci * m + cj) * m + ck).ravel().astype(np.int64)
    stride = np.array([m * m, m, 1], dtype=np.int64)
    tets = np.empty((base.size, 6, 4), dtype=np.int32)
    for p, perm in enumerate(_KUHN_PERMS):
        v = base.copy()
        tets[:, p, 0] = v
        for s, axis in enumerate(perm):
            v = v + stride[axis]
            tets[:, p, s + 1] = v
    tets = tets.reshape(-1, 4)

    tid = np.arange(tets.shape[0], dtype=np.uint64)
    ua = hash_uniform(seed, 11, tid)
    uq = hash_uniform(seed, 12, tid)
    if scalars == "uniform":
        alpha = 0.1 + 3.9 * ua
        q = 0.1 + 3.9 * uq
    elif scalars == "const":
        alpha = np.full(tets.shape[0], 1.5)
        q = np.full(tets.shape[0], 0.75)
    elif scalars == "sphere":
        cen = pts[tets].mean(axis=1)
        inside = np.linalg.norm(cen - np.asarray(centre), axis=1) < 0.25 * side
        alpha = np.where(inside, 3.5, 0.3) * (1.0 + 0.1 * (2.0 * ua - 1.0))
        q = np.where(inside, 2.0, 0.1)
    else:
        raise ValueError(f"unknown scalars mode {scalars!r}")

    if carve_sphere:
        cen = pts[tets].mean(axis=1)
        keep = np.linalg.norm(cen - np.asarray(centre), axis=1) >= 0.25 * side
        tets, alpha, q = tets[keep], alpha[keep], q[keep]

    return TetMesh(points=pts, tets=np.ascontiguousarray(tets),
                   alpha=np.ascontiguousarray(alpha, dtype=np.float64),
                   q=np.ascontiguousarray(q, dtype=np.float64))


# The named configurations of BASELINE.json / SURVEY.md §8d --------------------------------

CONFIGS = {
    # name: (lattice n, seed, generator kwargs, view flags)
    "C1": dict(n=32, seed=1, gen={}, res=(600, 450),
               view=dict(X=0.0, Y=0.0, D=0.0, I=0.0, alpha_limit=2.5)),
    "C2": dict(n=55, seed=2, gen=dict(scalars="sphere"), res=(1200, 900),
               view=dict(X=0.0, Y=0.0, D=0.0, I=0.0, alpha_limit=2.5)),
    "C2b": dict(n=55, seed=2, gen=dict(scalars="sphere", carve_sphere=True), res=(1200, 900),
                view=dict(X=0.0, Y=0.0, D=0.0, I=0.0, alpha_limit=2.5)),
    "C3": dict(n=110, seed=3, gen={}, res=(2400, 1800),
               view=dict(X=0.5, Y=0.0, D=0.0, I=0.0, alpha_limit=3.0)),
    "C4": dict(n=110, seed=3, gen={}, res=(1200, 900),
               view=dict(X=0.4, Y=0.0, D=0.1, I=-0.03, alpha_limit=2.5)),
    "C5": dict(n=203, seed=5, gen=dict(grade_beta=1.5), res=(4800, 3600),
               view=dict(X=0.4, Y=0.3, D=0.0, I=0.0, alpha_limit=2.5)),
    # north_star's own target workload: "a 2400 x 1800 view of a 50M-tet mesh" (C5's mesh, C3's resolution)
    "C5t": dict(n=203, seed=5, gen=dict(grade_beta=1.5), res=(2400, 1800),
                view=dict(X=0.4, Y=0.3, D=0.0, I=0.0, alpha_limit=2.5)),
}


def make_config(name: str, *, n: int | None = None) -> tuple[TetMesh, dict]:
    """Mesh + view for a named configuration; `n` overrides the lattice size (tests)."""
    cfg = CONFIGS[name]
    mesh = kuhn_cube(n if n is not None else cfg["n"], cfg["seed"], **cfg["gen"])
    return mesh, dict(res_x=cfg["res"][0], res_y=cfg["res"][1], **cfg["view"])


# Legacy VTK writer (what `course -f` reads) ------------------------------------------------

def write_legacy_vtk(path: str, mesh: TetMesh, *, binary: bool = False) -> None:
    """UNSTRUCTURED_GRID, cell type 10, CELL_DATA scalars AbsorpCoef / radEnLooseRate."""
    n_p, n_t = mesh.n_points, mesh.n_tets
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 3.0\ncourse5 synthetic tetrahedral grid\n")
        f.write(b"BINARY\n" if binary else b"ASCII\n")
        f.write(b"DATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {n_p} double\n".encode())
        if binary:
            f.write(mesh.points.astype(">f8").tobytes())
            f.write(b"\n")
        else:
            np.savetxt(f, mesh.points, fmt="%.17g")
        f.write(f"CELLS {n_t} {5 * n_t}\n".encode())
        cells = np.empty((n_t, 5), dtype=np.int32)
        cells[:, 0] = 4
        cells[:, 1:] = mesh.tets
        if binary:
            f.write(cells.astype(">i4").tobytes())
            f.write(b"\n")
        else:
            np.savetxt(f, cells, fmt="%d")
        f.write(f"CELL_TYPES {n_t}\n".encode())
        if binary:
            f.write(np.full(n_t, 10, dtype=">i4").tobytes())
            f.write(b"\n")
        else:
            np.savetxt(f, np.full(n_t, 10, dtype=np.int32), fmt="%d")
        f.write(f"CELL_DATA {n_t}\n".encode())
        for name, arr in (("AbsorpCoef", mesh.alpha), ("radEnLooseRate", mesh.q)):
            f.write(f"SCALARS {name} double 1\nLOOKUP_TABLE default\n".encode())
            if binary:
                f.write(arr.astype(">f8").tobytes())
                f.write(b"\n")
            else:
                np.savetxt(f, arr, fmt="%.17g")
