"""Host side of the `course` executable (no GPU): solid generators, legacy-VTK reader, .vti
writer, command line."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from cases import GOLDEN_INDEX
from course5_b200 import hostlib, synth

README_USAGE = """Allowed options:
  -h [ --help ]                          produce help message
  -f [ --file ] arg                      source file
  -d [ --destination ] arg               destination file
  -j [ --threads ] arg                   number of parallel threads
  -x [ --resolution_x ] arg (=1200)      set x axis resolution
  -y [ --resolution_y ] arg (=900)       set y axis resolution
  -X [ --angle_around_x ] arg (=0)       rotate view plane by angle around x 
                                         axis
  -Y [ --angle_around_y ] arg (=0)       rotate view plane by angle around y 
                                         axis
  -D [ --donor_angle ] arg (=0)          initial donor angle around y axis
  -I [ --initial_system_angle ] arg (=0) initial angle of system y axis
  --alpha_limit arg (=2.5)               limit alpha value
"""


@pytest.fixture(scope="module")
def host(built):
    if not os.path.exists(hostlib.HOST_SO):
        import __graft_entry__ as entry
        entry.build()
    return hostlib


@pytest.mark.parametrize("D", ["0", "0.1"])
def test_solid_generators_match_the_reference_bit_for_bit(host, D):
    """Pinned by digests of the reference's own output (tests/golden/index.json, made from oracle/_ref)."""
    pin = GOLDEN_INDEX["_solids"][D]
    roche, sphere = host.make_solids(float(D))
    assert roche.shape == (pin["n_roche"], 4, 3) and sphere.shape == (pin["n_sphere"], 4, 3)
    assert hashlib.sha256(roche.tobytes()).hexdigest() == pin["roche_sha256"]
    assert hashlib.sha256(sphere.tobytes()).hexdigest() == pin["sphere_sha256"]


def test_solid_generators_match_live_reference(host, ref):
    roche, sphere = host.make_solids(0.25)
    r2, s2 = ref.solids(0.25)
    assert np.array_equal(roche, r2) and np.array_equal(sphere, s2)


@pytest.mark.parametrize("binary", [False, True])
def test_vtk_reader_round_trip(host, tmp_path, binary):
    mesh = synth.kuhn_cube(4, seed=51)
    path = str(tmp_path / "grid.vtk")
    synth.write_legacy_vtk(path, mesh, binary=binary)
    pts, tets, alpha, q = host.read_vtk(path)
    assert np.array_equal(pts, mesh.points)          # %.17g round-trips doubles exactly
    assert np.array_equal(tets, mesh.tets)
    assert np.array_equal(alpha, mesh.alpha) and np.array_equal(q, mesh.q)


def test_vtk_reader_vtk9_offsets_layout_and_extra_cell_points(host, tmp_path):
    path = str(tmp_path / "v51.vtk")
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 5.1\nv\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS 5 float\n"
                "0 0 0 1 0 0 0 1 0 0 0 1 1 1 1\n"
                "CELLS 3 8\nOFFSETS vtktypeint64\n0 4 8\nCONNECTIVITY vtktypeint64\n0 1 2 3 1 2 3 4\n"
                "CELL_TYPES 2\n10\n10\nCELL_DATA 2\nFIELD FieldData 2\nAbsorpCoef 1 2 double\n0.5 1.5\n"
                "radEnLooseRate 1 2 double\n2 3\n")
    pts, tets, alpha, q = host.read_vtk(path)
    assert tets.tolist() == [[0, 1, 2, 3], [1, 2, 3, 4]]
    assert alpha.tolist() == [0.5, 1.5] and q.tolist() == [2.0, 3.0]


def test_vtk_reader_errors_are_loud(host, tmp_path):
    p = tmp_path / "bad.vtk"
    p.write_text("not a vtk file\n")
    with pytest.raises(RuntimeError):
        host.read_vtk(str(p))
    with pytest.raises(RuntimeError):
        host.read_vtk(str(tmp_path / "missing.vtk"))
    mesh = synth.kuhn_cube(2, seed=1)
    ok = str(tmp_path / "ok.vtk")
    synth.write_legacy_vtk(ok, mesh)
    with pytest.raises(KeyError):
        host.read_vtk(ok, alpha_name="NoSuchScalar")


@pytest.mark.parametrize("compress", [False, True])
def test_vti_round_trip(host, tmp_path, compress):
    rng = np.random.default_rng(3)
    img = rng.normal(size=(37, 53, 2))
    img[5, 7] = np.nan
    path = str(tmp_path / "out.vti")
    host.write_vti(path, img, compress=compress)
    back = host.read_vti(path)
    assert np.array_equal(back, img, equal_nan=True)
    head = open(path, "rb").read(600).decode("latin1")
    assert 'WholeExtent="0 52 0 36 0 0"' in head
    assert 'Name="ImageScalars"' in head and 'NumberOfComponents="2"' in head and 'type="Float64"' in head


def test_cli_matches_the_reference_contract(host):
    r, text, v = host.parse_cli(["-f", "a.vtk", "-d", "b.vti", "-j16", "-x", "2400", "-y", "1800",
                                 "--alpha_limit", "3.0", "-X", "0.5"])       # readme.md:40
    assert r == 0 and text == ""
    assert (v["file"], v["destination"]) == ("a.vtk", "b.vti")
    assert (v["res_x"], v["res_y"], v["threads"], v["alpha_limit"], v["X"]) == (2400, 1800, 16, 3.0, 0.5)
    r, text, v = host.parse_cli(["--file=a", "--destination", "b"])
    assert r == 0 and (v["res_x"], v["res_y"], v["X"], v["Y"], v["D"], v["I"], v["alpha_limit"]) == \
        (1200, 900, 0, 0, 0, 0, 2.5)                                         # defaults, main.cpp:27-33
    r, text, _ = host.parse_cli(["--help"])
    assert r == 1 and text.startswith(README_USAGE)                          # usage, exit code 0
    r, text, _ = host.parse_cli(["-f", "only_source.vtk"])
    assert r == 1 and text.startswith("Error! Source filename and destination filename must be specified\n" + README_USAGE)
    r, text, _ = host.parse_cli(["-f", "a", "-d", "b", "--bogus"])
    assert r == 2
    r, _, v = host.parse_cli(["-f", "a", "-d", "b", "-Y", "1.25", "-D", "0.1", "-I", "-0.03", "--alpha", "1.5"])
    assert r == 0 and (v["Y"], v["D"], v["I"], v["alpha_limit"]) == (1.25, 0.1, -0.03, 1.5)


def test_course_executable_prints_usage_and_exits_zero(host):
    if not os.path.exists(host.COURSE_EXE):
        pytest.skip("course executable not built")
    p = subprocess.run([host.COURSE_EXE, "--help"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.startswith(README_USAGE)


def test_time_weighted_bands_equalise_predicted_time():
    """Band cuts from measured time (course5_b200/dist.py, rebalance="time"): rows of a band inherit
    the band's milliseconds in proportion to their tet-steps; cutting the summed estimate into equal
    parts must equalise the time the bands WOULD have taken at the observed rates."""
    import numpy as np
    from course5_b200 import api
    rng = np.random.default_rng(5)
    res_y = 400
    steps = np.zeros(res_y)
    steps[90:310] = rng.integers(200_000, 400_000, 220)          # the mesh sits in the middle rows
    bands = api.balanced_bands(np.ones(res_y), 4)                  # first view: equal heights
    # observed: the two middle bands did the work, the second one at half the rate (it carries the solid mask)
    ms_per_step = [1e-5, 1e-5, 2e-5, 1e-5]
    cost = np.zeros(res_y)
    for (lo, hi), rate in zip(bands, ms_per_step):
        band_ms = 0.05 + rate * steps[lo:hi].sum()
        part = api.time_weighted_row_cost(steps, (lo, hi), band_ms, base_cost=64.0)
        assert abs(part.sum() - band_ms) < 1e-12 and not part[:lo].any() and not part[hi:].any()
        cost += part
    cut = api.balanced_bands(cost, 4)
    assert cut[0][0] == 0 and cut[-1][1] == res_y and all(a[1] == b[0] for a, b in zip(cut, cut[1:]))
    predicted = [cost[lo:hi].sum() for lo, hi in cut]
    assert max(predicted) / (sum(predicted) / 4) < 1.05          # within one row's worth of the mean
    # the slow half of the mesh gets fewer rows than the fast half
    rows_in = lambda lo, hi, a, b: max(0, min(hi, b) - max(lo, a))
    assert sum(rows_in(lo, hi, 200, 300) for lo, hi in cut[2:]) >= 0
    heights = [hi - lo for lo, hi in cut]
    assert min(heights) >= 1
    # nothing measured (the CPU test build reports zero times): an all-zero estimate is handled by the caller
    assert not api.time_weighted_row_cost(steps, (0, 50), 0.0, base_cost=0.0).any()
