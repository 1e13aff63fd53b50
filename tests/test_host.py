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


FIXTURES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vtk")


@pytest.mark.parametrize("name", ["ascii_v42_float_lookup.vtk", "binary_v30_float_int.vtk", "binary_v51_int64.vtk"])
def test_vtk_reader_against_fixtures_written_from_the_format_specification(host, name):
    """Files made by hand from the VTK file-formats document (not by this repo's writer): float points,
    vtktypeint64 OFFSETS / CONNECTIVITY, custom and default LOOKUP_TABLEs, POINT_DATA / VECTORS /
    METADATA sections to skip, a 5-point cell of which the first four points count
    (object3d_base.cpp:39-42), scalars as SCALARS and as FIELD arrays. Expected values typed here."""
    pts, tets, alpha, q = host.read_vtk(os.path.join(FIXTURES, name))
    f32 = lambda v: float(np.float32(v))
    assert pts.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1],
                            [-0.5, 0.225 if name.startswith("ascii") else f32(0.225), 3.0]]
    assert tets.tolist() == [[0, 1, 2, 3], [1, 2, 3, 4], [4, 3, 2, 1]]
    assert alpha.tolist() == [0.25, 1.5, 4.0]
    assert q.tolist() == [1e-3, 2.0, 3.0000000000000004]


def test_binary_fixtures_are_what_their_script_writes():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_binary_fixtures", os.path.join(FIXTURES, "make_binary_fixtures.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert open(os.path.join(FIXTURES, "binary_v30_float_int.vtk"), "rb").read() == mod.v30()
    assert open(os.path.join(FIXTURES, "binary_v51_int64.vtk"), "rb").read() == mod.v51()


@pytest.mark.parametrize("damage", ["negative_offset", "decreasing_offset", "offset_past_end", "huge_count",
                                    "truncated_binary", "point_id_out_of_range", "char_type"])
def test_vtk_reader_rejects_malformed_files(host, tmp_path, damage):
    """A malformed or truncated file is an error message, never an out-of-bounds read."""
    head = "# vtk DataFile Version 5.1\nv\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS 5 float\n0 0 0 1 0 0 0 1 0 0 0 1 1 1 1\n"
    tail = "CELL_TYPES 2\n10\n10\nCELL_DATA 2\nFIELD FieldData 2\nAbsorpCoef 1 2 double\n0.5 1.5\nradEnLooseRate 1 2 double\n2 3\n"
    cells = {"negative_offset": "CELLS 3 8\nOFFSETS vtktypeint64\n-4 0 4\nCONNECTIVITY vtktypeint64\n0 1 2 3 1 2 3 4\n",
             "decreasing_offset": "CELLS 3 8\nOFFSETS vtktypeint64\n4 0 8\nCONNECTIVITY vtktypeint64\n0 1 2 3 1 2 3 4\n",
             "offset_past_end": "CELLS 3 8\nOFFSETS vtktypeint64\n0 4 12\nCONNECTIVITY vtktypeint64\n0 1 2 3 1 2 3 4\n",
             "huge_count": "CELLS 3 9000000000000000000\nOFFSETS vtktypeint64\n0 4 8\nCONNECTIVITY vtktypeint64\n0 1 2 3 1 2 3 4\n",
             "point_id_out_of_range": "CELLS 3 8\nOFFSETS vtktypeint64\n0 4 8\nCONNECTIVITY vtktypeint64\n0 1 2 3 1 2 3 4294967297\n"}
    path = tmp_path / "bad.vtk"
    if damage in cells:
        path.write_text(head + cells[damage] + tail)
    elif damage == "truncated_binary":
        good = open(os.path.join(FIXTURES, "binary_v51_int64.vtk"), "rb").read()
        path.write_bytes(good[: good.index(b"CONNECTIVITY") + 40])
    else:  # a 1-byte integer type where ids are expected: read as 1-byte values (no over-read), then rejected by content
        path.write_bytes(b"# vtk DataFile Version 5.1\nv\nBINARY\nDATASET UNSTRUCTURED_GRID\nPOINTS 2 float\n" + bytes(24) +
                         b"\nCELLS 2 4\nOFFSETS char\n" + bytes([0, 4]) + b"\nCONNECTIVITY char\n" + bytes([0, 1, 1, 9]) + b"\n")
    with pytest.raises(RuntimeError) as e:
        host.read_vtk(str(path))
    assert "bad.vtk" in str(e.value)


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


def _decode_vti_independently(path):
    """The .vti decoded with the standard library only (xml.etree, base64, zlib) following the VTK XML
    format: header_type UInt64; appended data after '_'; uncompressed = [n_bytes][data]; compressed =
    [n_blocks][block_size][last_block_size][compressed sizes...] then the zlib blocks; with
    encoding="base64" the table and the data are two separate base64 streams."""
    import base64
    import struct
    import xml.etree.ElementTree as ET
    import zlib
    raw = open(path, "rb").read()
    cut = raw.index(b"<AppendedData")
    start = raw.index(b"_", raw.index(b">", cut)) + 1
    end = raw.rindex(b"</AppendedData>")
    root = ET.fromstring(raw[:cut] + b"</VTKFile>")             # the XML part parses as XML
    assert root.tag == "VTKFile" and root.attrib["type"] == "ImageData" and root.attrib["byte_order"] == "LittleEndian"
    assert root.attrib["header_type"] == "UInt64"
    image = root.find("ImageData")
    x0, x1, y0, y1, z0, z1 = (int(v) for v in image.attrib["WholeExtent"].split())
    piece = image.find("Piece")
    assert piece.attrib["Extent"] == image.attrib["WholeExtent"]
    pd = piece.find("PointData")
    assert pd.attrib["Scalars"] == "ImageScalars"
    arr = pd.find("DataArray")
    assert (arr.attrib["type"], arr.attrib["Name"], arr.attrib["NumberOfComponents"], arr.attrib["format"],
            arr.attrib["offset"]) == ("Float64", "ImageScalars", "2", "appended", "0")
    encoding = raw[cut: raw.index(b">", cut)].decode().split('encoding="')[1].split('"')[0]
    payload = raw[start:end]
    compressed = root.attrib.get("compressor") == "vtkZLibDataCompressor"
    n_values = (x1 - x0 + 1) * (y1 - y0 + 1) * 2
    if encoding == "base64":
        text = payload.strip()

        def stream(at, n_bytes):                                  # one padded base64 stream of n_bytes
            n_chars = (n_bytes + 2) // 3 * 4
            return base64.b64decode(text[at: at + n_chars])[:n_bytes], at + n_chars
    if not compressed:
        assert encoding == "raw"
        (n_bytes,) = struct.unpack("<Q", payload[:8])
        data = payload[8: 8 + n_bytes]
    else:
        if encoding == "raw":
            n_blocks, block, last = struct.unpack("<3Q", payload[:24])
            sizes = struct.unpack(f"<{n_blocks}Q", payload[24: 24 + 8 * n_blocks])
            body = payload[24 + 8 * n_blocks:]
        else:
            first, _ = stream(0, 24)
            n_blocks, block, last = struct.unpack("<3Q", first)
            table, at = stream(0, 24 + 8 * n_blocks)
            sizes = struct.unpack(f"<{n_blocks}Q", table[24:])
            body, _ = stream(at, sum(sizes))
        data, at = b"", 0
        for b, sz in enumerate(sizes):
            chunk = zlib.decompress(body[at: at + sz])
            assert len(chunk) == (last if (b == n_blocks - 1 and last) else block)
            data += chunk
            at += sz
    assert len(data) == 8 * n_values
    return np.frombuffer(data, dtype="<f8").reshape(y1 - y0 + 1, x1 - x0 + 1, 2)


@pytest.mark.parametrize("mode", ["raw", "zlib", "zlib_base64"])
def test_vti_writer_against_the_vtk_xml_format(host, tmp_path, mode):
    """What the writer produces, decoded by an independent reader built from the format description
    (not by this repo's read_vti): layout (x fastest, ImageScalars, 2 components,
    object2d.cpp:11-21), NaNs, and all three encodings — zlib_base64 is what the reference's
    vtkXMLImageDataWriter writes by default."""
    rng = np.random.default_rng(3)
    img = rng.normal(size=(37, 53, 2))
    img[5, 7] = np.nan
    big = rng.normal(size=(300, 260, 2))           # > 1 MiB: more than one compressed block
    for k, im in enumerate((img, big)):
        path = str(tmp_path / f"out{k}.vti")
        host.write_vti(path, im, compress=mode != "raw", base64=mode == "zlib_base64")
        back = _decode_vti_independently(path)
        assert np.array_equal(back, im, equal_nan=True)
        assert back[0, 1, 0] == im[0, 1, 0] and back[1, 0, 1] == im[1, 0, 1]      # x fastest, component last
        assert np.array_equal(host.read_vti(path), im, equal_nan=True)          # and the repo's own reader agrees
    head = open(str(tmp_path / "out0.vti"), "rb").read(600).decode("latin1")
    assert 'WholeExtent="0 52 0 36 0 0"' in head


@pytest.mark.parametrize("b64", [False, True])
def test_vti_block_compressor_threads_do_not_change_the_file(host, tmp_path, b64, monkeypatch):
    """The compressed blocks are independent zlib streams: deflated by one thread or by many, the file
    is the same, byte for byte (vti_writer.cpp compressor_threads)."""
    rng = np.random.default_rng(5)
    img = rng.normal(size=(700, 500, 2)).round(3)        # 5.6 MB: six blocks, the last one short
    img[3, 4] = np.nan
    files = []
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("C5_VTI_THREADS", threads)
        path = str(tmp_path / f"t{threads}.vti")
        host.write_vti(path, img, compress=True, base64=b64)
        files.append(open(path, "rb").read())
    assert files[0] == files[1] == files[2]
    assert np.array_equal(_decode_vti_independently(str(tmp_path / "t16.vti")), img, equal_nan=True)


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
