"""Writes the BINARY legacy-VTK fixtures byte by byte from the VTK file-formats document (big-endian
payloads after each header line, one newline after each payload) — independently of
course5_b200.synth.write_legacy_vtk, whose files the reader is also tested with.

    python tests/golden/vtk/make_binary_fixtures.py      # rewrites the two .vtk files next to it

binary_v30_float_int.vtk   what VTK < 9 writes: "CELLS n size" with int32 [count, ids...] records,
                           float POINTS, float and double SCALARS with LOOKUP_TABLE default
binary_v51_int64.vtk       what VTK >= 9 writes (file version 5.1): "CELLS n_offsets n_connectivity",
                           OFFSETS / CONNECTIVITY as vtktypeint64, a METADATA block after POINTS, the
                           cell scalars as FIELD arrays
Both hold the same three cells as ascii_v42_float_lookup.vtk (two tets and a 5-point pyramid of which
only the first four points count, object3d_base.cpp:39-42).
"""
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
POINTS = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1), (-0.5, 0.225, 3.0)]
CELLS = [(0, 1, 2, 3), (1, 2, 3, 4), (4, 3, 2, 1, 5)]
TYPES = [10, 10, 14]
ALPHA = [0.25, 1.5, 4.0]
Q = [1e-3, 2.0, 3.0000000000000004]


def be(fmt, values):
    return struct.pack(">" + fmt * len(values), *values)


def v30():
    out = b"# vtk DataFile Version 3.0\nbinary, VTK < 9 layout\nBINARY\nDATASET UNSTRUCTURED_GRID\n"
    out += b"POINTS 6 float\n" + be("f", [c for p in POINTS for c in p]) + b"\n"
    flat = [v for c in CELLS for v in (len(c), *c)]
    out += f"CELLS {len(CELLS)} {len(flat)}\n".encode() + be("i", flat) + b"\n"
    out += b"CELL_TYPES 3\n" + be("i", TYPES) + b"\n"
    out += b"CELL_DATA 3\nSCALARS AbsorpCoef float 1\nLOOKUP_TABLE default\n" + be("f", ALPHA) + b"\n"
    out += b"SCALARS radEnLooseRate double\nLOOKUP_TABLE default\n" + be("d", Q) + b"\n"
    return out


def v51():
    out = b"# vtk DataFile Version 5.1\nbinary, VTK 9 layout\nBINARY\nDATASET UNSTRUCTURED_GRID\n"
    out += b"POINTS 6 float\n" + be("f", [c for p in POINTS for c in p]) + b"\n"
    out += b"METADATA\nINFORMATION 0\n\n"
    offsets, conn = [0], []
    for c in CELLS:
        conn += list(c)
        offsets.append(len(conn))
    out += f"CELLS {len(offsets)} {len(conn)}\n".encode()
    out += b"OFFSETS vtktypeint64\n" + be("q", offsets) + b"\n"
    out += b"CONNECTIVITY vtktypeint64\n" + be("q", conn) + b"\n"
    out += b"CELL_TYPES 3\n" + be("i", TYPES) + b"\n"
    out += b"CELL_DATA 3\nFIELD FieldData 2\n"
    out += b"AbsorpCoef 1 3 double\n" + be("d", ALPHA) + b"\n"
    out += b"radEnLooseRate 1 3 double\n" + be("d", Q) + b"\n"
    return out


if __name__ == "__main__":
    for name, data in (("binary_v30_float_int.vtk", v30()), ("binary_v51_int64.vtk", v51())):
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(data)
        print(name, len(data), "bytes")
