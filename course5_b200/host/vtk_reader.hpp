// vtk_reader.hpp — legacy-VTK UNSTRUCTURED_GRID reader producing shared points + tet connectivity
// + named per-cell scalars (what c5_upload_mesh takes).
//
// Replaces object3d_base::read_vtk_file (object3d_base.cpp:13-52), which goes through
// vtkUnstructuredGridReader ("VTK library is very slow", main.cpp:99-102), copies the first four
// points of every cell into a private per-tet record (:37-43) and throws the connectivity away.
// This reader maps the file, parses ASCII or big-endian BINARY sections directly, keeps the
// connectivity (the face-neighbour table needs it) and, like the reference, uses the first four
// point ids of each cell and the scalars "AbsorpCoef" / "radEnLooseRate"
// (object3d_accretion_disk.cpp:4).
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace c5host {

struct tet_grid {
    std::vector<double> points;       // xyz per point
    std::vector<int32_t> tets;        // 4 ids per cell
    std::map<std::string, std::vector<double>> cell_scalars;
    std::size_t n_points() const { return points.size() / 3; }
    std::size_t n_tets() const { return tets.size() / 4; }
};

// Throws std::runtime_error with a message naming the file and the problem.
tet_grid read_legacy_vtk(const std::string& filename);

} // namespace c5host
