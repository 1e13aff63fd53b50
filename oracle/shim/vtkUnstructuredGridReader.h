// Stand-in for vtkUnstructuredGridReader: a small ASCII legacy-VTK
// ("# vtk DataFile Version x.y") UNSTRUCTURED_GRID parser covering POINTS,
// CELLS (classic "n id id id id" layout), CELL_TYPES and CELL_DATA SCALARS.
// TEST INFRASTRUCTURE ONLY (see vtkSmartPointer.h). Used by the reference at
// object3d_base.cpp:3-11 (SetFileName, SetReadAllScalars, Update, GetOutput).
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vtkUnstructuredGrid.h>

class vtkUnstructuredGridReader {
public:
    void SetFileName(const char* f) { _file = f; }
    void SetReadAllScalars(bool) {}
    // Real VTK hands out a reference-counted object that outlives the reader (the reference relies
    // on that at object3d_base.cpp:3-11); the stand-in simply never frees it.
    vtkUnstructuredGrid* GetOutput() { return _out; }

    void Update() {
        std::ifstream in(_file);
        if (!in) {
            throw std::runtime_error("vtk shim: cannot open " + _file);
        }
        std::string tok;
        long long n_cells = 0;
        while (in >> tok) {
            if (tok == "POINTS") {
                long long n;
                std::string type;
                in >> n >> type;
                _grid.coords.resize(static_cast<size_t>(3 * n));
                for (auto& c : _grid.coords) in >> c;
            } else if (tok == "CELLS") {
                long long total;
                in >> n_cells >> total;
                _grid.offsets.assign(1, 0);
                for (long long c = 0; c < n_cells; c++) {
                    long long k;
                    in >> k;
                    for (long long i = 0; i < k; i++) {
                        long long id;
                        in >> id;
                        _grid.connectivity.push_back(id);
                    }
                    _grid.offsets.push_back(static_cast<long long>(_grid.connectivity.size()));
                }
            } else if (tok == "CELL_TYPES") {
                long long n, t;
                in >> n;
                for (long long c = 0; c < n; c++) in >> t;
            } else if (tok == "SCALARS") {
                std::string name, type, maybe;
                in >> name >> type;
                // optional numComp, then "LOOKUP_TABLE default"
                in >> maybe;
                if (maybe != "LOOKUP_TABLE") in >> maybe;
                in >> maybe; // table name
                auto& arr = _grid.cell_data.arrays[name];
                arr.values.resize(static_cast<size_t>(n_cells));
                for (auto& v : arr.values) in >> v;
            }
        }
    }

private:
    std::string _file;
    vtkUnstructuredGrid* _out = new vtkUnstructuredGrid();
    vtkUnstructuredGrid& _grid = *_out;
};
