// Stand-in for vtkUnstructuredGrid: points + tetrahedral cells + cell data.
// TEST INFRASTRUCTURE ONLY (see vtkSmartPointer.h). Used by the reference at
// object3d_base.cpp:14-43 (GetCellData, GetNumberOfCells,
// GetCell(k)->GetPoints()->GetPoint(i)).
#pragma once
#include <vector>
#include <vtkCellData.h>
#include <vtkSmartPointer.h>

class vtkPoints {
public:
    const std::vector<double>* coords = nullptr;
    long long ids[8] = {0};
    double scratch[3] = {0, 0, 0};
    double* GetPoint(long long i) {
        const double* src = coords->data() + 3 * ids[i];
        scratch[0] = src[0];
        scratch[1] = src[1];
        scratch[2] = src[2];
        return scratch;
    }
};

class vtkCell {
public:
    vtkPoints pts;
    vtkPoints* GetPoints() { return &pts; }
};

class vtkUnstructuredGrid {
public:
    std::vector<double> coords;          // xyz per point
    std::vector<long long> connectivity; // cell vertex ids, concatenated
    std::vector<long long> offsets;      // start of each cell in connectivity, plus end
    vtkCellData cell_data;
    vtkCell cell;

    vtkCellData* GetCellData() { return &cell_data; }
    long long GetNumberOfCells() const {
        return offsets.empty() ? 0 : static_cast<long long>(offsets.size()) - 1;
    }
    vtkCell* GetCell(long long k) {
        cell.pts.coords = &coords;
        const long long b = offsets[static_cast<size_t>(k)];
        const long long e = offsets[static_cast<size_t>(k) + 1];
        for (long long i = 0; i < e - b && i < 8; i++) {
            cell.pts.ids[i] = connectivity[static_cast<size_t>(b + i)];
        }
        return &cell;
    }
};
