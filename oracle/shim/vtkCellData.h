// Stand-in for vtkCellData / vtkDataArray: named per-cell scalar arrays.
// TEST INFRASTRUCTURE ONLY (see vtkSmartPointer.h). Used by the reference at
// object3d_base.cpp:17-28,48 (GetScalars(name)->GetTuple(k)).
#pragma once
#include <map>
#include <string>
#include <vector>

class vtkDataArray {
public:
    std::vector<double> values;
    double* GetTuple(long long k) { return &values[static_cast<size_t>(k)]; }
};

class vtkCellData {
public:
    std::map<std::string, vtkDataArray> arrays;
    vtkDataArray* GetScalars(const char* name) {
        auto it = arrays.find(name);
        return it == arrays.end() ? nullptr : &it->second;
    }
};
