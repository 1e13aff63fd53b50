// Stand-in for vtkImageData: a dense (X,Y,Z) image with N double components.
// TEST INFRASTRUCTURE ONLY (see vtkSmartPointer.h). Used by the reference at
// object2d.cpp:11-23 (SetDimensions, AllocateScalars, GetDimensions,
// GetScalarPointer).
#pragma once
#include <vector>
#include <vtkSmartPointer.h>

#ifndef VTK_DOUBLE
#define VTK_DOUBLE 11
#endif

class vtkImageData {
public:
    void SetDimensions(int x, int y, int z) {
        _dims[0] = x;
        _dims[1] = y;
        _dims[2] = z;
    }
    void AllocateScalars(int /*type*/, int comps) {
        _comps = comps;
        _data.assign(static_cast<size_t>(_dims[0]) * _dims[1] * _dims[2] * comps, 0.0);
    }
    int* GetDimensions() { return _dims; }
    void* GetScalarPointer(int x, int y, int z) {
        const size_t idx = (static_cast<size_t>(z) * _dims[1] + y) * _dims[0] + x;
        return &_data[idx * _comps];
    }
    int components() const { return _comps; }
    const std::vector<double>& raw() const { return _data; }

private:
    int _dims[3] = {0, 0, 0};
    int _comps = 1;
    std::vector<double> _data;
};
