// Stand-in header: the reference includes <vtkCellArray.h> (object3d_base.hpp:11)
// but uses nothing from it. TEST INFRASTRUCTURE ONLY.
#pragma once
