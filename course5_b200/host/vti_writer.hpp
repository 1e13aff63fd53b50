// vti_writer.hpp — the .vti the reference exports (object2d.cpp:7-29): vtkImageData with
// dimensions (res_x, res_y, 1), one PointData array "ImageScalars" (utility/screen.py:12 colours
// by ('POINTS', 'ImageScalars', 'Y')), Float64, 2 components {tau, I}, x fastest.
#pragma once

#include <cstddef>
#include <string>

namespace c5host {

// image: res_y * res_x * 2 doubles, x fastest. compress: zlib-compressed appended blocks (what
// VTK's writer does by default) instead of raw appended data; both are standard VTK XML.
void write_vti(const std::string& filename, const double* image, std::size_t res_x, std::size_t res_y,
               bool compress = false);

// Reads back what write_vti (or the oracle's shim writer) wrote: raw-appended or zlib-appended
// Float64 ImageScalars. Used by tests and by `course --compare`.
void read_vti(const std::string& filename, std::size_t& res_x, std::size_t& res_y, std::size_t& comps,
              double*& image_out);

} // namespace c5host
