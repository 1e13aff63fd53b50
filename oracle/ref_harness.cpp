// Parity-oracle harness around the UNMODIFIED reference translation units.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product (course5_b200/, the `course`
// CLI, libc5gpu.so) may link, import or execute this file or its output.
// Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline
// legs load oracle/_ref/libc5ref.so.
//
// The recipe in oracle/Makefile compiles this file together with
// /root/reference/project/src/{plane,line,tetra,object3d_base,object3d_sphere,
// object3d_roche_lobe,object3d_accretion_disk,object2d}.cpp, in place, with the
// reference's own flags (-std=c++17 -O3 -fopenmp, CMakeLists.txt:4-7) plus
// -Dnone=shared (GCC >= 9 rejects `default(none)` at plane.cpp:161,187 and
// object3d_base.cpp:205,214,237 because const locals are not listed) and the
// VTK stand-in headers in oracle/shim/. main.cpp is not compiled (Boost is
// absent); this harness replays main.cpp:93-137 from the same flag values.
//
// What it exposes beyond the reference's own output: the PRE-float-cast
// doubles of both channels (the reference casts at plane.cpp:165-166), the
// per-pixel record count (line::number_of_intersections, the "tet-steps"
// unit, plane.cpp:3-12) and the solid mask (line.cpp:246-249).

#include <algorithm>
#include <array>
#include <bitset>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <omp.h>

// Open up the reference classes so the harness can (i) fill object3d_base::_data
// without a file, (ii) read plane::_lines and drive the public `line` methods
// itself (reproducing plane.cpp:162-168 without the float cast).
#define private public
#define protected public
#include <config.hpp>
#include <line.hpp>
#include <object2d.hpp>
#include <object3d_accretion_disk.hpp>
#include <object3d_base.hpp>
#include <object3d_roche_lobe.hpp>
#include <object3d_sphere.hpp>
#include <plane.hpp>
#include <tetra.hpp>
#undef private
#undef protected

namespace {

using clk = std::chrono::steady_clock;

double seconds_since(clk::time_point t0) {
    return std::chrono::duration<double>(clk::now() - t0).count();
}

// main.cpp:96,105-107 / 112-114: the three view rotations applied to an object.
void apply_view_rotations(object3d_base& obj, double X, double Y, double I) {
    const double a0 = -I * PI + PI / 2.;
    obj.rotate_around_x_axis(a0);
    obj.rotate_around_y_axis(Y * PI, ACC_X0);
    obj.rotate_around_x_axis(-a0 + X * PI);
}

void fill_grid(object3d_base& obj, const double* tet_pts, const double* alpha, const double* q,
               long long n_tets) {
    obj._data->reserve(static_cast<size_t>(n_tets));
    for (long long t = 0; t < n_tets; t++) {
        std::array<std::array<double, 3>, 4> p{};
        for (int v = 0; v < 4; v++) {
            for (int c = 0; c < 3; c++) {
                p[v][c] = tet_pts[(t * 4 + v) * 3 + c];
            }
        }
        obj._data->emplace_back(p, alpha[t], q[t], tetra_type::transparent);
    }
}

void fill_solids(object3d_base& obj, const double* pts, long long n) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    obj._data->reserve(static_cast<size_t>(n));
    for (long long t = 0; t < n; t++) {
        std::array<std::array<double, 3>, 4> p{};
        for (int v = 0; v < 4; v++) {
            for (int c = 0; c < 3; c++) {
                p[v][c] = pts[(t * 4 + v) * 3 + c];
            }
        }
        obj._data->emplace_back(p, nan, 0., tetra_type::solid);
    }
}

long long dump_object(object3d_base& obj, double* out, long long cap_tets, long long at) {
    const auto& v = *obj._data;
    for (size_t t = 0; t < v.size(); t++) {
        if (at + static_cast<long long>(t) >= cap_tets) break;
        for (int k = 0; k < 4; k++) {
            for (int c = 0; c < 3; c++) {
                out[((at + t) * 4 + k) * 3 + c] = v[t][k][c];
            }
        }
    }
    return static_cast<long long>(v.size());
}

} // namespace

extern "C" {

// Number of solid tets the reference generates (Roche lobe, then sphere:
// the object order of main.cpp:127) and, if `out` is non-null, their points in
// the PRE-view frame (the Roche lobe already carries its donor rotation,
// object3d_roche_lobe.cpp:48). out is [n][4][3] doubles, cap_tets its capacity.
long long c5ref_solids(double donor_angle_pi, double* out, long long cap_tets, long long* n_roche,
                       long long* n_sphere) {
    auto roche = object3d_roche_lobe{{ACC_X0, ACC_Y0, ACC_Z0}, L, donor_angle_pi * PI, M_ACC, M_DONOR, OMEGA};
    auto sphere = object3d_sphere{{ACC_X0, ACC_Y0, ACC_Z0}, ACC_DISK_R};
    const long long nr = static_cast<long long>(roche._data->size());
    const long long ns = static_cast<long long>(sphere._data->size());
    if (n_roche) *n_roche = nr;
    if (n_sphere) *n_sphere = ns;
    if (out) {
        dump_object(roche, out, cap_tets, 0);
        dump_object(sphere, out, cap_tets, nr);
    }
    return nr + ns;
}

// Replays main.cpp:93-130.
//   tet_pts [n_tets][4][3], alpha/q [n_tets]  : the grid ("accretion disk") in the file frame
//   solids  : 0 = none, 1 = reference's own Roche lobe + sphere, 2 = caller-supplied soup
//   mode    : 0 = the reference's timed flow (plane ctor + find_intersections +
//                 trace_rays, values are float-rounded like plane.cpp:165-166)
//             1 = raw: same flow but the per-pixel loop is driven from here so the
//                 doubles are captured before the float cast
//   tau/inten [res_y][res_x] (x fastest, the .vti order of object2d.cpp:17-21)
//   steps     [res_y][res_x] records per pixel (may be null)
//   solid     [res_y][res_x] 1 where line::_marked_solid (may be null)
//   timings   [4] seconds: rotations, plane ctor, find_intersections, trace_rays
// Returns 0, or -1 on bad arguments. Reference exceptions are thrown inside
// OpenMP regions and terminate the process (plane.cpp:39-41): keep inputs
// inside the window.
int c5ref_render(const double* tet_pts, const double* alpha, const double* q, long long n_tets,
                 int solids, const double* solid_pts, long long n_solid, int res_x, int res_y, double X,
                 double Y, double D, double I, double alpha_limit, int threads, int mode, double* tau,
                 double* inten, uint32_t* steps, uint8_t* solid, double* timings,
                 unsigned long long* total_steps) {
    if (!tet_pts || n_tets <= 0 || res_x < 2 || res_y < 2 || threads < 1 || threads > MAX_NUMBER_OF_THREADS) {
        return -1;
    }
    app::instance().config.limit_alpha_value = alpha_limit;
    app::instance().config.threads = threads;
    omp_set_num_threads(threads);

    std::vector<double> domain = {2.2, -0.2, 0.9, -0.9}; // main.cpp:83

    auto t0 = clk::now();
    object3d_base grid{};
    fill_grid(grid, tet_pts, alpha, q, n_tets);
    std::vector<object3d_base> objects;
    t0 = clk::now();
    apply_view_rotations(grid, X, Y, I);
    objects.push_back(grid);
    if (solids == 1) {
        auto roche = object3d_roche_lobe{{ACC_X0, ACC_Y0, ACC_Z0}, L, D * PI, M_ACC, M_DONOR, OMEGA};
        apply_view_rotations(roche, X, Y, I);
        auto sphere = object3d_sphere{{ACC_X0, ACC_Y0, ACC_Z0}, ACC_DISK_R};
        objects.push_back(roche);
        objects.push_back(sphere); // main.cpp:116: the sphere is NOT rotated
    } else if (solids == 2 && solid_pts && n_solid > 0) {
        object3d_base soup{};
        fill_solids(soup, solid_pts, n_solid);
        apply_view_rotations(soup, X, Y, I);
        objects.push_back(soup);
    }
    const double t_rot = seconds_since(t0);

    t0 = clk::now();
    plane base_plane{static_cast<size_t>(res_x), static_cast<size_t>(res_y), objects, domain};
    const double t_ctor = seconds_since(t0);

    t0 = clk::now();
    base_plane.find_intersections();
    const double t_find = seconds_since(t0);

    // per-pixel bookkeeping the reference itself never exports (outside the timers)
    unsigned long long sum = 0;
    for (int i = 0; i < res_x; i++) {
        for (int j = 0; j < res_y; j++) {
            auto& ln = base_plane._lines[i][j];
            const size_t n = ln._marked_solid ? 0 : ln.number_of_intersections();
            sum += n;
            if (steps) steps[static_cast<size_t>(j) * res_x + i] = static_cast<uint32_t>(n);
            if (solid) solid[static_cast<size_t>(j) * res_x + i] = ln._marked_solid ? 1 : 0;
        }
    }
    if (total_steps) *total_steps = sum;

    t0 = clk::now();
    if (mode == 0) {
        object2d result = base_plane.trace_rays(tetra_value::alpha, tetra_value::Q);
        const double t_trace = seconds_since(t0);
        if (timings) {
            timings[0] = t_rot;
            timings[1] = t_ctor;
            timings[2] = t_find;
            timings[3] = t_trace;
        }
        for (int i = 0; i < res_x; i++) {
            for (int j = 0; j < res_y; j++) {
                if (tau) tau[static_cast<size_t>(j) * res_x + i] = result._object2d_data.first[i][j];
                if (inten) inten[static_cast<size_t>(j) * res_x + i] = result._object2d_data.second[i][j];
            }
        }
        return 0;
    }

    // mode 1: plane.cpp:161-169 with the casts removed
    auto& data = *base_plane._data;
#pragma omp parallel for schedule(dynamic, 8) collapse(2)
    for (int i = 0; i < res_x; i++) {
        for (int j = 0; j < res_y; j++) {
            auto& ln = base_plane._lines[i][j];
            ln.calculate_intersections(data);
            const double a = ln.direct_calculate_ray_value(data, tetra_value::alpha);
            const double b = ln.integrate_ray_value_by_i(data, tetra_value::alpha, tetra_value::Q);
            ln.free_memory();
            if (tau) tau[static_cast<size_t>(j) * res_x + i] = a;
            if (inten) inten[static_cast<size_t>(j) * res_x + i] = b;
        }
    }
    if (timings) {
        timings[0] = t_rot;
        timings[1] = t_ctor;
        timings[2] = t_find;
        timings[3] = seconds_since(t0);
    }
    return 0;
}

// The accumulated pixel coordinates the reference builds at plane.cpp:304-314.
void c5ref_pixel_coords(int res_x, int res_y, double* xs, double* ys) {
    const double x_max = 2.2, x_min = -0.2, y_max = 0.9, y_min = -0.9;
    const double step_x = (x_max - x_min) / (res_x - 1.);
    const double step_y = (y_max - y_min) / (res_y - 1.);
    double cx = x_min;
    for (int i = 0; i < res_x; i++) {
        xs[i] = cx;
        cx = cx + step_x;
    }
    double cy = y_min;
    for (int j = 0; j < res_y; j++) {
        ys[j] = cy;
        cy = cy + step_y;
    }
}

// Full file-to-file replay (read_vtk_file via the shim reader, export_to_vti via
// the shim writer): main.cpp:96-137 minus Boost and the banner.
int c5ref_run_files(const char* src_vtk, const char* dst_vti, int res_x, int res_y, double X, double Y,
                    double D, double I, double alpha_limit, int threads) {
    app::instance().config.limit_alpha_value = alpha_limit;
    omp_set_num_threads(threads);
    std::vector<double> domain = {2.2, -0.2, 0.9, -0.9};
    object3d_accretion_disk acc_disk{std::string(src_vtk)};
    apply_view_rotations(acc_disk, X, Y, I);
    auto roche = object3d_roche_lobe{{ACC_X0, ACC_Y0, ACC_Z0}, L, D * PI, M_ACC, M_DONOR, OMEGA};
    apply_view_rotations(roche, X, Y, I);
    auto sphere = object3d_sphere{{ACC_X0, ACC_Y0, ACC_Z0}, ACC_DISK_R};
    plane base_plane{static_cast<size_t>(res_x), static_cast<size_t>(res_y), {acc_disk, roche, sphere}, domain};
    base_plane.find_intersections();
    object2d result = base_plane.trace_rays(tetra_value::alpha, tetra_value::Q);
    result.export_to_vti(dst_vti);
    return 0;
}

} // extern "C"
