// solids.cpp — see solids.hpp. Host-side, one-off (< 1 s); only the resulting NaN mask is on the
// hot path (c5_exact.cu: solid_mask).
#include "solids.hpp"

#include <cmath>
#include <limits>

#include "config.hpp"

namespace c5host {

namespace {

double norm2(const point& v) { // vector_2_norm, object3d_roche_lobe.cpp:3-9
    double sum{};
    for (double c : v) sum += c * c;
    return std::sqrt(sum);
}

point cross(const point& a, const point& b) { // vector_multiplication, object3d_roche_lobe.cpp:11-18
    return {a[1] * b[2] - a[2] * b[1], -(a[0] * b[2]) + (a[2] * b[0]), a[0] * b[1] - a[1] * b[0]};
}

point turn_about_z(const point& v, double a) { // object3d_base.cpp:66-75
    return {v[0] * std::cos(a) + v[1] * std::sin(a), -v[0] * std::sin(a) + v[1] * std::cos(a), v[2]};
}

point turn_about_y(const point& v, double a) { // object3d_base.cpp:55-64
    return {v[0] * std::cos(a) + v[2] * std::sin(a), v[1], -v[0] * std::sin(a) + v[2] * std::cos(a)};
}

// marches from `from` in steps of `dir` until the potential reaches the level (do-while: the
// first sample is one step away from the start) — object3d_base.cpp:103-107,132-136
point march(const std::function<double(const point&)>& f, point from, const point& dir, double level) {
    double value{};
    do {
        for (int c = 0; c < 3; c++) from[c] += dir[c];
        value = f(from);
    } while (value < level);
    return from;
}

} // namespace

void rotate_point_x(point& p, double angle) {
    const double y = p[1];
    p[1] = p[1] * std::cos(angle) - p[2] * std::sin(angle);
    p[2] = y * std::sin(angle) + p[2] * std::cos(angle);
}

void rotate_point_y(point& p, double angle, double x0) {
    p[0] -= x0;
    const double x = p[0];
    p[0] = p[0] * std::cos(angle) - p[2] * std::sin(angle);
    p[2] = x * std::sin(angle) + p[2] * std::cos(angle);
    p[0] += x0;
}

std::vector<tet_points> init_polar(const std::function<double(const point&)>& potential, double x0, double y0,
                                   double z0, double level_value, double step, double angle_step) {
    const double eps = std::numeric_limits<double>::epsilon();
    const point center{x0, y0, z0};
    const point unit_ray{0.001, 0, 0};
    const double dangle = PI / angle_step;

    const point top = march(potential, center, {0, 0, step}, level_value);
    const point bottom = march(potential, center, {0, 0, -step}, level_value);

    // rings[latitude][longitude]; both angles advance by repeated addition
    std::vector<std::vector<point>> rings;
    double lat = -PI + dangle;
    double lon = 0;
    while (lat < (PI - dangle + eps)) {
        const point lat_dir = turn_about_z(unit_ray, lat);
        lat += dangle;
        std::vector<point> ring;
        while (lon < 2 * PI - dangle + eps) {
            ring.push_back(march(potential, center, turn_about_y(lat_dir, lon), level_value));
            lon += dangle;
        }
        rings.push_back(std::move(ring));
        lon = 0;
    }

    std::vector<tet_points> out;
    const std::size_t n_lon = rings[0].size();
    const std::size_t last = rings.size() - 1;
    out.reserve(2 * n_lon * (last + 1));
    auto emit = [&](const point& a, const point& b, const point& c) { out.push_back({center, a, b, c}); };

    // bottom cap (object3d_base.cpp:156-161)
    for (std::size_t i = 1; i < n_lon; i++) emit(bottom, rings[0][i], rings[0][i - 1]);
    emit(bottom, rings[0][0], rings[0][n_lon - 1]);
    // top cap; its closing tet takes ring 0's last point, as in the reference (:163-174)
    for (std::size_t i = 1; i < n_lon; i++) emit(top, rings[last][i], rings[last][i - 1]);
    emit(top, rings[last][0], rings[0][n_lon - 1]);
    // bands between consecutive rings, two tets per quad (:176-193)
    for (std::size_t i = 1; i <= last; i++) {
        const auto& lo = rings[i - 1];
        const auto& hi = rings[i];
        for (std::size_t j = 1; j < n_lon; j++) {
            emit(lo[j - 1], lo[j], hi[j - 1]);
            emit(hi[j - 1], hi[j], lo[j]);
        }
        emit(lo[n_lon - 1], lo[0], hi[n_lon - 1]);
        emit(hi[n_lon - 1], hi[0], lo[0]);
    }
    return out;
}

std::vector<tet_points> make_roche_lobe(const point& pos_accretor, double dist, double donor_angle_around_y,
                                        double m_accretor, double m_donor, double def_omega) {
    const double donor_pos_x = pos_accretor[0] - dist;
    const double mass_center_pos_x =
        (donor_pos_x * m_donor + pos_accretor[0] * m_accretor) / (m_accretor + m_donor);
    // the reference computes the analytic L1 and then overrides it (object3d_roche_lobe.cpp:25-30)
    const double lagrange1_pos_x = 0.35515;

    auto potential = [&](const point& r) -> double {
        const double acc_den = norm2({r[0] - pos_accretor[0], r[1], r[2]});
        const double donor_den = norm2({r[0] - donor_pos_x, r[1], r[2]});
        const double w = norm2(cross({r[0] - mass_center_pos_x, r[1], r[2]}, {0, def_omega, 0}));
        const double centrifugal = (1. / 2.) * w * w;
        // G_SOL is long double: the sum is formed in extended precision and rounded once (:42)
        const long double F = -((G_SOL * m_accretor) / acc_den) - ((G_SOL * m_donor) / donor_den) - centrifugal;
        return static_cast<double>(F);
    };

    const double level = potential({lagrange1_pos_x, 0, 0});
    std::vector<tet_points> tets = init_polar(potential, donor_pos_x, 0, 0, level, 0.001, 128);
    for (auto& t : tets) {
        for (auto& p : t) rotate_point_y(p, donor_angle_around_y, ACC_X0);
    }
    return tets;
}

std::vector<tet_points> make_sphere(const point& center, double R) {
    auto distance = [&](const point& p) { return norm2({p[0] - center[0], p[1] - center[1], p[2] - center[2]}); };
    return init_polar(distance, center[0], center[1], center[2], R, 0.001, 256);
}

} // namespace c5host
