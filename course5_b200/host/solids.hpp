// solids.hpp — the two procedurally generated solid objects the reference always renders next to
// the grid: the donor's Roche lobe and the accretor sphere (main.cpp:110-116).
#pragma once

#include <array>
#include <functional>
#include <vector>

namespace c5host {

using point = std::array<double, 3>;
using tet_points = std::array<point, 4>;

// Level-set surface sampler -> fan of tets around (x0,y0,z0).
// Restates object3d_base::init_polar (object3d_base.cpp:84-196) including its quirks: the ray
// step length is the literal 0.001 whatever `step` says (`step` only drives the two polar rays,
// :87,99-100), angles are accumulated by repeated addition (:123-143), and the cap at the top
// mixes one point of ring 0 (:171-174).
std::vector<tet_points> init_polar(const std::function<double(const point&)>& potential, double x0, double y0,
                                   double z0, double level_value, double step, double angle_step);

// object3d_roche_lobe (object3d_roche_lobe.cpp:20-49): equipotential through the hard-coded L1
// position x = 0.35515 (:30), sampled with angle_step 128, then rotated by the donor angle about
// the axis x = ACC_X0 (:48).
std::vector<tet_points> make_roche_lobe(const point& pos_accretor, double dist, double donor_angle_around_y,
                                        double m_accretor, double m_donor, double def_omega);

// object3d_sphere (object3d_sphere.cpp:11-18): radius R, angle_step 256.
std::vector<tet_points> make_sphere(const point& center, double R);

// In-place rotations of a point, tetra.cpp:44-62.
void rotate_point_x(point& p, double angle);
void rotate_point_y(point& p, double angle, double x0);

} // namespace c5host
