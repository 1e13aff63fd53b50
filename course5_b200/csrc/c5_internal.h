// c5_internal.h — context and the entry points shared between the library's translation units.
#pragma once

#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/c5gpu.h"
#include "c5_types.h"

namespace c5 {

constexpr int kMaxRot = C5_MAX_ROT;
constexpr int kTimelinePhases = 6; // start, rotated, refitted, masked, pixel kernel done, grazing-ray kernel done

// Counters the walk kernel accumulates (one 64-bit atomic per warp).
// kDeferred = rays the pixel kernel handed to the grazing-ray kernel, kTicket = how many of those
// the grazing-ray kernel's warps have drawn.
enum Counter { kSteps = 0, kHitPixels = 1, kSolidPixels = 2, kWalkErrors = 3, kDeferred = 4, kTicket = 5,
               kNumCounters = 8 };

struct StaticMaskKey {  // what the footprint of a solid that is never rotated depends on (all of it, bit for bit)
    int res_x, res_y;
    double x_min, y_min, step_x, step_y;
};

struct SolidSet {
    DevBuf<double> pts0;      // [n][4][3] pre-view frame
    DevBuf<double> pts_view;  // [n][4][3] view frame (rotated copy, or == pts0 content for static ones)
    DevBuf<uint32_t> faces;   // unique faces (4 * tet + k): a fan face shared by two tets is scanned once
    int64_t n = 0, n_faces = 0;
    double extent = 0.0;      // largest edge length of any solid tet (rotation invariant)
};

// Everything resident on ONE device.
struct DeviceState {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};

    // mesh (uploaded once)
    int64_t n_pts = 0, n_tets = 0, n_bfaces = 0;
    double mesh_lo[3] = {0, 0, 0}, mesh_hi[3] = {0, 0, 0}; // bounding box in the file frame
    DevBuf<double> px, py, pz;       // Morton-ordered file-frame coordinates (SoA: the rotate kernel streams them)
    DevBuf<Cell> cells;
    DevBuf<double> q0;               // Q per tet (Morton order); cells[t].s is derived from it
    double cells_limit = 0.0;        // the alpha_limit cells[].s was last prepared for
    bool cells_limit_valid = false;
    DeviceState* origin = nullptr;   // sibling context: the device state that owns the mesh (and these two flags)
    uint64_t mesh_version = 0;       // bumped by every upload; a sibling re-aliases when it falls behind
    bool mesh_shared = false;        // some sibling aliases this state's arrays
#ifdef C5_EXPERIMENTS
    DevBuf<StepRec> recs;            // experimental "rec" walk variant: 4 step records per tet, built on first use
    double recs_limit = 0.0;
    bool recs_valid = false;
    // C5_TRACE_FILE: start / end time and SM of every block of the first pixel-kernel launches,
    // written to that file when the context is destroyed (scripts/trace_blocks.py reads it)
    DevBuf<unsigned long long> trace;
    int trace_launches = 0;
    unsigned trace_grid[64] = {};
#endif
    DevBuf<BFace> bfaces;            // Morton-sorted boundary faces (BVH leaves)
    DevBuf<BvhNode> nodes;           // n_bfaces - 1 internal nodes, BFS order (root = 0)
    DevBuf<int32_t> node_parent;     // per internal node: (parent << 1) | which child, -1 for the root
    DevBuf<int32_t> leaf_parent;     // per leaf: (parent << 1) | which child
    DevBuf<uint32_t> refit_flags;    // per internal node arrival counter

    // solids
    SolidSet solid_follow, solid_static;

    // per view
    DevBuf<Vtx> vrot;
    DevBuf<double> xs, ys;
    int xs_res = 0, ys_res = 0;
    double xs_win[2] = {0, 0}, ys_win[2] = {0, 0};
    DevBuf<uint8_t> mask;
    DevBuf<uint32_t> mask_tall;      // solid mask: faces taller than a few rows, listed by the first pass
    DevBuf<unsigned> mask_counts;    // their number per solid set
    DevBuf<uint8_t> mask_tiles;      // "every pixel of this 16 x 8 tile is solid already"
    DevBuf<uint8_t> mask_static;     // whole-image mask of the solids that do not follow the view (cached)
    StaticMaskKey mask_static_key{}; // the pixel grid mask_static was scan-converted for
    bool mask_static_valid = false;  // reset by every upload / clear of solids
    DevBuf<double> out;              // band output, {tau, I} per pixel
    DevBuf<uint32_t> steps;
    DevBuf<unsigned long long> counters;
    DevBuf<unsigned long long> row_cost;
    DevBuf<DeferredRay> queue;       // grazing rays of the current view (capacity: pixels of the band)
    int sm_count = 0;
    cudaEvent_t ev_walk = nullptr;   // between the pixel kernel and the grazing-ray kernel (phase times)
    cudaEvent_t ev_done = nullptr;   // after the view's last copy: what c5_render_wait waits for
    // results of the view in flight, copied by the stream into pinned host memory
    unsigned long long* h_counters = nullptr;   // [kNumCounters]
    uint64_t* h_row_cost = nullptr;             // [h_row_cost_n]
    size_t h_row_cost_n = 0;
    // c5_debug_set (tests, diagnostics); 0 = default
    int opt_graze_list = 0, opt_query_budget = 0, opt_serial_list = 0, opt_graze_blocks = 0, opt_mask_per_face = 0, opt_mask_tile = 0;
    bool opt_no_zero_copy = false, opt_no_static_mask = false;
    int opt_prep_priority = 0;                  // bit 0: rotate / refit / mask, bit 1: the grazing-ray kernel, on prep_stream
    cudaStream_t prep_stream = nullptr;         // high priority (see opt_prep_priority)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // c5_debug_set("timeline", n): phase events of the last n views, read by c5_timeline_read
    std::vector<cudaEvent_t> tl_events;         // [n][kTimelinePhases]
    uint64_t tl_views = 0;                      // views recorded since the timeline was enabled

    uint64_t launches = 0;
};

struct MeshHost;

} // namespace c5

namespace c5 {
struct ViewPlan {      // a validated c5_view: trig evaluated, band and pixel steps resolved
    Rot rot[kMaxRot];
    int n_rot;
    int row_begin, row_end;
    double x_min, y_min, step_x, step_y;
};
struct PendingView {   // the view a context has in flight between c5_render_submit and c5_render_wait
    bool active = false;
    uint64_t ticket = 0;
    c5_view view{};
    ViewPlan plan{};
    bool d2h_copy = false; // the image went through the device band buffer and a copy (not stored in place)
};
} // namespace c5

struct c5_ctx {
    std::vector<std::unique_ptr<c5::DeviceState>> dev;
    std::string err;
    bool has_mesh = false;
    c5_mesh_info info{};
    std::vector<uint64_t> last_row_cost;
    void* nccl = nullptr; // NcclGroup*, multi-device contexts only
    c5_ctx* parent = nullptr;        // sibling context (c5_create_sibling): shares parent's mesh and solids
    std::vector<c5_ctx*> siblings;   // contexts created from this one (the caller's, and our own lanes)
    std::vector<c5_ctx*> lanes;      // lanes of c5_render_submit beyond this context itself (owned; also in siblings)
    int views_in_flight = 3;
    uint64_t next_ticket = 1;
    unsigned next_lane = 0;
    c5::PendingView pending;         // this context's own view in flight
    std::vector<std::pair<void*, bool>> images;   // c5_image_create (true) / c5_image_open (false) pointers
    std::vector<std::pair<void*, void*>> image_offsets; // imported images: (pointer handed out, mapping base)
    std::vector<void*> registered;                // c5_host_register pointers
};

namespace c5 {

// c5_topology.cu — builds every mesh-resident array of `d` from host input.
void build_mesh(DeviceState& d, const double* pts, int64_t n_pts, const int32_t* tets, int64_t n_tets,
                const double* alpha, const double* q);

// c5_exact.cu — kernels whose arithmetic must match the reference's host code bit for bit
// (compiled with -fmad=false).
void launch_rotate_vertices(DeviceState& d, const Rot* rot, int n_rot);
void launch_rotate_solids(DeviceState& d, const Rot* rot, int n_rot);
void launch_solid_mask(DeviceState& d, int res_x, int res_y, double x_min, double y_min, double step_x,
                       double step_y, int row_begin, int row_end);
void launch_bvh_refit(DeviceState& d);
void dedupe_solid_faces(DeviceState& d, SolidSet& ss); // fills ss.faces / ss.n_faces from ss.pts0
void launch_prepare_cells(DeviceState& d, double alpha_limit); // cells[t].s = q0[t] / min(alpha, limit)

// c5_walk.cu
constexpr int kTraceLaunches = 64, kTraceBlocks = 40960;
struct WalkLaunch {
    int res_x, res_y, row_begin, row_end;
    int i_begin, i_end, j_begin, j_end; // pixel rectangle that can see the mesh (clipped to the band by the launch)
    double alpha_limit;
    int round_through_float;
    int use_mask;
    int write_steps;
    int precision;
    double* out; // {tau, I} per pixel of the band, x fastest
    cudaEvent_t mark_walk_done;    // recorded between the pixel kernel and the grazing-ray kernel (may be null)
    cudaEvent_t mark_walk_done_tl; // the same moment for the timeline ring (may be null)
    cudaStream_t graze_stream;     // non-null: the grazing-ray kernel goes there (after mark_walk_done) and is joined back
    cudaEvent_t graze_join;
};
void launch_walk(DeviceState& d, const WalkLaunch& w);

} // namespace c5
