// c5_api.cu — the C ABI of include/c5gpu.h: context, mesh/solid upload, the per-view pipeline.
#include <cmath>
#include <cstring>
#include <algorithm>
#include <exception>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include <dlfcn.h>

#include "c5_internal.h"

using namespace c5;

namespace {

std::mutex g_create_err_mu;
std::string g_create_err;

const double kPi = 3.14159265358979323846; // config.hpp:45

template <class Fn>
int guarded(c5_ctx* ctx, Fn&& fn) {
    try {
        if (ctx) g_launch_counter = &ctx->dev[0]->launches;
        fn();
        return C5_OK;
    } catch (const Error& e) {
        if (ctx) ctx->err = e.text;
        return e.code;
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->err = "host allocation failed";
        return C5_E_NOMEM;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = e.what();
        return C5_E_INVALID;
    }
}

void use_device(const DeviceState& d) {
    if (!kHostSim) C5_CUDA(cudaSetDevice(d.device));
}

// Phase events of the view being enqueued: 0 start, 1 rotated, 2 BVH refitted, 3 mask done,
// 4 walk done (pixel + grazing-ray kernel), 5 band gathered (multi-device), 6 image on the host;
// ev_walk sits between the two walk kernels.
void record(DeviceState& d, int k) {
    if (!kHostSim) C5_CUDA(cudaEventRecord(d.ev[k], d.stream));
}

// c5_debug_set("timeline", n): the same moments kept for the last n views (c5_timeline_read).
void timeline_mark(DeviceState& d, int phase) {
    if (kHostSim || d.tl_events.empty()) return;
    const size_t n = d.tl_events.size() / kTimelinePhases;
    C5_CUDA(cudaEventRecord(d.tl_events[(d.tl_views % n) * kTimelinePhases + static_cast<size_t>(phase)], d.stream));
}

float elapsed(DeviceState& d, int a, int b) {
    if (kHostSim) return 0.f;
    float ms = 0.f;
    C5_CUDA(cudaEventElapsedTime(&ms, d.ev[a], d.ev[b]));
    return ms;
}

ViewPlan plan_view(const c5_view* v) {
    if (!v) fail(C5_E_INVALID, "render: view is NULL");
    if (v->res_x < 2 || v->res_y < 2) fail(C5_E_INVALID, "render: resolution must be at least 2 x 2");
    if (static_cast<int64_t>(v->res_x) * v->res_y >= (int64_t{1} << 31)) fail(C5_E_INVALID, "render: image too large");
    if (v->n_rot < 0 || v->n_rot > kMaxRot) fail(C5_E_INVALID, "render: n_rot out of range");
    if (!(v->window[0] > v->window[1]) || !(v->window[2] > v->window[3])) {
        fail(C5_E_INVALID, "render: window must be {x_max, x_min, y_max, y_min} with max > min");
    }
    if (v->precision != 64 && v->precision != 32 && v->precision != 0) {
        fail(C5_E_INVALID, "render: precision must be 64 or 32");
    }
    ViewPlan p{};
    p.n_rot = v->n_rot;
    for (int k = 0; k < v->n_rot; k++) {
        if (v->rot[k].axis != 0 && v->rot[k].axis != 1) fail(C5_E_INVALID, "render: rotation axis must be 0 or 1");
        // the trig is evaluated here by the host libm, exactly like the reference's host loop
        p.rot[k] = Rot{v->rot[k].axis, 0, std::cos(v->rot[k].angle), std::sin(v->rot[k].angle), v->rot[k].x0};
    }
    p.row_begin = v->row_begin;
    p.row_end = v->row_end;
    if (p.row_begin == 0 && p.row_end == 0) p.row_end = v->res_y;
    if (p.row_begin < 0 || p.row_end > v->res_y || p.row_begin >= p.row_end) {
        fail(C5_E_INVALID, "render: bad row band");
    }
    p.x_min = v->window[1];
    p.y_min = v->window[3];
    p.step_x = (v->window[0] - v->window[1]) / (v->res_x - 1.); // plane.cpp:298-302
    p.step_y = (v->window[2] - v->window[3]) / (v->res_y - 1.);
    return p;
}

// Pixel coordinates by repeated addition (plane.cpp:304-314) — not x_min + i * step.
void ensure_pixel_tables(DeviceState& d, const c5_view* v, const ViewPlan& p) {
    if (d.xs_res != v->res_x || d.xs_win[0] != v->window[0] || d.xs_win[1] != v->window[1]) {
        std::vector<double> xs(static_cast<size_t>(v->res_x));
        double c = p.x_min;
        for (int i = 0; i < v->res_x; i++) {
            xs[static_cast<size_t>(i)] = c;
            c = c + p.step_x;
        }
        d.xs.alloc(xs.size());
        h2d(d.xs.p, xs.data(), d.xs.bytes(), d.stream);
        stream_sync(d.stream);
        d.xs_res = v->res_x;
        d.xs_win[0] = v->window[0];
        d.xs_win[1] = v->window[1];
    }
    if (d.ys_res != v->res_y || d.ys_win[0] != v->window[2] || d.ys_win[1] != v->window[3]) {
        std::vector<double> ys(static_cast<size_t>(v->res_y));
        double c = p.y_min;
        for (int j = 0; j < v->res_y; j++) {
            ys[static_cast<size_t>(j)] = c;
            c = c + p.step_y;
        }
        d.ys.alloc(ys.size());
        h2d(d.ys.p, ys.data(), d.ys.bytes(), d.stream);
        stream_sync(d.stream);
        d.ys_res = v->res_y;
        d.ys_win[0] = v->window[2];
        d.ys_win[1] = v->window[3];
    }
}

// Where on the screen the mesh can be: the view-frame box of the 8 corners of its file-frame box,
// widened by a pixel. The walk's grid covers only that rectangle (the rest of the band is filled as
// background); a conservative hint — inside it every tile is still tested against the BVH root.
void mesh_pixel_rect(const DeviceState& d, const c5_view* v, const ViewPlan& p, WalkLaunch& w) {
    double lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
    for (int corner = 0; corner < 8; corner++) {
        double x = (corner & 1) ? d.mesh_hi[0] : d.mesh_lo[0];
        double y = (corner & 2) ? d.mesh_hi[1] : d.mesh_lo[1];
        double z = (corner & 4) ? d.mesh_hi[2] : d.mesh_lo[2];
        for (int k = 0; k < p.n_rot; k++) { // tetra.cpp:44-62
            const double c = p.rot[k].c, s = p.rot[k].s;
            if (p.rot[k].axis == 0) {
                const double y0 = y;
                y = y * c - z * s;
                z = y0 * s + z * c;
            } else {
                x -= p.rot[k].x0;
                const double x1 = x;
                x = x * c - z * s;
                z = x1 * s + z * c;
                x += p.rot[k].x0;
            }
        }
        lo[0] = std::min(lo[0], x);
        hi[0] = std::max(hi[0], x);
        lo[1] = std::min(lo[1], y);
        hi[1] = std::max(hi[1], y);
    }
    auto pixel = [](double c, double c_min, double step, int res, bool up) {
        double r = (c - c_min) / step + (up ? 2.0 : -2.0); // one pixel for rounding, one for the repeated-addition tables
        if (!(r > -1.0)) r = -1.0;
        if (!(r < res + 1.0)) r = res + 1.0;
        return static_cast<int>(up ? std::ceil(r) : std::floor(r));
    };
    w.i_begin = pixel(lo[0], p.x_min, p.step_x, v->res_x, false);
    w.i_end = pixel(hi[0], p.x_min, p.step_x, v->res_x, true) + 1;
    w.j_begin = pixel(lo[1], p.y_min, p.step_y, v->res_y, false);
    w.j_end = pixel(hi[1], p.y_min, p.step_y, v->res_y, true) + 1;
}

// Enqueues one view's kernels for rows [row_begin,row_end) on device d. Events:
// 0 start, 1 rotated, 2 bvh, 3 mask, 4 walk.
void enqueue_view(DeviceState& d, const c5_view* v, const ViewPlan& p, bool want_steps, double* out_override = nullptr) {
    use_device(d);
    g_launch_counter = &d.launches;
    ensure_pixel_tables(d, v, p);
    // Per-view buffers are sized for the WHOLE image, not the band: cost-balanced bands move from view to
    // view, and a buffer that had to grow in the middle of a sweep would cost a cudaFree (a device-wide
    // synchronisation) and a cudaMalloc on the spot.
    const size_t n_pix = static_cast<size_t>(v->res_y) * v->res_x;
    const bool solids = v->use_solids && (d.solid_follow.n + d.solid_static.n) > 0;
    if (!out_override) d.out.ensure(2 * n_pix);
    if (want_steps) d.steps.ensure(n_pix);
    d.counters.ensure(kNumCounters);
    d.queue.ensure(n_pix); // every ray may be deferred (reserved, hardly ever touched)
    d.row_cost.ensure(static_cast<size_t>(v->res_y));
    if (solids) d.mask.ensure(n_pix);

    // c5_debug_set("prep_priority", 1): rotate / refit / mask go to a high-priority stream of their own, so
    // that their (small, latency-bound) blocks are dispatched ahead of the queued blocks of other lanes'
    // pixel kernels instead of behind them; the walk waits for them by event.
    struct StreamSwap {
        DeviceState& d;
        cudaStream_t main;
        ~StreamSwap() { d.stream = main; }
    } swap{d, d.stream};
    const bool prep = !kHostSim && (d.opt_prep_priority & 1) && d.prep_stream != nullptr;
    if (prep) {
        C5_CUDA(cudaEventRecord(d.ev_fork, swap.main)); // after whatever the caller's stream holds (this lane's previous view)
        C5_CUDA(cudaStreamWaitEvent(d.prep_stream, d.ev_fork, 0));
        d.stream = d.prep_stream;
    }
    record(d, 0);
    timeline_mark(d, 0);
    launch_prepare_cells(d, v->alpha_limit); // no-op unless --alpha_limit changed since the last view
    launch_rotate_vertices(d, p.rot, p.n_rot);
    if (solids) launch_rotate_solids(d, p.rot, p.n_rot);
    record(d, 1);
    timeline_mark(d, 1);
    launch_bvh_refit(d);
    record(d, 2);
    timeline_mark(d, 2);
    if (solids) {
        launch_solid_mask(d, v->res_x, v->res_y, p.x_min, p.y_min, p.step_x, p.step_y, p.row_begin, p.row_end);
    }
    record(d, 3);
    timeline_mark(d, 3);
    if (prep) {
        C5_CUDA(cudaEventRecord(d.ev_join, d.prep_stream));
        d.stream = swap.main;
        C5_CUDA(cudaStreamWaitEvent(d.stream, d.ev_join, 0));
    }
    dev_zero(d.counters.p, kNumCounters * sizeof(unsigned long long), d.stream);
    dev_zero(d.row_cost.p, static_cast<size_t>(v->res_y) * sizeof(unsigned long long), d.stream);
    WalkLaunch w{};
    w.res_x = v->res_x;
    w.res_y = v->res_y;
    w.row_begin = p.row_begin;
    w.row_end = p.row_end;
    w.alpha_limit = v->alpha_limit;
    w.round_through_float = v->round_through_float;
    w.use_mask = solids ? 1 : 0;
    w.write_steps = want_steps ? 1 : 0;
    w.precision = v->precision ? v->precision : 64;
    w.out = out_override ? out_override : d.out.p;
    w.mark_walk_done = d.ev_walk;
    w.mark_walk_done_tl = nullptr;
    w.graze_stream = (!kHostSim && (d.opt_prep_priority & 2) && d.prep_stream) ? d.prep_stream : nullptr;
    w.graze_join = d.ev_join;
    if (!kHostSim && !d.tl_events.empty()) {
        const size_t n = d.tl_events.size() / kTimelinePhases;
        w.mark_walk_done_tl = d.tl_events[(d.tl_views % n) * kTimelinePhases + 4];
    }
    mesh_pixel_rect(d, v, p, w);
    if (!kHostSim) { // a band the mesh does not reach launches no walk kernels: the markers are recorded anyway
        C5_CUDA(cudaEventRecord(d.ev_walk, d.stream));
        if (w.mark_walk_done_tl) C5_CUDA(cudaEventRecord(w.mark_walk_done_tl, d.stream));
    }
    launch_walk(d, w);
    record(d, 4);
    timeline_mark(d, 5);
    if (!d.tl_events.empty()) d.tl_views++;
}

// ---- NCCL, loaded lazily: only multi-device contexts need it ---------------------------------------
struct NcclApi {
    void* handle = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat64 = 8; // ncclDouble

struct NcclGroup {
    NcclApi api;
    std::vector<void*> comms;
};

void nccl_check(const NcclApi& api, int rc, const char* what) {
    if (rc != 0) fail(C5_E_NCCL, std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(rc) : "error"));
}

NcclGroup* nccl_open(const std::vector<int>& devices) {
    auto g = std::make_unique<NcclGroup>();
    NcclApi& a = g->api;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) fail(C5_E_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) {
        void* p = dlsym(a.handle, n);
        if (!p) fail(C5_E_NCCL, std::string("libnccl lacks ") + n);
        return p;
    };
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(sym("ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    g->comms.assign(devices.size(), nullptr);
    nccl_check(a, a.CommInitAll(g->comms.data(), static_cast<int>(devices.size()), devices.data()), "ncclCommInitAll");
    return g.release();
}

void nccl_close(void* p) {
    auto* g = static_cast<NcclGroup*>(p);
    if (!g) return;
    for (void* c : g->comms) {
        if (c) g->api.CommDestroy(c);
    }
    delete g;
}

// Contiguous row bands of [row_begin,row_end) with equal estimated cost (previous view's per-row
// tet-steps plus a constant per row), one per device; equal heights when there is no history.
std::vector<std::pair<int, int>> cut_bands(const c5_ctx* ctx, const c5_view* v, const ViewPlan& p, int n) {
    const int rows = p.row_end - p.row_begin;
    std::vector<double> cost(static_cast<size_t>(rows), 1.0);
    if (ctx->last_row_cost.size() == static_cast<size_t>(v->res_y)) {
        double total = 0;
        for (int j = 0; j < rows; j++) total += static_cast<double>(ctx->last_row_cost[static_cast<size_t>(p.row_begin + j)]);
        if (total > 0) {
            const double base = 64.0 + 0.02 * total / rows; // empty rows are not free (BVH miss, stores)
            for (int j = 0; j < rows; j++) {
                cost[static_cast<size_t>(j)] = base + static_cast<double>(ctx->last_row_cost[static_cast<size_t>(p.row_begin + j)]);
            }
        }
    }
    double total = 0;
    for (double c : cost) total += c;
    std::vector<std::pair<int, int>> bands;
    int lo = 0;
    double acc = 0;
    int j = 0;
    for (int b = 0; b < n; b++) {
        const double target = total * (b + 1) / n;
        while (j < rows && (acc + cost[static_cast<size_t>(j)] <= target || j == lo) && rows - (j + 1) >= n - b - 1) {
            acc += cost[static_cast<size_t>(j)];
            j++;
        }
        if (b == n - 1) j = rows;
        bands.emplace_back(p.row_begin + lo, p.row_begin + j);
        lo = j;
    }
    return bands;
}

struct Counters {
    unsigned long long c[kNumCounters] = {};
};

void follow_parent(c5_ctx* ctx);

// Multi-device contexts (c5_create with several devices, one process): plane ctor + find_intersections
// + trace_rays for one view into a HOST buffer; device r renders band r, the bands land in device 0's
// image by one grouped NCCL send/recv, device 0 copies the image to the host. Synchronous.
void render_multi_device(c5_ctx* ctx, const c5_view* v, double* out, uint32_t* steps, uint8_t* solid_mask, c5_stats* st) {
    if (!ctx->has_mesh) fail(C5_E_STATE, "render: no mesh uploaded");
    if (!out) fail(C5_E_INVALID, "render: out is NULL");
    const ViewPlan p = plan_view(v);
    const int n_dev = static_cast<int>(ctx->dev.size());
    if ((p.row_end - p.row_begin) < n_dev) fail(C5_E_INVALID, "render: fewer rows than devices");
    const auto bands = cut_bands(ctx, v, p, n_dev);
    DeviceState& d0 = *ctx->dev[0];
    const size_t row_doubles = static_cast<size_t>(v->res_x) * 2;
    const size_t n_pix_view = static_cast<size_t>(p.row_end - p.row_begin) * v->res_x;
    use_device(d0);
    d0.out.ensure(2 * n_pix_view); // device 0 assembles the whole view here
    const size_t view_off = static_cast<size_t>(p.row_begin) * v->res_x;

    // enqueue every device's band (asynchronous; devices run concurrently). With several devices
    // one host thread per device issues the launches, so launch latency is not serialised.
    auto enqueue_band = [&](int r) {
        DeviceState& d = *ctx->dev[static_cast<size_t>(r)];
        ViewPlan pr = p;
        pr.row_begin = bands[static_cast<size_t>(r)].first;
        pr.row_end = bands[static_cast<size_t>(r)].second;
        double* target = nullptr;
        if (r == 0) target = d0.out.p + static_cast<size_t>(pr.row_begin - p.row_begin) * row_doubles;
        enqueue_view(d, v, pr, steps != nullptr, target);
    };
    if (kHostSim) {
        for (int r = 0; r < n_dev; r++) enqueue_band(r);
    } else {
        std::vector<std::thread> workers;
        std::vector<std::exception_ptr> errors(static_cast<size_t>(n_dev));
        for (int r = 0; r < n_dev; r++) {
            workers.emplace_back([&, r] {
                try {
                    enqueue_band(r);
                } catch (...) {
                    errors[static_cast<size_t>(r)] = std::current_exception();
                }
            });
        }
        for (auto& w : workers) w.join();
        for (auto& e : errors) {
            if (e) std::rethrow_exception(e);
        }
        g_launch_counter = &d0.launches;
    }

    // gather-v of the bands into device 0's image
    {
        if (kHostSim) {
            for (int r = 1; r < n_dev; r++) {
                const auto& b = bands[static_cast<size_t>(r)];
                std::memcpy(d0.out.p + static_cast<size_t>(b.first - p.row_begin) * row_doubles,
                            ctx->dev[static_cast<size_t>(r)]->out.p, static_cast<size_t>(b.second - b.first) * row_doubles * sizeof(double));
            }
        } else {
            auto* g = static_cast<NcclGroup*>(ctx->nccl);
            nccl_check(g->api, g->api.GroupStart(), "ncclGroupStart");
            for (int r = 1; r < n_dev; r++) {
                const auto& b = bands[static_cast<size_t>(r)];
                const size_t count = static_cast<size_t>(b.second - b.first) * row_doubles;
                DeviceState& d = *ctx->dev[static_cast<size_t>(r)];
                nccl_check(g->api, g->api.Send(d.out.p, count, kNcclFloat64, 0, g->comms[static_cast<size_t>(r)], d.stream), "ncclSend");
                nccl_check(g->api, g->api.Recv(d0.out.p + static_cast<size_t>(b.first - p.row_begin) * row_doubles, count, kNcclFloat64, r,
                                               g->comms[0], d0.stream), "ncclRecv");
            }
            nccl_check(g->api, g->api.GroupEnd(), "ncclGroupEnd");
        }
    }
    use_device(d0);
    record(d0, 5);
    d2h(out + 2 * view_off, d0.out.p, 2 * n_pix_view * sizeof(double), d0.stream);
    record(d0, 6);

    // per-device extras (steps, mask), counters, row costs
    const bool solids = v->use_solids && (d0.solid_follow.n + d0.solid_static.n) > 0;
    std::vector<Counters> counters(static_cast<size_t>(n_dev));
    std::vector<std::vector<uint64_t>> row_cost(static_cast<size_t>(n_dev));
    for (int r = 0; r < n_dev; r++) {
        DeviceState& d = *ctx->dev[static_cast<size_t>(r)];
        use_device(d);
        const auto& b = bands[static_cast<size_t>(r)];
        const size_t band_off = static_cast<size_t>(b.first) * v->res_x;
        const size_t n_pix_band = static_cast<size_t>(b.second - b.first) * v->res_x;
        if (steps) d2h(steps + band_off, d.steps.p, n_pix_band * sizeof(uint32_t), d.stream);
        if (solid_mask) {
            if (solids) d2h(solid_mask + band_off, d.mask.p + band_off, n_pix_band, d.stream);
            else std::memset(solid_mask + band_off, 0, n_pix_band);
        }
        d2h(counters[static_cast<size_t>(r)].c, d.counters.p, sizeof(Counters), d.stream);
        row_cost[static_cast<size_t>(r)].assign(static_cast<size_t>(v->res_y), 0);
        d2h(row_cost[static_cast<size_t>(r)].data(), d.row_cost.p, static_cast<size_t>(v->res_y) * sizeof(uint64_t), d.stream);
    }
    for (auto& dp : ctx->dev) {
        use_device(*dp);
        stream_sync(dp->stream);
    }
    use_device(d0);

    Counters total;
    ctx->last_row_cost.assign(static_cast<size_t>(v->res_y), 0);
    for (int r = 0; r < n_dev; r++) {
        for (int k = 0; k < kNumCounters; k++) total.c[k] += counters[static_cast<size_t>(r)].c[k];
        for (int j = 0; j < v->res_y; j++) ctx->last_row_cost[static_cast<size_t>(j)] += row_cost[static_cast<size_t>(r)][static_cast<size_t>(j)];
    }
    if (st) {
        std::memset(st, 0, sizeof(*st));
        st->pixels = static_cast<uint64_t>(n_pix_view);
        st->tet_steps = total.c[kSteps];
        st->hit_pixels = total.c[kHitPixels];
        st->solid_pixels = total.c[kSolidPixels];
        st->walk_errors = total.c[kWalkErrors];
        st->grazing_rays = static_cast<int32_t>(total.c[kDeferred]);
        for (auto& dp : ctx->dev) { // phase times: the slowest device
            use_device(*dp);
            st->ms_rotate = std::max(st->ms_rotate, elapsed(*dp, 0, 1));
            st->ms_bvh = std::max(st->ms_bvh, elapsed(*dp, 1, 2));
            st->ms_mask = std::max(st->ms_mask, elapsed(*dp, 2, 3));
            st->ms_walk = std::max(st->ms_walk, elapsed(*dp, 3, 4));
            if (!kHostSim) {
                float g = 0.f;
                C5_CUDA(cudaEventElapsedTime(&g, dp->ev_walk, dp->ev[4]));
                st->ms_graze = std::max(st->ms_graze, g);
            }
        }
        use_device(d0);
        st->ms_gather = elapsed(d0, 4, 5);
        st->ms_d2h = elapsed(d0, 5, 6);
        st->ms_total = elapsed(d0, 0, 6);
        st->n_devices = n_dev;
    }
    if (total.c[kWalkErrors]) {
        fail(C5_E_WALK, "render: " + std::to_string(total.c[kWalkErrors]) + " ray(s) exceeded the step cap");
    }
}

// Single-device, device-resident output (the caller's buffer), no host copy of the image.
void collect_stats(c5_ctx* ctx, DeviceState& d, const c5_view* v, const ViewPlan& p, c5_stats* st, int ev_last) {
    unsigned long long c[kNumCounters] = {};
    d2h(c, d.counters.p, sizeof(c), d.stream);
    ctx->last_row_cost.assign(static_cast<size_t>(v->res_y), 0);
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "row cost width");
    d2h(ctx->last_row_cost.data(), d.row_cost.p, static_cast<size_t>(v->res_y) * sizeof(uint64_t), d.stream);
    stream_sync(d.stream);
    if (st) {
        std::memset(st, 0, sizeof(*st));
        st->pixels = static_cast<uint64_t>(p.row_end - p.row_begin) * static_cast<uint64_t>(v->res_x);
        st->tet_steps = c[kSteps];
        st->hit_pixels = c[kHitPixels];
        st->solid_pixels = c[kSolidPixels];
        st->walk_errors = c[kWalkErrors];
        st->grazing_rays = static_cast<int32_t>(c[kDeferred]);
        st->ms_rotate = elapsed(d, 0, 1);
        st->ms_bvh = elapsed(d, 1, 2);
        st->ms_mask = elapsed(d, 2, 3);
        st->ms_walk = elapsed(d, 3, 4);
        if (!kHostSim) C5_CUDA(cudaEventElapsedTime(&st->ms_graze, d.ev_walk, d.ev[4]));
        st->ms_total = elapsed(d, 0, ev_last);
        st->n_devices = 1;
    }
    if (c[kWalkErrors]) {
        fail(C5_E_WALK, "render: " + std::to_string(c[kWalkErrors]) + " ray(s) exceeded the step cap");
    }
}

// ---- single-device contexts: submit / wait -----------------------------------------------------------
// A view is enqueued on the context's stream together with the copies that bring its results to the
// host (image unless it is stored in place, counters, per-row costs -> page-locked staging), then an
// event; waiting for the view is waiting for that event. Nothing in between touches the host, so
// several contexts that share a mesh (the lanes of c5_render_submit) overlap on the device.

void ensure_host_staging(DeviceState& d, size_t rows) {
    if (!d.h_counters) d.h_counters = static_cast<unsigned long long*>(host_pinned_alloc(kNumCounters * sizeof(unsigned long long)));
    if (rows > d.h_row_cost_n) {
        host_pinned_free(d.h_row_cost);
        d.h_row_cost = nullptr;
        d.h_row_cost_n = 0;
        d.h_row_cost = static_cast<uint64_t*>(host_pinned_alloc(rows * sizeof(uint64_t)));
        d.h_row_cost_n = rows;
    }
}

void submit_view(c5_ctx* ctx, const c5_view* v, double* out, uint32_t* steps, uint8_t* solid_mask, uint64_t ticket,
                 bool in_place_allowed = true) {
    follow_parent(ctx);
    if (ctx->pending.active) fail(C5_E_STATE, "render: this context still has a view in flight (c5_render_wait it first)");
    if (!ctx->has_mesh) fail(C5_E_STATE, "render: no mesh uploaded");
    if (!out) fail(C5_E_INVALID, "render: out is NULL");
    const ViewPlan p = plan_view(v);
    DeviceState& d = *ctx->dev[0];
    use_device(d);
    const size_t n_pix_band = static_cast<size_t>(p.row_end - p.row_begin) * v->res_x;
    const size_t band_off = static_cast<size_t>(p.row_begin) * v->res_x;
    // A PAGE-LOCKED host buffer is device-addressable: the walk can store its pixels straight into it
    // (posted 128-byte writes over PCIe, spread over the kernel's run time) and the separate
    // device-to-host copy of the image disappears. That is the shorter path for ONE view at a time
    // (c5_render). With several views in flight the copy engine is the better one: it moves view k's
    // image while the SMs walk view k+1, and the walk kernels are not held up by PCIe write
    // back-pressure (C3, 4 views in flight: 4.76 ms per view against 5.53 in place and 4.59 with no host
    // image at all; profiles/r02_bench_c3_n1_e2e_*.json) — so c5_render_submit passes
    // in_place_allowed = false. Pixels are stored as one 128-bit word, so an in-place buffer must be
    // 16-byte aligned; pageable or misaligned buffers take the copy.
    double* host_direct = nullptr;
    if (!kHostSim && in_place_allowed && !d.opt_no_zero_copy && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, out) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) {
            host_direct = static_cast<double*>(attr.devicePointer) + 2 * band_off;
        } else {
            cudaGetLastError(); // pageable memory reports an error on some drivers: not ours
        }
    }
    ensure_host_staging(d, static_cast<size_t>(v->res_y));
    enqueue_view(d, v, p, steps != nullptr, host_direct);
    record(d, 5);
    if (!host_direct) d2h(out + 2 * band_off, d.out.p, 2 * n_pix_band * sizeof(double), d.stream);
    record(d, 6);
    const bool solids = v->use_solids && (d.solid_follow.n + d.solid_static.n) > 0;
    if (steps) d2h(steps + band_off, d.steps.p, n_pix_band * sizeof(uint32_t), d.stream);
    if (solid_mask) {
        if (solids) d2h(solid_mask + band_off, d.mask.p + band_off, n_pix_band, d.stream);
        else std::memset(solid_mask + band_off, 0, n_pix_band);
    }
    d2h(d.h_counters, d.counters.p, kNumCounters * sizeof(unsigned long long), d.stream);
    d2h(d.h_row_cost, d.row_cost.p, static_cast<size_t>(v->res_y) * sizeof(uint64_t), d.stream);
    if (!kHostSim) C5_CUDA(cudaEventRecord(d.ev_done, d.stream));
    ctx->pending.active = true;
    ctx->pending.ticket = ticket;
    ctx->pending.view = *v;
    ctx->pending.plan = p;
    ctx->pending.d2h_copy = host_direct == nullptr;
}

// Waits for the view `lane` has in flight; row costs and errors are reported through `owner`
// (the context the caller holds: lane itself, or the context whose lane it is).
void finish_view(c5_ctx* owner, c5_ctx* lane, c5_stats* st) {
    if (!lane->pending.active) fail(C5_E_STATE, "render_wait: no view in flight");
    DeviceState& d = *lane->dev[0];
    use_device(d);
    lane->pending.active = false; // whatever happens below, the lane is free again
    if (!kHostSim) C5_CUDA(cudaEventSynchronize(d.ev_done));
    const c5_view& v = lane->pending.view;
    const ViewPlan& p = lane->pending.plan;
    owner->last_row_cost.assign(d.h_row_cost, d.h_row_cost + v.res_y);
    const unsigned long long* c = d.h_counters;
    if (st) {
        std::memset(st, 0, sizeof(*st));
        st->pixels = static_cast<uint64_t>(p.row_end - p.row_begin) * static_cast<uint64_t>(v.res_x);
        st->tet_steps = c[kSteps];
        st->hit_pixels = c[kHitPixels];
        st->solid_pixels = c[kSolidPixels];
        st->walk_errors = c[kWalkErrors];
        st->grazing_rays = static_cast<int32_t>(c[kDeferred]);
        st->ms_rotate = elapsed(d, 0, 1);
        st->ms_bvh = elapsed(d, 1, 2);
        st->ms_mask = elapsed(d, 2, 3);
        st->ms_walk = elapsed(d, 3, 4);
        if (!kHostSim) C5_CUDA(cudaEventElapsedTime(&st->ms_graze, d.ev_walk, d.ev[4]));
        st->ms_d2h = elapsed(d, 5, 6);
        st->ms_total = elapsed(d, 0, 6);
        st->n_devices = 1;
    }
    if (c[kWalkErrors]) {
        fail(C5_E_WALK, "render: " + std::to_string(c[kWalkErrors]) + " ray(s) exceeded the step cap");
    }
}

// The lanes of c5_render_submit: lane 0 is the context itself, the others are sibling contexts the
// library makes (and owns) on first use.
c5_ctx* lane_of(c5_ctx* ctx, unsigned k) {
    return k == 0 ? ctx : ctx->lanes[k - 1];
}

void ensure_lanes(c5_ctx* ctx) {
    while (static_cast<int>(ctx->lanes.size()) + 1 < ctx->views_in_flight) {
        c5_ctx* lane = nullptr;
        const int rc = c5_create_sibling(ctx, &lane);
        if (rc != C5_OK) fail(rc, "render_submit: cannot create a lane: " + ctx->err);
        ctx->lanes.push_back(lane);
    }
}

void upload_solids(c5_ctx* ctx, const double* pts, int64_t n, int follows) {
    if (n < 0 || (n > 0 && !pts)) fail(C5_E_INVALID, "upload_solids: bad arguments");
    if (n == 0) return;
    for (auto& dp : ctx->dev) {
        DeviceState& d = *dp;
        use_device(d);
        SolidSet& ss = follows ? d.solid_follow : d.solid_static;
        const size_t old_n = static_cast<size_t>(ss.n), add = static_cast<size_t>(n);
        DevBuf<double> merged;
        merged.alloc((old_n + add) * 12);
        if (old_n) d2d(merged.p, ss.pts0.p, old_n * 12 * sizeof(double), d.stream);
        h2d(merged.p + old_n * 12, pts, add * 12 * sizeof(double), d.stream);
        stream_sync(d.stream);
        ss.pts0 = std::move(merged);
        ss.n = static_cast<int64_t>(old_n + add);
        ss.pts_view.alloc(static_cast<size_t>(ss.n) * 12);
        // static solids are never rotated: their view-frame copy is the upload itself
        if (!follows) {
            d2d(ss.pts_view.p, ss.pts0.p, ss.pts0.bytes(), d.stream);
            stream_sync(d.stream);
        }
        // longest tet edge of this upload: bounds how many rows one face can span in any view
        for (size_t t = 0; t < add; t++) {
            const double* q = pts + 12 * t;
            for (int a = 0; a < 4; a++) {
                for (int b = a + 1; b < 4; b++) {
                    const double dx = q[3 * a] - q[3 * b], dy = q[3 * a + 1] - q[3 * b + 1], dz = q[3 * a + 2] - q[3 * b + 2];
                    ss.extent = std::max(ss.extent, std::sqrt(dx * dx + dy * dy + dz * dz));
                }
            }
        }
        g_launch_counter = &d.launches;
        dedupe_solid_faces(d, ss);
        d.mask_static_valid = false;
    }
    ctx->info.n_solid_tets += n;
}

// Points a sibling's device state at the mesh and solids its parent owns; per-view arrays are its own.
void alias_mesh(DeviceState& s, DeviceState& o) {
    use_device(s);
    s.origin = &o;
    o.mesh_shared = true;
    s.n_pts = o.n_pts;
    s.n_tets = o.n_tets;
    s.n_bfaces = o.n_bfaces;
    for (int a = 0; a < 3; a++) {
        s.mesh_lo[a] = o.mesh_lo[a];
        s.mesh_hi[a] = o.mesh_hi[a];
    }
    s.px.alias(o.px);
    s.py.alias(o.py);
    s.pz.alias(o.pz);
    s.cells.alias(o.cells);
    s.q0.alias(o.q0);
    s.bfaces.alias(o.bfaces);
    s.node_parent.alias(o.node_parent);
    s.leaf_parent.alias(o.leaf_parent);
    s.nodes.alloc(o.nodes.n); // child links are static, the boxes are refitted per view
    if (o.nodes.n) d2d(s.nodes.p, o.nodes.p, o.nodes.bytes(), s.stream);
    s.refit_flags.alloc(o.refit_flags.n);
    s.vrot.alloc(static_cast<size_t>(o.n_pts));
    const SolidSet* from[2] = {&o.solid_follow, &o.solid_static};
    SolidSet* to[2] = {&s.solid_follow, &s.solid_static};
    for (int k = 0; k < 2; k++) {
        to[k]->n = from[k]->n;
        to[k]->n_faces = from[k]->n_faces;
        to[k]->extent = from[k]->extent;
        to[k]->pts0.alias(from[k]->pts0);
        to[k]->faces.alias(from[k]->faces);
        if (k == 0) to[k]->pts_view.alloc(from[k]->pts_view.n); // rotated with the view
        else to[k]->pts_view.alias(from[k]->pts_view);         // static solids are never rotated
    }
    s.mask_static_valid = false;
    stream_sync(s.stream);
    s.mesh_version = o.mesh_version;
}

// A sibling follows its parent's uploads lazily, at its next render.
void follow_parent(c5_ctx* ctx) {
    if (!ctx->parent) return;
    c5_ctx* par = ctx->parent;
    DeviceState& s = *ctx->dev[0];
    DeviceState& o = *par->dev[0];
    if (par->has_mesh && (s.origin != &o || s.mesh_version != o.mesh_version)) alias_mesh(s, o);
    ctx->has_mesh = par->has_mesh;
    ctx->info = par->info;
}

// cudaMalloc may carve an allocation out of a larger block, and a CUDA IPC handle names the BLOCK:
// the importer gets the block's base address. The driver knows the base (cuMemGetAddressRange);
// the offset travels with the handle. libcuda is loaded lazily so that the library still loads on
// machines without a driver (the CPU-only build container).
uint64_t offset_in_allocation(void* p) {
    using Fn = int (*)(unsigned long long*, size_t*, unsigned long long);
    static Fn fn = [] {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
        return h ? reinterpret_cast<Fn>(dlsym(h, "cuMemGetAddressRange_v2")) : nullptr;
    }();
    if (!fn) fail(C5_E_CUDA, "image_create: cuMemGetAddressRange not available (libcuda.so.1)");
    unsigned long long base = 0;
    size_t size = 0;
    const int rc = fn(&base, &size, static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(p)));
    if (rc != 0) fail(C5_E_CUDA, "image_create: cuMemGetAddressRange failed (" + std::to_string(rc) + ")");
    return static_cast<uint64_t>(reinterpret_cast<uintptr_t>(p)) - base;
}

} // namespace

extern "C" {

int c5_abi_version(void) {
    return C5_ABI_VERSION;
}

void c5_view_from_flags(c5_view* v, int32_t res_x, int32_t res_y, double X_pi, double Y_pi, double I_pi,
                        double alpha_limit) {
    if (!v) return;
    std::memset(v, 0, sizeof(*v));
    v->res_x = res_x;
    v->res_y = res_y;
    v->window[0] = 2.2; // main.cpp:83
    v->window[1] = -0.2;
    v->window[2] = 0.9;
    v->window[3] = -0.9;
    const double a0 = -I_pi * kPi + kPi / 2.; // main.cpp:96
    v->n_rot = 3;
    v->rot[0] = c5_rotation{0, 0, a0, 0.0};                 // main.cpp:105
    v->rot[1] = c5_rotation{1, 0, Y_pi * kPi, 1.0};         // main.cpp:106, ACC_X0 = 1 (config.hpp:55)
    v->rot[2] = c5_rotation{0, 0, -a0 + X_pi * kPi, 0.0};   // main.cpp:107
    v->alpha_limit = alpha_limit;
    v->precision = 64;
    v->round_through_float = 1;
    v->use_solids = 1;
}

int c5_create(const int32_t* devices, int32_t n_dev, c5_ctx** out) {
    if (!out) return C5_E_INVALID;
    *out = nullptr;
    c5_ctx* ctx = nullptr;
    auto failed = [&](int code, const std::string& text) {
        {
            std::lock_guard<std::mutex> lock(g_create_err_mu);
            g_create_err = text;
        }
        c5_destroy(ctx); // the same teardown as a finished context: streams, events, NCCL, device memory
        return code;
    };
    try {
        if (n_dev < 1 || !devices) fail(C5_E_INVALID, "c5_create: need at least one device");
        ctx = new c5_ctx();
        if (!kHostSim) {
            int count = 0;
            cudaError_t e = cudaGetDeviceCount(&count);
            if (e != cudaSuccess || count == 0) {
                fail(C5_E_CUDA, std::string("c5_create: no CUDA device (") + cudaGetErrorString(e) +
                                    "); this library has no CPU path");
            }
            for (int k = 0; k < n_dev; k++) {
                if (devices[k] < 0 || devices[k] >= count) fail(C5_E_INVALID, "c5_create: device ordinal out of range");
            }
        }
        for (int k = 0; k < n_dev; k++) {
            ctx->dev.push_back(std::make_unique<DeviceState>());
            DeviceState* d = ctx->dev.back().get(); // in the context before anything is created: a failure below is torn down
            d->device = devices[k];
            if (!kHostSim) {
                C5_CUDA(cudaSetDevice(d->device));
                C5_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
                for (auto& e : d->ev) C5_CUDA(cudaEventCreate(&e));
                C5_CUDA(cudaEventCreate(&d->ev_walk));
                C5_CUDA(cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming));
                int prio_lo = 0, prio_hi = 0;
                C5_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
                C5_CUDA(cudaStreamCreateWithPriority(&d->prep_stream, cudaStreamNonBlocking, prio_hi));
                C5_CUDA(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
                C5_CUDA(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
            }
        }
        if (n_dev > 1 && !kHostSim) {
            ctx->nccl = nccl_open(std::vector<int>(devices, devices + n_dev));
            // peers are only touched by NCCL; each device keeps its own replica of the mesh
        }
        *out = ctx;
        return C5_OK;
    } catch (const Error& e) {
        return failed(e.code, e.text);
    } catch (const std::exception& e) {
        return failed(C5_E_NOMEM, e.what());
    }
}

int c5_create_sibling(c5_ctx* parent, c5_ctx** out) {
    if (!parent || !out) return C5_E_INVALID;
    *out = nullptr;
    if (parent->parent || parent->dev.size() != 1) {
        parent->err = "create_sibling: the parent must be a single-device context that is not itself a sibling";
        return C5_E_INVALID;
    }
    const int32_t device = parent->dev[0]->device;
    c5_ctx* ctx = nullptr;
    const int rc = c5_create(&device, 1, &ctx);
    if (rc != C5_OK) return rc;
    ctx->parent = parent;
    parent->siblings.push_back(ctx);
    DeviceState& from = *parent->dev[0];
    DeviceState& to = *ctx->dev[0];
    to.opt_graze_list = from.opt_graze_list;
    to.opt_query_budget = from.opt_query_budget;
    to.opt_serial_list = from.opt_serial_list;
    to.opt_graze_blocks = from.opt_graze_blocks;
    to.opt_prep_priority = from.opt_prep_priority;
    to.opt_mask_per_face = from.opt_mask_per_face;
    to.opt_no_static_mask = from.opt_no_static_mask;
    to.opt_mask_tile = from.opt_mask_tile;
    to.opt_no_zero_copy = from.opt_no_zero_copy;
    *out = ctx;
    return C5_OK;
}

void c5_destroy(c5_ctx* ctx) {
    if (!ctx) return;
    for (c5_ctx* lane : ctx->lanes) c5_destroy(lane); // our own lanes go first (each leaves ctx->siblings)
    ctx->lanes.clear();
    if (ctx->parent) { // a sibling leaves its parent's list
        auto& sib = ctx->parent->siblings;
        sib.erase(std::remove(sib.begin(), sib.end(), ctx), sib.end());
    }
    for (c5_ctx* s : ctx->siblings) { // a parent going first takes the shared mesh with it
        if (!kHostSim && !s->dev.empty()) {
            cudaSetDevice(s->dev[0]->device);
            cudaDeviceSynchronize();
        }
        s->parent = nullptr;
        s->has_mesh = false;
        if (!s->dev.empty()) s->dev[0]->origin = nullptr;
    }
    ctx->siblings.clear();
    nccl_close(ctx->nccl);
    ctx->nccl = nullptr;
    if (!kHostSim && !ctx->dev.empty() && ctx->dev[0]->stream) {
        cudaSetDevice(ctx->dev[0]->device);
        cudaStreamSynchronize(ctx->dev[0]->stream);
    }
    for (void* p : ctx->registered) {
        if (!kHostSim) cudaHostUnregister(p);
    }
    for (auto& im : ctx->images) {
        if (im.second) dev_free(im.first);
        else if (!kHostSim) cudaIpcCloseMemHandle(im.first);
    }
    for (auto& dp : ctx->dev) {
        DeviceState& d = *dp;
        if (!kHostSim) {
            cudaSetDevice(d.device);
            if (d.stream) cudaStreamSynchronize(d.stream);
        }
#ifdef C5_EXPERIMENTS
        if (!kHostSim && d.trace_launches > 0) {
            if (const char* path = std::getenv("C5_TRACE_FILE")) {
                cudaDeviceSynchronize();
                std::vector<unsigned long long> h(d.trace.n);
                if (cudaMemcpy(h.data(), d.trace.p, d.trace.bytes(), cudaMemcpyDeviceToHost) == cudaSuccess) {
                    if (FILE* f = std::fopen(path, "a")) {
                        for (int l = 0; l < d.trace_launches; l++) {
                            for (unsigned b = 0; b < d.trace_grid[l]; b++) {
                                const unsigned long long* r = h.data() + (static_cast<size_t>(l) * kTraceBlocks + b) * 4;
                                std::fprintf(f, "%p %d %u %llu %llu %llu\n", static_cast<void*>(ctx), l, b, r[2], r[0], r[1]);
                            }
                        }
                        std::fclose(f);
                    }
                }
            }
        }
#endif
        host_pinned_free(d.h_counters);
        host_pinned_free(d.h_row_cost);
        if (!kHostSim) {
            for (cudaEvent_t e : d.tl_events) cudaEventDestroy(e);
            if (d.ev_walk) cudaEventDestroy(d.ev_walk);
            if (d.ev_done) cudaEventDestroy(d.ev_done);
            if (d.ev_fork) cudaEventDestroy(d.ev_fork);
            if (d.ev_join) cudaEventDestroy(d.ev_join);
            if (d.prep_stream) {
                cudaStreamSynchronize(d.prep_stream);
                cudaStreamDestroy(d.prep_stream);
            }
            for (auto& e : d.ev) {
                if (e) cudaEventDestroy(e);
            }
            if (d.stream) cudaStreamDestroy(d.stream);
        }
        // DevBuf destructors free on the current device
        dp.reset();
    }
    delete ctx;
}

const char* c5_last_error(const c5_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lock(g_create_err_mu);
    static thread_local std::string copy;
    copy = g_create_err;
    return copy.c_str();
}

int c5_upload_mesh(c5_ctx* ctx, const double* points_xyz, int64_t n_points, const int32_t* tet_vertices,
                   int64_t n_tets, const double* alpha, const double* q) {
    if (!ctx) return C5_E_INVALID;
    return guarded(ctx, [&] {
        if (ctx->parent) fail(C5_E_STATE, "upload_mesh: a sibling context shares its parent's mesh; upload through the parent");
        if (!points_xyz || !tet_vertices || !alpha || !q) fail(C5_E_INVALID, "upload_mesh: NULL array");
        if (!ctx->siblings.empty() && !kHostSim) C5_CUDA(cudaDeviceSynchronize()); // siblings may be rendering from it
        ctx->dev[0]->mesh_version++;
        ctx->has_mesh = false;
        const size_t before = dev_bytes_in_use();
        for (auto& dp : ctx->dev) {
            use_device(*dp);
            g_launch_counter = &dp->launches;
            build_mesh(*dp, points_xyz, n_points, tet_vertices, n_tets, alpha, q);
        }
        ctx->info.n_points = n_points;
        ctx->info.n_tets = n_tets;
        ctx->info.n_boundary_faces = ctx->dev[0]->n_bfaces;
        ctx->info.n_bvh_nodes = ctx->dev[0]->n_bfaces - 1;
        ctx->info.device_bytes = static_cast<int64_t>((dev_bytes_in_use() - before) / ctx->dev.size());
        ctx->has_mesh = true;
    });
}

int c5_upload_solids(c5_ctx* ctx, const double* tet_points, int64_t n_tets, int32_t follows_view) {
    if (!ctx) return C5_E_INVALID;
    return guarded(ctx, [&] {
        if (ctx->parent) fail(C5_E_STATE, "upload_solids: upload through the parent context");
        if (!ctx->siblings.empty() && !kHostSim) C5_CUDA(cudaDeviceSynchronize());
        ctx->dev[0]->mesh_version++;
        upload_solids(ctx, tet_points, n_tets, follows_view);
    });
}

int c5_clear_solids(c5_ctx* ctx) {
    if (!ctx) return C5_E_INVALID;
    return guarded(ctx, [&] {
        if (ctx->parent) fail(C5_E_STATE, "clear_solids: clear through the parent context");
        if (!ctx->siblings.empty() && !kHostSim) C5_CUDA(cudaDeviceSynchronize());
        ctx->dev[0]->mesh_version++;
        for (auto& dp : ctx->dev) {
            use_device(*dp);
            for (SolidSet* ss : {&dp->solid_follow, &dp->solid_static}) {
                ss->pts0.release();
                ss->pts_view.release();
                ss->faces.release();
                ss->n = 0;
                ss->n_faces = 0;
                ss->extent = 0.0;
            }
            dp->mask_static_valid = false;
        }
        ctx->info.n_solid_tets = 0;
    });
}

int c5_mesh_info_get(const c5_ctx* ctx, c5_mesh_info* out) {
    if (!ctx || !out) return C5_E_INVALID;
    *out = ctx->info;
    return C5_OK;
}

int c5_render(c5_ctx* ctx, const c5_view* view, double* out, c5_stats* stats) {
    return c5_render_raw(ctx, view, out, nullptr, nullptr, stats);
}

int c5_render_raw(c5_ctx* ctx, const c5_view* view, double* out, uint32_t* steps, uint8_t* solid_mask,
                  c5_stats* stats) {
    if (!ctx) return C5_E_INVALID;
    return guarded(ctx, [&] {
        if (ctx->dev.size() > 1) {
            render_multi_device(ctx, view, out, steps, solid_mask, stats);
            return;
        }
        submit_view(ctx, view, out, steps, solid_mask, 0);
        finish_view(ctx, ctx, stats);
    });
}

int c5_set_views_in_flight(c5_ctx* ctx, int32_t n) {
    if (!ctx) return C5_E_INVALID;
    return guarded(ctx, [&] {
        if (n < 1 || n > C5_MAX_IN_FLIGHT) fail(C5_E_INVALID, "set_views_in_flight: 1 .. C5_MAX_IN_FLIGHT");
        if (ctx->parent || ctx->dev.size() != 1) fail(C5_E_INVALID, "set_views_in_flight: single-device contexts that are not siblings");
        for (unsigned k = 0; k <= ctx->lanes.size(); k++) {
            if (lane_of(ctx, k)->pending.active) fail(C5_E_STATE, "set_views_in_flight: views are in flight");
        }
        while (static_cast<int>(ctx->lanes.size()) + 1 > n) { // lanes that are no longer wanted
            c5_destroy(ctx->lanes.back());
            ctx->lanes.pop_back();
        }
        ctx->views_in_flight = n;
        ctx->next_lane = 0;
    });
}

int c5_render_submit(c5_ctx* ctx, const c5_view* view, double* out, uint64_t* ticket) {
    if (!ctx || !ticket) return C5_E_INVALID;
    return guarded(ctx, [&] {
        if (ctx->parent || ctx->dev.size() != 1) fail(C5_E_INVALID, "render_submit: single-device contexts that are not siblings");
        ensure_lanes(ctx);
        // the next lane in turn, or any free one
        const unsigned n = static_cast<unsigned>(ctx->lanes.size()) + 1;
        c5_ctx* lane = nullptr;
        for (unsigned k = 0; k < n && !lane; k++) {
            c5_ctx* cand = lane_of(ctx, (ctx->next_lane + k) % n);
            if (!cand->pending.active) {
                lane = cand;
                ctx->next_lane = (ctx->next_lane + k + 1) % n;
            }
        }
        if (!lane) fail(C5_E_STATE, "render_submit: " + std::to_string(n) + " views are in flight already (c5_render_wait one first)");
        const uint64_t t = ctx->next_ticket++;
        try {
            g_launch_counter = &lane->dev[0]->launches;
            submit_view(lane, view, out, nullptr, nullptr, t, /*in_place_allowed=*/ctx->views_in_flight == 1);
        } catch (const Error& e) {
            if (lane != ctx) ctx->err = e.text; // (guarded() reports through ctx)
            throw;
        }
        *ticket = t;
    });
}

int c5_render_wait(c5_ctx* ctx, uint64_t ticket, c5_stats* stats) {
    if (!ctx) return C5_E_INVALID;
    return guarded(ctx, [&] {
        for (unsigned k = 0; k <= ctx->lanes.size(); k++) {
            c5_ctx* lane = lane_of(ctx, k);
            if (lane->pending.active && lane->pending.ticket == ticket && ticket != 0) {
                finish_view(ctx, lane, stats);
                return;
            }
        }
        fail(C5_E_INVALID, "render_wait: no view with this ticket is in flight");
    });
}

int c5_render_device(c5_ctx* ctx, const c5_view* view, void* d_out, void* stream, c5_stats* stats) {
    if (!ctx) return C5_E_INVALID;
    return guarded(ctx, [&] {
        follow_parent(ctx);
        if (!ctx->has_mesh) fail(C5_E_STATE, "render: no mesh uploaded");
        if (!d_out) fail(C5_E_INVALID, "render_device: d_out is NULL");
        if (reinterpret_cast<uintptr_t>(d_out) & 15u) fail(C5_E_INVALID, "render_device: d_out must be 16-byte aligned");
        if (ctx->dev.size() != 1) fail(C5_E_INVALID, "render_device: single-device contexts only");
        if (ctx->pending.active) fail(C5_E_STATE, "render_device: this context has a submitted view in flight");
        const ViewPlan p = plan_view(view);
        DeviceState& d = *ctx->dev[0];
        // run on the CALLER's stream (NULL = the legacy default stream, which is what torch's
        // default "current stream" is) so the result is ordered with the caller's later work
        cudaStream_t own = d.stream;
        if (!kHostSim) d.stream = static_cast<cudaStream_t>(stream);
        try {
            enqueue_view(d, view, p, false, static_cast<double*>(d_out)); // the walk stores into the caller's buffer
            // stats == NULL: fire and forget — nothing is read back and the stream is not synchronised,
            // so successive views (and the caller's gather) pipeline on the device
            if (stats) collect_stats(ctx, d, view, p, stats, 4);
        } catch (...) {
            d.stream = own;
            throw;
        }
        d.stream = own;
    });
}

int c5_last_row_cost(c5_ctx* ctx, uint64_t* rows, int32_t n_rows) {
    if (!ctx || !rows || n_rows < 0) return C5_E_INVALID;
    for (int32_t j = 0; j < n_rows; j++) {
        rows[j] = static_cast<size_t>(j) < ctx->last_row_cost.size() ? ctx->last_row_cost[static_cast<size_t>(j)] : 0;
    }
    return C5_OK;
}

int c5_image_create(c5_ctx* ctx, uint64_t bytes, void** d_ptr, uint8_t handle[C5_IPC_HANDLE_BYTES]) {
    if (!ctx || !d_ptr || !handle) return C5_E_INVALID;
    return guarded(ctx, [&] {
        if (bytes == 0) fail(C5_E_INVALID, "image_create: zero bytes");
        use_device(*ctx->dev[0]);
        void* p = dev_alloc(bytes);
        std::memset(handle, 0, C5_IPC_HANDLE_BYTES);
        uint64_t offset = 0;
        if (kHostSim) {
            std::memcpy(handle, &p, sizeof(p)); // same-process stand-in (CPU logic tests only)
        } else {
            static_assert(sizeof(cudaIpcMemHandle_t) + 16 <= C5_IPC_HANDLE_BYTES, "IPC handle size");
            cudaIpcMemHandle_t h;
            try {
                offset = offset_in_allocation(p);
                C5_CUDA(cudaIpcGetMemHandle(&h, p));
            } catch (...) {
                dev_free(p);
                throw;
            }
            std::memcpy(handle, &h, sizeof(h));
        }
        std::memcpy(handle + 64, &offset, sizeof(offset));
        std::memcpy(handle + 72, &bytes, sizeof(bytes));
        ctx->images.emplace_back(p, true);
        *d_ptr = p;
    });
}

int c5_image_open(c5_ctx* ctx, const uint8_t handle[C5_IPC_HANDLE_BYTES], void** d_ptr) {
    if (!ctx || !d_ptr || !handle) return C5_E_INVALID;
    return guarded(ctx, [&] {
        use_device(*ctx->dev[0]);
        void* p = nullptr;
        if (kHostSim) {
            std::memcpy(&p, handle, sizeof(p));
        } else {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handle, sizeof(h));
            // maps the owner's allocation into this process and enables peer access to its device
            C5_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        }
        uint64_t offset = 0;
        std::memcpy(&offset, handle + 64, sizeof(offset));
        ctx->images.emplace_back(p, false); // the mapping's base: what cudaIpcCloseMemHandle wants back
        ctx->image_offsets.emplace_back(static_cast<char*>(p) + offset, p);
        *d_ptr = static_cast<char*>(p) + offset;
    });
}

int c5_image_close(c5_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return C5_E_INVALID;
    return guarded(ctx, [&] {
        for (size_t k = 0; k < ctx->image_offsets.size(); k++) { // an imported image is known by its offset pointer
            if (ctx->image_offsets[k].first != d_ptr) continue;
            d_ptr = ctx->image_offsets[k].second;
            ctx->image_offsets.erase(ctx->image_offsets.begin() + static_cast<long>(k));
            break;
        }
        for (size_t k = 0; k < ctx->images.size(); k++) {
            if (ctx->images[k].first != d_ptr) continue;
            const bool owner = ctx->images[k].second;
            ctx->images.erase(ctx->images.begin() + static_cast<long>(k));
            use_device(*ctx->dev[0]);
            if (owner) dev_free(d_ptr);
            else if (!kHostSim) C5_CUDA(cudaIpcCloseMemHandle(d_ptr));
            return;
        }
        fail(C5_E_INVALID, "image_close: not an image of this context");
    });
}

int c5_host_register(c5_ctx* ctx, void* ptr, uint64_t bytes) {
    if (!ctx || !ptr || bytes == 0) return C5_E_INVALID;
    return guarded(ctx, [&] {
        use_device(*ctx->dev[0]);
        if (!kHostSim) C5_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
        ctx->registered.push_back(ptr);
    });
}

int c5_host_unregister(c5_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return C5_E_INVALID;
    return guarded(ctx, [&] {
        for (size_t k = 0; k < ctx->registered.size(); k++) {
            if (ctx->registered[k] != ptr) continue;
            ctx->registered.erase(ctx->registered.begin() + static_cast<long>(k));
            if (!kHostSim) C5_CUDA(cudaHostUnregister(ptr));
            return;
        }
        fail(C5_E_INVALID, "host_unregister: not registered through this context");
    });
}

uint64_t c5_kernel_launches(const c5_ctx* ctx) {
    uint64_t n = 0;
    if (ctx) {
        for (const auto& d : ctx->dev) n += d->launches;
        for (const c5_ctx* lane : ctx->lanes) n += c5_kernel_launches(lane);
    }
    return n;
}

int c5_debug_set(c5_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return C5_E_INVALID;
    return guarded(ctx, [&] {
        const std::string k = key;
        std::vector<c5_ctx*> all{ctx};
        all.insert(all.end(), ctx->siblings.begin(), ctx->siblings.end()); // the caller's siblings and our lanes
        for (c5_ctx* c : all) {
            for (auto& dp : c->dev) {
                DeviceState& d = *dp;
                if (k == "graze_list") d.opt_graze_list = static_cast<int>(value);
                else if (k == "query_budget") d.opt_query_budget = static_cast<int>(value);
                else if (k == "serial_list") d.opt_serial_list = static_cast<int>(value);
                else if (k == "graze_blocks") d.opt_graze_blocks = static_cast<int>(value);
                else if (k == "prep_priority") d.opt_prep_priority = static_cast<int>(value & 3);
                else if (k == "mask_tile") d.opt_mask_tile = static_cast<int>(value);
                else if (k == "mask_per_face") d.opt_mask_per_face = static_cast<int>(value < 0 ? 0 : value > 6 ? 6 : value);
                else if (k == "no_static_mask") {
                    d.opt_no_static_mask = value != 0;
                    d.mask_static_valid = false;
                }
                else if (k == "no_zero_copy") d.opt_no_zero_copy = value != 0;
                else if (k == "timeline") {
                    if (value < 0 || value > 4096) fail(C5_E_INVALID, "debug_set: timeline 0 .. 4096 views");
                    if (kHostSim) continue;
                    use_device(d);
                    C5_CUDA(cudaStreamSynchronize(d.stream));
                    for (cudaEvent_t e : d.tl_events) cudaEventDestroy(e);
                    d.tl_events.clear();
                    d.tl_views = 0;
                    for (int64_t i = 0; i < value * kTimelinePhases; i++) {
                        cudaEvent_t e = nullptr;
                        C5_CUDA(cudaEventCreate(&e));
                        d.tl_events.push_back(e);
                    }
                } else {
                    fail(C5_E_INVALID, "debug_set: unknown key '" + k + "'");
                }
            }
        }
    });
}

int c5_timeline_read(c5_ctx* ctx, void* origin, float* ms_out, int32_t max_views, int32_t* n_views) {
    if (!ctx || !ms_out || !n_views || max_views < 0) return C5_E_INVALID;
    return guarded(ctx, [&] {
        *n_views = 0;
        if (kHostSim) return;
        if (!origin) fail(C5_E_INVALID, "timeline_read: origin event is NULL");
        DeviceState& d = *ctx->dev[0];
        use_device(d);
        const uint64_t cap = d.tl_events.size() / kTimelinePhases;
        if (cap == 0) return;
        const uint64_t have = std::min<uint64_t>({d.tl_views, cap, static_cast<uint64_t>(max_views)});
        for (uint64_t i = 0; i < have; i++) {
            const uint64_t view = d.tl_views - have + i;
            for (int ph = 0; ph < kTimelinePhases; ph++) {
                cudaEvent_t e = d.tl_events[(view % cap) * kTimelinePhases + static_cast<size_t>(ph)];
                C5_CUDA(cudaEventSynchronize(e));
                C5_CUDA(cudaEventElapsedTime(ms_out + i * kTimelinePhases + ph, static_cast<cudaEvent_t>(origin), e));
            }
        }
        *n_views = static_cast<int32_t>(have);
    });
}

} // extern "C"
