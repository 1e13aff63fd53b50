// c5_prefetch.cu — L2 slab prefetch for the tet walk.
//
// The walk reads every cell and vertex under the image once (DRAM traffic per view == the mesh
// once, profiles/r01_walk_full_c3*.csv) but at 2 % of the DRAM bandwidth: each ray is a pointer
// chase through the neighbour table, a warp-step waits for its slowest lane, and with 96 sectors
// per warp-step almost every step has a lane that misses L2. What the rays need is nevertheless
// known in bulk: blocks sweep the image in strips of rows, and the part of the mesh under a strip
// is a slab. This file finds, per view, which 4 KB chunks of the (Morton-ordered) cell and vertex
// arrays lie in which strip, so that the slab of the strip AHEAD of the one being started can be
// pulled into the 126 MB L2 with bulk prefetches (cp.async.bulk.prefetch.L2) at DRAM bandwidth,
// and the rays then find their cells at L2 latency. Consumed lines are never touched again, so
// LRU turns L2 into a sliding window: [consumed | in flight | prefetched ahead].
//
// Once per mesh: a bounding sphere per chunk in the file frame (rotation invariant radius).
// Once per view: chunk_rows[c] = first and last strip the sphere's y range touches, relative to
// the band (kChunkOutside if it misses the band or the window); chunks of the first strips are
// prefetched right here, the rest by the walk kernel's strip leaders (c5_walk.cu).
#include "c5_internal.h"

namespace c5 {

namespace {

struct ChunkSphereOp {
    const Cell* cells;
    const double *px, *py, *pz;
    int64_t n_tets;
    ChunkSphere* out;
    C5_HD void operator()(int64_t c) const {
        const int64_t t0 = c * kCellChunk;
        const int64_t t1 = t0 + kCellChunk < n_tets ? t0 + kCellChunk : n_tets;
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int64_t t = t0; t < t1; t++) {
            for (int k = 0; k < 4; k++) {
                const int v = cells[t].v[k];
                const double p[3] = {px[v], py[v], pz[v]};
                for (int a = 0; a < 3; a++) {
                    lo[a] = p[a] < lo[a] ? p[a] : lo[a];
                    hi[a] = p[a] > hi[a] ? p[a] : hi[a];
                }
            }
        }
        const double cx = 0.5 * (lo[0] + hi[0]), cy = 0.5 * (lo[1] + hi[1]), cz = 0.5 * (lo[2] + hi[2]);
        double r2 = 0.0;
        for (int64_t t = t0; t < t1; t++) {
            for (int k = 0; k < 4; k++) {
                const int v = cells[t].v[k];
                const double dx = px[v] - cx, dy = py[v] - cy, dz = pz[v] - cz;
                const double q = dx * dx + dy * dy + dz * dz;
                r2 = q > r2 ? q : r2;
            }
        }
        ChunkSphere s;
        s.x = static_cast<float>(cx);
        s.y = static_cast<float>(cy);
        s.z = static_cast<float>(cz);
        s.r = static_cast<float>(sqrt(r2) * 1.0001 + 1e-6); // float rounding of the centre stays inside
        out[c] = s;
    }
};

struct VertexChunkSphereOp {
    const double *px, *py, *pz;
    int64_t n_pts;
    ChunkSphere* out;
    C5_HD void operator()(int64_t c) const {
        const int64_t i0 = c * kVtxChunk;
        const int64_t i1 = i0 + kVtxChunk < n_pts ? i0 + kVtxChunk : n_pts;
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int64_t i = i0; i < i1; i++) {
            const double p[3] = {px[i], py[i], pz[i]};
            for (int a = 0; a < 3; a++) {
                lo[a] = p[a] < lo[a] ? p[a] : lo[a];
                hi[a] = p[a] > hi[a] ? p[a] : hi[a];
            }
        }
        const double cx = 0.5 * (lo[0] + hi[0]), cy = 0.5 * (lo[1] + hi[1]), cz = 0.5 * (lo[2] + hi[2]);
        double r2 = 0.0;
        for (int64_t i = i0; i < i1; i++) {
            const double dx = px[i] - cx, dy = py[i] - cy, dz = pz[i] - cz;
            const double q = dx * dx + dy * dy + dz * dz;
            r2 = q > r2 ? q : r2;
        }
        ChunkSphere s;
        s.x = static_cast<float>(cx);
        s.y = static_cast<float>(cy);
        s.z = static_cast<float>(cz);
        s.r = static_cast<float>(sqrt(r2) * 1.0001 + 1e-6);
        out[c] = s;
    }
};

struct RotF {
    int axis[kMaxRot];
    double c[kMaxRot], s[kMaxRot], x0[kMaxRot];
    int n;
};

struct StripParams {
    const ChunkSphere* spheres; // cell chunks, then vertex chunks
    uint32_t* chunk_rows;
    int64_t n_cell_chunks, n_chunks;
    const Cell* cells;
    const Vtx* vrot;
    int64_t n_tets, n_pts;
    RotF rot;
    double x_lo, x_hi;          // window in x
    double y_min, inv_step_y;   // pixel row of y: (y - y_min) * inv_step_y
    int row_begin, row_end;
    int strip_rows;             // pixel rows per strip
    int first_strips;           // chunks touching strips [0, first_strips) are prefetched here
};

} // namespace

__global__ void __launch_bounds__(256) classify_chunks(const StripParams S) {
    const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (c >= S.n_chunks) return;
    const ChunkSphere sp = S.spheres[c];
    double x = sp.x, y = sp.y, z = sp.z;
    for (int k = 0; k < S.rot.n; k++) { // tetra.cpp:44-62 (no bit-exactness needed: this is a hint)
        const double cc = S.rot.c[k], ss = S.rot.s[k];
        if (S.rot.axis[k] == 0) {
            const double y0 = y;
            y = y * cc - z * ss;
            z = y0 * ss + z * cc;
        } else {
            x -= S.rot.x0[k];
            const double x1 = x;
            x = x * cc - z * ss;
            z = x1 * ss + z * cc;
            x += S.rot.x0[k];
        }
    }
    const double r = sp.r;
    uint32_t packed = kChunkOutside;
    if (x + r >= S.x_lo && x - r <= S.x_hi) {
        const double row_lo = floor((y - r - S.y_min) * S.inv_step_y), row_hi = ceil((y + r - S.y_min) * S.inv_step_y);
        if (row_hi >= S.row_begin && row_lo < S.row_end) {
            const int lo = static_cast<int>(row_lo < S.row_begin ? S.row_begin : row_lo) - S.row_begin;
            const int hi = static_cast<int>(row_hi >= S.row_end ? S.row_end - 1 : row_hi) - S.row_begin;
            const int s_lo = lo / S.strip_rows, s_hi = hi / S.strip_rows;
            packed = static_cast<uint32_t>(s_lo < 0xFFFE ? s_lo : 0xFFFE) |
                     (static_cast<uint32_t>(s_hi < 0xFFFE ? s_hi : 0xFFFE) << 16);
            if (s_lo < S.first_strips) prefetch_chunk(S.cells, S.vrot, S.n_cell_chunks, S.n_tets, S.n_pts, c);
        }
    }
    S.chunk_rows[c] = packed;
}

void build_chunk_spheres(DeviceState& d) {
    d.n_cell_chunks = (d.n_tets + kCellChunk - 1) / kCellChunk;
    d.n_vtx_chunks = (d.n_pts + kVtxChunk - 1) / kVtxChunk;
    const size_t n = static_cast<size_t>(d.n_cell_chunks + d.n_vtx_chunks);
    d.chunk_spheres.alloc(n);
    d.chunk_rows.alloc(n);
    for_each(d.stream, d.n_cell_chunks, ChunkSphereOp{d.cells.p, d.px.p, d.py.p, d.pz.p, d.n_tets, d.chunk_spheres.p});
    for_each(d.stream, d.n_vtx_chunks,
             VertexChunkSphereOp{d.px.p, d.py.p, d.pz.p, d.n_pts, d.chunk_spheres.p + d.n_cell_chunks});
    stream_sync(d.stream);
}

void launch_classify_chunks(DeviceState& d, const Rot* rot, int n_rot, const SlabPlan& plan) {
    if (kHostSim) return; // a cache hint: nothing to simulate
    StripParams S{};
    S.spheres = d.chunk_spheres.p;
    S.chunk_rows = d.chunk_rows.p;
    S.n_cell_chunks = d.n_cell_chunks;
    S.n_chunks = d.n_cell_chunks + d.n_vtx_chunks;
    S.cells = d.cells.p;
    S.vrot = d.vrot.p;
    S.n_tets = d.n_tets;
    S.n_pts = d.n_pts;
    S.rot.n = n_rot;
    for (int k = 0; k < n_rot; k++) {
        S.rot.axis[k] = rot[k].axis;
        S.rot.c[k] = rot[k].c;
        S.rot.s[k] = rot[k].s;
        S.rot.x0[k] = rot[k].x0;
    }
    S.x_lo = plan.x_lo;
    S.x_hi = plan.x_hi;
    S.y_min = plan.y_min;
    S.inv_step_y = 1.0 / plan.step_y;
    S.row_begin = plan.row_begin;
    S.row_end = plan.row_end;
    S.strip_rows = plan.strip_rows;
    S.first_strips = plan.first_strips;
    count_launch();
    const int64_t n = S.n_chunks;
    classify_chunks<<<static_cast<unsigned>((n + 255) / 256), 256, 0, d.stream>>>(S);
    C5_CUDA(cudaGetLastError());
}

} // namespace c5
