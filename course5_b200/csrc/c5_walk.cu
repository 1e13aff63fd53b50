// c5_walk.cu — K3 + K4 + K4g: per-pixel ray entry (LBVH over boundary faces), the tet walk, and the
// warp-per-ray kernel for rays that graze the boundary.
//
// Replaces, for one view, the reference's
//   plane::find_intersections   plane.cpp:184-192 (scan-convert every face of every tet, one
//                               mutex-guarded push per tet-step, line.cpp:29-67,229-232)
//   line::calculate_intersections / direct_calculate_ray_value / integrate_ray_value_by_i
//                               line.cpp:84-227 (2 face-plane z's per record, per-pixel sort,
//                               tau = sum dz*alpha, I recurrence with the alpha clamp)
//   plane::trace_rays           plane.cpp:144-172 (driver, float cast)
// by one thread per pixel that (1) finds the lowest boundary face under the pixel whose outward
// normal points to -z, (2) walks tet to tet through the face-neighbour table in +z order — the
// order in which the reference integrates I (line.cpp:206: from the record with the lowest z
// to the highest) — and (3) on leaving the mesh asks the BVH for the next entries above (meshes
// with cavities or a bumpy silhouette are crossed several times). Rays with many crossings, or
// whose searches are expensive, are deferred to grazing_rays_* (one warp per ray, further down).
//
// File map: loads -> entry search (one thread) -> one crossing (FP64, FP32) -> one ray, in
// warp-synchronised phases (trace_ray) -> grazing rays: composition, serial form, warp-cooperative
// collection / sort / walk -> kernels (walk_block, fill_background, graze_block) -> launch_walk.
//
// Geometry of one step. All rays are parallel to z, so "which face does the ray leave through"
// is a 2-D question about the projected tet. The entry face (a,b,c) is kept counter-clockwise in
// projection, with coordinates relative to the pixel. With d the fourth vertex and
//     s_v = orient2(d, v)   (z-component of (d-p) x (v-p)),  v in {a,b,c},
// the ray leaves through face (d,a,b) iff s_a >= 0 > s_b, through (d,b,c) iff s_b >= 0 > s_c,
// through (d,c,a) iff s_c >= 0 > s_a. orient2 is exactly antisymmetric (c5_types.h), so two tets
// sharing an edge agree on the side the ray passes: the walk is watertight without epsilons.
// Two of the s values are also barycentric weights of the pixel in the exit face (the third is one
// more orient2 of the two vertices that stay), so the exit z costs one divide; it is the entry z
// of the next tet, i.e. each face plane is evaluated once per ray, not twice as in line.cpp:103-122.
//
// Per step the thread reads one 64-byte Cell and ONE new 32-byte vertex; the other three
// vertices and their ids stay in registers (the 72 B/step algorithmic figure of SURVEY.md §8d).
#include "c5_internal.h"

namespace c5 {

namespace {

constexpr int kStack = 64;        // private LBVH traversal stack; a search that would overflow it is an incomplete search
constexpr int kBlock = 128;       // 4 warps: 2 x 2 warp tiles of 8 x 4 pixels
constexpr int kTileX = 16, kTileY = 8;

struct WalkParams {
    const Cell* cells;
    const Vtx* vrot;
    const BFace* bfaces;
    const BvhNode* nodes;
    const double* xs;
    const double* ys;
    const uint8_t* mask;  // may be null
    double* out;          // band buffer: {tau, I} at ((j - row_begin) * res_x + i)
    uint32_t* steps;      // may be null; same indexing
    unsigned long long* counters;
    unsigned long long* row_cost; // [res_y]
    DeferredRay* queue;   // rays handed to the grazing-ray kernel; counters[kDeferred] of them
    int res_x, res_y, row_begin, row_end;
    int i0, i1, j0, j1;   // pixel rectangle [i0,i1) x [j0,j1) of the band that can see the mesh: the tiles cover it
    int n_tiles_x, n_tiles_y, n_macro_x;
    int graze_cap;        // entries one cooperative collection may hold (<= kGrazeList)
    int serial_cap;       // same for the serial (host loop) form (<= kSerialList)
    int query_budget;     // BVH nodes one thread of the pixel kernel may visit per search before deferring the ray
    int max_steps;
    int round_float;
    double alpha_limit;
#ifdef C5_EXPERIMENTS     // libc5gpu_exp.so only (scripts/): block timeline, shared-memory BVH top, step records
    unsigned long long* trace; // per block {start ns, end ns, sm, block}
    int top_nodes;        // BVH nodes [0, top_nodes) are staged in shared memory
    const StepRec* recs;
#endif
};

#ifdef C5_EXPERIMENTS
C5_HD int staged_nodes(const WalkParams& P) { return P.top_nodes; }
#else
C5_HD int staged_nodes(const WalkParams&) { return 0; }
#endif

// ---- loads through the read-only path ----------------------------------------------------------
struct CellData {
    int4 v, nbr, apex;
    double alpha, s;
};

// sm_100 has 256-bit global loads (SASS LDG.E.ENL2.256): a 64-byte cell is two of them and a
// 32-byte vertex one, i.e. 3 load instructions per tet-step instead of 6 LDG.128.
template <bool kWide>
C5_HD CellData load_cell(const Cell* cells, int t) {
    CellData c;
#ifdef __CUDA_ARCH__
    if (kWide) {
        const char* p = reinterpret_cast<const char*>(cells + t);
        long long as0, as1;
        asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(c.v.x), "=r"(c.v.y), "=r"(c.v.z), "=r"(c.v.w), "=r"(c.nbr.x), "=r"(c.nbr.y), "=r"(c.nbr.z),
                       "=r"(c.nbr.w)
                     : "l"(p));
        int a0, a1, a2, a3;
        int lo0, hi0, lo1, hi1;
        asm volatile("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3), "=r"(lo0), "=r"(hi0), "=r"(lo1), "=r"(hi1)
                     : "l"(p + 32));
        c.apex = make_int4(a0, a1, a2, a3);
        c.alpha = __hiloint2double(hi0, lo0);
        c.s = __hiloint2double(hi1, lo1);
        (void)as0;
        (void)as1;
    } else {
        const int4* p = reinterpret_cast<const int4*>(cells + t);
        c.v = __ldg(p);
        c.nbr = __ldg(p + 1);
        c.apex = __ldg(p + 2);
        const double2 as = __ldg(reinterpret_cast<const double2*>(p + 3));
        c.alpha = as.x;
        c.s = as.y;
    }
#else
    const Cell& s = cells[t];
    c.v = make_int4(s.v[0], s.v[1], s.v[2], s.v[3]);
    c.nbr = make_int4(s.nbr[0], s.nbr[1], s.nbr[2], s.nbr[3]);
    c.apex = make_int4(s.apex[0], s.apex[1], s.apex[2], s.apex[3]);
    c.alpha = s.alpha;
    c.s = s.s;
#endif
    return c;
}

C5_HD void load_vtx(const Vtx* vrot, int id, double& x, double& y, double& z) {
#ifdef __CUDA_ARCH__
    [[maybe_unused]] double w;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(vrot + id));
#else
    x = vrot[id].x;
    y = vrot[id].y;
    z = vrot[id].z;
#endif
}

#ifdef C5_EXPERIMENTS
C5_HD StepRec load_rec(const StepRec* recs, uint32_t face) {
    StepRec r;
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.w[0]), "=l"(r.w[1]), "=l"(r.w[2]), "=l"(r.w[3])
                 : "l"(recs + face));
#else
    r = recs[face];
#endif
    return r;
}
#endif

// Next step's cell and vertex are known as soon as the exit face is (Cell::nbr / Cell::apex), long
// before this step's divide and exp have retired: asking L1 for them now overlaps their L2/DRAM
// latency with that math at no register cost.
C5_HD void prefetch_l1(const void* p) {
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ---- entry search --------------------------------------------------------------------------------

// Entry list of one ray: the lowest entry faces above z_after, sorted by (z, leaf). A ray through a
// convex mesh has one entry; cavities add a few; a ray grazing a bumpy boundary (the jittered side
// walls of the synthetic grids, seen edge-on) has ~100. A thread walking such a ray alone is a chain
// of dependent BVH queries and short crossings, and a handful of them set the run time of a whole
// row band (measured: 1.2-1.3 ms for ANY band of the C3 README view, profiles/r01_exp_bands.jsonl).
// So the pixel kernel only ever asks for a short list (kInline entries); a ray whose list comes back
// full is handed to the grazing-ray kernel below, where a whole warp works on it.
template <int kCap>
struct EntryList {
    double z[kCap];
    int leaf[kCap];
    int n;
};

constexpr int kInline = 4;   // re-entry list of the pixel kernel: fewer entries than this are walked in place
constexpr int kSerialList = 64; // list of the serial (host loop) grazing path

// Depth at which the ray through (px, py) enters the mesh through boundary face `leaf`: inclusive
// point-in-triangle test with the same orientation predicate the walk uses, then the barycentric z.
// False if the face is not under the pixel, faces +z, or lies at or below z_after.
C5_HD bool entry_depth(const WalkParams& P, int leaf, double px, double py, double z_after, double& z_out) {
#ifdef __CUDA_ARCH__
    const int4 f = __ldg(reinterpret_cast<const int4*>(P.bfaces + leaf));
#else
    const BFace& bf = P.bfaces[leaf];
    const int4 f = make_int4(bf.a, bf.b, bf.c, bf.tet);
#endif
    double ax, ay, az, bx, by, bz, cx, cy, cz;
    load_vtx(P.vrot, f.x, ax, ay, az);
    load_vtx(P.vrot, f.y, bx, by, bz);
    load_vtx(P.vrot, f.z, cx, cy, cz);
    ax -= px; ay -= py;
    bx -= px; by -= py;
    cx -= px; cy -= py;
    const double o_ab = orient2(ax, ay, bx, by);
    const double o_bc = orient2(bx, by, cx, cy);
    const double o_ca = orient2(cx, cy, ax, ay);
    // outward normal towards -z  <=>  clockwise in projection  <=>  all three <= 0 inside
    if (!(o_ab <= 0 && o_bc <= 0 && o_ca <= 0)) return false;
    const double sum = o_ab + o_bc + o_ca;
    if (!(sum < 0)) return false;
    const double z = (o_bc * az + o_ca * bz + o_ab * cz) / sum;
    if (!(z > z_after)) return false;
    z_out = z;
    return true;
}

C5_HD bool entry_before(double z0, int leaf0, double z1, int leaf1) { // strict (z, leaf) order
    return z0 < z1 || (z0 == z1 && leaf0 < leaf1);
}

// Keeps the `cap` lowest entries, sorted; faces arrive roughly near-to-far, so the insertion from
// the back is close to linear.
template <int kCap>
C5_HD void insert_entry(EntryList<kCap>& L, int cap, double z, int leaf) {
    if (L.n == cap) {
        if (!entry_before(z, leaf, L.z[cap - 1], L.leaf[cap - 1])) return;
        L.n--; // the highest one falls off
    }
    int k = L.n++;
    while (k > 0 && entry_before(z, leaf, L.z[k - 1], L.leaf[k - 1])) {
        L.z[k] = L.z[k - 1];
        L.leaf[k] = L.leaf[k - 1];
        k--;
    }
    L.z[k] = z;
    L.leaf[k] = leaf;
}

// The (up to cap <= kCap) lowest entry faces strictly above z_after under pixel (px, py), by one
// thread with a private stack. cap = 1 is the classic nearest-hit search: once one face is found,
// every node whose box starts above it is pruned. L.n == cap on return means "there may be more
// above L.z[cap - 1]"; L.n < cap means the list is complete.
//
// `budget` bounds the nodes one thread may visit. A ray that runs along a boundary wall (inside or
// outside it) overlaps the boxes of every facet of that wall: proving "no further entry" then costs
// hundreds of dependent loads, and the few blocks holding such rays ran 60 % longer than all others
// (block timeline, profiles/r01_trace_band_c3.txt). A search that exceeds its budget returns false
// and the ray is handed to the grazing-ray kernel, where 32 lanes share the traversal.
template <int kCap>
C5_HD bool bvh_collect_entries(const WalkParams& P, const BvhNode* top, double px, double py, double z_after,
                               EntryList<kCap>& L, int cap, int budget) {
    int stack[kStack];
    int sp = 0;
    L.n = 0;
    int node = 0;
    const float fx_lo = f_round_down(px), fx_hi = f_round_up(px);
    const float fy_lo = f_round_down(py), fy_hi = f_round_up(py);
    // (one exit from the loop: with a `return` in the middle the compiler stopped reconverging the
    // warp after the search and the tet walk ran with half-empty warps — 8.3 instead of 4.7 ms on C3)
    bool complete = true;
    while (true) {
        if (--budget < 0) {
            complete = false;
            break;
        }
        const BvhNode* n = (top && node < staged_nodes(P)) ? (top + node) : (P.nodes + node);
#ifdef __CUDA_ARCH__
        const float4 bx = *reinterpret_cast<const float4*>(n->xlo); // xlo0 xlo1 xhi0 xhi1
        const float4 by = *reinterpret_cast<const float4*>(n->ylo);
        const float4 bz = *reinterpret_cast<const float4*>(n->zlo);
        const int2 ch = *reinterpret_cast<const int2*>(n->child);
#else
        const float4 bx = make_float4(n->xlo[0], n->xlo[1], n->xhi[0], n->xhi[1]);
        const float4 by = make_float4(n->ylo[0], n->ylo[1], n->yhi[0], n->yhi[1]);
        const float4 bz = make_float4(n->zlo[0], n->zlo[1], n->zhi[0], n->zhi[1]);
        const int2 ch = make_int2(n->child[0], n->child[1]);
#endif
        // once the list is full, nothing above its highest entry can get in
        const double z_cap = (L.n == cap) ? L.z[cap - 1] : INFINITY;
        // boxes are rounded outward and the pixel is widened to floats, so this never misses
        bool h0 = fx_hi >= bx.x && fx_lo <= bx.z && fy_hi >= by.x && fy_lo <= by.z &&
                  static_cast<double>(bz.z) > z_after && static_cast<double>(bz.x) <= z_cap;
        bool h1 = fx_hi >= bx.y && fx_lo <= bx.w && fy_hi >= by.y && fy_lo <= by.w &&
                  static_cast<double>(bz.w) > z_after && static_cast<double>(bz.y) <= z_cap;
        double z;
        if (h0 && ch.x < 0) {
            if (entry_depth(P, ~ch.x, px, py, z_after, z)) insert_entry(L, cap, z, ~ch.x);
            h0 = false;
        }
        if (h1 && ch.y < 0) {
            if (entry_depth(P, ~ch.y, px, py, z_after, z)) insert_entry(L, cap, z, ~ch.y);
            h1 = false;
        }
        if (h0 && h1) {
            const bool first0 = bz.x <= bz.y; // descend into the lower subtree first
            if (sp == kStack) { // a tree deeper than the private stack: never drop a subtree, hand the ray on
                complete = false;
                break;
            }
            stack[sp++] = first0 ? ch.y : ch.x;
            node = first0 ? ch.x : ch.y;
        } else if (h0) {
            node = ch.x;
        } else if (h1) {
            node = ch.y;
        } else {
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    return complete;
}

// ---- one crossing ----------------------------------------------------------------------------------

// One crossing of the mesh: the ray enters through boundary face `leaf` at depth z_in and goes from
// tet to tet until it leaves through another boundary face; returns that depth. tau, inten and steps
// are updated in place. kAffine: the caller does not know the intensity the ray arrives with (the
// grazing-ray kernel walks the crossings of one ray in parallel); inten then starts at 0 and gain at
// 1, and the crossing maps an arriving I to inten + gain * I (the recurrence of line.cpp:206-225 is
// affine in I: s - (s - I) e = s (1 - e) + e I).
template <bool kWide, int kPipe, bool kAffine>
C5_HD double crossing_f64(const WalkParams& P, double px, double py, int leaf, double z_in, double& tau,
                          double& inten, double& gain, uint32_t& steps, uint32_t& error) {
#ifdef __CUDA_ARCH__
    const int4 f = __ldg(reinterpret_cast<const int4*>(P.bfaces + leaf));
    int id = __ldg(&P.bfaces[leaf].apex);
#else
    const BFace& bf = P.bfaces[leaf];
    const int4 f = make_int4(bf.a, bf.b, bf.c, bf.tet);
    int id = bf.apex;
#endif
    // entry face (a, c, b) of the stored winding is counter-clockwise in projection
    int ia = f.x, ib = f.z, ic = f.y;
    double ax, ay, az, bx, by, bz, cx, cy, cz;
    load_vtx(P.vrot, ia, ax, ay, az);
    load_vtx(P.vrot, ib, bx, by, bz);
    load_vtx(P.vrot, ic, cx, cy, cz);
    ax -= px; ay -= py;
    bx -= px; by -= py;
    cx -= px; cy -= py;
    int t = f.w;
    double z_cur = z_in;
    // kPipe == 2 (software-pipelined loop): the cell and the vertex of step k+1 are requested as soon as
    // step k knows its exit face, BEFORE its divide and exp — the loads' latency overlaps that math.
    // Costs the registers of one cell + one vertex in flight next to the math's temporaries.
    CellData c;
    double dx, dy, dz;
    if (kPipe == 2) {
        c = load_cell<kWide>(P.cells, t);
        load_vtx(P.vrot, id, dx, dy, dz);
    }
    while (true) {
        if (steps >= static_cast<uint32_t>(P.max_steps)) {
            error = 1;
            break;
        }
        // id = the vertex of tet t that is not on the entry face. It is known BEFORE t's cell is
        // read (Cell::apex of the previous tet, BFace::apex at entry), so the cell load and the
        // vertex load of a step are independent and overlap: one memory latency per step, not two.
        if (kPipe != 2) {
            c = load_cell<kWide>(P.cells, t);
            load_vtx(P.vrot, id, dx, dy, dz);
        }
        dx -= px;
        dy -= py;
        const double sa = orient2(dx, dy, ax, ay);
        const double sb = orient2(dx, dy, bx, by);
        const double sc = orient2(dx, dy, cx, cy);
        // which face the ray leaves through, hence the next tet and its new vertex
        const bool drop_c = sa >= 0 && sb < 0;             // through (d, a, b)
        const bool drop_a = !drop_c && sb >= 0 && sc < 0;  // through (d, b, c)
        const int dropped = drop_c ? ic : drop_a ? ia : ib; // else through (d, c, a)
        const bool k0 = c.v.x == dropped, k1 = c.v.y == dropped, k2 = c.v.z == dropped;
        const int t_next = k0 ? c.nbr.x : k1 ? c.nbr.y : k2 ? c.nbr.z : c.nbr.w;
        const int id_next = k0 ? c.apex.x : k1 ? c.apex.y : k2 ? c.apex.z : c.apex.w;
        if (kPipe == 1 && t_next >= 0) {
            prefetch_l1(P.cells + t_next);
            prefetch_l1(reinterpret_cast<const char*>(P.cells + t_next) + 32);
            prefetch_l1(P.vrot + id_next);
        }
        // Barycentric weights of the pixel in the exit face (weight of a vertex = orient2 of the
        // other two, in cyclic order). Two of them are s values; the third belongs to d and is the
        // orient2 of the two vertices that stay — recomputed here (3 flops) rather than carried
        // from step to step (6 registers): same operands, same operations, so the same bits.
        double wa, wb, wc;
        if (drop_c) { // leaves through (d, a, b): c is replaced by d
            ic = id; cx = dx; cy = dy; cz = dz;
            wa = -sb;
            wb = sa;
            wc = orient2(ax, ay, bx, by);
        } else if (drop_a) { // through (d, b, c): a is replaced
            ia = id; ax = dx; ay = dy; az = dz;
            wb = -sc;
            wc = sb;
            wa = orient2(bx, by, cx, cy);
        } else { // through (d, c, a): b is replaced
            ib = id; bx = dx; by = dy; bz = dz;
            wc = -sa;
            wa = sc;
            wb = orient2(cx, cy, ax, ay);
        }
        const double alpha_t = c.alpha, s_t = c.s;
        if (kPipe == 2 && t_next >= 0) { // the next step's operands, on their way while this step finishes
            c = load_cell<kWide>(P.cells, t_next);
            load_vtx(P.vrot, id_next, dx, dy, dz);
        }
        const double wsum = wa + wb + wc;
        const double z_exit = (wsum != 0.0) ? (wa * az + wb * bz + wc * cz) / wsum : z_cur;
        const double dzv = fabs(z_exit - z_cur);
        // tau: line.cpp:176-193 (alpha not clamped)
        tau += dzv * alpha_t;
        // I: line.cpp:206-225 with s = Q / a^:  (Q - (Q - a^ I) e) / a^  ==  s - (s - I) e
        double a_c = alpha_t;
        if (a_c > P.alpha_limit) a_c = P.alpha_limit;
        if (!(a_c < DBL_EPSILON)) {
            const double e = exp(-a_c * dzv);
            inten = s_t - (s_t - inten) * e;
            if (kAffine) gain *= e;
        }
        steps++;
        z_cur = z_exit;
        t = t_next;
        id = id_next;
        if (t < 0) break;
    }
    return z_cur;
}

#ifdef C5_EXPERIMENTS
// Experimental variant "rec" (C5_WALK_VARIANT=rec, libc5gpu_exp.so only; measured: 3 % faster on the C3
// README view, 6 % slower on an oblique one, profiles/r02_exp_rec_vs_cells.jsonl — not adopted): the same crossing on step records
// (c5_types.h) — one 256-bit load for the tet instead of two, so two L1 data-pipe wavefronts per
// lane and step instead of three. The exit slot is the rank of the dropped vertex's id among the
// three entry-face ids. alpha and s arrive cut to 48 bits (7e-12 relative).
C5_HD double crossing_rec_f64(const WalkParams& P, double px, double py, int leaf, double z_in, double& tau,
                              double& inten, uint32_t& steps, uint32_t& error) {
#ifdef __CUDA_ARCH__
    const int4 f = __ldg(reinterpret_cast<const int4*>(P.bfaces + leaf));
    const int2 ae = __ldg(reinterpret_cast<const int2*>(&P.bfaces[leaf].apex)); // apex, entry-face index
    int id = ae.x;
    uint32_t face = 4u * static_cast<uint32_t>(f.w) + static_cast<uint32_t>(ae.y);
#else
    const BFace& bf = P.bfaces[leaf];
    const int4 f = make_int4(bf.a, bf.b, bf.c, bf.tet);
    int id = bf.apex;
    uint32_t face = 4u * static_cast<uint32_t>(bf.tet) + static_cast<uint32_t>(bf.pad[0]);
#endif
    int ia = f.x, ib = f.z, ic = f.y;
    double ax, ay, az, bx, by, bz, cx, cy, cz;
    load_vtx(P.vrot, ia, ax, ay, az);
    load_vtx(P.vrot, ib, bx, by, bz);
    load_vtx(P.vrot, ic, cx, cy, cz);
    ax -= px; ay -= py;
    bx -= px; by -= py;
    cx -= px; cy -= py;
    double z_cur = z_in;
    while (true) {
        if (steps >= static_cast<uint32_t>(P.max_steps)) {
            error = 1;
            break;
        }
        const StepRec r = load_rec(P.recs, face);
        double dx, dy, dz;
        load_vtx(P.vrot, id, dx, dy, dz);
        dx -= px;
        dy -= py;
        const double sa = orient2(dx, dy, ax, ay);
        const double sb = orient2(dx, dy, bx, by);
        const double sc = orient2(dx, dy, cx, cy);
        const bool drop_c = sa >= 0 && sb < 0;
        const bool drop_a = !drop_c && sb >= 0 && sc < 0;
        const int dropped = drop_c ? ic : drop_a ? ia : ib;
        const int slot = (ia < dropped ? 1 : 0) + (ib < dropped ? 1 : 0) + (ic < dropped ? 1 : 0);
        const uint64_t w = slot == 0 ? r.w[0] : slot == 1 ? r.w[1] : r.w[2];
        const uint32_t face_next = static_cast<uint32_t>(w) & kRecNoFace;
        const int id_next = static_cast<int>((w >> 28) & (kRecVtxLimit - 1));
        double wa, wb, wc;
        if (drop_c) {
            ic = id; cx = dx; cy = dy; cz = dz;
            wa = -sb;
            wb = sa;
            wc = orient2(ax, ay, bx, by);
        } else if (drop_a) {
            ia = id; ax = dx; ay = dy; az = dz;
            wb = -sc;
            wc = sb;
            wa = orient2(bx, by, cx, cy);
        } else {
            ib = id; bx = dx; by = dy; bz = dz;
            wc = -sa;
            wa = sc;
            wb = orient2(cx, cy, ax, ay);
        }
        const double wsum = wa + wb + wc;
        const double z_exit = (wsum != 0.0) ? (wa * az + wb * bz + wc * cz) / wsum : z_cur;
        const double dzv = fabs(z_exit - z_cur);
        const double alpha = rec_alpha(r);
        tau += dzv * alpha;
        double a_c = alpha;
        if (a_c > P.alpha_limit) a_c = P.alpha_limit;
        if (!(a_c < DBL_EPSILON)) {
            const double sf = rec_s(r);
            inten = sf - (sf - inten) * exp(-a_c * dzv);
        }
        steps++;
        z_cur = z_exit;
        face = face_next;
        id = id_next;
        if (face == kRecNoFace) break;
    }
    return z_cur;
}
#endif

// ---- FP32 variant -----------------------------------------------------------------------------------
// Same crossing with the per-step geometry in single precision: FP32 orientation tests (still exactly
// antisymmetric: two rounded products, one rounded difference), FP32 divide and expf — about half
// the issue slots and 56 instead of 72 registers. What stays in double: the ENTRY search (so the
// hit/miss set is the FP64 one, bit for bit), the vertex fetch and its subtraction of the pixel
// position / entry depth (rounding a coordinate ~1 to float would cost 1e-7, i.e. 1e-5 of a tet;
// rounding the DIFFERENCE costs 6e-8 of the tet size), and the accumulators tau and I. The walk
// is not bandwidth-bound, so keeping 32-byte vertices costs nothing measurable.
// Tolerance vs the reference: 1e-4 relative (tests/parity.py).
C5_HD float fmul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
C5_HD float fsub_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
C5_HD float orient2f(float ux, float uy, float vx, float vy) {
    return fsub_rn(fmul_rn(ux, vy), fmul_rn(uy, vx));
}

template <bool kWide, int kPipe, bool kAffine>
C5_HD double crossing_f32(const WalkParams& P, double px, double py, int leaf, double z_in, double& tau,
                          double& inten, double& gain, uint32_t& steps, uint32_t& error) {
    const float limit = static_cast<float>(P.alpha_limit);
    const double z0 = z_in; // depths are kept relative to the entry point of the crossing
#ifdef __CUDA_ARCH__
    const int4 f = __ldg(reinterpret_cast<const int4*>(P.bfaces + leaf));
    int id = __ldg(&P.bfaces[leaf].apex);
#else
    const BFace& bf = P.bfaces[leaf];
    const int4 f = make_int4(bf.a, bf.b, bf.c, bf.tet);
    int id = bf.apex;
#endif
    int ia = f.x, ib = f.z, ic = f.y;
    float ax, ay, az, bx, by, bz, cx, cy, cz;
    {
        double x, y, z;
        load_vtx(P.vrot, ia, x, y, z);
        ax = static_cast<float>(x - px); ay = static_cast<float>(y - py); az = static_cast<float>(z - z0);
        load_vtx(P.vrot, ib, x, y, z);
        bx = static_cast<float>(x - px); by = static_cast<float>(y - py); bz = static_cast<float>(z - z0);
        load_vtx(P.vrot, ic, x, y, z);
        cx = static_cast<float>(x - px); cy = static_cast<float>(y - py); cz = static_cast<float>(z - z0);
    }
    float wa = orient2f(bx, by, cx, cy);
    float wb = orient2f(cx, cy, ax, ay);
    float wc = orient2f(ax, ay, bx, by);
    int t = f.w;
    float z_cur = 0.f;
    while (true) {
        if (steps >= static_cast<uint32_t>(P.max_steps)) {
            error = 1;
            break;
        }
        const CellData c = load_cell<kWide>(P.cells, t);
        float dx, dy, dz;
        {
            double x, y, z;
            load_vtx(P.vrot, id, x, y, z);
            dx = static_cast<float>(x - px);
            dy = static_cast<float>(y - py);
            dz = static_cast<float>(z - z0);
        }
        const float sa = orient2f(dx, dy, ax, ay);
        const float sb = orient2f(dx, dy, bx, by);
        const float sc = orient2f(dx, dy, cx, cy);
        const bool drop_c = sa >= 0 && sb < 0;
        const bool drop_a = !drop_c && sb >= 0 && sc < 0;
        const int dropped = drop_c ? ic : drop_a ? ia : ib;
        const bool k0 = c.v.x == dropped, k1 = c.v.y == dropped, k2 = c.v.z == dropped;
        const int t_next = k0 ? c.nbr.x : k1 ? c.nbr.y : k2 ? c.nbr.z : c.nbr.w;
        const int id_next = k0 ? c.apex.x : k1 ? c.apex.y : k2 ? c.apex.z : c.apex.w;
        if (kPipe == 1 && t_next >= 0) {
            prefetch_l1(P.cells + t_next);
            prefetch_l1(reinterpret_cast<const char*>(P.cells + t_next) + 32);
            prefetch_l1(P.vrot + id_next);
        }
        if (drop_c) {
            ic = id; cx = dx; cy = dy; cz = dz;
            wa = -sb;
            wb = sa;
        } else if (drop_a) {
            ia = id; ax = dx; ay = dy; az = dz;
            wb = -sc;
            wc = sb;
        } else {
            ib = id; bx = dx; by = dy; bz = dz;
            wc = -sa;
            wa = sc;
        }
        const float wsum = wa + wb + wc;
        const float z_exit = (wsum != 0.0f) ? (wa * az + wb * bz + wc * cz) / wsum : z_cur;
        const float dzv = fabsf(z_exit - z_cur);
        tau += static_cast<double>(dzv) * c.alpha;
        float a_c = static_cast<float>(c.alpha);
        if (a_c > limit) a_c = limit;
        const double a_d = c.alpha > P.alpha_limit ? P.alpha_limit : c.alpha;
        if (!(a_d < DBL_EPSILON)) {
            const double e = static_cast<double>(expf(-a_c * dzv));
            inten = c.s - (c.s - inten) * e;
            if (kAffine) gain *= e;
        }
        steps++;
        z_cur = z_exit;
        t = t_next;
        id = id_next;
        if (t < 0) break;
    }
    // the next crossing must lie above this one's exit (and strictly above its entry)
    const double z_exit_abs = z0 + static_cast<double>(z_cur);
    return z_exit_abs > z0 ? z_exit_abs : z0;
}

template <bool kF32, bool kWide, int kPipe, bool kAffine, bool kRec = false>
C5_HD double crossing(const WalkParams& P, double px, double py, int leaf, double z_in, double& tau, double& inten,
                      double& gain, uint32_t& steps, uint32_t& error) {
#ifdef C5_EXPERIMENTS
    if (kRec) return crossing_rec_f64(P, px, py, leaf, z_in, tau, inten, steps, error);
#endif
    return kF32 ? crossing_f32<kWide, kPipe, kAffine>(P, px, py, leaf, z_in, tau, inten, gain, steps, error)
                : crossing_f64<kWide, kPipe, kAffine>(P, px, py, leaf, z_in, tau, inten, gain, steps, error);
}

#ifdef C5_EXPERIMENTS
// Step record of (tet, entry face) f = 4 t + e: the three faces the ray can leave through, ordered by
// the global id of the entry-face vertex each one drops.
C5_HD StepRec make_step_record(const Cell* cells, int64_t f) {
    const Cell& c = cells[f >> 2];
    const int e = static_cast<int>(f & 3);
    int k[3], n = 0;
    for (int j = 0; j < 4; j++) {
        if (j != e) k[n++] = j;
    }
    // sort the three local indices by global vertex id
    if (c.v[k[1]] < c.v[k[0]]) { const int t = k[0]; k[0] = k[1]; k[1] = t; }
    if (c.v[k[2]] < c.v[k[1]]) { const int t = k[1]; k[1] = k[2]; k[2] = t; }
    if (c.v[k[1]] < c.v[k[0]]) { const int t = k[0]; k[0] = k[1]; k[1] = t; }
    uint32_t next_face[3], next_apex[3];
    for (int i = 0; i < 3; i++) {
        const int nb = c.nbr[k[i]]; // across the face that drops vertex k[i]
        if (nb < 0) {
            next_face[i] = kRecNoFace;
            next_apex[i] = 0;
        } else {
            const Cell& o = cells[nb];
            const int apex = c.apex[k[i]];
            const int e2 = o.v[0] == apex ? 0 : o.v[1] == apex ? 1 : o.v[2] == apex ? 2 : 3;
            next_face[i] = 4u * static_cast<uint32_t>(nb) + static_cast<uint32_t>(e2);
            next_apex[i] = static_cast<uint32_t>(apex);
        }
    }
    return pack_rec(next_face, next_apex, c.alpha, c.s);
}
#endif

// ---- one ray (pixel kernel) ------------------------------------------------------------------------

struct RayResult {
    double tau, inten;
    uint32_t steps;
    uint32_t error;
    uint32_t deferred; // handed to the grazing-ray kernel: the pixel is stored there
};

// A ray's FIRST query is a nearest-hit search (cap 1: most rays enter once and the search stays as
// cheap as it can be). After the first exit ONE more query asks for the next kInline entries above:
// for a convex mesh it finds nothing after visiting a handful of nodes; a cavity gives one or two,
// which are walked here. A full list, or a search that runs out of budget, means a grazing ray: its
// state goes to the deferred queue.
//
// The four phases are written out and separated by __syncwarp over the lanes that trace (`tracing`,
// a ballot taken where the warp is still whole): the searches and the walks must each run with the
// warp CONVERGED — a search is hundreds of dependent loads and a walk hundreds of steps; executed
// lane group by lane group they cost a multiple. Left to the compiler's reconvergence analysis this
// was fragile: an innocent change to the search loop (a second exit) made it stop reconverging
// after the search and the whole kernel ran 1.75 x slower (8.3 vs 4.7 ms on C3).
C5_HD void sync_lanes(unsigned lanes) {
#ifdef __CUDA_ARCH__
    __syncwarp(lanes);
#else
    (void)lanes;
#endif
}

template <bool kF32, bool kWide, int kPipe, bool kRec = false>
C5_HD RayResult trace_ray(const WalkParams& P, const BvhNode* top, double px, double py, uint32_t pixel,
                          unsigned tracing) {
    RayResult r;
    r.tau = 0.0;
    r.inten = 0.0;
    r.steps = 0;
    r.error = 0;
    r.deferred = 0;
    double z_after = -INFINITY;
    double gain_unused = 1.0;
    EntryList<kInline> L;

    // A: nearest entry
    bool defer = !bvh_collect_entries(P, top, px, py, z_after, L, 1, P.query_budget);
    sync_lanes(tracing);
    // B: first crossing
    const bool entered = !defer && L.n > 0;
    if (entered) {
        z_after = crossing<kF32, kWide, kPipe, false, kRec>(P, px, py, L.leaf[0], L.z[0], r.tau, r.inten, gain_unused,
                                                            r.steps, r.error);
    }
    sync_lanes(tracing);
    // C: anything above the exit?
    L.n = 0;
    if (entered && !r.error) {
        const bool searched = bvh_collect_entries(P, top, px, py, z_after, L, kInline, P.query_budget);
        defer = !searched || L.n == kInline; // too many boxes in the way, or too many entries
    }
    sync_lanes(tracing);
    // D: the few re-entries of a cavity or a dent
    for (int e = 0; e < kInline - 1; e++) {
        // an entry at or below the ray's position is already behind it (two boundary faces sharing
        // the edge the ray passes through report the same depth)
        if (!defer && e < L.n && L.z[e] > z_after && !r.error) {
            z_after = crossing<kF32, kWide, kPipe, false, kRec>(P, px, py, L.leaf[e], L.z[e], r.tau, r.inten, gain_unused,
                                                                r.steps, r.error);
        }
        sync_lanes(tracing);
    }
    if (defer) {
#ifdef __CUDA_ARCH__
        const unsigned long long slot = atomicAdd(&P.counters[kDeferred], 1ull);
#else
        const unsigned long long slot = P.counters[kDeferred]++;
#endif
        // read by the grazing-ray kernel, which is launched after this one on the same stream
        DeferredRay& q = P.queue[slot];
        q.tau = r.tau;
        q.inten = r.inten;
        q.z_after = z_after;
        q.pixel = pixel;
        q.steps = r.steps;
        r.deferred = 1;
    }
    return r;
}

// ---- grazing rays ----------------------------------------------------------------------------------
// A deferred ray has many crossings (~110 on the C3 README view, each a few tets long). Two facts
// make it parallel: (1) its crossings are independent once their entry faces are known — tau is a
// sum and I an affine recurrence, so crossing k yields (tau_k, A_k, B_k) and the ray's values are
// their composition in z order; (2) collecting the entry faces is a BVH traversal with a wide
// frontier. So one WARP takes one ray: the lanes expand 32 BVH nodes per round from a shared stack,
// rank-sort the entries found, walk 32 crossings at a time and fold the results in order.

struct RayAcc {
    double tau, inten, z_after;
    uint32_t steps, error;
};

// Folds crossing (z_in -> z_out) into the ray unless the ray is already past its entry.
C5_HD void compose_crossing(RayAcc& a, double z_in, double z_out, double tau_k, double inten_k, double gain_k,
                            uint32_t steps_k, uint32_t error_k) {
    if (!(z_in > a.z_after)) return;
    a.tau += tau_k;
    a.inten = inten_k + gain_k * a.inten;
    a.steps += steps_k;
    a.error |= error_k;
    a.z_after = z_out;
}

// Serial form (one thread does everything): the host-loop build, and the reference the warp
// version is tested against. Same crossing function, same composition.
template <bool kF32>
C5_HD RayAcc graze_ray_serial(const WalkParams& P, const DeferredRay& q, double px, double py) {
    RayAcc a;
    a.tau = q.tau;
    a.inten = q.inten;
    a.z_after = q.z_after;
    a.steps = q.steps;
    a.error = 0;
    EntryList<kSerialList> L;
    int crossings = 0;
    while (!a.error) {
        if (!bvh_collect_entries(P, nullptr, px, py, a.z_after, L, P.serial_cap, 0x7FFFFFFF)) a.error = 1; // stack overflow
        for (int e = 0; e < L.n; e++) {
            if (!(L.z[e] > a.z_after)) continue;
            double tau_k = 0.0, inten_k = 0.0, gain_k = 1.0;
            uint32_t steps_k = 0, error_k = 0;
            const double z_out = crossing<kF32, false, 0, true>(P, px, py, L.leaf[e], L.z[e], tau_k, inten_k, gain_k,
                                                                steps_k, error_k);
            compose_crossing(a, L.z[e], z_out, tau_k, inten_k, gain_k, steps_k, error_k);
            if (++crossings > 65536) a.error = 1;
        }
        if (L.n < P.serial_cap) break;
    }
    return a;
}

C5_HD double __longlong_as_double_hd(long long bits) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(bits);
#else
    double v;
    memcpy(&v, &bits, sizeof(v));
    return v;
#endif
}

C5_HD void store_pixel(const WalkParams& P, int i, int j, double tau, double inten, uint32_t steps) {
    const size_t o = static_cast<size_t>(j - P.row_begin) * P.res_x + i;
    if (P.round_float) { // plane.cpp:165-166 then object2d.cpp:19-20
        tau = static_cast<double>(static_cast<float>(tau));
        inten = static_cast<double>(static_cast<float>(inten));
    }
#ifdef __CUDA_ARCH__
    reinterpret_cast<double2*>(P.out)[o] = make_double2(tau, inten);
#else
    P.out[2 * o] = tau;
    P.out[2 * o + 1] = inten;
#endif
    if (P.steps) P.steps[o] = steps;
}

// Pixels of the band outside the rectangle the walk covers: background (or solid).
C5_HD bool background_pixel(const WalkParams& P, int i, int j) {
    const bool solid = P.mask && P.mask[static_cast<size_t>(j) * P.res_x + i];
    const double v = solid ? __longlong_as_double_hd(0x7FF8000000000000ll) : 0.0; // quiet NaN (config.hpp:26-27)
    store_pixel(P, i, j, v, v, 0);
    return solid;
}

// ---- warp-cooperative form ---------------------------------------------------------------------------
constexpr int kGrazeWarps = 2;    // warps (= rays in flight) per block: small blocks fit into the gaps the pixel kernels leave
constexpr int kGrazeBlocksPerSm = 12; // grid of the grazing-ray kernel (persistent warps drawing tickets). The kernel is
                                      // latency-bound: C3 whole view alone 0.80 ms with 4 blocks per SM, 0.57 with 8, 0.50 with 12,
                                      // 0.51 with 16 (profiles/r02_exp_graze_blocks.jsonl); blocks that find the queue empty leave at once
constexpr int kGrazeList = 256;   // entries per collection; more are fetched by another round
constexpr int kGrazeStack = 512;  // shared traversal stack per warp ...
constexpr int kGrazeSlack = 128;  // ... plus room for one wide round and a depth-first tail
constexpr int kGrazePerLane = kGrazeList / 32;

struct GrazeSmem {
    double z[kGrazeList];
    int leaf[kGrazeList];
    int stack[kGrazeStack + kGrazeSlack];
};

// Sorts S.z / S.leaf [0, n) by (z, leaf): every lane ranks its own <= 8 entries against all n
// (broadcast reads), then scatters them. n <= 256, so this is a few thousand compares per lane.
__device__ __forceinline__ void graze_sort(GrazeSmem& S, int n, int lane) {
    double zr[kGrazePerLane];
    int lr[kGrazePerLane], rank[kGrazePerLane];
#pragma unroll
    for (int q = 0; q < kGrazePerLane; q++) {
        const int k = q * 32 + lane;
        zr[q] = k < n ? S.z[k] : INFINITY;
        lr[q] = k < n ? S.leaf[k] : 0x7FFFFFFF;
        rank[q] = 0;
    }
    for (int m = 0; m < n; m++) {
        const double zm = S.z[m];
        const int lm = S.leaf[m];
#pragma unroll
        for (int q = 0; q < kGrazePerLane; q++) rank[q] += entry_before(zm, lm, zr[q], lr[q]) ? 1 : 0;
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < kGrazePerLane; q++) {
        if (q * 32 + lane < n) {
            S.z[rank[q]] = zr[q];
            S.leaf[rank[q]] = lr[q];
        }
    }
    __syncwarp();
}

// All entry faces above z_after under (px, py), collected by the whole warp into S (sorted).
// Returns their number; `truncated` is set if the list overflowed and only the lowest part was kept
// (everything up to the last kept entry is present: the caller walks it and asks again).
__device__ __forceinline__ int graze_collect(const WalkParams& P, GrazeSmem& S, double px, double py, double z_after,
                                             int lane, bool& truncated) {
    const unsigned full = 0xFFFFFFFFu;
    const unsigned lt = (1u << lane) - 1u;
    const float fx_lo = f_round_down(px), fx_hi = f_round_up(px);
    const float fy_lo = f_round_down(py), fy_hi = f_round_up(py);
    int sp = 1, n = 0;
    double z_cap = INFINITY;
    truncated = false;
    if (lane == 0) S.stack[0] = 0;
    __syncwarp();
    while (sp > 0) {
        if (n > P.graze_cap - 64) { // a round appends at most 64: keep the lowest half, forget the rest
            graze_sort(S, n, lane);
            n = P.graze_cap / 2;
            z_cap = S.z[n - 1];
            truncated = true;
            __syncwarp();
        }
        // 32 nodes per round while the stack has room for their 64 children; else depth first
        const int take = sp <= kGrazeStack - 64 ? (sp < 32 ? sp : 32) : 1;
        const int node = lane < take ? S.stack[sp - 1 - lane] : -1;
        sp -= take;
        __syncwarp();
        bool h0 = false, h1 = false;
        int2 ch = make_int2(0, 0);
        if (node >= 0) {
            const BvhNode* nd = P.nodes + node;
            const float4 bx = *reinterpret_cast<const float4*>(nd->xlo);
            const float4 by = *reinterpret_cast<const float4*>(nd->ylo);
            const float4 bz = *reinterpret_cast<const float4*>(nd->zlo);
            ch = *reinterpret_cast<const int2*>(nd->child);
            h0 = fx_hi >= bx.x && fx_lo <= bx.z && fy_hi >= by.x && fy_lo <= by.z &&
                 static_cast<double>(bz.z) > z_after && static_cast<double>(bz.x) <= z_cap;
            h1 = fx_hi >= bx.y && fx_lo <= bx.w && fy_hi >= by.y && fy_lo <= by.w &&
                 static_cast<double>(bz.w) > z_after && static_cast<double>(bz.y) <= z_cap;
        }
        double z0 = 0.0, z1 = 0.0;
        const bool e0 = h0 && ch.x < 0 && entry_depth(P, ~ch.x, px, py, z_after, z0) && z0 <= z_cap;
        const bool e1 = h1 && ch.y < 0 && entry_depth(P, ~ch.y, px, py, z_after, z1) && z1 <= z_cap;
        unsigned m = __ballot_sync(full, e0);
        if (e0) {
            const int pos = n + __popc(m & lt);
            S.z[pos] = z0;
            S.leaf[pos] = ~ch.x;
        }
        n += __popc(m);
        m = __ballot_sync(full, e1);
        if (e1) {
            const int pos = n + __popc(m & lt);
            S.z[pos] = z1;
            S.leaf[pos] = ~ch.y;
        }
        n += __popc(m);
        const bool i0 = h0 && ch.x >= 0, i1 = h1 && ch.y >= 0;
        m = __ballot_sync(full, i0);
        if (i0) S.stack[sp + __popc(m & lt)] = ch.x;
        sp += __popc(m);
        m = __ballot_sync(full, i1);
        if (i1) S.stack[sp + __popc(m & lt)] = ch.y;
        sp += __popc(m);
        __syncwarp();
    }
    graze_sort(S, n, lane);
    return n;
}

template <bool kF32>
__device__ __forceinline__ void graze_block(const WalkParams& P) {
    __shared__ GrazeSmem smem[kGrazeWarps];
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    GrazeSmem& S = smem[threadIdx.x >> 5];
    // Launched after the pixel kernel on the same stream, so the queue and its length are final.
    // The length lives on the device: a fixed grid of persistent warps draws tickets until they run
    // out (an empty queue costs one short launch).
    const unsigned long long n_rays = P.counters[kDeferred];
    while (true) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(&P.counters[kTicket], 1ull);
        ticket = __shfl_sync(full, ticket, 0);
        if (ticket >= n_rays) break;
        const DeferredRay* qp = P.queue + ticket;
        const uint32_t q_steps = qp->steps;
        const uint32_t q_pixel = qp->pixel;
        const int i = static_cast<int>(q_pixel % static_cast<uint32_t>(P.res_x));
        const int j = static_cast<int>(q_pixel / static_cast<uint32_t>(P.res_x));
        const double px = P.xs[i], py = P.ys[j];
        RayAcc a;
        a.tau = qp->tau;
        a.inten = qp->inten;
        a.z_after = qp->z_after;
        a.steps = q_steps;
        a.error = 0;
        int crossings = 0;
        bool truncated = true;
        while (truncated && !a.error) {
            const int n = graze_collect(P, S, px, py, a.z_after, lane, truncated);
            for (int base = 0; base < n; base += 32) {
                const int k = base + lane;
                const bool have = k < n;
                const double z_in = have ? S.z[k] : INFINITY;
                double z_out = z_in, tau_k = 0.0, inten_k = 0.0, gain_k = 1.0;
                uint32_t steps_k = 0, error_k = 0;
                if (have && z_in > a.z_after) { // a is the same in every lane
                    z_out = crossing<kF32, true, 0, true>(P, px, py, S.leaf[k], z_in, tau_k, inten_k, gain_k, steps_k,
                                                          error_k);
                }
                const int cnt = n - base < 32 ? n - base : 32;
                for (int l = 0; l < cnt; l++) {
                    compose_crossing(a, __shfl_sync(full, z_in, l), __shfl_sync(full, z_out, l),
                                     __shfl_sync(full, tau_k, l), __shfl_sync(full, inten_k, l),
                                     __shfl_sync(full, gain_k, l), __shfl_sync(full, steps_k, l),
                                     __shfl_sync(full, error_k, l));
                }
            }
            crossings += n;
            if (crossings > 65536) a.error = 1;
            __syncwarp();
        }
        if (lane == 0) {
            store_pixel(P, i, j, a.tau, a.inten, a.steps);
            const unsigned long long more = a.steps - q_steps;
            if (more) {
                atomicAdd(&P.counters[kSteps], more);
                atomicAdd(&P.row_cost[j], more);
                if (q_steps == 0) atomicAdd(&P.counters[kHitPixels], 1ull); // deferred before its first step
            }
            if (a.error) atomicAdd(&P.counters[kWalkErrors], 1ull);
        }
    }
}

// ---- kernel ----------------------------------------------------------------------------------------

// 3-bit Morton decode: bits 0,2,4 -> x, bits 1,3,5 -> y
__device__ __forceinline__ int compact3(int v) {
    return (v & 1) | ((v >> 1) & 2) | ((v >> 2) & 4);
}

template <bool kF32, bool kWide, int kPipe, int kWarpsX = 2, int kWarpsY = 2, bool kRec = false>
__device__ __forceinline__ void walk_block_body(const WalkParams& P) {
    constexpr int kTx = 8 * kWarpsX, kTy = 4 * kWarpsY;
    [[maybe_unused]] constexpr int kThreads = 32 * kWarpsX * kWarpsY;
#ifdef C5_EXPERIMENTS
    extern __shared__ __align__(64) unsigned char smem_raw[];
    BvhNode* top = reinterpret_cast<BvhNode*>(smem_raw);
#else
    // (staging the top BVH levels in shared memory was measured slower than letting L1 serve them:
    // C3 6.38 ms with 0 nodes, 6.54 with 63, 6.88 with 255; profiles/r01_exp_c3_variants_b.jsonl)
    const BvhNode* top = nullptr;
#endif

    // screen-space order: macro tiles of 8 x 8 block tiles in row-major order, Morton inside, so that
    // concurrently resident blocks cover a compact patch of the image and share tets in L2
    const int b = blockIdx.x;
    const int macro = b >> 6, r = b & 63;
    const int tile_x = (macro % P.n_macro_x) * 8 + compact3(r);
    const int tile_y = (macro / P.n_macro_x) * 8 + compact3(r >> 1);
    if (tile_x >= P.n_tiles_x || tile_y >= P.n_tiles_y) return;

    // Tiles that cannot see the mesh (outside the root's two boxes) skip the staging and the rays.
    const int i_lo = P.i0 + tile_x * kTx, j_lo = P.j0 + tile_y * kTy;
    const int i_hi = min(i_lo + kTx, P.i1) - 1, j_hi = min(j_lo + kTy, P.j1) - 1;
    bool tile_sees_mesh;
    {
        const BvhNode* root = P.nodes;
        const float x0 = f_round_down(P.xs[i_lo]), x1 = f_round_up(P.xs[i_hi]);
        const float y0 = f_round_down(P.ys[j_lo]), y1 = f_round_up(P.ys[j_hi]);
        const bool s0 = x1 >= root->xlo[0] && x0 <= root->xhi[0] && y1 >= root->ylo[0] && y0 <= root->yhi[0];
        const bool s1 = x1 >= root->xlo[1] && x0 <= root->xhi[1] && y1 >= root->ylo[1] && y0 <= root->yhi[1];
        tile_sees_mesh = s0 || s1;
    }
#ifdef C5_EXPERIMENTS
    if (tile_sees_mesh && P.top_nodes > 0) {
        const int4* src = reinterpret_cast<const int4*>(P.nodes);
        int4* dst = reinterpret_cast<int4*>(top);
        for (int k = threadIdx.x; k < P.top_nodes * 4; k += kThreads) dst[k] = src[k];
        __syncthreads();
    }
#endif

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = P.i0 + tile_x * kTx + (warp % kWarpsX) * 8 + (lane & 7);
    const int j = P.j0 + tile_y * kTy + (warp / kWarpsX) * 4 + (lane >> 3);
    const bool live = i < P.i1 && j < P.j1;

    RayResult res;
    res.tau = 0.0;
    res.inten = 0.0;
    res.steps = 0;
    res.error = 0;
    res.deferred = 0;
    bool solid = false;
    if (live) solid = P.mask && P.mask[static_cast<size_t>(j) * P.res_x + i];
    // the lanes that will trace a ray, agreed on while the warp is whole (trace_ray synchronises on it)
    const unsigned tracing = __ballot_sync(0xFFFFFFFFu, live && !solid && tile_sees_mesh);
    if (live) {
        if (solid) {
            const double nan = __longlong_as_double(0x7FF8000000000000ll); // quiet NaN (config.hpp:26-27)
            store_pixel(P, i, j, nan, nan, 0);
        } else {
            if (tile_sees_mesh) {
                res = trace_ray<kF32, kWide, kPipe, kRec>(P, top, P.xs[i], P.ys[j],
                                                    static_cast<uint32_t>(j) * static_cast<uint32_t>(P.res_x) + static_cast<uint32_t>(i),
                                                    tracing);
            }
            if (!res.deferred) store_pixel(P, i, j, res.tau, res.inten, res.steps);
        }
    }

    // statistics: one atomic per warp per counter, one per warp-row for the row costs
    const unsigned full = 0xFFFFFFFFu;
    unsigned row_steps = res.steps;
    row_steps += __shfl_xor_sync(full, row_steps, 1);
    row_steps += __shfl_xor_sync(full, row_steps, 2);
    row_steps += __shfl_xor_sync(full, row_steps, 4);
    if ((lane & 7) == 0 && live && row_steps) atomicAdd(&P.row_cost[j], static_cast<unsigned long long>(row_steps));
    unsigned warp_steps = row_steps;
    warp_steps += __shfl_xor_sync(full, warp_steps, 8);
    warp_steps += __shfl_xor_sync(full, warp_steps, 16);
    const unsigned hits = __popc(__ballot_sync(full, live && res.steps > 0));
    const unsigned solids = __popc(__ballot_sync(full, solid));
    const unsigned errors = __popc(__ballot_sync(full, res.error != 0));
    if (lane == 0) {
        if (warp_steps) atomicAdd(&P.counters[kSteps], static_cast<unsigned long long>(warp_steps));
        if (hits) atomicAdd(&P.counters[kHitPixels], static_cast<unsigned long long>(hits));
        if (solids) atomicAdd(&P.counters[kSolidPixels], static_cast<unsigned long long>(solids));
        if (errors) atomicAdd(&P.counters[kWalkErrors], static_cast<unsigned long long>(errors));
    }
}

#ifdef C5_EXPERIMENTS
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

template <bool kF32, bool kWide, int kPipe, int kWarpsX = 2, int kWarpsY = 2, bool kRec = false>
__device__ __forceinline__ void walk_block(const WalkParams& P) {
#ifdef C5_EXPERIMENTS   // block timeline (C5_TRACE_FILE)
    if (P.trace && threadIdx.x == 0) P.trace[4ull * blockIdx.x] = global_ns(); // stored at once: nothing stays live
#endif
    walk_block_body<kF32, kWide, kPipe, kWarpsX, kWarpsY, kRec>(P);
#ifdef C5_EXPERIMENTS
    if (P.trace) {
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            unsigned long long* rec = P.trace + 4ull * blockIdx.x;
            rec[1] = global_ns();
            rec[2] = sm;
            rec[3] = blockIdx.x;
        }
    }
#endif
}

} // namespace

// The product kernels. 7 blocks of 128 threads per SM = 72 registers: one more resident block than the
// 80 the compiler picks on its own, measured faster (5.41 vs 5.64 ms on C3) despite a few spilled words
// outside the step loop.
__global__ void __launch_bounds__(kBlock, 7) tet_walk_fp64(const WalkParams P) { walk_block<false, true, 0>(P); }
// optional single-precision step geometry (north-star item (d)); entry search and accumulators stay FP64
__global__ void __launch_bounds__(kBlock, 8) tet_walk_fp32(const WalkParams P) { walk_block<true, true, 0>(P); }

#ifdef C5_EXPERIMENTS
// Variants measured and NOT adopted (C5_WALK_VARIANT selects one in libc5gpu_exp.so; results under
// profiles/): 128-bit loads, L1 prefetch of the next step, 64-thread blocks, other register caps,
// step records.
__global__ void __launch_bounds__(kBlock) tet_walk_fp64_l128(const WalkParams P) { walk_block<false, false, 0>(P); }
__global__ void __launch_bounds__(kBlock) tet_walk_fp64_pf(const WalkParams P) { walk_block<false, true, 1>(P); }
__global__ void __launch_bounds__(64) tet_walk_fp64_b64(const WalkParams P) { walk_block<false, true, 0, 1, 2>(P); }
__global__ void __launch_bounds__(kBlock, 6) tet_walk_fp64_r80(const WalkParams P) { walk_block<false, true, 0>(P); }
__global__ void __launch_bounds__(kBlock, 8) tet_walk_fp64_r64(const WalkParams P) { walk_block<false, true, 0>(P); }
__global__ void __launch_bounds__(kBlock, 5) tet_walk_fp64_r96(const WalkParams P) { walk_block<false, true, 0>(P); }
__global__ void __launch_bounds__(kBlock, 7) tet_walk_fp64_rec(const WalkParams P) { walk_block<false, true, 0, 2, 2, true>(P); }
// software-pipelined step loop (kPipe == 2) at 5, 6 and 7 resident blocks per SM (96 / 80 / 72 registers)
__global__ void __launch_bounds__(kBlock, 5) tet_walk_fp64_sp5(const WalkParams P) { walk_block<false, true, 2>(P); }
__global__ void __launch_bounds__(kBlock, 6) tet_walk_fp64_sp6(const WalkParams P) { walk_block<false, true, 2>(P); }
__global__ void __launch_bounds__(kBlock, 7) tet_walk_fp64_sp7(const WalkParams P) { walk_block<false, true, 2>(P); }
// Builds the four step records of every tet from its Cell (after prepare_cells: s depends on --alpha_limit).
__global__ void __launch_bounds__(256) build_step_records(int64_t n_faces, const Cell* __restrict__ cells, StepRec* __restrict__ recs) {
    const int64_t f = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (f < n_faces) recs[f] = make_step_record(cells, f);
}
#endif

// The walk's grid only covers the pixel rectangle that can see the mesh (a one-wave row band whose
// grid is 60 % empty tiles fills the SMs unevenly: measured 1.60 M busy cycles on the fullest SM
// against 1.07 M on average, profiles/r01_walk_band_c3.csv). Everything else in the band is background
// or solid and is written by this streaming kernel.
__global__ void __launch_bounds__(256) fill_background(const WalkParams P) {
    const int n_rows = P.row_end - P.row_begin;
    const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    bool solid = false;
    if (k < static_cast<int64_t>(n_rows) * P.res_x) {
        const int j = P.row_begin + static_cast<int>(k / P.res_x), i = static_cast<int>(k % P.res_x);
        if (!(i >= P.i0 && i < P.i1 && j >= P.j0 && j < P.j1)) solid = background_pixel(P, i, j);
    }
    const unsigned solids = __popc(__ballot_sync(0xFFFFFFFFu, solid));
    if ((threadIdx.x & 31) == 0 && solids) atomicAdd(&P.counters[kSolidPixels], static_cast<unsigned long long>(solids));
}

// Grazing rays (deferred by the pixel kernel): persistent warps, one ray per warp at a time.
__global__ void __launch_bounds__(32 * kGrazeWarps) grazing_rays_fp64(const WalkParams P) { graze_block<false>(P); }
__global__ void __launch_bounds__(32 * kGrazeWarps) grazing_rays_fp32(const WalkParams P) { graze_block<true>(P); }

namespace {

void graze_on_host(const WalkParams& P, bool f32) {
    const unsigned long long n_rays = P.counters[kDeferred];
    for (unsigned long long k = 0; k < n_rays; k++) {
        const DeferredRay& q = P.queue[k];
        const int i = static_cast<int>(q.pixel % static_cast<uint32_t>(P.res_x));
        const int j = static_cast<int>(q.pixel / static_cast<uint32_t>(P.res_x));
        const RayAcc a = f32 ? graze_ray_serial<true>(P, q, P.xs[i], P.ys[j]) : graze_ray_serial<false>(P, q, P.xs[i], P.ys[j]);
        store_pixel(P, i, j, a.tau, a.inten, a.steps);
        P.counters[kSteps] += a.steps - q.steps;
        P.row_cost[j] += a.steps - q.steps;
        if (q.steps == 0 && a.steps > 0) P.counters[kHitPixels]++;
        if (a.error) P.counters[kWalkErrors]++;
    }
    P.counters[kTicket] = n_rays;
}

void background_on_host(const WalkParams& P) {
    for (int j = P.row_begin; j < P.row_end; j++) {
        for (int i = 0; i < P.res_x; i++) {
            if (i >= P.i0 && i < P.i1 && j >= P.j0 && j < P.j1) continue;
            if (background_pixel(P, i, j)) P.counters[kSolidPixels]++;
        }
    }
}

void walk_on_host(const WalkParams& P, bool f32) {
    for (int j = P.j0; j < P.j1; j++) {
        for (int i = P.i0; i < P.i1; i++) {
            if (P.mask && P.mask[static_cast<size_t>(j) * P.res_x + i]) {
                store_pixel(P, i, j, NAN, NAN, 0);
                P.counters[kSolidPixels]++;
                continue;
            }
            const uint32_t pixel = static_cast<uint32_t>(j) * static_cast<uint32_t>(P.res_x) + static_cast<uint32_t>(i);
            RayResult r;
#ifdef C5_EXPERIMENTS
            if (!f32 && P.recs) r = trace_ray<false, false, 0, true>(P, nullptr, P.xs[i], P.ys[j], pixel, 0);
            else
#endif
            r = f32 ? trace_ray<true, false, 0>(P, nullptr, P.xs[i], P.ys[j], pixel, 0)
                    : trace_ray<false, false, 0>(P, nullptr, P.xs[i], P.ys[j], pixel, 0);
            if (!r.deferred) store_pixel(P, i, j, r.tau, r.inten, r.steps);
            P.counters[kSteps] += r.steps;
            P.row_cost[j] += r.steps;
            if (r.steps) P.counters[kHitPixels]++;
            if (r.error) P.counters[kWalkErrors]++;
        }
    }
}

} // namespace

#ifdef C5_EXPERIMENTS
// The step records of the mesh `d` renders (its own, or its parent's for a sibling context), (re)built when
// --alpha_limit changed.
const StepRec* step_records(DeviceState& dd, double alpha_limit) {
    DeviceState& d = dd.origin ? *dd.origin : dd;
    if (d.recs_valid && d.recs_limit == alpha_limit && d.recs.n == static_cast<size_t>(4 * d.n_tets)) return d.recs.p;
    if (4 * d.n_tets >= static_cast<int64_t>(kRecNoFace) || d.n_pts >= static_cast<int64_t>(kRecVtxLimit)) {
        fail(C5_E_INVALID, "walk variant 'rec': mesh too large for 28-bit record / 25-bit vertex indices");
    }
    const bool shared = dd.origin != nullptr || d.mesh_shared;
    if (shared && !kHostSim) C5_CUDA(cudaDeviceSynchronize()); // other lanes may be walking the old records
    d.recs.ensure(static_cast<size_t>(4 * d.n_tets));
    const int64_t n = 4 * d.n_tets;
    count_launch();
    if (kHostSim) {
        for (int64_t f = 0; f < n; f++) d.recs.p[f] = make_step_record(d.cells.p, f);
    } else {
        build_step_records<<<static_cast<unsigned>((n + 255) / 256), 256, 0, dd.stream>>>(n, d.cells.p, d.recs.p);
        C5_CUDA(cudaGetLastError());
        if (shared) C5_CUDA(cudaDeviceSynchronize());
    }
    d.recs_limit = alpha_limit;
    d.recs_valid = true;
    return d.recs.p;
}
#endif

// One view's rays on d.stream: background fill, the pixel kernel, then the grazing-ray kernel for the
// rays it deferred — three launches in stream order, no kernel ever waits for another one. (Round 1
// ran the grazing-ray kernel BESIDE the pixel kernel, spinning on the queue; CUDA promises no forward
// progress between kernels, and with three or four views in flight the ordered form is also the faster
// one: profiles/r02_exp_lanes_graze_after.jsonl against r02_exp_lanes_graze_beside.jsonl.)
void launch_walk(DeviceState& d, const WalkLaunch& w) {
    if (w.precision != 64 && w.precision != 32) fail(C5_E_INVALID, "render: precision must be 64 or 32");
    const bool f32 = w.precision == 32;
    WalkParams P{};
    P.cells = d.cells.p;
    P.vrot = d.vrot.p;
    P.bfaces = d.bfaces.p;
    P.nodes = d.nodes.p;
    P.xs = d.xs.p;
    P.ys = d.ys.p;
    P.mask = w.use_mask ? d.mask.p : nullptr;
    P.out = w.out;
    P.steps = w.write_steps ? d.steps.p : nullptr;
    P.counters = d.counters.p;
    P.row_cost = d.row_cost.p;
    P.queue = d.queue.p;
    P.res_x = w.res_x;
    P.res_y = w.res_y;
    P.row_begin = w.row_begin;
    P.row_end = w.row_end;
    int tile_x = kTileX, tile_y = kTileY;
#ifdef C5_EXPERIMENTS
    const char* variant = std::getenv("C5_WALK_VARIANT");
    const std::string var = variant ? variant : "";
    const bool small_blocks = !f32 && var == "b64"; // 8 x 8 pixel tiles, 64 threads
    if (small_blocks) tile_x = tile_y = 8;
    P.recs = (!f32 && var == "rec") ? step_records(d, w.alpha_limit) : nullptr;
#endif
    // the walk covers [i0,i1) x [j0,j1): the caller's estimate of where the mesh can be, cut to the band
    P.i0 = w.i_begin < 0 ? 0 : w.i_begin;
    P.i1 = w.i_end > w.res_x ? w.res_x : w.i_end;
    P.j0 = w.j_begin < w.row_begin ? w.row_begin : w.j_begin;
    P.j1 = w.j_end > w.row_end ? w.row_end : w.j_end;
    const bool empty_rect = P.i0 >= P.i1 || P.j0 >= P.j1;
    if (empty_rect) P.i0 = P.i1 = P.j0 = P.j1 = 0;
    const bool whole_band = P.i0 == 0 && P.i1 == w.res_x && P.j0 == w.row_begin && P.j1 == w.row_end;
    P.n_tiles_x = (P.i1 - P.i0 + tile_x - 1) / tile_x;
    P.n_tiles_y = (P.j1 - P.j0 + tile_y - 1) / tile_y;
    P.n_macro_x = (P.n_tiles_x + 7) / 8;
    const int n_macro_y = (P.n_tiles_y + 7) / 8;
    // c5_debug_set (tests shrink these to reach the overflow paths on small meshes); defaults otherwise
    P.graze_cap = d.opt_graze_list >= 128 && d.opt_graze_list <= kGrazeList ? (d.opt_graze_list & ~1) : kGrazeList;
    P.query_budget = d.opt_query_budget >= 1 ? d.opt_query_budget : 64;
    P.serial_cap = d.opt_serial_list >= 2 && d.opt_serial_list <= kSerialList ? d.opt_serial_list : kSerialList;
    P.max_steps = static_cast<int>(d.n_tets < (1 << 20) ? d.n_tets : (1 << 20));
    P.round_float = w.round_through_float;
    P.alpha_limit = w.alpha_limit;

    if (!whole_band) {
        count_launch();
        if (kHostSim) {
            background_on_host(P);
        } else {
            const int64_t n = static_cast<int64_t>(w.row_end - w.row_begin) * w.res_x;
            fill_background<<<static_cast<unsigned>((n + 255) / 256), 256, 0, d.stream>>>(P);
            C5_CUDA(cudaGetLastError());
        }
    }
    if (empty_rect) return; // the mesh is not in this band: no rays
    count_launch();
    if (kHostSim) {
        walk_on_host(P, f32);
        count_launch();
        graze_on_host(P, f32);
        return;
    }
    const unsigned grid = static_cast<unsigned>(P.n_macro_x) * static_cast<unsigned>(n_macro_y) * 64u;
    size_t smem = 0;
#ifdef C5_EXPERIMENTS
    const int64_t n_nodes = d.n_bfaces - 1;
    int top = 0;
    if (const char* e = std::getenv("C5_TOP_NODES")) top = std::atoi(e);
    top = top < 0 ? 0 : top > 1023 ? 1023 : top;
    P.top_nodes = static_cast<int>(n_nodes < top ? n_nodes : top);
    smem = static_cast<size_t>(P.top_nodes) * sizeof(BvhNode);
    P.trace = nullptr;
    if (std::getenv("C5_TRACE_FILE") && d.trace_launches < kTraceLaunches) { // block timeline of the first launches
        if (d.trace.n == 0) {
            d.trace.alloc(static_cast<size_t>(kTraceLaunches) * kTraceBlocks * 4);
            dev_zero(d.trace.p, d.trace.bytes(), d.stream);
        }
        if (grid <= kTraceBlocks) {
            P.trace = d.trace.p + static_cast<size_t>(d.trace_launches) * kTraceBlocks * 4;
            d.trace_grid[d.trace_launches] = grid;
            d.trace_launches++;
        }
    }
    if (f32) {
        tet_walk_fp32<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (small_blocks) {
        tet_walk_fp64_b64<<<grid, 64, smem, d.stream>>>(P);
    } else if (P.recs) {
        tet_walk_fp64_rec<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "l128") {
        tet_walk_fp64_l128<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "pf") {
        tet_walk_fp64_pf<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "r80") {
        tet_walk_fp64_r80<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "r64") {
        tet_walk_fp64_r64<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "r96") {
        tet_walk_fp64_r96<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "sp5") {
        tet_walk_fp64_sp5<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "sp6") {
        tet_walk_fp64_sp6<<<grid, kBlock, smem, d.stream>>>(P);
    } else if (var == "sp7") {
        tet_walk_fp64_sp7<<<grid, kBlock, smem, d.stream>>>(P);
    } else {
        tet_walk_fp64<<<grid, kBlock, smem, d.stream>>>(P);
    }
#else
    if (f32) {
        tet_walk_fp32<<<grid, kBlock, smem, d.stream>>>(P);
    } else {
        tet_walk_fp64<<<grid, kBlock, smem, d.stream>>>(P);
    }
#endif
    C5_CUDA(cudaGetLastError());
    if (w.mark_walk_done) C5_CUDA(cudaEventRecord(w.mark_walk_done, d.stream));
    if (w.mark_walk_done_tl) C5_CUDA(cudaEventRecord(w.mark_walk_done_tl, d.stream));

    // Rays the pixel kernel deferred: a fixed grid of persistent warps that draw tickets.
    if (d.sm_count == 0) C5_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device));
    const unsigned graze_grid = static_cast<unsigned>(d.sm_count) * static_cast<unsigned>(d.opt_graze_blocks > 0 ? d.opt_graze_blocks : kGrazeBlocksPerSm);
    count_launch();
    cudaStream_t gs = d.stream;
    if (w.graze_stream && w.mark_walk_done) { // c5_debug_set("prep_priority", 2): on the high-priority stream, still AFTER the pixel kernel
        gs = w.graze_stream;
        C5_CUDA(cudaStreamWaitEvent(gs, w.mark_walk_done, 0));
    }
    if (f32) {
        grazing_rays_fp32<<<graze_grid, 32 * kGrazeWarps, 0, gs>>>(P);
    } else {
        grazing_rays_fp64<<<graze_grid, 32 * kGrazeWarps, 0, gs>>>(P);
    }
    C5_CUDA(cudaGetLastError());
    if (gs != d.stream) {
        C5_CUDA(cudaEventRecord(w.graze_join, gs));
        C5_CUDA(cudaStreamWaitEvent(d.stream, w.graze_join, 0));
    }
}

} // namespace c5
