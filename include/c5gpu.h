/* c5gpu.h — C ABI of the B200 ray pass that replaces course5's OpenMP loops.
 *
 * The reference has no plugin/FFI API; its hot path is reached by three C++ calls in one
 * place (/root/reference/project/src/main.cpp:127-129):
 *
 *     plane base_plane{res_x, res_y, {acc_disk, roche_lobe, acc_sphere}, domain};   // :127
 *     base_plane.find_intersections();                                              // :128
 *     object2d result = base_plane.trace_rays(tetra_value::alpha, tetra_value::Q);  // :129
 *
 * preceded by three rotations per object (main.cpp:104-107,112-114) that the reference
 * charges to its load timer. This header is what a binding for that path links against:
 * plain pointers and sizes, no C++ or torch types, no exceptions. Every call returns C5_OK
 * or a negative C5_E_* code; c5_last_error() gives the text. A context is thread-compatible
 * (one thread at a time), not thread-safe.
 *
 * The library is CUDA-only (sm_100a). There is no CPU path: without a usable device
 * c5_create() fails with C5_E_CUDA.
 */
#ifndef C5GPU_H
#define C5GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C5_ABI_VERSION 2
#define C5_MAX_IN_FLIGHT 8
#define C5_MAX_ROT 8

typedef struct c5_ctx c5_ctx;

enum c5_status {
    C5_OK = 0,
    C5_E_INVALID = -1,  /* bad argument */
    C5_E_CUDA = -2,     /* CUDA runtime error / no device */
    C5_E_NOMEM = -3,
    C5_E_TOPOLOGY = -4, /* mesh is not a conforming manifold partition (a face shared by >2 tets) */
    C5_E_STATE = -5,    /* call order (render before upload_mesh, ...) */
    C5_E_WALK = -6,     /* a ray exceeded the step cap (reported in c5_stats.walk_errors too) */
    C5_E_NCCL = -7
};

/* One rigid rotation, the unit of object3d_base::rotate_around_{x,y}_axis
 * (object3d_base.cpp:202-219 -> tetra.cpp:44-62). axis 0: about the x axis through the origin;
 * axis 1: about the axis parallel to y through (x0, 0, 0). angle in radians. */
typedef struct c5_rotation {
    int32_t axis;
    int32_t reserved;
    double angle;
    double x0;
} c5_rotation;

/* One view. Replaces the implicit inputs of the reference's ray pass: the plane constructor
 * arguments (plane.hpp:19), the window literal (main.cpp:83), the rotations applied to the
 * geometry beforehand (main.cpp:96,104-107) and app::instance().config.limit_alpha_value read
 * inside the loop (line.cpp:204). */
typedef struct c5_view {
    int32_t res_x, res_y;       /* plane(res_x, res_y, ...) */
    double window[4];           /* x_max, x_min, y_max, y_min */
    int32_t n_rot;              /* rotations applied in order to the grid and to view-following solids */
    int32_t reserved0;
    c5_rotation rot[C5_MAX_ROT];
    double alpha_limit;         /* --alpha_limit */
    int32_t precision;          /* 64 (default) or 32 */
    int32_t round_through_float;/* 1: (double)(float)v like plane.cpp:165-166 + object2d.cpp:19-20 */
    int32_t use_solids;         /* 1: pixels under uploaded solid tets become NaN (line.cpp:246-249) */
    int32_t row_begin, row_end; /* row band [row_begin, row_end); 0,0 = all rows */
    int32_t reserved1;
} c5_view;

/* Fills *v the way main.cpp:83-107 sets a view up from the CLI flags (angles in units of pi):
 * window {2.2,-0.2,0.9,-0.9}; rotations Rx(a0), Ry(Y*pi) about x0 = 1, Rx(-a0 + X*pi),
 * a0 = -I*pi + pi/2; precision 64; round_through_float 1; use_solids 1; all rows. */
void c5_view_from_flags(c5_view* v, int32_t res_x, int32_t res_y, double X_pi, double Y_pi, double I_pi,
                        double alpha_limit);

typedef struct c5_stats {
    uint64_t pixels;        /* pixels of the rendered band(s) */
    uint64_t tet_steps;     /* sum over pixels of tets crossed == plane::count_all_intersections (plane.cpp:3-12) */
    uint64_t hit_pixels;    /* pixels crossing >= 1 tet */
    uint64_t solid_pixels;  /* pixels under the solid mask */
    uint64_t walk_errors;   /* rays stopped by the step cap (0 on a valid mesh) */
    float ms_rotate;        /* device time per phase (CUDA events), max over devices */
    float ms_bvh;
    float ms_mask;
    float ms_walk;
    float ms_gather;        /* NCCL band gather (multi-device contexts) */
    float ms_d2h;
    float ms_total;         /* whole call on the device timeline */
    int32_t n_devices;
    int32_t grazing_rays;   /* rays with so many crossings that the grazing-ray kernel finished them */
    float ms_graze;         /* the grazing-ray kernel's share of ms_walk */
    int32_t reserved;
} c5_stats;

typedef struct c5_mesh_info {
    int64_t n_points, n_tets, n_boundary_faces, n_bvh_nodes, n_solid_tets;
    int64_t device_bytes;   /* per device */
} c5_mesh_info;

/* devices: CUDA ordinals, n_dev >= 1. With n_dev > 1 the mesh is replicated, each device
 * renders a contiguous band of rows and one NCCL gather assembles the image on devices[0]. */
int c5_create(const int32_t* devices, int32_t n_dev, c5_ctx** out);
void c5_destroy(c5_ctx* ctx);
const char* c5_last_error(const c5_ctx* ctx); /* ctx may be NULL: text of the last failed c5_create */
int c5_abi_version(void);

/* Replaces object3d_base::read_vtk_file's output (object3d_base.cpp:13-52: an AoS of private
 * per-tet point copies) by shared points + connectivity, which are flattened once into
 * device-resident arrays: Morton-reordered vertices and tets, a face-neighbour table, the
 * boundary-face list and an LBVH hierarchy over it. Caller keeps ownership of the host arrays;
 * the call blocks. alpha = "AbsorpCoef", q = "radEnLooseRate" (object3d_accretion_disk.cpp:4). */
int c5_upload_mesh(c5_ctx* ctx, const double* points_xyz, int64_t n_points, const int32_t* tet_vertices,
                   int64_t n_tets, const double* alpha, const double* q);

/* Solid tets (the reference's Roche lobe and sphere, tetra_type::solid): [n][4][3] doubles in the
 * frame BEFORE the view rotations. follows_view = 1 for objects main.cpp rotates with the view
 * (the Roche lobe, main.cpp:112-114), 0 for those it does not (the sphere, main.cpp:116). Calls
 * append; c5_clear_solids() empties the set. */
int c5_upload_solids(c5_ctx* ctx, const double* tet_points, int64_t n_tets, int32_t follows_view);
int c5_clear_solids(c5_ctx* ctx);

int c5_mesh_info_get(const c5_ctx* ctx, c5_mesh_info* out);

/* plane ctor + find_intersections + trace_rays for one view (main.cpp:127-129), including the
 * rotations main.cpp:104-107 does beforehand. out: caller-owned HOST buffer of
 * res_y * res_x * 2 doubles, x fastest, {tau, I} per pixel — the vtkImageData layout of
 * object2d.cpp:17-21 (with a row band: only rows [row_begin,row_end) are written, at their
 * final position). stats may be NULL. If `out` is page-locked (cudaHostAlloc, cudaHostRegister or
 * c5_host_register) and 16-byte aligned the walk kernel stores its pixels straight into it over
 * PCIe (the shortest path for one view at a time); any other buffer gets a device-to-host copy. */
int c5_render(c5_ctx* ctx, const c5_view* view, double* out, c5_stats* stats);

/* The same pass, asynchronous: c5_render_submit enqueues the view and returns a ticket at once;
 * c5_render_wait blocks until that view's image is complete in `out` and fills stats (may be NULL).
 * Up to c5_set_views_in_flight() views (default 3, at most C5_MAX_IN_FLIGHT) may be submitted before
 * the first is waited for; each is rendered by a lane of its own — per-view device state plus a
 * stream; the mesh is shared — so that consecutive views overlap on the device: the tail of one
 * view's walk and its grazing-ray kernel run beside the next view's rays, and the copy engine moves
 * one view's image to the host while the SMs walk the next. `out` (page-locked for any overlap: a
 * copy to pageable memory blocks the submitting thread) must stay valid and untouched
 * until the wait returns; *view is copied. One more submit than lanes returns C5_E_STATE. Tickets may
 * be waited for in any order, once each. A sweep is: submit k+1, k+2; wait k; write frame k; ...
 * Single-device contexts that are not siblings. */
int c5_render_submit(c5_ctx* ctx, const c5_view* view, double* out, uint64_t* ticket);
int c5_render_wait(c5_ctx* ctx, uint64_t ticket, c5_stats* stats);
int c5_set_views_in_flight(c5_ctx* ctx, int32_t n);

/* Test/diagnostic variant: same pass, but also returns per-pixel tets crossed and the solid
 * mask (each res_y * res_x, x fastest; either may be NULL). With view->round_through_float = 0
 * the doubles are the pre-cast values the parity gate is defined on. */
int c5_render_raw(c5_ctx* ctx, const c5_view* view, double* out, uint32_t* steps, uint8_t* solid_mask,
                  c5_stats* stats);

/* Device-resident variant for callers that own device memory (e.g. a torch tensor that an NCCL
 * gather will read): d_out is a DEVICE pointer on the context's first device to
 * (row_end - row_begin) * res_x * 2 doubles — the band only. Work is enqueued on `stream`
 * (a cudaStream_t; NULL = the legacy default stream). d_out must be 16-byte aligned (pixels are
 * stored as one 128-bit word; C5_E_INVALID otherwise). With stats != NULL the call returns
 * after the stream has been synchronised, so stats (and c5_last_row_cost) are final. With
 * stats == NULL nothing is read back and the call returns as soon as the work is enqueued: the
 * caller orders later work on the same stream (successive views and an NCCL gather then pipeline
 * on the device without host round trips). Single-device contexts only. */
int c5_render_device(c5_ctx* ctx, const c5_view* view, void* d_out, void* stream, c5_stats* stats);

/* Per-row tet-step totals of the last render on this context (res_y entries; rows outside the
 * rendered band are 0). Callers use it to cut cost-balanced row bands for the next view. */
int c5_last_row_cost(c5_ctx* ctx, uint64_t* rows, int32_t n_rows);

/* A second context on the same device that SHARES parent's uploaded mesh and solids (no copy) and
 * owns only per-view state (rotated vertices, BVH boxes, mask, counters). Two views can then be in
 * flight at once — c5_render_device(parent, ..., stream_a) and c5_render_device(sibling, ...,
 * stream_b) — so that the tail of one view's walk overlaps the start of the next one's: a sweep's
 * throughput is then set by the work, not by the last rays of every view. Uploads go through the
 * parent (a sibling picks them up at its next render); destroy siblings before the parent. */
int c5_create_sibling(c5_ctx* parent, c5_ctx** out);

/* ---- one image, several processes (one process per GPU) -----------------------------------------
 * Row bands are independent (plane.cpp:161-169 has no cross-pixel state), so N processes can
 * render N bands of one view straight into ONE image and nothing has to be gathered afterwards:
 *
 *  - a DEVICE image on one GPU that the other GPUs' walk kernels store into through NVLink peer
 *    mappings (CUDA IPC): the owner calls c5_image_create and passes the 64-byte handle to the
 *    other processes (any byte transport), they call c5_image_open and hand
 *    `(char*)ptr + row_begin * res_x * 16` to c5_render_device as d_out;
 *  - a HOST image in shared memory (shm_open / mmap, the caller's business) that every process
 *    registers with c5_host_register: c5_render() with a row band then writes the band in place
 *    over that GPU's own PCIe link (a registered buffer is device-addressable, so the walk stores
 *    into it directly).
 *
 * Ordering between processes (all bands of view k written before anyone reads view k) is the
 * caller's: one barrier per view on the streams involved. */
#define C5_IPC_HANDLE_BYTES 80 /* a cudaIpcMemHandle_t (64) + the image's offset in its allocation + its size */
int c5_image_create(c5_ctx* ctx, uint64_t bytes, void** d_ptr, uint8_t handle[C5_IPC_HANDLE_BYTES]);
int c5_image_open(c5_ctx* ctx, const uint8_t handle[C5_IPC_HANDLE_BYTES], void** d_ptr);
int c5_image_close(c5_ctx* ctx, void* d_ptr); /* frees (owner) or unmaps (importer) */
int c5_host_register(c5_ctx* ctx, void* ptr, uint64_t bytes);
int c5_host_unregister(c5_ctx* ctx, void* ptr);

/* Number of kernels this library has launched on behalf of ctx (and the lanes c5_render_submit
 * made for it) since creation. */
uint64_t c5_kernel_launches(const c5_ctx* ctx);

/* ---- diagnostics (tests, profiling scripts; no effect on results) ----------------------------------
 * c5_debug_set(ctx, key, value) — applies to ctx, its lanes and its siblings:
 *   "graze_list"    entries one cooperative collection of the grazing-ray kernel holds (128..256, 0 = default):
 *                   tests shrink it to reach the overflow path on small meshes
 *   "query_budget"  BVH nodes a thread of the pixel kernel may visit per search before it hands the
 *                   ray to the grazing-ray kernel (>= 1, 0 = default 64)
 *   "serial_list"   the same for the serial form of the CPU test build (2..64)
 *   "graze_blocks"  blocks per SM of the grazing-ray kernel's grid (tuning experiments; 0 = default)
 *   "mask_tile"     100 w + h: tile size of the solid mask's "already solid" flags (0 = default)
 *   "mask_per_face" n > 0: the solid mask's face passes run one face per thread, 2^(n-1) lanes per tall face
 *                   (only in the experiments build, libc5gpu_exp.so; the product library ignores it)
 *   "no_static_mask" 1: solids that do not follow the view are scan-converted in every view, like the
 *                   others, instead of once per pixel grid (tests compare the two)
 *   "prep_priority" bit 0: rotate / refit / mask of a view run on a high-priority stream of their own;
 *                   bit 1: so does the grazing-ray kernel (still after the pixel kernel, by event)
 *   "no_zero_copy"  1: page-locked output buffers get a device-to-host copy like pageable ones
 *   "timeline"      n > 0: keep the phase events of the last n views of every lane (0 = off)
 * Unknown keys return C5_E_INVALID.
 * c5_timeline_read: for the last views rendered by ctx itself (not its lanes; oldest first, at most
 * max_views), six times in milliseconds since `origin` (a cudaEvent_t the caller recorded earlier
 * on the same device): view start, vertices rotated, BVH refitted, solid mask done, pixel kernel
 * done, grazing-ray kernel done. Synchronises with the last of them. */
int c5_debug_set(c5_ctx* ctx, const char* key, int64_t value);
int c5_timeline_read(c5_ctx* ctx, void* origin, float* ms_out, int32_t max_views, int32_t* n_views);

#ifdef __cplusplus
}
#endif
#endif /* C5GPU_H */
