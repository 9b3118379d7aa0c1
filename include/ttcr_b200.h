/*
 * ttcr_b200 -- C ABI of the B200-native 3D rectilinear fast-sweeping (FSM) eikonal solver.
 *
 * This is the drop-in boundary for ONE hot path of groupeLIAMG/ttcr: the classes
 * Grid3Drnfs / Grid3Drcfs (and their OpenCL twins) behind ttcrpy.rgrid.Grid3d.  The
 * reference has no C ABI: its boundary is the abstract C++ class ttcr::Grid3D<T1,T2>
 * (ttcr/Grid3D.h:44-466) as declared for Cython in src/ttcrpy/rgrid.pxd:30-106.  Every
 * entry point below names the reference member it replaces.  A header-only adapter that
 * subclasses ttcr::Grid3D<T1,T2> and forwards to this ABI is in
 * include/Grid3Drfs_B200.h; the Cython / ctypes bindings are shown in INTEGRATION.md.
 *
 * Conventions
 *  - All pointers are HOST pointers owned by the caller; the library copies in and out
 *    (same ownership rule as the reference: std::vector by value/reference, rgrid.pyx:153).
 *    The library owns all device memory behind the handle.
 *  - Element type of every `const void*` / `void*` array is the grid dtype chosen at
 *    creation (TTCR_B200_F64 -> double, TTCR_B200_F32 -> float).
 *  - nx, ny, nz are CELL counts (reference constructor, ttcr/Grid3Drnfs.h:39-50); node
 *    counts are nx+1, ny+1, nz+1.
 *  - `order` selects the memory order of full-grid arrays:
 *      TTCR_B200_ORDER_X_FASTEST : n = (k*(ny+1)+j)*(nx+1)+i, the reference C++ order
 *                                  (ttcr/Grid3Drn.h:2823; cells: (k*ny+j)*nx+i, :404)
 *      TTCR_B200_ORDER_Z_FASTEST : n = (i*(ny+1)+j)*(nz+1)+k, numpy C order of an
 *                                  (nx+1,ny+1,nz+1) array -- lets the Python surface skip
 *                                  the order='F' flatten of rgrid.pyx:559-566 / :435.
 *  - Every function returns a status code; ttcr_b200_last_error() gives the message.
 *    The status codes map 1:1 onto the C++ exception types the reference throws, so the
 *    adapter can rethrow them and Cython's `except +` behaviour is unchanged.
 *  - Thread safety: concurrent ttcr_b200_raytrace() calls on one handle are allowed for
 *    DISTINCT slots (slot == the reference's threadNo; per-slot traveltime field and CUDA
 *    stream).  Slowness must not change during solves.
 */
#ifndef TTCR_B200_H
#define TTCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ttcr_b200_grid ttcr_b200_grid;

enum {
    TTCR_B200_OK = 0,
    TTCR_B200_ERR_RUNTIME = 1, /* std::runtime_error: point outside grid (Grid3Drn.h:784-786) */
    TTCR_B200_ERR_LENGTH = 2,  /* std::length_error: slowness size mismatch (Grid3Drn.h:82-85) */
    TTCR_B200_ERR_LOGIC = 3,   /* std::logic_error (Grid3Drnfs.h:111-113) */
    TTCR_B200_ERR_INVALID = 4, /* std::invalid_argument: bad handle / argument */
    TTCR_B200_ERR_CUDA = 5,    /* CUDA failure (no CPU fallback: the call fails) */
    TTCR_B200_ERR_UNSUPPORTED = 6
};

enum { TTCR_B200_F64 = 0, TTCR_B200_F32 = 1 };
enum { TTCR_B200_ORDER_X_FASTEST = 0, TTCR_B200_ORDER_Z_FASTEST = 1 };

/* Per-slot statistics of the last solve (extension; reference exposes only get_niter /
 * get_niterw, Grid3Drnfs.h:56-57, and the OpenCL profile printout, Grid3Drn_OpenCL.h:204-229). */
typedef struct {
    int niter;            /* first-order iterations (groups of 8 directional sweeps) */
    int niterw;           /* WENO iterations */
    double solve_ms;      /* device time, CUDA events: reinit + initFSM + sweeps + reductions */
    double sweep_ms;      /* device time of the sweep kernels only (CUDA events around each directional sweep) */
    long long launches;   /* kernels launched by the solve */
    long long sweep_launches; /* of which sweep kernels */
    int sweeps;           /* directional sweeps executed: 8 * (niter + niterw) */
    double last_change;   /* L1 change of the last iteration */
    int kernel;           /* sweep kernel actually used (TTCR_B200_KERNEL_*) */
} ttcr_b200_stats;

enum { TTCR_B200_KERNEL_AUTO = 0, TTCR_B200_KERNEL_PLANE = 1, TTCR_B200_KERNEL_TILE = 2, TTCR_B200_KERNEL_COOP = 6, TTCR_B200_KERNEL_MARCH = 7 };   /* (3..5: earlier marching kernels, retired) */

/* Replaces: new Grid3Drnfs<T,uint32_t>(nx,ny,nz,dx,xmin,ymin,zmin,eps,maxit,weno,ttrp,intVel,nt,
 * translateOrigin) (Grid3Drnfs.h:39-50, rgrid.pyx:256-261) and the Grid3Drcfs twin
 * (Grid3Drcfs.h:40-52, rgrid.pyx:221-227) when cell_slowness != 0.
 * n_slots == the reference's nt (one traveltime field per slot).  device: CUDA ordinal, or -1
 * for the current device. */
int ttcr_b200_create(ttcr_b200_grid** out, uint32_t nx, uint32_t ny, uint32_t nz, double dx,
                     double xmin, double ymin, double zmin, double eps, int maxit, int weno,
                     int tt_from_rp, int interp_vel, size_t n_slots, int translate_origin,
                     int cell_slowness, int dtype, int device);

/* Replaces: delete grid (rgrid.pyx:284-285). */
void ttcr_b200_destroy(ttcr_b200_grid* g);

/* Message of the last error on this handle (or of the last failed create if g == NULL). */
const char* ttcr_b200_last_error(const ttcr_b200_grid* g);

/* Replaces: Grid3Drn::setSlowness (Grid3Drn.h:82-89, n == nodes) or Grid3Drcfs::setSlowness
 * (Grid3Drcfs.h:88-171, n == cells; 8/4/2/1-cell average onto nodes, on the device).
 * TTCR_B200_ERR_LENGTH on size mismatch. */
int ttcr_b200_set_slowness(ttcr_b200_grid* g, const void* s, size_t n, int order);

/* Extension: same as ttcr_b200_set_slowness but `s_dev` is a DEVICE pointer on the grid's device
 * (e.g. the landing buffer of an NCCL broadcast; the reference's OpenCL upload,
 * Grid3Drn_OpenCL.h:675-703, has no such path).  Synchronous with respect to the caller's prior
 * work on that buffer only if the caller has synchronised; the call itself returns when done. */
int ttcr_b200_set_slowness_device(ttcr_b200_grid* g, const void* s_dev, size_t n, int order);

/* Extension: the x planes [i_first, i_first + i_count) of a NODE model that is arriving piecewise in the device buffer
 * `s_dev` (the whole (nx+1)(ny+1)(nz+1) array, numpy order = TTCR_B200_ORDER_Z_FASTEST only): lets the caller import each
 * chunk of a pipelined upload / NCCL broadcast while the next chunk is still travelling.  The caller has synchronised the
 * chunk's arrival; the call enqueues the import and returns, except for the chunk that ends at the last plane, which
 * waits for all of them and makes the model current.  Cell models: TTCR_B200_ERR_INVALID (use set_slowness_device). */
int ttcr_b200_set_slowness_device_planes(ttcr_b200_grid* g, const void* s_dev, size_t n, int i_first, int i_count);

/* Replaces: Grid3Drn::getSlowness (Grid3Drn.h:90-97): NODE slowness, (nx+1)(ny+1)(nz+1) values. */
int ttcr_b200_get_slowness(ttcr_b200_grid* g, void* out, int order);

/* Replaces: Grid3D::raytrace(Tx,t0,Rx,traveltimes,threadNo) (Grid3D.h:115-119, :470-502) ->
 * Grid3Drnfs::raytrace (Grid3Drnfs.h:84-155).  tx_xyz: ntx x 3, rx_xyz: nrx x 3 (row major),
 * t0: ntx, tt_out: nrx (may be NULL when nrx == 0).  All Tx points belong to ONE source.
 * Receiver traveltimes are Grid3Drn::getTraveltime (trilinear, Grid3Drn.h:794-930) when the grid was created
 * with tt_from_rp = 0, else Grid3Drn::getTraveltimeFromRaypath (Grid3Drn.h:1103-1243), as Grid3D.h:493-501.
 * TTCR_B200_ERR_RUNTIME if a point is outside the grid (checkPts, Grid3Drn.h:771-790). */
int ttcr_b200_raytrace(ttcr_b200_grid* g, const void* tx_xyz, const void* t0, size_t ntx,
                       const void* rx_xyz, size_t nrx, void* tt_out, size_t slot);

/* Replaces: Grid3D::raytrace(Tx,t0,Rx,traveltimes,r_data,threadNo) (Grid3D.h:545-586) -> Grid3Drnfs::raytrace, then
 * Grid3Drn::getRaypath per receiver (Grid3Drn.h:1339-1500): traveltimes integrated along the raypaths AND the raypaths.
 * ray_npts (nrx values): number of points of every ray (first = the receiver, last = the source point reached).  The
 * points themselves stay on the slot; fetch them with ttcr_b200_get_rays into a buffer of 3 * sum(ray_npts) elements
 * (std::vector<std::vector<sxyz<T>>> flattened: ray after ray, x y z per point). */
int ttcr_b200_raytrace_rays(ttcr_b200_grid* g, const void* tx_xyz, const void* t0, size_t ntx, const void* rx_xyz,
                            size_t nrx, void* tt_out, size_t* ray_npts, size_t slot);
int ttcr_b200_get_rays(ttcr_b200_grid* g, size_t slot, void* xyz_out);

/* Replaces: the m_data overloads of Grid3D::raytrace (Grid3D.h:85-95, :146-154, :187-194, :646-690, :743-780), i.e.
 * Grid3Drn::getRaypath with m_data (Grid3Drn.h:1500-1801, :2144-2448): the sensitivity matrix M of node-slowness grids.
 * With option "m_terms" = 1 (as the overload with m_data only) or 2 (as the overload with r_data and m_data: the reference's two
 * overloads differ, see raypath.cuh), ttcr_b200_raytrace_rays also walks the M terms; this call returns them RAW: 8 (column, value)
 * pairs per ray point, ray after ray, in the order the reference produces them (the 8 pairs of a ray's first point are zero:
 * the receiver closes no segment).  node_out / value_out: 8 * sum(ray_npts) elements.  The reference then merges the terms
 * of equal column of a ray in order of appearance (m_data[nm].v += m.v); so does the caller (ttcr_b200/rgrid.py). */
int ttcr_b200_get_m_terms(ttcr_b200_grid* g, size_t slot, unsigned long long* node_out, void* value_out);

/* Replaces: Grid3D::raytrace(vector<vector<sxyz>>&Tx, vector<vector<T>>&t0, vector<vector<sxyz>>&Rx,
 * vector<vector<T>>&tt) (Grid3D.h:172-175, :810-853).  Source s owns Tx points
 * [tx_off[s], tx_off[s+1]) and receivers [rx_off[s], rx_off[s+1]); tt_out is indexed like rx.
 * Sources are dealt round-robin to the slots, each slot on its own CUDA stream (the
 * reference's one-source-per-thread fan-out).  niter/niterw (may be NULL): per source. */
int ttcr_b200_raytrace_multi(ttcr_b200_grid* g, size_t nsrc, const size_t* tx_off, const void* tx_xyz,
                             const void* t0, const size_t* rx_off, const void* rx_xyz, void* tt_out,
                             int* niter, int* niterw);

/* Replaces: Grid3Drn::getTT(tt, threadNo) (Grid3Drn.h:102-108): the full traveltime field. */
int ttcr_b200_get_tt(ttcr_b200_grid* g, void* out, size_t slot, int order);

/* Extension: the full traveltime field into a DEVICE buffer of (nx+1)(ny+1)(nz+1) elements. */
int ttcr_b200_get_tt_device(ttcr_b200_grid* g, void* out_dev, size_t slot, int order);

/* Replaces: Grid3Drnfs::get_niter / get_niterw (Grid3Drnfs.h:56-57), per slot. */
int ttcr_b200_get_niter(ttcr_b200_grid* g, size_t slot, int* niter, int* niterw);

/* Replaces: Grid3D::setTraveltimeFromRaypath / setUsePool (Grid3D.h:287,302-309) and tuning knobs.
 * keys: "tt_from_rp" (0/1), "kernel" (TTCR_B200_KERNEL_*), "tile_rows" (flag chunk, rows),
 *       "ctas_per_sm", "march_nodes" (k_sweep_march: 2 or 4 nodes per thread and step, 0 = by grid size), "tile_warps" (compute warps per
 *       tile, 0 = by grid size), "m_terms" (0/1/2: ttcr_b200_raytrace_rays also produces the raw terms of the matrix M), "weno_kernel"
 *       (kernel of the WENO stage), "max_ctas" (cap on a marching kernel's grid), "plane_graph" / "plane_pdl" (0/1: replay the plane-per-launch sweeps of the WENO stage from a
 *       captured CUDA graph / chain them by programmatic dependent launch), "use_pool" (accepted, ignored). */
int ttcr_b200_set_option(ttcr_b200_grid* g, const char* key, double value);

/* Replaces: Grid3D::getNthreads (Grid3D.h:309). */
size_t ttcr_b200_n_slots(const ttcr_b200_grid* g);

/* Extension.  Solve only: sources already described by host Tx (tiny), slowness resident on the
 * device; the field stays on the device (no receiver extraction, no field copy).  This is the
 * region the Mnodes/s metric is defined on. */
int ttcr_b200_solve(ttcr_b200_grid* g, const void* tx_xyz, const void* t0, size_t ntx, size_t slot);

/* Extension: statistics of the last solve on a slot. */
int ttcr_b200_get_stats(ttcr_b200_grid* g, size_t slot, ttcr_b200_stats* out);

/* Extension: device memory held by the handle, in bytes. */
size_t ttcr_b200_device_bytes(const ttcr_b200_grid* g);

/* Library version string. */
const char* ttcr_b200_version(void);

/* ---- 2-D twins: Grid2Drnfs / Grid2Drcfs (ttcr/Grid2Drnfs.h, ttcr/Grid2Drcfs.h; OpenCL: ttcr/Grid2Drn_OpenCL.h:405-600) ----
 * Arrays are numpy C order of shape (nx+1, nz+1) nodes / (nx, nz) cells = the reference's own node / cell index
 * (n = i * (nz+1) + j, Grid2Drnfs.h buildGridNodes; cell i * nz + j).  Points are (x, z) pairs. */
typedef struct ttcr_b200_grid2d ttcr_b200_grid2d;

/* Replaces: the Grid2Drnfs / Grid2Drcfs constructors (Grid2Drnfs.h:44-58, Grid2Drcfs.h:42-56; selected in
 * src/ttcrpy/rgrid.pyx Grid2d.__cinit__ for method='FSM').  nx, nz are CELL counts; `weno` and `rotated_template` as there;
 * n_slots = the reference's nt (one traveltime field per slot). */
int ttcr_b200_create2d(ttcr_b200_grid2d** out, uint32_t nx, uint32_t nz, double dx, double dz, double xmin, double zmin, double eps,
                       int maxit, int weno, int rotated_template, size_t n_slots, int cell_slowness, int dtype, int device);
void ttcr_b200_destroy2d(ttcr_b200_grid2d* g);
/* Replaces: Grid2Drn::setSlowness / Grid2Drcfs::setSlowness (Grid2Drcfs.h:99-138: cell -> node averaging). */
int ttcr_b200_set_slowness2d(ttcr_b200_grid2d* g, const void* s, size_t n);
/* Replaces: Grid2Drn::getSlowness: NODE slowness. */
int ttcr_b200_get_slowness2d(ttcr_b200_grid2d* g, void* out);
/* Replaces: Grid2Drnfs::raytrace(Tx, t0, Rx, traveltimes, threadNo) (Grid2Drnfs.h:195-299) + getTraveltime (Grid2Drn.h:359-415). */
int ttcr_b200_raytrace2d(ttcr_b200_grid2d* g, const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt_out, size_t slot);
/* Replaces: Grid2D::raytrace over a vector of sources (the thread fan-out of ttcr/Grid2D.h): up to n_slots sources per launch,
 * one CTA each.  niter_out: (niter, niterw) per source, or NULL. */
int ttcr_b200_raytrace2d_multi(ttcr_b200_grid2d* g, size_t n_sources, const size_t* tx_off, const void* tx, const void* t0, const size_t* rx_off,
                               const void* rx, void* tt_out, int* niter_out);
/* Replaces: Grid2Drn::getTT. */
int ttcr_b200_get_tt2d(ttcr_b200_grid2d* g, void* out, size_t slot);
/* Replaces: Grid2Drnfs::get_niter / get_niterw. */
int ttcr_b200_get_niter2d(ttcr_b200_grid2d* g, size_t slot, int* niter, int* niterw);
/* Extension: device time (CUDA events) of the solves of the last raytrace call, milliseconds. */
double ttcr_b200_last_solve_ms2d(const ttcr_b200_grid2d* g);

#ifdef __cplusplus
}
#endif
#endif /* TTCR_B200_H */
