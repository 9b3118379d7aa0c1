// 2-D twins of the FSM path: Grid2Drnfs / Grid2Drcfs on the device.
//
// Reference: ttcr/Grid2Drnfs.h:195-299 (raytrace), ttcr/Grid2Drn.h:713-917 (the four Gauss-Seidel passes of the five node
// updates), :920-1357 (update_node, update_node45, update_node_xz, update_node_weno3, update_node_weno3_xz), :1360-1419
// (initFSM), :359-415 (getTraveltime), ttcr/Grid2Drcfs.h:99-138 (cell -> node slowness); OpenCL twins
// ttcr/Grid2Drn_OpenCL.h:405-600, ttcr/Grid2Drn_kernels.cl.
//
// Design.  A 2-D problem is small for a B200: 2000 x 2000 nodes (the largest case of the reference's published table) are
// 16 MB per array and live in L2, and a wavefront holds at most min(nx, nz) nodes.  So ONE CTA (1024 threads) solves one
// source completely in ONE launch -- reinit, initFSM, every sweep of every iteration, the L1 convergence sums and the
// decision to stop all stay on the device, a CTA barrier per wavefront -- and independent sources run on different SMs
// (grid = sources in flight: the source-parallel fan-out of ttcr/Grid2D.h inside one launch).  Wavefronts: anti-diagonals
// i + j = const for the axis-aligned stencils (first order, dx != dz, WENO: the nodes a lexicographic Gauss-Seidel pass has
// already updated are exactly those on earlier diagonals), ROWS for the stencil rotated by pi/4 (it only reads the rows
// i - 1 and i + 1).  Any order that respects these dependencies reproduces the reference's lexicographic pass bit for bit.
//
// Arithmetic: T = double evaluates the reference's expressions operation for operation (this TU is compiled with
// -fmad=false); T = float evaluates the same expressions in float (the reference's float build promotes through its double
// literals; tolerance 1e-4, tests/).
#pragma once
#include <cooperative_groups.h>

#include "update.cuh"

namespace ttcrb200 {

template <typename T>
struct P2 {
    int ncx, ncz;          // cells
    T dx, dz, xmin, zmin;
    T eps_total;           // eps * node count (Grid2Drnfs.h:92)
    int maxit, weno, rotated;
    const T* s;            // node slowness, n = i * (ncz + 1) + j
    T* tt;                 // [source in flight][node]
    unsigned char* frozen; // [source in flight][node]
    const T* tx;           // [source][ntx_max][2]
    const T* t0;           // [source][ntx_max]
    const int* ntx;        // [source]
    int ntx_max;
    int* niter;            // [source][2]
    int* err;              // [source]: 1 = Tx outside the grid
};

// A traveltime read.  CG: the field is shared by the CTAs of a cluster (k2d_solve<T, true>): the L1 of an SM is not coherent
// with the stores of the other SMs, so the read goes to L2 (ld.global.cg); stores write through anyway.
template <bool CG, typename T>
__device__ __forceinline__ T ld2(const T* p) {
    if (CG) return __ldcg(p);
    return *p;
}
// first-order one-sided minimum along one axis (Grid2Drn.h:924-943)
template <bool CG, typename T>
__device__ __forceinline__ T ax1_2d(const T* tt, size_t n, int q, int nc, size_t st) {
    if (q == 0) return ld2<CG>(tt + n + st);
    if (q == nc) return ld2<CG>(tt + n - st);
    const T a = ld2<CG>(tt + n - st), t = ld2<CG>(tt + n + st);
    return a < t ? a : t;
}
// per-axis WENO estimate: the branch order of Grid2Drn.h:1080-1190 is the one of axis_weno (update.cuh)
template <bool CG, typename T>
__device__ __forceinline__ T axw_2d(const T* tt, size_t n, int q, int nc, size_t st, T d) {
    const T vm2 = q >= 2 ? ld2<CG>(tt + n - 2 * st) : T(0), vm1 = q >= 1 ? ld2<CG>(tt + n - st) : T(0);
    const T vp1 = q <= nc - 1 ? ld2<CG>(tt + n + st) : T(0), vp2 = q <= nc - 2 ? ld2<CG>(tt + n + 2 * st) : T(0);
    return axis_weno<T>(vm2, vm1, ld2<CG>(tt + n), vp1, vp2, q, nc, d);
}
// the two local solvers: square cells (Grid2Drn.h:945-953) and dx != dz (:1041-1057); return the new value (or the old one)
template <typename T>
__device__ __forceinline__ T solve_sq_2d(T old, T a, T b, T fh) {
    T t;
    T d = a - b;
    d = d < 0 ? -d : d;
    if (d >= fh) t = (a < b ? a : b) + fh;
    else t = T(0.5) * (a + b + sqrt(T(2.) * fh * fh - (a - b) * (a - b)));
    return t < old ? t : old;
}
template <typename T>
__device__ __forceinline__ T solve_xz_2d(T old, T a, T b, T s, T dx, T dz) {
    T t;
    if (a < b && ((b - a) / dx) > s) {
        t = a + s * dx;
    } else if (a > b && ((a - b) / dz) > s) {
        t = b + s * dz;
    } else {
        const T dx2 = dx * dx, dz2 = dz * dz, s2 = s * s;
        t = (b * dx2 + a * dz2) / (dx2 + dz2) +
            sqrt((T(2.0) * a * b * dx2 * dz2 - a * a * dx2 * dz2 - b * b * dx2 * dz2 + dx2 * dx2 * dz2 * s2 + dx2 * dz2 * dz2 * s2) /
                 ((dx2 + dz2) * (dx2 + dz2)));
    }
    return t < old ? t : old;
}

// kind: 0 update_node, 1 update_node45, 2 update_node_xz, 3 update_node_weno3, 4 update_node_weno3_xz.  Returns old - new.
template <bool CG, typename T>
__device__ __forceinline__ T update_2d(const P2<T>& p, T* tt, int i, int j, int kind) {
    const size_t st = (size_t)p.ncz + 1, n = (size_t)i * st + j;
    const int ncx = p.ncx, ncz = p.ncz;
    const T old = ld2<CG>(tt + n);
    T a, b, t, nw;
    switch (kind) {
        case 0:
            a = ax1_2d<CG>(tt, n, i, ncx, st);
            b = ax1_2d<CG>(tt, n, j, ncz, 1);
            nw = solve_sq_2d(old, a, b, p.s[n] * p.dx);
            break;
        case 1: {   // stencil rotated by pi/4 (Grid2Drn.h:957-1015): +MAX off the grid
            const T M = Lim<T>::max();
            const T pp = (i != ncx && j != ncz) ? ld2<CG>(tt + n + st + 1) : M, mm = (i != 0 && j != 0) ? ld2<CG>(tt + n - st - 1) : M;
            const T pm = (i != ncx && j != 0) ? ld2<CG>(tt + n + st - 1) : M, mp = (i != 0 && j != ncz) ? ld2<CG>(tt + n - st + 1) : M;
            if (i == 0) { a = pp; b = pm; }
            else if (i == ncx) { a = mm; b = mp; }
            else { a = pp; t = mm; a = a < t ? a : t; b = pm; t = mp; b = b < t ? b : t; }
            nw = solve_sq_2d(old, a, b, T(1.414213562373095) * p.s[n] * p.dx);
            break;
        }
        case 2:
            a = ax1_2d<CG>(tt, n, i, ncx, st);
            b = ax1_2d<CG>(tt, n, j, ncz, 1);
            nw = solve_xz_2d(old, a, b, p.s[n], p.dx, p.dz);
            break;
        case 3:
            a = axw_2d<CG>(tt, n, i, ncx, st, p.dx);
            b = axw_2d<CG>(tt, n, j, ncz, 1, p.dx);   // (sic: dx on both axes, the scheme requires dx == dz)
            nw = solve_sq_2d(old, a, b, p.s[n] * p.dx);
            break;
        default:
            a = axw_2d<CG>(tt, n, i, ncx, st, p.dx);
            b = axw_2d<CG>(tt, n, j, ncz, 1, p.dz);
            nw = solve_xz_2d(old, a, b, p.s[n], p.dx, p.dz);
    }
    if (nw < old) { tt[n] = nw; return old - nw; }
    return T(0);
}

// initFSM (Grid2Drn.h:1360-1419), one thread, Tx points in order
template <typename T>
__device__ void init_2d(const P2<T>& p, T* tt, unsigned char* frozen, const T* tx, const T* t0, int ntx, int npts) {
    const double small = 1.e-4;
    const int ncx = p.ncx, ncz = p.ncz;
    const size_t st = (size_t)ncz + 1;
    const T dx = p.dx, dz = p.dz, xmin = p.xmin, zmin = p.zmin;
    const T xmax = xmin + ncx * dx, zmax = zmin + ncz * dz;
    for (int n = 0; n < ntx; ++n) {
        const T px = tx[2 * n], pz = tx[2 * n + 1];
        int fi = -1, fj = -1;
        {   // first node in index order within `small` (only the nodes next to the point can match)
            const int ci = (int)floor(((double)px - (double)xmin) / (double)dx + 0.5), cj = (int)floor(((double)pz - (double)zmin) / (double)dz + 0.5);
            for (int i = max(0, ci - 2); i <= min(ncx, ci + 2) && fi < 0; ++i) { const T x = xmin + i * dx; if (fabs((double)(x - px)) < small) fi = i; }
            for (int j = max(0, cj - 2); j <= min(ncz, cj + 2) && fj < 0; ++j) { const T z = zmin + j * dz; if (fabs((double)(z - pz)) < small) fj = j; }
        }
        if (fi >= 0 && fj >= 0) {
            const size_t nn = (size_t)fi * st + fj;
            tt[nn] = t0[n];
            frozen[nn] = 1;
            for (int ii = fi - npts; ii <= fi + npts; ++ii) {
                if (ii < 0 || ii > ncx) continue;
                for (int jj = fj - npts; jj <= fj + npts; ++jj) {
                    if (jj < 0 || jj > ncz || (ii == fi && jj == fj)) continue;
                    const size_t nnn = (size_t)ii * st + jj;
                    const T X = xmin + ii * dx, Z = zmin + jj * dz;
                    const T dist = sqrt((X - px) * (X - px) + (Z - pz) * (Z - pz));
                    tt[nnn] = t0[n] + dist * T(0.5) * (p.s[nnn] + p.s[nn]);
                    frozen[nnn] = 1;
                }
            }
        } else {
            const T x = (double)(xmax - px) < small ? T(xmax - T(.5) * dx) : px;
            const T z = (double)(zmax - pz) < small ? T(zmax - T(.5) * dz) : pz;
            const int i = (int)(unsigned)(small + (double)((x - xmin) / dx)), j = (int)(unsigned)(small + (double)((z - zmin) / dz));
            for (int ii = i - (npts - 1); ii <= i + npts; ++ii) {
                if (ii < 0 || ii > ncx) continue;
                for (int jj = j - (npts - 1); jj <= j + npts; ++jj) {
                    if (jj < 0 || jj > ncz) continue;
                    const size_t nnn = (size_t)ii * st + jj;
                    const T X = xmin + ii * dx, Z = zmin + jj * dz;
                    const T dist = sqrt((X - px) * (X - px) + (Z - pz) * (Z - pz));
                    tt[nnn] = t0[n] + dist * p.s[nnn];
                    frozen[nnn] = 1;
                }
            }
        }
    }
}

// One CTA = one source, start to finish (CL = false, 1024 threads), or -- wide grids -- one CLUSTER of K2D_CLUSTER CTAs of 256
// threads per source (CL = true): the nodes of a wavefront are dealt over the cluster's threads, the CTA barrier becomes the
// hardware cluster barrier (barrier.cluster, release / acquire), traveltimes are read through L2.  With one CTA a wavefront
// of a 2001 x 2001 grid costs 4.6 us -- every node of a diagonal lies in a different 128-byte line, and ~22000 sector requests
// per diagonal go through ONE SM's L1 -- with eight SMs sharing them it is bound by the L2 round trip and the barrier.
constexpr int K2D_CLUSTER = 8;

template <typename T, bool CL>
__global__ void __launch_bounds__(1024, 1) k2d_solve(P2<T> p, int first_source) {
    namespace cg = cooperative_groups;
    __shared__ double red[32];
    __shared__ double cred[K2D_CLUSTER];   // (rank 0's copy collects the CTAs' sums)
    __shared__ int go;
    __shared__ int fbox[4];                // bounding box of the nodes initFSM may freeze: the frozen bytes are only read inside
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = CL ? (int)cluster.block_rank() : 0, ncta = CL ? (int)cluster.num_blocks() : 1;   // (<= K2D_CLUSTER)
    const int b = (int)blockIdx.x / ncta, src = first_source + b;
    const size_t N = (size_t)(p.ncx + 1) * (p.ncz + 1);
    T* tt = p.tt + (size_t)b * N;
    unsigned char* frozen = p.frozen + (size_t)b * N;
    const T* tx = p.tx + (size_t)src * p.ntx_max * 2;
    const T* t0 = p.t0 + (size_t)src * p.ntx_max;
    const int ntx = p.ntx[src];
    const int ltid = threadIdx.x;
    const int tid = rank * blockDim.x + ltid, nth = ncta * blockDim.x;   // the thread's place among all threads of the source
    const int ncx = p.ncx, ncz = p.ncz;
    auto sync_all = [&]() {
        if (CL) cluster.sync(); else __syncthreads();
    };

    if (ltid == 0) {   // checkPts (Grid2Drn.h:333-342); every CTA of a cluster comes to the same verdict
        const T xmax = p.xmin + ncx * p.dx, zmax = p.zmin + ncz * p.dz;
        int bad = 0;
        for (int n = 0; n < ntx; ++n)
            if (tx[2 * n] < p.xmin || tx[2 * n] > xmax || tx[2 * n + 1] < p.zmin || tx[2 * n + 1] > zmax) bad = 1;
        if (rank == 0) p.err[src] = bad;
        go = !bad;
        // initFSM freezes the (2 npts + 1)^2 nodes around an on-node Tx point and the (2 npts)^2 around the cell of an off-node
        // one (Grid2Drn.h:1360-1419); one node of margin for the rounding of the cell index
        const int np = p.weno ? 2 : 1;
        int ilo = 1 << 30, ihi = -1, jlo = 1 << 30, jhi = -1;
        for (int n = 0; n < ntx; ++n) {
            const int ci = (int)floor(((double)tx[2 * n] - (double)p.xmin) / (double)p.dx), cj = (int)floor(((double)tx[2 * n + 1] - (double)p.zmin) / (double)p.dz);
            ilo = min(ilo, ci - np - 1); ihi = max(ihi, ci + np + 2);
            jlo = min(jlo, cj - np - 1); jhi = max(jhi, cj + np + 2);
        }
        fbox[0] = ilo; fbox[1] = ihi; fbox[2] = jlo; fbox[3] = jhi;
    }
    __syncthreads();
    if (!go) return;
    for (size_t n = tid; n < N; n += nth) { tt[n] = Lim<T>::max(); frozen[n] = 0; }
    sync_all();
    if (tid == 0) init_2d(p, tt, frozen, tx, t0, ntx, p.weno ? 2 : 1);
    sync_all();

    const bool square = p.dx == p.dz;
    const int f_ilo = fbox[0], f_ihi = fbox[1], f_jlo = fbox[2], f_jhi = fbox[3];
    auto is_frozen = [&](int i, int j) -> bool {   // (no load, and no second round trip to L2 before the update's loads, outside the box)
        if (i < f_ilo || i > f_ihi || j < f_jlo || j > f_jhi) return false;
        const size_t n = (size_t)i * (ncz + 1) + j;
        return CL ? __ldcg(frozen + n) != 0 : frozen[n] != 0;
    };
    auto sweeps = [&](int kind) -> double {   // the four passes (i up, j up), (i down, j up), (i down, j down), (i up, j down)
        double acc = 0.0;
        for (int d = 0; d < 4; ++d) {
            const bool iu = d == 0 || d == 3, ju = d < 2;
            if (kind == 1) {
                for (int ii = 0; ii <= ncx; ++ii) {
                    const int i = iu ? ii : ncx - ii;
                    for (int j = tid; j <= ncz; j += nth)
                        if (!is_frozen(i, j)) acc += (double)update_2d<CL>(p, tt, i, j, 1);
                    sync_all();
                }
            } else {
                for (int ds = 0; ds <= ncx + ncz; ++ds) {
                    const int lo = max(0, ds - ncz), hi = min(ncx, ds);
                    for (int ii = lo + tid; ii <= hi; ii += nth) {
                        const int jj = ds - ii;
                        const int i = iu ? ii : ncx - ii, j = ju ? jj : ncz - jj;
                        if (!is_frozen(i, j)) acc += (double)update_2d<CL>(p, tt, i, j, kind);
                    }
                    sync_all();
                }
            }
        }
        return acc;
    };
    auto converged = [&](double acc) -> bool {   // L1 change of the iteration = sum of the decreases (tt only decreases)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((ltid & 31) == 0) red[ltid >> 5] = acc;
        __syncthreads();
        if (ltid == 0) {
            double c = 0.0;
            for (int w = 0; w < ((int)blockDim.x + 31) / 32; ++w) c += red[w];
            if (CL) *cluster.map_shared_rank(&cred[rank], 0) = c;   // fixed order: warps of a CTA, then CTAs by rank
            else cred[0] = c;
        }
        sync_all();
        if (ltid == 0) {
            double c = 0.0;
            for (int r = 0; r < ncta; ++r) c += CL ? *cluster.map_shared_rank(&cred[r], 0) : cred[r];
            const T change = c > (double)Lim<T>::max() ? Lim<T>::max() : (T)c;
            go = change >= p.eps_total;
        }
        __syncthreads();
        const bool r = !go;
        sync_all();   // (rank 0's sums have been read by everybody before anybody starts the next iteration)
        return r;
    };
    int niter = 0, niterw = 0;
    if (p.weno) {
        for (bool done = false; !done && niter < p.maxit;) { done = converged(sweeps(square ? 0 : 2)); ++niter; }
        for (bool done = false; !done && niterw < p.maxit;) { done = converged(sweeps(square ? 3 : 4)); ++niterw; }
    } else {
        for (bool done = false; !done && niter < p.maxit;) {
            double a = sweeps(square ? 0 : 2);
            if (square && p.rotated) a += sweeps(1);
            done = converged(a);
            ++niter;
        }
    }
    if (tid == 0) { p.niter[2 * src] = niter; p.niter[2 * src + 1] = niterw; }
}

// Grid2Drcfs::setSlowness (Grid2Drcfs.h:99-138): node slowness = mean of the 1 / 2 / 4 adjacent cells, additions in the source's order
template <typename T>
__global__ void k2d_cell_to_node(const T* __restrict__ s, int nx, int nz, T* __restrict__ sn) {
    const size_t st = (size_t)nz + 1, N = (size_t)(nx + 1) * st;
    for (size_t n = blockIdx.x * (size_t)blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
        const size_t i = n / st, j = n % st;
        const bool i0 = i == 0, i1 = i == (size_t)nx, j0 = j == 0, j1 = j == (size_t)nz;
        T v;
        if ((i0 || i1) && (j0 || j1)) v = s[(i0 ? 0 : nx - 1) * (size_t)nz + (j0 ? 0 : nz - 1)];
        else if (i0 || i1) { const size_t r = (i0 ? 0 : nx - 1) * (size_t)nz; v = T(0.5) * (s[r + j] + s[r + j - 1]); }
        else if (j0) v = T(0.5) * (s[i * nz] + s[(i - 1) * nz]);
        else if (j1) v = T(0.5) * (s[(i + 1) * nz - 1] + s[i * nz - 1]);
        else v = T(0.25) * (s[i * nz + j] + s[i * nz + j - 1] + s[(i - 1) * nz + j] + s[(i - 1) * nz + j - 1]);
        sn[n] = v;
    }
}

// Grid2Drn::getTraveltime (Grid2Drn.h:359-415): bilinear, x first then z, on-node / on-edge cases at tolerance 1e-4
template <typename T>
__global__ void k2d_interp(P2<T> p, const T* __restrict__ tt, const T* __restrict__ rx, int nrx, T* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrx) return;
    const double small = 1.e-4;
    const size_t nnz = (size_t)p.ncz + 1;
    const T px = rx[2 * r], pz = rx[2 * r + 1], xmin = p.xmin, zmin = p.zmin, dx = p.dx, dz = p.dz;
    const size_t i = (unsigned)(small + (double)((px - xmin) / dx)), j = (unsigned)(small + (double)((pz - zmin) / dz));
    const bool onx = fabs((double)(px - (xmin + i * dx))) < small, onz = fabs((double)(pz - (zmin + j * dz))) < small;
    T t;
    if (onx && onz) {
        t = tt[i * nnz + j];
    } else if (onx) {
        const T t1 = tt[i * nnz + j], t2 = tt[i * nnz + j + 1];
        const T w1 = (zmin + (j + 1) * dz - pz) / dz, w2 = (pz - (zmin + j * dz)) / dz;
        t = t1 * w1 + t2 * w2;
    } else if (onz) {
        const T t1 = tt[i * nnz + j], t2 = tt[(i + 1) * nnz + j];
        const T w1 = (xmin + (i + 1) * dx - px) / dx, w2 = (px - (xmin + i * dx)) / dx;
        t = t1 * w1 + t2 * w2;
    } else {
        T t1 = tt[i * nnz + j], t2 = tt[(i + 1) * nnz + j];
        const T t3 = tt[i * nnz + j + 1], t4 = tt[(i + 1) * nnz + j + 1];
        T w1 = (xmin + (i + 1) * dx - px) / dx, w2 = (px - (xmin + i * dx)) / dx;
        t1 = t1 * w1 + t2 * w2;
        t2 = t3 * w1 + t4 * w2;
        w1 = (zmin + (j + 1) * dz - pz) / dz;
        w2 = (pz - (zmin + j * dz)) / dz;
        t = t1 * w1 + t2 * w2;
    }
    out[r] = t;
}

}  // namespace ttcrb200
