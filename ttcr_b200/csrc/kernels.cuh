// CUDA kernels of the B200 FSM solver (everything except the persistent tile sweep, which
// lives in sweep_tile.cuh).  sm_100a only.
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"
#include "update.cuh"

namespace ttcrb200 {

// ---------------------------------------------------------------------------------------
// geometry of the grid in the solver's arithmetic type (reference members dx, xmin, ...,
// Grid3Drn.h:68-77; xmax = xmin + nx*dx evaluated in T1)
template <typename T>
struct Geom {
    T dx, xmin, ymin, zmin, xmax, ymax, zmax;
    int ncx, ncy, ncz;
};

// ---------------------------------------------------------------------------------------
// fills
template <typename T>
__global__ void k_fill(T* __restrict__ p, size_t n, T v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// tt = MAX on the valid slots of layout L1 (Node3Dn::reinit, Node3Dn.h:103-105).  One thread per
// slot of L1; padding slots already hold MAX and are left alone.
template <typename T>
__global__ void k_reinit_l1(T* __restrict__ tt, Dims d) {
    const size_t n = d.elems();
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride) {
        const int k = (int)(e % d.kpad);
        const int q = (int)((e / d.kpad) % d.qs) - GUARD;
        const int j = q - k;
        if (q >= 0 && q < d.q && k < d.nk && j >= 0 && j < d.nj) tt[e] = Lim<T>::max();
    }
}

// ---------------------------------------------------------------------------------------
// linear <-> sheared.  `order` 0: x fastest (reference C++), 1: z fastest (numpy C order).
__device__ __forceinline__ size_t lin_index(const Dims& d, int order, int i, int j, int k) {
    return order == 0 ? ((size_t)k * d.nj + j) * d.ni + i : ((size_t)i * d.nj + j) * d.nk + k;
}

// one thread per slot of the destination layout (coalesced stores, gathered loads)
template <typename T>
__global__ void k_import(const T* __restrict__ lin, int order, T* __restrict__ dst, int layout, Dims d, int i0 = 0, int i1 = -1) {
    // planes i0 <= i < i1 of the destination (default: all); a model arriving from the host is imported chunk by chunk
    // behind its own copy (Grid::set_slowness_any)
    const size_t n = (size_t)(i1 < 0 ? d.ni : i1) * d.qs * d.kpad;
    size_t e = (size_t)i0 * d.qs * d.kpad + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride) {
        const int k = (int)(e % d.kpad);
        const size_t row = e / d.kpad;
        const int r = (int)(row % d.qs) - GUARD;
        const int i = (int)(row / d.qs);
        const int j = layout ? r + k - (d.nk - 1) : r - k;
        if (r >= 0 && r < d.q && k < d.nk && j >= 0 && j < d.nj) dst[e] = lin[lin_index(d, order, i, j, k)];
    }
}

// one thread per node of the linear array (coalesced stores)
template <typename T>
__global__ void k_export(const T* __restrict__ src, int layout, T* __restrict__ lin, int order, Dims d) {
    const size_t n = d.nodes();
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride) {
        int i, j, k;
        if (order == 0) {
            i = (int)(e % d.ni); j = (int)((e / d.ni) % d.nj); k = (int)(e / ((size_t)d.ni * d.nj));
        } else {
            k = (int)(e % d.nk); j = (int)((e / d.nk) % d.nj); i = (int)(e / ((size_t)d.nk * d.nj));
        }
        lin[e] = src[d.at(layout, i, j, k)];
    }
}

// L1 <-> L2 through a 32(j) x 32(k) shared-memory tile: both sides move whole row segments.
// grid: (ceil(nk/32), ceil(nj/32), ni), block: (32, 8).
template <typename T>
__global__ void k_relayout(const T* __restrict__ src, int src_layout, T* __restrict__ dst, Dims d) {
    __shared__ T tile[32][33];
    const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32, i = blockIdx.z;
    const int lane = threadIdx.x;
    const int k = k0 + lane;
    // source rows that intersect the tile: row id - row0 = 0 .. 62
    //   L1: row = j + k        -> row0 = j0 + k0,               jl = rr - lane
    //   L2: row = j - k + nk-1 -> row0 = j0 - (k0+31) + nk-1,   jl = rr + lane - 31
    {
        const int row0 = src_layout ? j0 - (k0 + 31) + d.nk - 1 : j0 + k0;
        for (int rr = threadIdx.y; rr < 63; rr += 8) {
            const int jl = src_layout ? rr + lane - 31 : rr - lane;
            const int j = j0 + jl;
            if (jl >= 0 && jl < 32 && j < d.nj && k < d.nk)
                tile[jl][lane] = src[d.row(i, row0 + rr) + k];
        }
    }
    __syncthreads();
    {
        const int dl = 1 - src_layout;
        const int row0 = dl ? j0 - (k0 + 31) + d.nk - 1 : j0 + k0;
        for (int rr = threadIdx.y; rr < 63; rr += 8) {
            const int jl = dl ? rr + lane - 31 : rr - lane;
            const int j = j0 + jl;
            if (jl >= 0 && jl < 32 && j < d.nj && k < d.nk)
                dst[d.row(i, row0 + rr) + k] = tile[jl][lane];
        }
    }
}

// L1 <-> L2 without shared memory: for a fixed lane k the two layouts differ by a shift of nk-1-2k rows, so a
// destination row reads, lane by lane, source rows two apart; a block walks a band of RB destination rows of one
// 32-lane column, and the 128-byte source lines it touches (RB+62 of them) are reused through L1.
// grid: (kpad/32, ceil(Q/RB), ni), block: (32, 8).
template <typename T, int RB>
__global__ void __launch_bounds__(256) k_relayout2(const T* __restrict__ src, int src_layout, T* __restrict__ dst, Dims d) {
    const int k = blockIdx.x * 32 + threadIdx.x;
    const int i = blockIdx.z;
    const int r0 = blockIdx.y * RB;
    if (k >= d.nk) return;
    const int shift = d.nk - 1 - 2 * k;                 // L2 row = L1 row + shift
    const size_t plane = (size_t)i * d.qs + GUARD;
#pragma unroll 8
    for (int y = threadIdx.y; y < RB; y += 8) {
        const int rd = r0 + y;                           // destination row
        if (rd >= d.q) break;
        const int rs = src_layout ? rd + shift : rd - shift;   // source row (src L2 -> dst L1: L2 row = L1 row + shift)
        const int j = src_layout ? rd - k : rd + k - (d.nk - 1);
        if (j >= 0 && j < d.nj) dst[(plane + rd) * d.kpad + k] = src[(plane + rs) * d.kpad + k];
    }
}

// 16-byte fill (n elements, n * sizeof(T) a multiple of 16)
template <typename T>
__global__ void k_fill16(T* __restrict__ p, size_t n, T v) {
    constexpr int V = 16 / sizeof(T);
    struct alignas(16) Pack { T x[V]; };
    Pack pk;
#pragma unroll
    for (int e = 0; e < V; ++e) pk.x[e] = v;
    Pack* const q = reinterpret_cast<Pack*>(p);
    const size_t m = n / V;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) q[i] = pk;
}

// ---------------------------------------------------------------------------------------
// Grid3Drcfs::setSlowness (Grid3Drcfs.h:88-171): node slowness = mean of the adjacent cells.
// Summation order as in the source: k outer, j, i inner; j outer, k inner on x faces.
// Input and output are linear arrays in the same `order`.
template <typename T>
__global__ void k_cell_to_node(const T* __restrict__ sc, T* __restrict__ sn, int order, int ncx, int ncy, int ncz) {
    const int nx1 = ncx + 1, ny1 = ncy + 1, nz1 = ncz + 1;
    const size_t n = (size_t)nx1 * ny1 * nz1;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < n; e += stride) {
        int i, j, k;
        if (order == 0) {
            i = (int)(e % nx1); j = (int)((e / nx1) % ny1); k = (int)(e / ((size_t)nx1 * ny1));
        } else {
            k = (int)(e % nz1); j = (int)((e / nz1) % ny1); i = (int)(e / ((size_t)nz1 * ny1));
        }
        int ks[2], js[2], is[2], nk = 0, nj = 0, ni = 0;
        if (k < ncz) ks[nk++] = k;
        if (k > 0) ks[nk++] = k - 1;
        if (j < ncy) js[nj++] = j;
        if (j > 0) js[nj++] = j - 1;
        if (i < ncx) is[ni++] = i;
        if (i > 0) is[ni++] = i - 1;
        auto cell = [&](int ci, int cj, int ck) -> T {
            return order == 0 ? sc[((size_t)ck * ncy + cj) * ncx + ci] : sc[((size_t)ci * ncy + cj) * ncz + ck];
        };
        T sum = 0;
        bool first = true;
        if (ni == 1 && nj == 2 && nk == 2) {
            for (int b = 0; b < 2; ++b)
                for (int a = 0; a < 2; ++a) {
                    const T v = cell(is[0], js[b], ks[a]);
                    sum = first ? v : sum + v;
                    first = false;
                }
        } else {
            for (int a = 0; a < nk; ++a)
                for (int b = 0; b < nj; ++b)
                    for (int c = 0; c < ni; ++c) {
                        const T v = cell(is[c], js[b], ks[a]);
                        sum = first ? v : sum + v;
                        first = false;
                    }
        }
        const int cnt = nk * nj * ni;
        const T w = cnt == 1 ? T(1) : cnt == 2 ? T(0.5) : cnt == 4 ? T(0.25) : T(0.125);
        sn[e] = w * sum;
    }
}

// ---------------------------------------------------------------------------------------
// initFSM (Grid3Drn.h:3487-3556).  Sequential over Tx points (later points overwrite earlier
// ones, as in the reference); one thread.  Writes tt (layout L1) and the frozen bit masks of
// both layouts.  Arithmetic is T with individually rounded operations (no FMA in this TU).
template <typename T>
__global__ void k_init_fsm(Geom<T> g, Dims d, const T* __restrict__ tx, const T* __restrict__ t0, int ntx, int npts,
                           T* __restrict__ tt_l1, const T* __restrict__ s_l1, uint32_t* __restrict__ m1,
                           uint32_t* __restrict__ m2) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const double small = 1.e-4, small2 = small * small;
    auto freeze = [&](int i, int j, int k, T val) {
        const size_t e1 = d.l1(i, j, k), e2 = d.l2(i, j, k);
        tt_l1[e1] = val;
        m1[e1 >> 5] |= 1u << (e1 & 31);
        m2[e2 >> 5] |= 1u << (e2 & 31);
    };
    // first node index on one axis within `small` of p (Node3Dn.h:147-149); -1 if none
    auto match = [&](T p, T mn, int nc) -> int {
        int c = (int)floor(((double)p - (double)mn) / (double)g.dx + 0.5);
        int lo = c - 2 < 0 ? 0 : c - 2, hi = c + 2 > nc ? nc : c + 2;
        for (int a = lo; a <= hi; ++a) {
            const T x = mn + T(a) * g.dx;
            if (fabs((double)(x - p)) < small) return a;
        }
        return -1;
    };
    for (int n = 0; n < ntx; ++n) {
        const T px = tx[3 * n], py = tx[3 * n + 1], pz = tx[3 * n + 2];
        int i = match(px, g.xmin, g.ncx), j = match(py, g.ymin, g.ncy), k = match(pz, g.zmin, g.ncz);
        int lo;
        if (i >= 0 && j >= 0 && k >= 0) {
            lo = npts;
            freeze(i, j, k, t0[n]);
        } else {
            // getCellNo, Grid3Drn.h:207-215
            const T x = (double)(g.xmax - px) < small2 ? T((double)g.xmax - .5 * (double)g.dx) : px;
            const T y = (double)(g.ymax - py) < small2 ? T((double)g.ymax - .5 * (double)g.dx) : py;
            const T z = (double)(g.zmax - pz) < small2 ? T((double)g.zmax - .5 * (double)g.dx) : pz;
            i = (int)(unsigned)(small2 + (double)((x - g.xmin) / g.dx));
            j = (int)(unsigned)(small2 + (double)((y - g.ymin) / g.dx));
            k = (int)(unsigned)(small2 + (double)((z - g.zmin) / g.dx));
            lo = npts - 1;
        }
        for (int kk = k - lo; kk <= k + npts; ++kk) {
            if (kk < 0 || kk > g.ncz) continue;
            for (int jj = j - lo; jj <= j + npts; ++jj) {
                if (jj < 0 || jj > g.ncy) continue;
                for (int ii = i - lo; ii <= i + npts; ++ii) {
                    if (ii < 0 || ii > g.ncx || (ii == i && jj == j && kk == k)) continue;
                    const T X = g.xmin + T(ii) * g.dx, Y = g.ymin + T(jj) * g.dx, Z = g.zmin + T(kk) * g.dx;
                    const T ex = X - px, ey = Y - py, ez = Z - pz;
                    const T dist = sqrt(ex * ex + ey * ey + ez * ez);   // Node3Dn.h:138-140
                    freeze(ii, jj, kk, t0[n] + dist * s_l1[d.l1(ii, jj, kk)]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Grid3Drn::getTraveltime (Grid3Drn.h:794-930): receiver traveltimes by trilinear interpolation
// with the on-node / edge / face cases at tolerance small2; one thread per receiver.
template <typename T>
__global__ void k_interp(Geom<T> g, Dims d, const T* __restrict__ tt_l1, const T* __restrict__ rx, int nrx,
                         T* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrx) return;
    const double small2 = 1.e-8;
    const T px = rx[3 * r], py = rx[3 * r + 1], pz = rx[3 * r + 2];
    const T dx = g.dx;
    const int i = (int)(unsigned)(small2 + (double)((px - g.xmin) / dx));
    const int j = (int)(unsigned)(small2 + (double)((py - g.ymin) / dx));
    const int k = (int)(unsigned)(small2 + (double)((pz - g.zmin) / dx));
    const bool onx = fabs((double)(px - (g.xmin + T(i) * dx))) < small2;
    const bool ony = fabs((double)(py - (g.ymin + T(j) * dx))) < small2;
    const bool onz = fabs((double)(pz - (g.zmin + T(k) * dx))) < small2;
    auto TT = [&](int a, int b, int c) -> T { return tt_l1[d.l1(a, b, c)]; };
    // weights (Grid3Drn.h:817-818 etc.), evaluated only where the corresponding node exists
    T v;
    if (onx && ony && onz) {
        v = TT(i, j, k);
    } else {
        const T wz1 = (g.zmin + T(k + 1) * dx - pz) / dx, wz2 = (pz - (g.zmin + T(k) * dx)) / dx;
        const T wy1 = (g.ymin + T(j + 1) * dx - py) / dx, wy2 = (py - (g.ymin + T(j) * dx)) / dx;
        const T wx1 = (g.xmin + T(i + 1) * dx - px) / dx, wx2 = (px - (g.xmin + T(i) * dx)) / dx;
        if (onx && ony) {
            v = TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
        } else if (onx && onz) {
            v = TT(i, j, k) * wy1 + TT(i, j + 1, k) * wy2;
        } else if (ony && onz) {
            v = TT(i, j, k) * wx1 + TT(i + 1, j, k) * wx2;
        } else if (onx) {
            const T t1 = TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
            const T t2 = TT(i, j + 1, k) * wz1 + TT(i, j + 1, k + 1) * wz2;
            v = t1 * wy1 + t2 * wy2;
        } else if (ony) {
            const T t1 = TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
            const T t2 = TT(i + 1, j, k) * wz1 + TT(i + 1, j, k + 1) * wz2;
            v = t1 * wx1 + t2 * wx2;
        } else if (onz) {
            const T t1 = TT(i, j, k) * wy1 + TT(i, j + 1, k) * wy2;
            const T t2 = TT(i + 1, j, k) * wy1 + TT(i + 1, j + 1, k) * wy2;
            v = t1 * wx1 + t2 * wx2;
        } else {
            T t1 = TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
            T t2 = TT(i, j + 1, k) * wz1 + TT(i, j + 1, k + 1) * wz2;
            const T t3 = TT(i + 1, j, k) * wz1 + TT(i + 1, j, k + 1) * wz2;
            const T t4 = TT(i + 1, j + 1, k) * wz1 + TT(i + 1, j + 1, k + 1) * wz2;
            t1 = t1 * wy1 + t2 * wy2;
            t2 = t3 * wy1 + t4 * wy2;
            v = t1 * wx1 + t2 * wx2;
        }
    }
    out[r] = v;
}

// ---------------------------------------------------------------------------------------
// Sweep kernel "PLANE": one launch per wavefront p = u + m (all nodes of one diagonal plane
// i'+j'+k' = p in parallel), coalesced through the sheared layout.  This is the design the
// reference's OpenCL path uses (one launch per (direction, level), Grid3Drn_OpenCL.h:858-862,
// kernels Grid3Drn_kernels.cl:111-237 / :280-715) minus its index list and frozen bytes.  It is
// launch-latency bound on large grids; the persistent tile kernel (sweep_tile.cuh) replaces it
// there, and this one remains as the small-grid path and as the in-library cross-check.
//
// block (32, 8): x = lane v, y = u.  grid (kpad/32, ceil(nu_range/8)).
struct FrozenBox {   // conservative bounding box (true i,j,k) of all frozen nodes of the source
    int ilo, ihi, jlo, jhi, klo, khi;
};

// Clears the frozen bits of the previous source: they all lie inside its (conservative) bounding box, so a solve does not
// have to zero the two whole bit masks (2 x 37 MB at 512^3: 0.37 ms of a 12 ms solve).
static __global__ void k_clear_frozen_box(Dims d, FrozenBox fb, uint32_t* __restrict__ m1, uint32_t* __restrict__ m2) {
    const int nx = fb.ihi - fb.ilo + 1, ny = fb.jhi - fb.jlo + 1, nz = fb.khi - fb.klo + 1;
    const long long n = (long long)nx * ny * nz;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int k = fb.klo + (int)(t % nz), j = fb.jlo + (int)((t / nz) % ny), i = fb.ilo + (int)(t / ((long long)nz * ny));
        const size_t e1 = d.l1(i, j, k), e2 = d.l2(i, j, k);
        atomicAnd(&m1[e1 >> 5], ~(1u << (e1 & 31)));
        atomicAnd(&m2[e2 >> 5], ~(1u << (e2 & 31)));
    }
}

// The update of node (u, p - u, v) of wavefront plane p in oriented coordinates, first order (Grid3Drn.h:2902-2959) or
// WENO (:3078-3484); returns the decrease of its traveltime (0 if the slot is no node, frozen or unchanged).
template <typename T, bool WENO>
__device__ __forceinline__ double plane_node(const SweepView& w, const Dims& d, T* __restrict__ tt, const T* __restrict__ slo,
                                             const uint32_t* __restrict__ frozen, const FrozenBox& fb, int p, int u, int u_hi, int v,
                                             T dx) {
    const int m = p - u;
    const int jo = m - v + w.joff;   // oriented j
    const int ko = v - w.vlo;        // oriented k
    double delta = 0.0;
    const bool valid = u <= u_hi && m >= 0 && m < w.nm && ko >= 0 && v < w.vhi && jo >= 0 && jo < w.nj;
    if (valid) {
        const long long e = w.base + (long long)u * w.su + (long long)m * w.sm + (long long)v * w.sv;
        const int it = w.ri ? d.ni - 1 - u : u, jt = w.rj ? d.nj - 1 - jo : jo, kt = w.rk ? d.nk - 1 - ko : ko;
        // The next plane updates (u, m+1, v): everything it reads has been touched by this plane or the ones before,
        // except its farthest downwind row and its slowness.  Pull those into L2 now (guard rows keep the addresses
        // inside the arrays), so that the next plane waits for L2, not for DRAM.
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tt + e + (WENO ? 3 : 2) * w.sm));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(slo + e + w.sm));
        bool frz = false;
        if (it >= fb.ilo && it <= fb.ihi && jt >= fb.jlo && jt <= fb.jhi && kt >= fb.klo && kt <= fb.khi)
            frz = (frozen[e >> 5] >> (e & 31)) & 1u;
        if (!frz) {
            const T MAXV = Lim<T>::max();
            const T told = tt[e];
            T au, aj, ak;
            if (!WENO) {
                const T um = u > 0 ? tt[e - w.su] : MAXV, up = u < w.nu - 1 ? tt[e + w.su] : MAXV;
                const T jm = jo > 0 ? tt[e - w.sm] : MAXV, jp = jo < w.nj - 1 ? tt[e + w.sm] : MAXV;
                const T km = ko > 0 ? tt[e - w.sm - w.sv] : MAXV, kp = ko < d.nk - 1 ? tt[e + w.sm + w.sv] : MAXV;
                au = tmin(um, up); aj = tmin(jm, jp); ak = tmin(km, kp);
            } else {
                // values at oriented offsets -2..+2, then swapped into TRUE axis order
                auto ld = [&](bool ok, long long off) -> T { return ok ? tt[e + off] : T(0); };
                {
                    T a = ld(u > 1, -2 * w.su), b = ld(u > 0, -w.su), c = ld(u < w.nu - 1, w.su), dd = ld(u < w.nu - 2, 2 * w.su);
                    if (w.ri) { T s = a; a = dd; dd = s; s = b; b = c; c = s; }
                    au = axis_weno(a, b, told, c, dd, it, d.ni - 1, dx);
                }
                {
                    T a = ld(jo > 1, -2 * w.sm), b = ld(jo > 0, -w.sm), c = ld(jo < w.nj - 1, w.sm), dd = ld(jo < w.nj - 2, 2 * w.sm);
                    if (w.rj) { T s = a; a = dd; dd = s; s = b; b = c; c = s; }
                    aj = axis_weno(a, b, told, c, dd, jt, d.nj - 1, dx);
                }
                {
                    const long long s1 = w.sm + w.sv;
                    T a = ld(ko > 1, -2 * s1), b = ld(ko > 0, -s1), c = ld(ko < d.nk - 1, s1), dd = ld(ko < d.nk - 2, 2 * s1);
                    if (w.rk) { T s = a; a = dd; dd = s; s = b; b = c; c = s; }
                    ak = axis_weno(a, b, told, c, dd, kt, d.nk - 1, dx);
                }
            }
            const T fh = slo[e] * dx;
            const T t = godunov(ak, aj, au, fh);   // a1 = k axis, a2 = j, a3 = i as in Grid3Drn.h:2906-2934
            if (t < told) {
                tt[e] = t;
                delta = (double)told - (double)t;
            }
        }
    }
    return delta;
}

// sum of a (32, 8) block's per-thread values, warps in order, into *out
__device__ __forceinline__ void block_partial(double v, double* out) {
    __shared__ double sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) sh[threadIdx.y] = v;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += sh[w];
        *out = a;
    }
}

template <typename T, bool WENO>
__global__ void __launch_bounds__(256) k_sweep_plane(SweepView w, Dims d, T* __restrict__ tt, const T* __restrict__ slo,
                                                     const uint32_t* __restrict__ frozen, const FrozenBox* __restrict__ fbp, int p,
                                                     int u_lo, int u_hi, T dx, double* __restrict__ partial) {
    // Programmatic dependent launch (grid.cu launches the planes of a sweep with programmatic stream serialisation):
    // let plane p+1 be scheduled now, and touch memory only when plane p-1 has completed and flushed.  Both are no-ops
    // in a plain launch.  The frozen box sits in device memory so that the launch arguments of a plane do not depend
    // on the source: the planes of a sweep are captured once into a CUDA graph and replayed.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const FrozenBox fb = *fbp;
    const int v = blockIdx.x * 32 + threadIdx.x;
    const int u = u_lo + blockIdx.y * 8 + threadIdx.y;
    double delta = plane_node<T, WENO>(w, d, tt, slo, frozen, fb, p, u, u_hi, v, dx);
    // block reduction of the L1 change (Grid3Drnfs.h:144-150; tt only ever decreases, so the per-sweep decreases telescope to
    // sum |times - tt| of the iteration).  Every block writes its sum to its own slot of `partial` (the plane's slice of the
    // sweep's buffer); k_sum_partials adds the slots in a fixed order, so the iteration count does not depend on the order in
    // which blocks happen to finish (an atomicAdd of doubles would).
    block_partial(delta, partial + (size_t)blockIdx.y * gridDim.x + blockIdx.x);
}

// All wavefront planes of a directional sweep in ONE cooperative launch: the grid walks the planes and meets at a
// grid-wide barrier after each (a kernel boundary costs ~3 us even inside a CUDA graph, the barrier well under 1 us).
// Plane p is cut into the same (32 lanes x 8 planes of u) blocks as k_sweep_plane's grid; CTA c takes blocks c, c + G, ...
// `bar` is a monotone arrival counter owned by the launch (zeroed by the host before it): after plane p every CTA has
// added 1, so the release value is (p + 1) * G.  The fence before the arrival publishes the CTA's stores, the acquire
// load after it (plus the CTA barrier) makes everybody else's visible, L1 included.
template <typename T, bool WENO>
__global__ void __launch_bounds__(256) k_sweep_planes_coop(SweepView w, Dims d, T* __restrict__ tt, const T* __restrict__ slo,
                                                           const uint32_t* __restrict__ frozen, const FrozenBox* __restrict__ fbp,
                                                           T dx, double* __restrict__ partial, unsigned* __restrict__ bar) {
    const FrozenBox fb = *fbp;
    const int np = w.nu + w.nm - 1;
    const int vb = d.kpad / 32;
    const bool leader = threadIdx.x == 0 && threadIdx.y == 0;
    double delta = 0.0;
    for (int p = 0; p < np; ++p) {
        const int u_lo = max(0, p - w.nm + 1), u_hi = min(w.nu - 1, p);
        const int nblk = vb * ((u_hi - u_lo + 1 + 7) / 8);
        for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
            const int by = b / vb, bx = b - by * vb;
            delta += plane_node<T, WENO>(w, d, tt, slo, frozen, fb, p, u_lo + by * 8 + (int)threadIdx.y, u_hi, bx * 32 + (int)threadIdx.x, dx);
        }
        if (p + 1 < np) {
            __syncthreads();
            if (leader) {
                __threadfence();
                atomicAdd(bar, 1u);
                const unsigned want = (unsigned)(p + 1) * gridDim.x;
                unsigned seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
                } while (seen < want);
                __threadfence();
            }
            __syncthreads();
        }
    }
    block_partial(delta, partial + blockIdx.x);   // (one slot per CTA, summed in CTA order by k_sum_partials)
}

}  // namespace ttcrb200
