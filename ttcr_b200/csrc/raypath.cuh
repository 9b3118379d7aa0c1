// Receiver traveltimes integrated along raypaths (the reference's default, tt_from_rp = 1):
// Grid3Drn::getTraveltimeFromRaypath (ttcr/Grid3Drn.h:1103-1243) with its helpers grad (:1032-1100), getIJK
// (:239-243), getTraveltime (:794-930) and computeSlowness (:2451-2676, processVel == false; Interpolator.h:37-85).
//
// One thread per receiver walks from the receiver against the traveltime gradient, from grid plane to grid plane,
// integrating 0.5 (s1 + s2) |segment| until it is within one cell diagonal of a source point.  The walk is
// sequential by nature and receivers are few, so this kernel is latency bound and small; what matters is that it
// evaluates the reference's expressions operation for operation (T variables, double literals, no contraction:
// the translation unit is compiled with -fmad=false), so that fp64 AND fp32 results are bit-identical to the
// reference's (tests/test_gpu_parity.py::test_tt_from_raypath_*).
//
// Differences from the reference, none of them observable on inputs the reference survives: array indices are
// clamped to the grid (the reference reads out of bounds, and crashes, for receivers on some grid corners), and
// the walk is bounded (status 2) instead of looping forever when the gradient vanishes.
#pragma once
#include "kernels.cuh"

namespace ttcrb200 {

template <typename T> __device__ __forceinline__ T rp_abs(T v) { return v < T(0) ? -v : v; }
__device__ __forceinline__ float rp_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double rp_sqrt(double v) { return sqrt(v); }
template <typename T> __device__ __forceinline__ int rp_sgn(T v) { return v > T(0) ? 1 : (v < T(0) ? -1 : 0); }   // boost::math::sign

// Grid3Drn::getTraveltime at one point (same arithmetic as k_interp)
template <typename T>
__device__ T rp_tt_at(const Geom<T>& g, const Dims& d, const T* __restrict__ tt_l1, T px, T py, T pz) {
    const double small2 = 1.e-8;
    const T dx = g.dx;
    const int i = (int)(unsigned)(small2 + (double)((px - g.xmin) / dx));
    const int j = (int)(unsigned)(small2 + (double)((py - g.ymin) / dx));
    const int k = (int)(unsigned)(small2 + (double)((pz - g.zmin) / dx));
    const bool onx = fabs((double)(px - (g.xmin + T(i) * dx))) < small2;
    const bool ony = fabs((double)(py - (g.ymin + T(j) * dx))) < small2;
    const bool onz = fabs((double)(pz - (g.zmin + T(k) * dx))) < small2;
    auto TT = [&](int a, int b, int c) -> T { return tt_l1[d.l1(min(a, d.ni - 1), min(b, d.nj - 1), min(c, d.nk - 1))]; };
    if (onx && ony && onz) return TT(i, j, k);
    const T wz1 = (g.zmin + T(k + 1) * dx - pz) / dx, wz2 = (pz - (g.zmin + T(k) * dx)) / dx;
    const T wy1 = (g.ymin + T(j + 1) * dx - py) / dx, wy2 = (py - (g.ymin + T(j) * dx)) / dx;
    const T wx1 = (g.xmin + T(i + 1) * dx - px) / dx, wx2 = (px - (g.xmin + T(i) * dx)) / dx;
    if (onx && ony) return TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
    if (onx && onz) return TT(i, j, k) * wy1 + TT(i, j + 1, k) * wy2;
    if (ony && onz) return TT(i, j, k) * wx1 + TT(i + 1, j, k) * wx2;
    if (onx) {
        const T t1 = TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
        const T t2 = TT(i, j + 1, k) * wz1 + TT(i, j + 1, k + 1) * wz2;
        return t1 * wy1 + t2 * wy2;
    }
    if (ony) {
        const T t1 = TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
        const T t2 = TT(i + 1, j, k) * wz1 + TT(i + 1, j, k + 1) * wz2;
        return t1 * wx1 + t2 * wx2;
    }
    if (onz) {
        const T t1 = TT(i, j, k) * wy1 + TT(i, j + 1, k) * wy2;
        const T t2 = TT(i + 1, j, k) * wy1 + TT(i + 1, j + 1, k) * wy2;
        return t1 * wx1 + t2 * wx2;
    }
    T t1 = TT(i, j, k) * wz1 + TT(i, j, k + 1) * wz2;
    T t2 = TT(i, j + 1, k) * wz1 + TT(i, j + 1, k + 1) * wz2;
    const T t3 = TT(i + 1, j, k) * wz1 + TT(i + 1, j, k + 1) * wz2;
    const T t4 = TT(i + 1, j + 1, k) * wz1 + TT(i + 1, j + 1, k + 1) * wz2;
    t1 = t1 * wy1 + t2 * wy2;
    t2 = t3 * wy1 + t4 * wy2;
    return t1 * wx1 + t2 * wx2;
}

// on which node plane of an axis does p lie (tolerance small2), or -1: the reference scans all nodes of the axis
template <typename T>
__device__ __forceinline__ int rp_on_node(T p, T pmin, T dx, int nn) {
    const double small2 = 1.e-8;
    const int c = (int)floor((double)(p - pmin) / (double)dx);
    for (int n = max(c - 1, 0); n <= c + 2 && n < nn; ++n)
        if ((double)rp_abs(p - (pmin + T(n) * dx)) < small2) return n;
    return -1;
}

// Grid3Drn::computeSlowness(pt, true); pv = processVel (interp_vel): the node VELOCITIES are interpolated and the result
// inverted (Grid3Drn.h:2489-2669; 1.0 is a double literal there, so the divisions are done in double)
template <typename T>
__device__ T rp_slow_at(const Geom<T>& g, const Dims& d, const T* __restrict__ s_l1, T px, T py, T pz, bool pv) {
    const double small = 1.e-4;
    const T dx = g.dx;
    const int onX = rp_on_node(px, g.xmin, dx, d.ni), onY = rp_on_node(py, g.ymin, dx, d.nj), onZ = rp_on_node(pz, g.zmin, dx, d.nk);
    auto SN0 = [&](int a, int b, int c) -> T { return s_l1[d.l1(min(a, d.ni - 1), min(b, d.nj - 1), min(c, d.nk - 1))]; };
    auto SN = [&](int a, int b, int c) -> T { const T v = SN0(a, b, c); return pv ? T(1.0 / (double)v) : v; };
    auto RET = [&](T r) -> T { return pv ? T(1.0 / (double)r) : r; };
    if (onX != -1 && onY != -1 && onZ != -1) return SN0(onX, onY, onZ);
    const int i = (int)(unsigned)(small + (double)((px - g.xmin) / dx));
    const int j = (int)(unsigned)(small + (double)((py - g.ymin) / dx));
    const int k = (int)(unsigned)(small + (double)((pz - g.zmin) / dx));
    T x[3], y[3], z[3], s[8];
    if (onX != -1 && onY != -1) {
        s[0] = SN(onX, onY, k); s[1] = SN(onX, onY, k + 1);
        x[0] = pz; x[1] = g.zmin + T(k) * dx; x[2] = g.zmin + T(k + 1) * dx;
        return RET((s[0] * (x[2] - x[0]) + s[1] * (x[0] - x[1])) / (x[2] - x[1]));
    }
    if (onX != -1 && onZ != -1) {
        s[0] = SN(onX, j, onZ); s[1] = SN(onX, j + 1, onZ);
        x[0] = py; x[1] = g.ymin + T(j) * dx; x[2] = g.ymin + T(j + 1) * dx;
        return RET((s[0] * (x[2] - x[0]) + s[1] * (x[0] - x[1])) / (x[2] - x[1]));
    }
    if (onY != -1 && onZ != -1) {
        s[0] = SN(i, onY, onZ); s[1] = SN(i + 1, onY, onZ);
        x[0] = px; x[1] = g.xmin + T(i) * dx; x[2] = g.xmin + T(i + 1) * dx;
        return RET((s[0] * (x[2] - x[0]) + s[1] * (x[0] - x[1])) / (x[2] - x[1]));
    }
    if (onX != -1 || onY != -1 || onZ != -1) {
        if (onX != -1) {
            s[0] = SN(onX, j, k); s[1] = SN(onX, j, k + 1); s[2] = SN(onX, j + 1, k); s[3] = SN(onX, j + 1, k + 1);
            x[0] = py; y[0] = pz; x[1] = g.ymin + T(j) * dx; y[1] = g.zmin + T(k) * dx; x[2] = g.ymin + T(j + 1) * dx; y[2] = g.zmin + T(k + 1) * dx;
        } else if (onY != -1) {
            s[0] = SN(i, onY, k); s[1] = SN(i, onY, k + 1); s[2] = SN(i + 1, onY, k); s[3] = SN(i + 1, onY, k + 1);
            x[0] = px; y[0] = pz; x[1] = g.xmin + T(i) * dx; y[1] = g.zmin + T(k) * dx; x[2] = g.xmin + T(i + 1) * dx; y[2] = g.zmin + T(k + 1) * dx;
        } else {
            s[0] = SN(i, j, onZ); s[1] = SN(i, j + 1, onZ); s[2] = SN(i + 1, j, onZ); s[3] = SN(i + 1, j + 1, onZ);
            x[0] = px; y[0] = py; x[1] = g.xmin + T(i) * dx; y[1] = g.ymin + T(j) * dx; x[2] = g.xmin + T(i + 1) * dx; y[2] = g.ymin + T(j + 1) * dx;
        }
        return RET((s[0] * (x[2] - x[0]) * (y[2] - y[0]) + s[1] * (x[2] - x[0]) * (y[0] - y[1]) + s[2] * (x[0] - x[1]) * (y[2] - y[0]) +
                    s[3] * (x[0] - x[1]) * (y[0] - y[1])) /
                   ((x[2] - x[1]) * (y[2] - y[1])));
    }
    s[0] = SN(i, j, k); s[1] = SN(i, j, k + 1); s[2] = SN(i, j + 1, k); s[3] = SN(i, j + 1, k + 1);
    s[4] = SN(i + 1, j, k); s[5] = SN(i + 1, j, k + 1); s[6] = SN(i + 1, j + 1, k); s[7] = SN(i + 1, j + 1, k + 1);
    x[0] = px; y[0] = py; z[0] = pz;
    x[1] = g.xmin + T(i) * dx; y[1] = g.ymin + T(j) * dx; z[1] = g.zmin + T(k) * dx;
    x[2] = g.xmin + T(i + 1) * dx; y[2] = g.ymin + T(j + 1) * dx; z[2] = g.zmin + T(k + 1) * dx;
    return RET((s[0] * (x[2] - x[0]) * (y[2] - y[0]) * (z[2] - z[0]) + s[1] * (x[2] - x[0]) * (y[2] - y[0]) * (z[0] - z[1]) +
                s[2] * (x[2] - x[0]) * (y[0] - y[1]) * (z[2] - z[0]) + s[3] * (x[2] - x[0]) * (y[0] - y[1]) * (z[0] - z[1]) +
                s[4] * (x[0] - x[1]) * (y[2] - y[0]) * (z[2] - z[0]) + s[5] * (x[0] - x[1]) * (y[2] - y[0]) * (z[0] - z[1]) +
                s[6] * (x[0] - x[1]) * (y[0] - y[1]) * (z[2] - z[0]) + s[7] * (x[0] - x[1]) * (y[0] - y[1]) * (z[0] - z[1])) /
               ((x[2] - x[1]) * (y[2] - y[1]) * (z[2] - z[1])));
}

// one axis of grad(): stencil points p1..p4 (first = p - off), shifted inwards at the grid faces
template <typename T>
__device__ __forceinline__ void rp_stencil(T p, T off, T d, T lo, T hi, T q[4]) {
    T p1 = p - off;
    T p2 = p1 + 0.5 * d, p3 = p1 + 1.5 * d, p4 = p1 + 2.0 * d;
    if (p1 <= lo) {
        p1 = lo; p2 = p1 + 0.5 * d; p3 = p1 + 1.5 * d; p4 = p1 + 2.0 * d;
    } else if (p4 >= hi) {
        p4 = hi; p3 = p4 - 0.5 * d; p2 = p4 - 1.5 * d; p1 = p4 - 2.0 * d;
    }
    q[0] = p1; q[1] = p2; q[2] = p3; q[3] = p4;
}

// the move to the next grid plane along g, from (cx, cy, cz); shared by the walk and by the last leg towards Tx
template <typename T>
__device__ __forceinline__ void rp_advance(const Geom<T>& g, T gx, T gy, T gz, T& cx, T& cy, T& cz) {
    const double small2 = 1.e-8;
    const T dx = g.dx;
    const long long i = (long long)(small2 + (double)((cx - g.xmin) / dx));
    const long long j = (long long)(small2 + (double)((cy - g.ymin) / dx));
    const long long k = (long long)(small2 + (double)((cz - g.zmin) / dx));
    T xp = g.xmin + dx * ((double)i + (rp_sgn(gx) > 0.0 ? 1.0 : 0.0));
    T yp = g.ymin + dx * ((double)j + (rp_sgn(gy) > 0.0 ? 1.0 : 0.0));
    T zp = g.zmin + dx * ((double)k + (rp_sgn(gz) > 0.0 ? 1.0 : 0.0));
    if ((double)rp_abs(xp - cx) < small2) xp += dx * T(rp_sgn(gx));
    if ((double)rp_abs(yp - cy) < small2) yp += dx * T(rp_sgn(gy));
    if ((double)rp_abs(zp - cz) < small2) zp += dx * T(rp_sgn(gz));
    const T ax = gx != T(0) ? (xp - cx) / gx : Lim<T>::max();
    const T ay = gy != T(0) ? (yp - cy) / gy : Lim<T>::max();
    const T az = gz != T(0) ? (zp - cz) / gz : Lim<T>::max();
    if (ax < ay && ax < az) {
        cx += ax * gx; cy += ax * gy; cz += ax * gz; cx = xp;
    } else if (ay < az) {
        cx += ay * gx; cy += ay * gy; cz += ay * gz; cy = yp;
    } else {
        cx += az * gx; cy += az * gy; cz += az * gz; cz = zp;
    }
}

template <typename T>
__device__ __forceinline__ T rp_dist(T ax, T ay, T az, T bx, T by, T bz) {
    return rp_sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by) + (az - bz) * (az - bz));
}

// The eight terms one ray segment adds to the matrix M (Grid3Drn::getRaypath with m_data, Grid3Drn.h:1590-1626, :1676-1709,
// :1718-1751, :1760-1793): -s^2 ds times the trilinear weight of each node of the cell around the segment's mid point, with
// the reference's arithmetic -- T variables, the "1. -" and the product of the three factors in double, node positions
// iv*dx WITHOUT the grid's origin, column index (kv nny + jv) nnx + iv.  Raw terms: the reference merges the terms of equal
// column in order of appearance; the host does that (ttcr_b200/rgrid.py).
template <typename T>
__device__ void rp_m_terms(const Geom<T>& g, const Dims& d, const T* __restrict__ s_l1, T ax, T ay, T az, T bx, T by, T bz, bool interp_vel,
                           unsigned long long* __restrict__ node, T* __restrict__ val) {
    const T dx = g.dx;
    const T mx = T(0.5) * (ax + bx), my = T(0.5) * (ay + by), mz = T(0.5) * (az + bz);   // mid_pt = 0.5 * (a + b)
    T s = rp_slow_at(g, d, s_l1, mx, my, mz, interp_vel);
    s *= s;
    const T ds = rp_dist(ax, ay, az, bx, by, bz);
    const unsigned long long ix = (unsigned long long)((mx - g.xmin) / dx), iy = (unsigned long long)((my - g.ymin) / dx),
                             iz = (unsigned long long)((mz - g.zmin) / dx);
    const unsigned long long nnx = (unsigned long long)d.ni, nny = (unsigned long long)d.nj;
    int e = 0;
    for (unsigned long long ii = 0; ii < 2; ++ii)
        for (unsigned long long jj = 0; jj < 2; ++jj)
            for (unsigned long long kk = 0; kk < 2; ++kk) {
                const unsigned long long iv = ix + ii, jv = iy + jj, kv = iz + kk;
                const T dvdv = T((1. - (double)(rp_abs(mx - T(iv) * dx) / dx)) * (1. - (double)(rp_abs(my - T(jv) * dx) / dx)) *
                                 (1. - (double)(rp_abs(mz - T(kv) * dx) / dx)));
                node[e] = (kv * nny + jv) * nnx + iv;
                val[e] = -s * ds * dvdv;
                ++e;
            }
}

// status: 0 ok, 1 the ray left the grid (the reference throws), 2 it did not reach a source
template <typename T>
__global__ void k_tt_from_rp(Geom<T> g, Dims d, const T* __restrict__ tt_l1, const T* __restrict__ s_l1, const T* __restrict__ tx,
                             const T* __restrict__ t0, int ntx, const T* __restrict__ rx, int nrx, T* __restrict__ out,
                             T* __restrict__ status, int* __restrict__ ray_n, const unsigned long long* __restrict__ ray_off,
                             T* __restrict__ ray_xyz, bool interp_vel, int m_mode = 0, unsigned long long* __restrict__ m_node = nullptr,
                             T* __restrict__ m_val = nullptr) {
    // m_mode: the reference's two getRaypath overloads with m_data differ (both are reproduced as they are):
    //   1 = m_data only (Grid3Drn.h:1500-1801): a walk step sets prev_pt = curr_pt BEFORE its M terms, so the terms of the
    //       steps carry ds = 0 (their columns still enter the matrix, with value -0); the last legs are right
    //   2 = r_data and m_data (:2144-2448): the steps are right (prev_pt = r_data.back() before the push); on the leg to the
    //       plane in front of Tx the point is pushed first, so THAT leg carries ds = 0
    //   both return tt = 0 (not t0) for a receiver that coincides with a Tx point
    // ray_n != nullptr: the points of the raypath are wanted too (Grid3Drn::getRaypath, Grid3Drn.h:1339-1500: the same walk
    // with r_data.push_back).  First launch with ray_xyz == nullptr counts them, the second one (ray_off = exclusive prefix
    // sum of the counts) stores them.
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrx) return;
    int npt = 0;
    T* const rp = ray_xyz ? ray_xyz + 3 * ray_off[r] : nullptr;
    auto push = [&](T x, T y, T z) {
        if (rp) { rp[3 * npt] = x; rp[3 * npt + 1] = y; rp[3 * npt + 2] = z; }
        ++npt;
    };
    // m_node != nullptr (store pass only): the 8 raw M terms of the segment that ENDS at the point about to be pushed go to
    // slot 8 * (ray_off[r] + its index); every point but the receiver closes exactly one segment
    auto mterms = [&](T ax, T ay, T az, T bx, T by, T bz) {
        if (m_mode && m_node && rp) rp_m_terms(g, d, s_l1, ax, ay, az, bx, by, bz, interp_vel, m_node + 8 * (ray_off[r] + npt), m_val + 8 * (ray_off[r] + npt));
    };
    const T dx = g.dx;
    const T k1 = 1. / 24., k2 = 9. / 8.;
    const T maxDist = rp_sqrt(dx * dx + dx * dx + dx * dx);
    const T Rx = rx[3 * r], Ry = rx[3 * r + 1], Rz = rx[3 * r + 2];
    push(Rx, Ry, Rz);
    for (int ns = 0; ns < ntx; ++ns)
        if (Rx == tx[3 * ns] && Ry == tx[3 * ns + 1] && Rz == tx[3 * ns + 2]) {
            out[r] = m_mode ? T(0) : t0[ns];
            status[r] = T(0);
            if (ray_n) ray_n[r] = npt;
            return;
        }
    T ttr = 0.0;
    T px = Rx, py = Ry, pz = Rz;   // prev_pt
    T cx = Rx, cy = Ry, cz = Rz;   // curr_pt
    T s1 = rp_slow_at(g, d, s_l1, cx, cy, cz, interp_vel), s2;
    bool reached = false;
    const long long guard_max = 16ll * (g.ncx + g.ncy + g.ncz) + 1024;
    for (long long it = 0; !reached; ++it) {
        if (it >= guard_max) { out[r] = ttr; status[r] = T(2); if (ray_n) ray_n[r] = npt; return; }
        T q[4], gx, gy, gz;
        rp_stencil(cx, dx, dx, g.xmin, g.xmax, q);   // x: first point at pt.x - dx (sic, Grid3Drn.h:1041)
        gx = (k1 * rp_tt_at(g, d, tt_l1, q[0], cy, cz) - k2 * rp_tt_at(g, d, tt_l1, q[1], cy, cz) + k2 * rp_tt_at(g, d, tt_l1, q[2], cy, cz) -
              k1 * rp_tt_at(g, d, tt_l1, q[3], cy, cz)) / dx;
        rp_stencil(cy, T(dx / 2.0), dx, g.ymin, g.ymax, q);
        gy = (k1 * rp_tt_at(g, d, tt_l1, cx, q[0], cz) - k2 * rp_tt_at(g, d, tt_l1, cx, q[1], cz) + k2 * rp_tt_at(g, d, tt_l1, cx, q[2], cz) -
              k1 * rp_tt_at(g, d, tt_l1, cx, q[3], cz)) / dx;
        rp_stencil(cz, T(dx / 2.0), dx, g.zmin, g.zmax, q);
        gz = (k1 * rp_tt_at(g, d, tt_l1, cx, cy, q[0]) - k2 * rp_tt_at(g, d, tt_l1, cx, cy, q[1]) + k2 * rp_tt_at(g, d, tt_l1, cx, cy, q[2]) -
              k1 * rp_tt_at(g, d, tt_l1, cx, cy, q[3])) / dx;
        gx *= T(-1.0); gy *= T(-1.0); gz *= T(-1.0);
        rp_advance(g, gx, gy, gz, cx, cy, cz);
        if (cx < g.xmin || cx > g.xmax || cy < g.ymin || cy > g.ymax || cz < g.zmin || cz > g.zmax) {
            out[r] = ttr;
            status[r] = T(1);
            if (ray_n) ray_n[r] = npt;
            return;
        }
        s2 = rp_slow_at(g, d, s_l1, cx, cy, cz, interp_vel);
        ttr += 0.5 * (s1 + s2) * rp_dist(px, py, pz, cx, cy, cz);
        s1 = s2;
        if (m_mode == 2) mterms(cx, cy, cz, px, py, pz);
        px = cx; py = cy; pz = cz;
        if (m_mode == 1) mterms(cx, cy, cz, px, py, pz);   // (sic: prev_pt = curr_pt already, Grid3Drn.h:1587-1594)
        push(cx, cy, cz);
        // close enough to one of the Tx points?  (the reference does not leave this loop early)
        for (int ns = 0; ns < ntx; ++ns) {
            const T Tx = tx[3 * ns], Ty = tx[3 * ns + 1], Tz = tx[3 * ns + 2];
            const T dist = rp_dist(cx, cy, cz, Tx, Ty, Tz);
            if (dist < maxDist) {
                rp_advance(g, T(Tx - cx), T(Ty - cy), T(Tz - cz), cx, cy, cz);
                if (rp_dist(cx, cy, cz, px, py, pz) > dist || (cx == Tx && cy == Ty && cz == Tz)) {
                    s2 = rp_slow_at(g, d, s_l1, Tx, Ty, Tz, interp_vel);
                    ttr += t0[ns] + 0.5 * (s1 + s2) * rp_dist(px, py, pz, Tx, Ty, Tz);
                    mterms(Tx, Ty, Tz, px, py, pz);
                    push(Tx, Ty, Tz);
                } else {
                    s2 = rp_slow_at(g, d, s_l1, cx, cy, cz, interp_vel);
                    ttr += 0.5 * (s1 + s2) * rp_dist(px, py, pz, cx, cy, cz);
                    if (m_mode == 2) mterms(cx, cy, cz, cx, cy, cz);   // (sic: pushed first, prev_pt = r_data.back(), :2357-2366)
                    else mterms(cx, cy, cz, px, py, pz);
                    push(cx, cy, cz);
                    s1 = s2;
                    s2 = rp_slow_at(g, d, s_l1, Tx, Ty, Tz, interp_vel);
                    ttr += t0[ns] + 0.5 * (s1 + s2) * rp_dist(cx, cy, cz, Tx, Ty, Tz);
                    mterms(Tx, Ty, Tz, cx, cy, cz);
                    push(Tx, Ty, Tz);
                }
                reached = true;
            }
        }
    }
    out[r] = ttr;
    status[r] = T(0);
    if (ray_n) ray_n[r] = npt;
}

}  // namespace ttcrb200
