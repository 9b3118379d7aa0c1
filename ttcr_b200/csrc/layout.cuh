// Device data layout of the B200 FSM solver.
//
// The reference stores nodes as an array of structs, x fastest (ttcr/Grid3Drn.h:2823,
// ttcr/Node3Dn.h:36-177).  A Gauss-Seidel sweep in direction (si,sj,sk) may update in
// parallel any set of nodes that are mutually unordered in the dependency DAG
// (i,j,k) -> (i+si,j,k), (i,j+sj,k), (i,j,k+sk).  For a warp to touch one 128-byte line per
// access, 32 such nodes must also be CONTIGUOUS in memory.  Nodes along the direction
// (0,-1,+1) are unordered whenever sj == sk, and nodes along (0,+1,+1) whenever sj != sk.
// Hence two sheared ("diagonal-major") layouts, both with k (z) as the contiguous lane axis:
//
//   L1[i][q][k],  q = j + k              used by sweeps with sj == sk   (1,2,7,8)
//   L2[i][r][k],  r = j - k + (nk-1)     used by sweeps with sj != sk   (3,4,5,6)
//
// Each has Q = nj + nk - 1 rows per i and rows padded to KPAD = roundup(nk, 32) lanes; every
// i-plane additionally carries GUARD never-written rows before and after its Q rows, so that a
// marching kernel can prefetch a few rows past either end without bounds checks.
// Slots that correspond to no node (guard rows, j out of range, k >= nk) hold +MAX forever
// (0 in the slowness arrays), which makes the reference's one-sided face stencils
// (Grid3Drn.h:2906-2934) fall out of a plain min().  With the reference's sweep order (+++,-++,+-+,--+,++-,-+-,+--,---; :2819-2898) the
// layout changes only twice per iteration (after sweep 2 and after sweep 6).
//
// A sweep is expressed in ORIENTED coordinates (u, m, v): u along i, m along the row axis,
// v along the lane axis, each running in the sweep's own direction, so that every sweep
// has the same dependency pattern
//     new values : (u-1, m, v)   (u, m-1, v)   (u, m-1, v-1)
//     old values : (u+1, m, v)   (u, m+1, v)   (u, m+1, v+1)
// and differs only by the affine map SweepView below.
#pragma once
#include <cstddef>
#include <cstdint>

namespace ttcrb200 {

constexpr int GUARD = 48;   // guard rows on each side of an i-plane (>= NW*R + 2*D + 2 of the tile kernel)

struct Dims {
    int ni, nj, nk;   // node counts along x, y, z
    int kpad;         // lanes per row (multiple of 32)
    int q;            // node rows per i: nj + nk - 1
    int qs;           // allocated rows per i: q + 2*GUARD
    __host__ __device__ size_t rows() const { return (size_t)ni * qs; }
    __host__ __device__ size_t elems() const { return rows() * kpad; }
    __host__ __device__ size_t nodes() const { return (size_t)ni * nj * nk; }
    // element offset of node row r (0 <= r < q) of plane i
    __host__ __device__ size_t row(int i, int r) const { return ((size_t)i * qs + GUARD + r) * kpad; }
    __host__ __device__ size_t l1(int i, int j, int k) const { return row(i, j + k) + k; }
    __host__ __device__ size_t l2(int i, int j, int k) const { return row(i, j - k + nk - 1) + k; }
    __host__ __device__ size_t at(int layout, int i, int j, int k) const { return layout ? l2(i, j, k) : l1(i, j, k); }
};

inline Dims make_dims(int ni, int nj, int nk) {
    Dims d;
    d.ni = ni; d.nj = nj; d.nk = nk;
    d.kpad = (nk + 31) / 32 * 32;
    d.q = nj + nk - 1;
    d.qs = d.q + 2 * GUARD;
    return d;
}

// Affine map oriented (u,m,v) -> element offset, plus validity of a slot.
struct SweepView {
    long long base;   // offset of (0,0,0)
    long long su, sm; // element strides along u and m (signed)
    int sv;           // +1 / -1
    int nu, nm;       // extents: ni, q
    int vlo, vhi;     // valid lanes: vlo <= v < vhi  (vhi - vlo == nk)
    int joff, nj;     // oriented j = m - v + joff must be in [0, nj)
    int layout;       // 0 = L1, 1 = L2
    int ri, rj, rk;   // 1 if the sweep runs the axis downwards
};

// direction d = 0..7 in the reference's order: bit0 = i reversed, bit1 = j reversed, bit2 = k reversed
inline SweepView make_view(const Dims& d, int dir) {
    SweepView w;
    w.ri = dir & 1; w.rj = (dir >> 1) & 1; w.rk = (dir >> 2) & 1;
    w.layout = (w.rj == w.rk) ? 0 : 1;
    const long long rowlen = d.kpad, plane = (long long)d.qs * d.kpad;
    w.su = w.ri ? -plane : plane;
    w.sm = w.rj ? -rowlen : rowlen;
    w.sv = w.rk ? -1 : 1;
    w.base = (w.ri ? (long long)(d.ni - 1) * plane : 0) + (long long)GUARD * rowlen +
             (w.rj ? (long long)(d.q - 1) * rowlen : 0) + (w.rk ? (long long)(d.kpad - 1) : 0);
    w.nu = d.ni; w.nm = d.q;
    w.vlo = w.rk ? d.kpad - d.nk : 0;
    w.vhi = w.vlo + d.nk;
    w.joff = w.vlo;
    w.nj = d.nj;
    return w;
}

}  // namespace ttcrb200
