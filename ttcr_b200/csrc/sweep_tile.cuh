// Sweep kernel "TILE": one persistent launch per directional sweep.
//
// Why: a directional Gauss-Seidel sweep of an N^3 grid has 3N-2 dependent wavefronts; at 512^3 the
// HBM roofline leaves ~0.2 us per wavefront, far below any kernel-launch or grid-barrier latency, so
// "one launch per wavefront" (k_sweep_plane, and the reference's OpenCL path) is launch bound.
//
// How: in oriented coordinates (layout.cuh) the grid is cut into tiles of NW*R rows of u by 32 lanes
// of v.  A CTA owns a tile and MARCHES along m (the sheared row axis).  Compute warp wq, register row
// r handles u = u0 + wq*R + r; lane l handles v = v0 + l; at step s that node row is on
// m = m_first + s - (wq*R + r): consecutive u lag one step, which is exactly the Gauss-Seidel
// dependency.  All 32 lanes of a warp touch one contiguous 128-byte row segment per array per step
// (tt read, slowness read, tt write: the 12 algorithmic bytes per node), streamed through register
// queues D rows deep.  Everything else a node needs comes from registers, shuffles or shared memory:
//     (u, m-1, v)   own previous result             register
//     (u, m-1, v-1) lane l-1's previous result      __shfl_up    (lane 0: halo of tile V-1, from L2)
//     (u, m+1, v)   own next old value              register queue
//     (u, m+1, v+1) lane l+1's next old value       __shfl_down  (lane 31: halo of tile V+1)
//     (u-1, m, v)   register row r-1 (previous step) / warp wq-1 via shared memory
//                                                   (first row of the tile: halo of tile U-1, from L2)
//     (u+1, m, v)   register row r+1's queue / warp wq+1 via shared memory
//                                                   (last row of the tile: halo of tile U+1)
// Tiles depend on their U-1 and V-1 neighbours row by row.  Each tile publishes "rows completed" in a
// global flag every `chunk` rows.  Two helper warps keep all fences off the compute warps:
//   * the POLLER polls the two upstream flags (relaxed loads, one acquire fence per observed advance)
//     and mirrors them into shared memory, which is all the compute warps ever look at;
//   * the PUBLISHER is released by the compute warps through a named barrier (bar.arrive, which does
//     not block them), then executes the release fence and stores the flag.
// Tiles are handed out through an atomic ticket in an order that is a linear extension of the
// dependency order, so a CTA can only wait for tiles that are already running or finished: no
// co-residency requirement, no deadlock.  Every spin is bounded; on timeout the kernel raises an
// abort flag and the host reports an error.
//
// Results are bit-identical to k_sweep_plane and to the lexicographic CPU order: all three execute the
// same dependency DAG with the same arithmetic (update.cuh).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace ttcrb200 {

struct TileOptions {
    int chunk = 8;         // rows between progress-flag publications (rounded down to a power of two)
    int ctas_per_sm = 0;   // 0 = as many as fit
    int warps = 8;         // compute warps per tile: 4 or 8
    int rows = 1;          // u rows per thread (R): 1 or 2
    int depth = 8;         // register queue depth (rows of loads in flight): 4 or 8
    long long spin_limit = 1ll << 22;   // polls before a wait is declared dead
    int max_ctas = 0;      // k_sweep_march: cap on the grid size (0 = none); tests use it to make every CTA run several tiles
    int nodes = 0;         // k_sweep_march: nodes per thread and step: 2 (sweep_march.cuh, tiles of 16 x 32), 4 (sweep_march4.cuh, 16 x 64), 0 = by grid size
};

struct TileState {
    int* d_order = nullptr;      // ticket -> tile id
    int* d_flags = nullptr;      // per tile: rows completed
    int* d_ctrl = nullptr;       // [0] ticket counter, [1] abort flag
    double* d_partial = nullptr; // per tile: sum of (old - new)
    int* h_abort = nullptr;      // pinned copy of the abort flag
    int cap_tiles = 0;
    int order_key = -1;
    int ntiles = 0;
    // mailboxes of k_sweep_tile4 (allocated on first use)
    unsigned long long* d_mbu = nullptr;
    unsigned long long* d_mbv = nullptr;
    int mb_rows = 0, mb_tiles = 0;
    unsigned serial = 0;
};

struct TileParams {
    SweepView w;
    Dims d;
    FrozenBox fb;
    int nU, nV, ntiles;
    int chunk;
    long long spin_limit;
    const int* order;
    int* flags;
    int* ctrl;
    double* partial;
    long long* trace;   // debug (TTCR_B200_TRACE): 8 timestamps per tile, or nullptr
};

// ---- small PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// Streaming accesses of the traveltime field: weak, not allocating in L1 (each value is touched once
// per sweep, and a halo written by another SM must come from L2; the poller's acquire fence
// additionally invalidates L1).
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream(double* p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// keep a loop-invariant value in a register (stops the compiler re-reading the constant bank)
__device__ __forceinline__ int pin(int x) {
    asm volatile("" : "+r"(x));
    return x;
}
template <typename P>
__device__ __forceinline__ P* pin_ptr(P* x) {
    asm volatile("" : "+l"(x));
    return x;
}
// shared-memory accesses through a pinned 32-bit shared address plus a compile-time byte offset
template <int OFF> __device__ __forceinline__ float lds_at(unsigned a, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF> __device__ __forceinline__ double lds_at(unsigned a, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF> __device__ __forceinline__ void sts_at(unsigned a, float v) {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory");
}
template <int OFF> __device__ __forceinline__ void sts_at(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}
// named barriers.  1: compute warps, every step.  2,3: compute warps (arrive) -> publisher (sync).
template <int N>
__device__ __forceinline__ void bar_compute() {
    asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ int bar_compute_or(int pred) {
    int out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.or.pred q, 1, %2, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(out)
        : "r"(pred), "n"(N)
        : "memory");
    return out;
}
template <int N>
__device__ __forceinline__ void bar_pub_arrive(int id) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bar_pub_sync(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
}

// ---- the kernel ---------------------------------------------------------------------------------
// block = NW compute warps + poller warp + publisher warp.
template <typename T, int NW, int R, int D>
__global__ void __launch_bounds__((NW + 2) * 32) k_sweep_tile(TileParams p, T* __restrict__ tt, const T* __restrict__ slo,
                                                             const uint32_t* __restrict__ frozen, T dx) {
    static_assert(D >= 3, "the exchange protocol publishes old values two rows ahead");
    static_assert(NW * R + 2 * D + 2 <= GUARD, "guard rows too few for this tile shape");
    constexpr int NU = NW * R;          // u rows per tile
    constexpr int NC = NW * 32;         // compute threads
    constexpr int NP = (NW + 1) * 32;   // compute threads + publisher warp
    __shared__ T xnew[2][NW][32];   // result of the tile's last register row of each warp, previous step
    __shared__ T xold[2][NW][32];   // old values two rows ahead of each warp's first register row
    __shared__ double sred[NW];
    __shared__ int sm_tile;
    __shared__ volatile int sm_known_u, sm_known_v, sm_pubseq, sm_fin, sm_abort;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const T MAXV = Lim<T>::max();

    for (;;) {
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sm_tile = (ab || t >= p.ntiles) ? -1 : t;
            sm_known_u = 0; sm_known_v = 0; sm_pubseq = 0; sm_fin = 0; sm_abort = 0;
        }
        __syncthreads();
        const int ticket = sm_tile;
        if (ticket < 0) break;
        const int tile = p.order[ticket];
        const int U = tile / p.nV, V = tile - U * p.nV;
        const int u0 = U * NU, v0 = V * 32;
        const int va = max(v0, w.vlo), vb = min(v0 + 32, w.vhi);
        const int m_first = va - w.joff;
        const int nrows = (vb - va) + w.nj - 1;
        const int nsteps = nrows + NU - 1;
        const bool has_u = U > 0, has_v = V > 0;
        // tile (U, V-1): its first row and row count
        const int va_p = max(v0 - 32, w.vlo);
        const int nrows_p = (v0 - va_p) + w.nj - 1;
        const int dmf = m_first - (va_p - w.joff);
        const int chunk = p.chunk;                       // power of two
        const int nch = (nrows - 1) / chunk;             // in-loop publications (rows chunk, 2*chunk, ... < nrows)
        if (p.trace && threadIdx.x == 0) {
            p.trace[tile * 8 + 0] = gtime();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[tile * 8 + 7] = smid;
        }

        if (warp == NW) {
            // ================= poller =================
            if (lane == 0) {
                int ku = has_u ? 0 : nrows, kv = has_v ? 0 : nrows_p;
                long long spins = 0;
                const int* fu = &p.flags[has_u ? tile - p.nV : tile];
                const int* fv = &p.flags[has_v ? tile - 1 : tile];
                while (ku < nrows || kv < nrows_p) {
                    bool progress = false;
                    if (ku < nrows) {
                        const int k2 = ld_relaxed_gpu(fu);
                        if (k2 > ku) { fence_acq_rel_gpu(); ku = k2; sm_known_u = ku; progress = true; }
                    }
                    if (kv < nrows_p) {
                        const int k2 = ld_relaxed_gpu(fv);
                        if (k2 > kv) { fence_acq_rel_gpu(); kv = k2; sm_known_v = kv; progress = true; }
                    }
                    if (sm_fin || sm_abort) break;
                    if (progress) {
                        spins = 0;
                    } else {
                        __nanosleep(20);
                        if (++spins > p.spin_limit) {
                            atomicExch(&p.ctrl[1], 1);
                            sm_abort = 1;
                            break;
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == NW + 1) {
            // ================= publisher =================
            // chunk c (0-based) is complete when the compute warps have arrived on barrier 2+(c&1);
            // the last arrival (c == nch) announces the whole tile.
            for (int c = 0; c <= nch; ++c) {
                bar_pub_sync<NP>(2 + (c & 1));
                if (lane == 0) {
                    __threadfence();
                    st_relaxed_gpu(&p.flags[tile], c < nch ? (c + 1) * chunk : nrows);
                    sm_pubseq = c + 1;
                }
                __syncwarp();
            }
        } else {
            // ================= compute warps =================
            const int wq = warp;
            const int v = v0 + lane;
            const bool v_ok = v >= w.vlo && v < w.vhi;
            const bool first_w = wq == 0;
            const bool lane_lo = lane == 0, lane_hi = lane == 31;
            const bool ld_hk = (lane_lo && has_v) || (lane_hi && (v0 + 32 < w.vhi));
            const int sm32 = pin((int)w.sm);
            const int nj = pin(w.nj);
            const int chunk_mask = pin(chunk - 1);
            const int hrow_off = lane_lo ? -((int)w.sm + w.sv) : ((int)w.sm + w.sv);   // lane 0: (row-1, v-1); lane 31: (row+1, v+1)
            const int ulast = w.nu - 1;

            // Addressing: one running 32-bit element offset `off` (the row that is D steps ahead of the
            // current step, relative to the row of step 0) plus per-row base pointers, so that every
            // access is base + off (one IMAD.WIDE).
            const T* ldb[R];   // tt loads (queue refill): row of step 0
            const T* sbb[R];   // slowness loads
            const T* hbb[R];   // lane halo loads: lane 0 -> (row-1, v-1), lane 31 -> (row+1, v+1)
            T* stb[R];         // tt stores: D rows behind the refill row
            int jo0[R];        // oriented j at step 0, or a large negative number if the row does not exist
            bool fz[R];        // this u row intersects the frozen box
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int uu = u0 + wq * R + r;
                const int uc = min(uu, ulast);   // rows past the grid: read something harmless, never valid
                const int row0 = m_first - (wq * R + r);
                const long long e = w.base + (long long)uc * w.su + (long long)v * w.sv + (long long)row0 * w.sm;
                ldb[r] = pin_ptr(tt + e);
                sbb[r] = pin_ptr(slo + e);
                hbb[r] = pin_ptr(tt + e + hrow_off);
                stb[r] = pin_ptr(tt + e - (long long)D * w.sm);
                jo0[r] = (uu <= ulast && v_ok) ? row0 - v + w.joff : -(1 << 30);
                const int it = w.ri ? ulast - uc : uc;
                fz[r] = it >= p.fb.ilo && it <= p.fb.ihi;
            }
            // the tile's first row reads (u-1) from tile U-1; its last existing row reads (u+1) from tile U+1
            const int r_last = min(NU - 1, ulast - u0);            // last existing row of the tile (tile-local)
            const bool own_last = (r_last / R) == wq;              // this warp holds it ...
            const int rl = r_last - wq * R;                        // ... in register row rl
            const bool ld_un = first_w && has_u;
            const bool ld_uo = own_last && (u0 + r_last < ulast);
            const T* const eb_un = pin_ptr(ldb[0] - w.su);
            const T* const eb_uo = pin_ptr(ldb[own_last ? rl : 0] + w.su);
            // frozen nodes can only be in tiles that intersect the source box (k range of the tile)
            bool any_fz = false;
            {
                const int ka0 = va - w.vlo, kb0 = vb - 1 - w.vlo;
                const int ka = w.rk ? p.d.nk - 1 - kb0 : ka0, kb = w.rk ? p.d.nk - 1 - ka0 : kb0;
                const bool tile_k = !(kb < p.fb.klo || ka > p.fb.khi);
#pragma unroll
                for (int r = 0; r < R; ++r) { fz[r] = fz[r] && tile_k; any_fz = any_fz || fz[r]; }
            }

            T tq[R][D], sq[R][D], hq[R][D], unq[D], uoq[D];
            T t_prev[R];
            T acc = T(0);
#pragma unroll
            for (int r = 0; r < R; ++r) t_prev[r] = MAXV;

            // highest target step whose halo loads may be issued with what is known to be complete
            int ready_until = (has_u || has_v) ? -1 : (1 << 30);
            auto slow_wait = [&](int st) {
                long long spins = 0;
                for (;;) {
                    const int ku = sm_known_u, kv = sm_known_v;
                    int ru = (1 << 30), rv = (1 << 30);
                    if (has_u && first_w && ku < nrows) ru = ku - 1;   // need st + 1 <= ku
                    if (has_v && kv < nrows_p) rv = kv - dmf;          // need st + dmf <= kv
                    ready_until = min(ru, rv);
                    if (ready_until >= st) break;
                    if (sm_abort) break;
                    if (++spins > p.spin_limit) {
                        atomicExch(&p.ctrl[1], 1);
                        sm_abort = 1;
                        break;
                    }
                }
            };
            // loads of the row at element offset `o` (relative to the step-0 row) into queue slot q
            auto issue = [&](int q, int o) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    tq[r][q] = ld_stream(ldb[r] + o);
                    sq[r][q] = __ldg(sbb[r] + o);
                    hq[r][q] = MAXV;
                    if (ld_hk) hq[r][q] = ld_stream(hbb[r] + o);
                }
                unq[q] = MAXV;
                uoq[q] = MAXV;
                if (ld_un) unq[q] = ld_stream(eb_un + o);
                if (ld_uo) uoq[q] = ld_stream(eb_uo + o);
            };

            // ---- prologue: fill the queues for steps 0 .. D-1
            if (D - 1 > ready_until) slow_wait(D - 1);
#pragma unroll
            for (int q = 0; q < D; ++q) issue(q, q * sm32);
            int off = pin(D * sm32);   // row of step D
            // exchange protocol: at step s a warp publishes the old value of its first register row two
            // rows ahead (the row warp wq-1's last register row works on at step s+1) into xold[(s+1)&1].
            // Step 0 needs a priming publication: that row is my first row's step-1 row.
            const unsigned a_xnew = pin((int)__cvta_generic_to_shared(&xnew[0][wq][lane]));
            const unsigned a_xold = pin((int)__cvta_generic_to_shared(&xold[0][wq][lane]));
            constexpr int XS = NW * 32 * (int)sizeof(T);   // bytes between the two parity buffers
            constexpr int XW = 32 * (int)sizeof(T);        // bytes between neighbouring warps
            sts_at<0>(a_xold, tq[0][1]);
            bar_compute<NC>();
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 1] = gtime();
            int dead = 0;
            int narrive = 0;                                 // publisher arrivals done so far
            const int ngroups = (nsteps + D - 1) / D;        // trailing steps past nsteps touch no valid node
            static_assert(D % 2 == 0, "step parity must be a compile-time constant in the unrolled loop");

            for (int g = 0; g < ngroups && !dead; ++g) {
                if (p.trace && threadIdx.x == 0) {
                    if (g == ngroups / 4) p.trace[tile * 8 + 2] = gtime();
                    if (g == ngroups / 2) p.trace[tile * 8 + 3] = gtime();
                    if (g == (3 * ngroups) / 4) p.trace[tile * 8 + 4] = gtime();
                }
#pragma unroll
                for (int q = 0; q < D; ++q) {
                    const int s = g * D + q;
                    T told[R], sl[R], hv[R], jp[R], tp_old[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        told[r] = tq[r][q];
                        sl[r] = sq[r][q];
                        hv[r] = hq[r][q];
                        jp[r] = tq[r][(q + 1) % D];      // old (u_r, m+1, v)
                        tp_old[r] = t_prev[r];
                    }
                    const T un_g = unq[q], uo_g = uoq[q];
                    const T pub_old = tq[0][(q + 2) % D];   // old two rows ahead of my first row
                    // values from the neighbouring warps (issued early: shared-memory latency)
                    T um0 = un_g, upL = MAXV;
                    if (!first_w) um0 = (q & 1) ? lds_at<-XW>(a_xnew, T(0)) : lds_at<XS - XW>(a_xnew, T(0));
                    if (wq != NW - 1) upL = (q & 1) ? lds_at<XS + XW>(a_xold, T(0)) : lds_at<XW>(a_xold, T(0));
                    // refill slot q with the loads for step s + D
                    if (s + D > ready_until) slow_wait(s + D);
                    issue(q, off);

                    // is any lane of this warp possibly on a frozen node at this step?  (rare)
                    unsigned fzmask = 0;
                    if (any_fz) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (fz[r] && (unsigned)(jo0[r] + s) < (unsigned)nj) {
                                const long long e = (long long)(stb[r] + off - tt);
                                if ((frozen[e >> 5] >> (e & 31)) & 1u) fzmask |= 1u << r;
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        T km = __shfl_up_sync(0xffffffffu, tp_old[r], 1);
                        if (lane_lo) km = hv[r];
                        T kp = __shfl_down_sync(0xffffffffu, jp[r], 1);
                        if (lane_hi) kp = hv[r];
                        const T um = (r == 0) ? um0 : tp_old[r > 0 ? r - 1 : 0];
                        T up = (r == R - 1) ? upL : jp[r + 1 < R ? r + 1 : r];
                        if (own_last && r == rl) up = uo_g;   // last existing row of the tile
                        const T t = godunov(tmin(km, kp), tmin(tp_old[r], jp[r]), tmin(um, up), sl[r] * dx);
                        const bool valid = (unsigned)(jo0[r] + s) < (unsigned)nj && !((fzmask >> r) & 1u);
                        T tnew = told[r];
                        if (valid && t < told[r]) {
                            tnew = t;
                            st_stream(stb[r] + off, t);
                            acc += told[r] - t;
                        }
                        t_prev[r] = tnew;
                    }
                    off = pin(off + sm32);
                    if (q & 1) { sts_at<XS>(a_xnew, t_prev[R - 1]); sts_at<0>(a_xold, pub_old); }
                    else       { sts_at<0>(a_xnew, t_prev[R - 1]); sts_at<XS>(a_xold, pub_old); }
                    if (q == D - 1)
                        dead = bar_compute_or<NC>(sm_abort);
                    else
                        bar_compute<NC>();
                    {   // rows finished by every row of the tile after this step: release the publisher per chunk
                        const int rd = s - NU + 2;
                        if (rd > 0 && rd < nrows && (rd & chunk_mask) == 0) {
                            // barrier ids alternate; do not lap the publisher by more than one chunk
                            if (narrive >= 2) {
                                long long spins = 0;
                                while (sm_pubseq < narrive - 1 && !sm_abort && ++spins < p.spin_limit) {}
                            }
                            bar_pub_arrive<NP>(2 + (narrive & 1));
                            ++narrive;
                        }
                    }
                }
            }
            // remaining arrivals: the final one, plus any skipped because of an abort
            for (; narrive <= nch; ++narrive) {
                if (narrive >= 2) {
                    long long spins = 0;
                    while (sm_pubseq < narrive - 1 && ++spins < p.spin_limit) {}
                }
                bar_pub_arrive<NP>(2 + (narrive & 1));
            }
            // per-tile change: fixed-order reduction (deterministic)
            double dacc = (double)acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
            if (lane == 0) sred[wq] = dacc;
            bar_compute<NC>();
            if (threadIdx.x == 0) {
                double ssum = 0.0;
                for (int i = 0; i < NW; ++i) ssum += sred[i];
                p.partial[tile] = ssum;
                sm_fin = 1;
                if (p.trace) p.trace[tile * 8 + 5] = gtime();
            }
        }
        __syncthreads();
    }
}

// sum of the per-tile partial changes in tile order (deterministic), added to *change
static __global__ void k_sum_partials(const double* __restrict__ partial, int n, double* __restrict__ change) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) a += partial[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *change += sh[0];
}

// ---- host side -----------------------------------------------------------------------------------
#define TCK(call)                                                                        \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " (" #call ")"); \
    } while (0)

inline void tile_alloc(TileState& s, const Dims& d, size_t& bytes) {
    s.cap_tiles = ((d.ni + 3) / 4) * (d.kpad / 32);   // smallest tile height is 4 rows of u
    TCK(cudaMalloc(&s.d_order, s.cap_tiles * sizeof(int)));
    TCK(cudaMalloc(&s.d_flags, s.cap_tiles * sizeof(int)));
    TCK(cudaMalloc(&s.d_ctrl, 8 * sizeof(int)));
    TCK(cudaMalloc(&s.d_partial, s.cap_tiles * sizeof(double)));
    TCK(cudaMallocHost(&s.h_abort, sizeof(int)));
    TCK(cudaMemset(s.d_ctrl, 0, 8 * sizeof(int)));
    *s.h_abort = 0;
    bytes += (size_t)s.cap_tiles * (2 * sizeof(int) + sizeof(double));
}

inline void tile_free(TileState& s) {
    cudaFree(s.d_order); cudaFree(s.d_flags); cudaFree(s.d_ctrl); cudaFree(s.d_partial);
    cudaFree(s.d_mbu); cudaFree(s.d_mbv);
    cudaFreeHost(s.h_abort);
    s = TileState{};
}

// called after a stream synchronize: did any tile kernel give up waiting?
inline void tile_check(TileState& s) {
    if (s.h_abort && *s.h_abort) {
        int info[8] = {0};
        cudaMemcpy(info, s.d_ctrl, sizeof(info), cudaMemcpyDeviceToHost);
        *s.h_abort = 0;
        cudaMemset(s.d_ctrl, 0, 8 * sizeof(int));
        throw std::runtime_error("tile sweep kernel aborted: a dependency wait exceeded spin_limit (reason " +
                                 std::to_string(info[1]) + ", tile " + std::to_string(info[2]) + ", chunk/step " +
                                 std::to_string(info[3]) + ", a " + std::to_string(info[4]) + ", b " + std::to_string(info[5]) +
                                 ", c " + std::to_string(info[6]) + ")");
    }
}

template <typename T> inline bool tile_supported(bool weno_stage) { return !weno_stage; }

template <typename T, int NW, int R, int D>
inline int tile_launch(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, T* tt,
                       const T* slo, const uint32_t* frozen, const FrozenBox& fb, T dx, double* d_change, cudaStream_t st) {
    constexpr int NU = NW * R;
    TileParams p;
    p.w = w; p.d = d; p.fb = fb;
    p.nV = d.kpad / 32;
    p.nU = (w.nu + NU - 1) / NU;
    p.ntiles = p.nU * p.nV;
    p.chunk = 1;
    while (p.chunk * 2 <= o.chunk) p.chunk *= 2;   // power of two
    p.spin_limit = o.spin_limit;
    p.order = s.d_order; p.flags = s.d_flags; p.ctrl = s.d_ctrl; p.partial = s.d_partial;
    static long long* d_trace = nullptr;   // debug only, not thread safe
    const char* trace_path = getenv("TTCR_B200_TRACE");
    if (trace_path && !d_trace) TCK(cudaMalloc(&d_trace, (size_t)s.cap_tiles * 8 * sizeof(long long)));
    p.trace = trace_path ? d_trace : nullptr;
    // ticket order: any linear extension of (U-1,V) < (U,V), (U,V-1) < (U,V); sorted by estimated
    // start time so that running CTAs are the ones whose inputs are about to be ready
    const int key = NU * 1000 + D * 10 + p.chunk * 100000;
    if (s.order_key != key || s.ntiles != p.ntiles) {
        std::vector<std::pair<long long, int>> k(p.ntiles);
        const long long lag_u = NU + D + p.chunk + 4;
        for (int U = 0; U < p.nU; ++U)
            for (int V = 0; V < p.nV; ++V) {
                const int va = std::max(V * 32, w.vlo);
                k[U * p.nV + V] = {U * lag_u + (long long)(va - w.joff), U * p.nV + V};
            }
        std::stable_sort(k.begin(), k.end());
        std::vector<int> order(p.ntiles);
        for (int i = 0; i < p.ntiles; ++i) order[i] = k[i].second;
        TCK(cudaMemcpyAsync(s.d_order, order.data(), p.ntiles * sizeof(int), cudaMemcpyHostToDevice, st));
        TCK(cudaStreamSynchronize(st));   // `order` is a local
        s.order_key = key;
        s.ntiles = p.ntiles;
    }
    TCK(cudaMemsetAsync(s.d_flags, 0, p.ntiles * sizeof(int), st));
    TCK(cudaMemsetAsync(s.d_ctrl, 0, sizeof(int), st));   // ticket counter only; the abort flag is sticky
    // per instantiation; a pure query (static shared memory, no per-device function attribute), so every device of a box and
    // every slot thread computes the same value: the atomic only makes the concurrent first use well defined
    static std::atomic<int> occ_cache{0};
    if (!occ_cache.load(std::memory_order_relaxed)) {
        int q = 0;
        TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_sweep_tile<T, NW, R, D>, (NW + 2) * 32, 0));
        if (q < 1) throw std::runtime_error("tile kernel does not fit on an SM");
        occ_cache.store(q, std::memory_order_relaxed);
    }
    int occ = occ_cache;
    if (o.ctas_per_sm > 0) occ = std::min(occ, o.ctas_per_sm);
    const int grid = std::min(p.ntiles, occ * sm_count);
    k_sweep_tile<T, NW, R, D><<<grid, (NW + 2) * 32, 0, st>>>(p, tt, slo, frozen, dx);
    k_sum_partials<<<1, 256, 0, st>>>(s.d_partial, p.ntiles, d_change);
    TCK(cudaMemcpyAsync(s.h_abort, s.d_ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    TCK(cudaGetLastError());
    if (trace_path) {
        std::vector<long long> h((size_t)p.ntiles * 8);
        TCK(cudaStreamSynchronize(st));
        TCK(cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        FILE* f = fopen(trace_path, "ab");
        if (f) {
            const int hdr[4] = {p.ntiles, p.nU, p.nV, NU};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h.data(), sizeof(long long), h.size(), f);
            fclose(f);
        }
    }
    return 2;
}

template <typename T>
inline int tile_sweep(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, T* tt,
                      const T* slo, const uint32_t* frozen, const FrozenBox& fb, T dx, bool weno_stage, double* d_change,
                      cudaStream_t st) {
    if (weno_stage) throw std::runtime_error("tile kernel: WENO stage not supported");
#define TTCR_TILE_CASE(NW_, R_, D_) \
    if (o.warps == NW_ && o.rows == R_ && o.depth == D_) \
        return tile_launch<T, NW_, R_, D_>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    TTCR_TILE_CASE(8, 2, 4)
    TTCR_TILE_CASE(8, 1, 4)
    TTCR_TILE_CASE(4, 2, 4)
    TTCR_TILE_CASE(8, 2, 8)
    TTCR_TILE_CASE(8, 1, 8)
    TTCR_TILE_CASE(4, 1, 4)
#undef TTCR_TILE_CASE
    // not an instantiated combination: use the default shape
    return tile_launch<T, 8, 1, 8>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
}

}  // namespace ttcrb200
