// Sweep kernel "TILE": one persistent launch per directional sweep.
//
// Why: a directional Gauss-Seidel sweep of an N^3 grid has 3N-2 dependent wavefronts; at 512^3 the
// HBM roofline leaves ~0.2 us per wavefront, far below any kernel-launch or grid-barrier latency, so
// "one launch per wavefront" (k_sweep_plane, and the reference's OpenCL path) is launch bound.
//
// How: in oriented coordinates (layout.cuh) the grid is cut into tiles of NW rows of u by 32 lanes
// of v.  A CTA owns a tile and MARCHES along m (the sheared row axis): warp wq handles u = u0+wq,
// lane l handles v = v0+l, and at step s warp wq updates row m = m_first + s - wq.  All 32 lanes of
// a warp touch one contiguous 128-byte row segment per array per step (tt read, slowness read, tt
// write: the 12 algorithmic bytes per node), streamed through a register queue D rows deep.
// Everything else a node needs comes from registers, shuffles or shared memory:
//     (u, m-1, v)   own previous result            register
//     (u, m-1, v-1) lane l-1's previous result     __shfl_up    (lane 0: halo from tile V-1, L2)
//     (u, m+1, v)   own next old value             register queue
//     (u, m+1, v+1) lane l+1's next old value      __shfl_down  (lane 31: halo from tile V+1)
//     (u-1, m, v)   warp wq-1's result of step s-1 shared memory (warp 0: halo from tile U-1, L2)
//     (u+1, m, v)   warp wq+1's old value          shared memory (last warp: halo from tile U+1)
// Tiles depend on their U-1 and V-1 neighbours row by row.  Each tile publishes "rows completed"
// in a global flag every `chunk` rows; a dedicated SYNC WARP per CTA polls the two upstream flags
// (ld.acquire.gpu) and publishes the tile's own (fence + st.release.gpu), so the compute warps only
// ever read two shared-memory words and never stall on a fence.  Tiles are handed out through an
// atomic ticket in an order that is a linear extension of the dependency order, so a CTA can only
// wait for tiles that are already running or finished: no co-residency requirement, no deadlock.
// Every spin is bounded; on timeout the kernel raises an abort flag and the host reports an error.
//
// Results are bit-identical to k_sweep_plane and to the lexicographic CPU order: all three execute
// the same dependency DAG with the same arithmetic (update.cuh).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace ttcrb200 {

struct TileOptions {
    int chunk = 4;         // rows between progress-flag publications
    int ctas_per_sm = 0;   // 0 = as many as fit
    int warps = 8;         // compute warps (u rows) per tile: 4, 8 or 16
    long long spin_limit = 1ll << 22;   // polls before a wait is declared dead
};

struct TileState {
    int* d_order = nullptr;      // ticket -> tile id
    int* d_flags = nullptr;      // per tile: rows completed
    int* d_ctrl = nullptr;       // [0] ticket counter, [1] abort flag
    double* d_partial = nullptr; // per tile: sum of (old - new)
    int* h_abort = nullptr;      // pinned copy of the abort flag
    int cap_tiles = 0;
    int order_key = -1;          // (dir, warps) the order table was built for
    int ntiles = 0;
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
};

// ---- small PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
    int v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(int* p, int v) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_shared(const int* p) {
    int v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
template <int N>
__device__ __forceinline__ void bar_compute() {   // named barrier 1: the compute warps only
    asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}
// same barrier, OR-reducing a predicate so that every compute thread takes the same decision
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

// ---- the kernel ---------------------------------------------------------------------------------
// block = (NW + 1) warps: NW compute warps + 1 sync warp.  D = register queue depth (rows in flight).
template <typename T, int NW, int D>
__global__ void __launch_bounds__((NW + 1) * 32) k_sweep_tile(TileParams p, T* __restrict__ tt, const T* __restrict__ slo,
                                                             const uint32_t* __restrict__ frozen, T dx) {
    static_assert(D >= 3, "the exchange protocol publishes old values two rows ahead");
    __shared__ T xnew[2][NW][32];   // new values of the step just finished, per warp
    __shared__ T xold[2][NW][32];   // old values two rows ahead, per warp
    __shared__ double sred[NW];
    __shared__ int sm_tile, sm_known_u, sm_known_v, sm_done, sm_abort;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const T MAXV = Lim<T>::max();

    for (;;) {
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sm_tile = (ab || t >= p.ntiles) ? -1 : t;
            sm_known_u = 0; sm_known_v = 0; sm_done = 0; sm_abort = 0;
        }
        __syncthreads();
        const int ticket = sm_tile;
        if (ticket < 0) break;
        const int tile = p.order[ticket];
        const int U = tile / p.nV, V = tile - U * p.nV;
        const int u0 = U * NW, v0 = V * 32;
        const int va = max(v0, w.vlo), vb = min(v0 + 32, w.vhi);
        const int m_first = va - w.joff;
        const int nrows = (vb - va) + w.nj - 1;
        const int nsteps = nrows + NW - 1;
        const bool has_u = U > 0, has_v = V > 0;
        // tile (U, V-1): its first row and row count
        const int va_p = max(v0 - 32, w.vlo);
        const int mf_p = va_p - w.joff;
        const int nrows_p = (v0 - va_p) + w.nj - 1;
        const int dmf = m_first - mf_p;

        if (warp == NW) {
            // ================= sync warp =================
            if (lane == 0) {
                int pub = 0;
                int ku = has_u ? 0 : nrows, kv = has_v ? 0 : nrows_p;
                long long spins = 0;
                int* myflag = &p.flags[tile];
                const int* fu = &p.flags[has_u ? tile - p.nV : tile];
                const int* fv = &p.flags[has_v ? tile - 1 : tile];
                for (;;) {
                    const int done = ld_acquire_cta_shared(&sm_done);
                    if (done > pub && (done >= pub + p.chunk || done >= nrows)) {
                        __threadfence();
                        st_release_gpu(myflag, done);
                        pub = done;
                        spins = 0;
                    }
                    if (ku < nrows) {
                        const int k2 = ld_acquire_gpu(fu);
                        if (k2 > ku) { ku = k2; st_release_cta_shared(&sm_known_u, ku); spins = 0; }
                    }
                    if (kv < nrows_p) {
                        const int k2 = ld_acquire_gpu(fv);
                        if (k2 > kv) { kv = k2; st_release_cta_shared(&sm_known_v, kv); spins = 0; }
                    }
                    if (pub >= nrows) break;
                    if (ld_volatile_shared(&sm_abort)) break;
                    if (++spins > p.spin_limit) {
                        atomicExch(&p.ctrl[1], 1);
                        st_release_cta_shared(&sm_abort, 1);
                        break;
                    }
                }
            }
            __syncwarp();
        } else {
            // ================= compute warps =================
            const int wq = warp;
            const int u = u0 + wq;
            const int v = v0 + lane;
            const bool u_ok = u < w.nu;
            const bool v_ok = v >= w.vlo && v < w.vhi;
            const bool first_w = wq == 0;
            const bool last_w = (wq == NW - 1) || (u == w.nu - 1);
            const bool ld_un = first_w && has_u;                     // (u-1) new halo from global
            const bool ld_uo = last_w && u_ok && (u + 1 < w.nu);     // (u+1) old halo from global
            const bool lane_lo = lane == 0, lane_hi = lane == 31;
            const bool ld_hk = u_ok && ((lane_lo && has_v) || (lane_hi && (v0 + 32 < w.vhi)));
            // element offset of (u, row 0, v); rows advance by w.sm
            const long long e0 = w.base + (long long)u * w.su + (long long)v * w.sv;
            // halo offsets: lane 0 reads (row-1, v-1), lane 31 reads (row+1, v+1)
            const long long hk_off = lane_lo ? (-w.sm - w.sv) : (w.sm + w.sv);
            // frozen nodes can only be in tiles that intersect the source box
            bool tile_frozen;
            {
                const int ia = w.ri ? w.nu - 1 - min(u0 + NW - 1, w.nu - 1) : u0, ib = w.ri ? w.nu - 1 - u0 : min(u0 + NW - 1, w.nu - 1);
                const int ka0 = va - w.vlo, kb0 = vb - 1 - w.vlo;
                const int ka = w.rk ? p.d.nk - 1 - kb0 : ka0, kb = w.rk ? p.d.nk - 1 - ka0 : kb0;
                tile_frozen = !(ib < p.fb.ilo || ia > p.fb.ihi || kb < p.fb.klo || ka > p.fb.khi);
            }
            const int it = w.ri ? w.nu - 1 - u : u;
            const bool row_frozen_u = tile_frozen && it >= p.fb.ilo && it <= p.fb.ihi;

            T tq[D], sq[D], hq[D], nq[D], oq[D];
            T t_prev = MAXV;
            double acc = 0.0;

            auto wait_deps = [&](int st) {   // before issuing halo loads that target step st
                if (!(has_u || has_v)) return;
                const int need_u = (has_u && first_w) ? min(st + 1, nrows) : 0;
                const int need_v = has_v ? max(0, min(st + dmf, nrows_p)) : 0;
                long long spins = 0;
                while (ld_acquire_cta_shared(&sm_known_u) < need_u || ld_acquire_cta_shared(&sm_known_v) < need_v) {
                    if (ld_volatile_shared(&sm_abort)) break;
                    if (++spins > p.spin_limit) {
                        atomicExch(&p.ctrl[1], 1);
                        st_release_cta_shared(&sm_abort, 1);
                        break;
                    }
                }
            };
            // loads that feed step st (row index relative to m_first: st - wq)
            auto issue = [&](int st, T& tv, T& sv_, T& hv, T& nv, T& ov) {
                const int m = m_first + st - wq;
                const bool row_ok = m >= 0 && m < w.nm;
                const long long e = e0 + (long long)m * w.sm;
                tv = (u_ok && row_ok) ? __ldcg(&tt[e]) : MAXV;
                sv_ = (u_ok && row_ok) ? __ldg(&slo[e]) : T(0);
                const int mh = lane_lo ? m - 1 : m + 1;
                hv = (ld_hk && mh >= 0 && mh < w.nm) ? __ldcg(&tt[e + hk_off]) : MAXV;
                nv = (ld_un && row_ok) ? __ldcg(&tt[e - w.su]) : MAXV;
                ov = (ld_uo && row_ok) ? __ldcg(&tt[e + w.su]) : MAXV;
            };

            // ---- prologue: fill the queues for steps 0 .. D-1
            wait_deps(D - 1);
#pragma unroll
            for (int r = 0; r < D; ++r) issue(r, tq[r], sq[r], hq[r], nq[r], oq[r]);
            // old values two rows ahead must be visible to warp wq-1 at its step: publish for step 0, 1
            // exchange protocol: at step s a warp publishes its old value of step s+2's row (the row warp
            // wq-1 works on at step s+1) into xold[(s+1)&1].  Step 0 needs a priming publication: warp
            // wq-1 at step 0 is on my step-1 row.
            xold[0][wq][lane] = tq[1];
            bar_compute<NW * 32>();
            bool dead = false;

            for (int sb = 0; sb < nsteps && !dead; sb += D) {
#pragma unroll
                for (int r = 0; r < D; ++r) {
                    const int s = sb + r;
                    if (s < nsteps && !dead) {   // uniform
                        const int par = s & 1;
                        const int m = m_first + s - wq;
                        const T told = tq[r];
                        const T sl = sq[r];
                        const T hv = hq[r];
                        const T un_g = nq[r];
                        const T uo_g = oq[r];
                        const T jp = tq[(r + 1) % D];   // old (u, m+1, v)
                        const T pub_old = tq[(r + 2) % D];   // old (u, m+2, v): what warp wq-1 needs next step
                        // refill slot r with the loads for step s + D
                        wait_deps(s + D);
                        issue(s + D, tq[r], sq[r], hq[r], nq[r], oq[r]);

                        // neighbours
                        T km = __shfl_up_sync(0xffffffffu, t_prev, 1);
                        if (lane_lo) km = hv;
                        T kp = __shfl_down_sync(0xffffffffu, jp, 1);
                        if (lane_hi) kp = hv;
                        const T um = first_w ? un_g : xnew[par ^ 1][first_w ? 0 : wq - 1][lane];
                        const T up = last_w ? uo_g : xold[par][last_w ? 0 : wq + 1][lane];
                        const T au = tmin(um, up), aj = tmin(t_prev, jp), ak = tmin(km, kp);
                        const T fh = sl * dx;
                        const T t = godunov(ak, aj, au, fh);

                        const int jo = m - v + w.joff;
                        bool valid = u_ok && v_ok && jo >= 0 && jo < w.nj;
                        if (row_frozen_u && valid) {
                            const long long e = e0 + (long long)m * w.sm;
                            if ((frozen[e >> 5] >> (e & 31)) & 1u) valid = false;
                        }
                        T tnew = told;
                        if (valid && t < told) {
                            tnew = t;
                            __stcg(&tt[e0 + (long long)m * w.sm], t);
                            acc += (double)told - (double)t;
                        }
                        t_prev = tnew;
                        xnew[par][wq][lane] = tnew;
                        xold[par ^ 1][wq][lane] = pub_old;
                        dead = bar_compute_or<NW * 32>(ld_volatile_shared(&sm_abort)) != 0;
                        {   // rows finished by every warp after this step; hand over to the sync warp per chunk
                            const int rd = s - NW + 2;
                            if (threadIdx.x == 0 && rd > 0 && rd < nrows && rd % p.chunk == 0) st_release_cta_shared(&sm_done, rd);
                        }
                    }
                }
            }
            // per-tile change: fixed-order reduction (deterministic)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) sred[wq] = acc;
            bar_compute<NW * 32>();
            if (threadIdx.x == 0) {
                double ssum = 0.0;
                for (int i = 0; i < NW; ++i) ssum += sred[i];
                p.partial[tile] = ssum;
                st_release_cta_shared(&sm_done, nrows);
            }
        }
        __syncthreads();
    }
}

// sum of the per-tile partial changes in tile order (deterministic), added to *change
__global__ void k_sum_partials(const double* __restrict__ partial, int n, double* __restrict__ change) {
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
    TCK(cudaMalloc(&s.d_ctrl, 2 * sizeof(int)));
    TCK(cudaMalloc(&s.d_partial, s.cap_tiles * sizeof(double)));
    TCK(cudaMallocHost(&s.h_abort, sizeof(int)));
    TCK(cudaMemset(s.d_ctrl, 0, 2 * sizeof(int)));
    *s.h_abort = 0;
    bytes += (size_t)s.cap_tiles * (2 * sizeof(int) + sizeof(double));
}

inline void tile_free(TileState& s) {
    cudaFree(s.d_order); cudaFree(s.d_flags); cudaFree(s.d_ctrl); cudaFree(s.d_partial);
    cudaFreeHost(s.h_abort);
    s = TileState{};
}

// called after a stream synchronize: did any tile kernel give up waiting?
inline void tile_check(TileState& s) {
    if (s.h_abort && *s.h_abort) {
        *s.h_abort = 0;
        cudaMemset(s.d_ctrl, 0, 2 * sizeof(int));
        throw std::runtime_error("tile sweep kernel aborted: a dependency wait exceeded spin_limit");
    }
}

template <typename T> inline bool tile_supported(bool weno_stage) { return !weno_stage; }

template <typename T, int NW, int D>
inline int tile_launch(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, T* tt,
                       const T* slo, const uint32_t* frozen, const FrozenBox& fb, T dx, double* d_change, cudaStream_t st) {
    TileParams p;
    p.w = w; p.d = d; p.fb = fb;
    p.nV = d.kpad / 32;
    p.nU = (w.nu + NW - 1) / NW;
    p.ntiles = p.nU * p.nV;
    p.chunk = o.chunk;
    p.spin_limit = o.spin_limit;
    p.order = s.d_order; p.flags = s.d_flags; p.ctrl = s.d_ctrl; p.partial = s.d_partial;
    // ticket order: any linear extension of (U-1,V) < (U,V), (U,V-1) < (U,V); sorted by estimated
    // start time so that running CTAs are the ones whose inputs are about to be ready
    const int key = NW * 1000 + (w.rk ? 1 : 0);
    if (s.order_key != key || s.ntiles != p.ntiles) {
        std::vector<std::pair<long long, int>> k(p.ntiles);
        const long long lag_u = NW + D + o.chunk + 6;
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
    TCK(cudaMemsetAsync(s.d_ctrl, 0, sizeof(int), st));   // ticket counter only; abort flag is sticky
    int occ = 0;
    TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_tile<T, NW, D>, (NW + 1) * 32, 0));
    if (occ < 1) throw std::runtime_error("tile kernel does not fit on an SM");
    if (o.ctas_per_sm > 0) occ = std::min(occ, o.ctas_per_sm);
    const int grid = std::min(p.ntiles, occ * sm_count);
    k_sweep_tile<T, NW, D><<<grid, (NW + 1) * 32, 0, st>>>(p, tt, slo, frozen, dx);
    k_sum_partials<<<1, 256, 0, st>>>(s.d_partial, p.ntiles, d_change);
    TCK(cudaMemcpyAsync(s.h_abort, s.d_ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    TCK(cudaGetLastError());
    return 2;
}

template <typename T>
inline int tile_sweep(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, T* tt,
                      const T* slo, const uint32_t* frozen, const FrozenBox& fb, T dx, bool weno_stage, double* d_change,
                      cudaStream_t st) {
    if (weno_stage) throw std::runtime_error("tile kernel: WENO stage not supported");
    constexpr int D = 4;
    switch (o.warps) {
        case 4: return tile_launch<T, 4, D>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
        case 16: return tile_launch<T, 16, D>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
        default: return tile_launch<T, 8, D>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    }
}

}  // namespace ttcrb200
