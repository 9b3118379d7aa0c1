// Sweep kernel "TILE3": the tile-marching sweep of sweep_tile.cuh with warp-specialised data movement.
//
// Same decomposition, dependency protocol and arithmetic as k_sweep_tile (read its header first); what
// changes is who moves the data.  In k_sweep_tile every compute thread streams its own rows through
// register queues, and a step of a compute warp is ~170 mostly serial instructions (address arithmetic,
// queue rotation, five global loads) of which ~40 are the Godunov update.  Here:
//
//   * a LOADER warp (which is also the poller of the upstream flags) issues, per row group, one bulk
//     asynchronous copy per stream (cp.async.bulk global -> shared, SASS UBLKCP: the 1-D form of TMA),
//     completing on an mbarrier per chunk of 4 rows.  Streams of a tile with NU rows of u: NU+1
//     traveltime rings (the tile's rows and row u0+NU; 40 floats per row = 32 lanes + the v0-1 / v0+32
//     halo lanes), NU slowness rings (32 floats per row), and one ring for row u0-1.
//   * the COMPUTE warps (one row of u each) read everything from shared memory with immediate offsets:
//     own next row (jp), the v+1 lane of it (kp), the u+1 ring (up), slowness, the u-1 result of the
//     neighbouring warp (exchange buffer) and the v-1 halo; one shuffle; godunov(); one predicated
//     global store; one exchange store; one named barrier.  No global loads, no queues.
//   * the PUBLISHER warp is unchanged (bar.arrive hand-off, release fence, flag store).
//
// Ring rows are indexed by the LOCAL row index l of a u row (the step at which that row of u works on
// it): row i of the tile at step s is on local index s; it reads local s+1 of its own ring (jp, kp) and of
// ring i+1 (up), slowness local s, and the v-1 halo of local s-1.  The loader numbers its row groups
// G = l + 1 (group 0 is local -1, needed by the v-1 halo of step 0); slot = G mod 16.
//
// fp32 only (fp64 rings would need opt-in shared memory); other types use k_sweep_tile.
#pragma once
#include "sweep_tile.cuh"

namespace ttcrb200 {

__device__ __forceinline__ void mbar_init(unsigned a, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned a) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
// one try_wait attempt (the hardware suspends the thread for a bounded time); 1 = that phase has completed
__device__ __forceinline__ int mbar_test(unsigned a, unsigned parity) {
    int ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.s32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
    return ok;
}
// global -> shared bulk copy (bytes multiple of 16, both addresses 16-byte aligned), completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <int NW>
struct Tile3Smem {
    static constexpr int RING = 16;   // row slots per ring
    static constexpr int TW = 40;     // floats per traveltime ring row: 4 + 32 + 4
    float T[NW + 1][RING][TW];        // ring i: plane u0+i (old values)
    float Tlo[RING][TW];              // plane u0-1 (new values of tile U-1)
    float S[NW][RING][32];
    float xnew[2][NW][32];
    unsigned long long full[4];       // one mbarrier per chunk slot
};

template <int NW>
__global__ void __launch_bounds__((NW + 2) * 32) k_sweep_tile3(TileParams p, float* __restrict__ tt, const float* __restrict__ slo,
                                                              const uint32_t* __restrict__ frozen, float dx) {
    using SM = Tile3Smem<NW>;
    constexpr int RING = SM::RING, TW = SM::TW;
    constexpr int C = 4, NCH = RING / C;          // rows per mbarrier chunk, chunk slots
    constexpr int NU = NW, NC = NW * 32, NP = (NW + 1) * 32;
    constexpr int TROW = TW * 4, SROW = 32 * 4;   // bytes per ring row
    static_assert(NCH == 4, "parity arithmetic assumes 4 chunk slots of 4 rows");
    static_assert(NW + RING + 8 <= GUARD, "guard rows too few");
    static_assert(2 * NW + 2 <= 32, "one loader lane per stream");
    __shared__ __align__(128) SM sm;
    __shared__ double sred[NW];
    __shared__ int sm_tile;
    __shared__ volatile int sm_pubseq, sm_abort, sm_step_done;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const float MAXV = FLT_MAX;
    const unsigned a_full = (unsigned)__cvta_generic_to_shared(&sm.full[0]);
    bool first_tile = true;

    for (;;) {
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sm_tile = (ab || t >= p.ntiles) ? -1 : t;
            sm_pubseq = 0; sm_abort = 0; sm_step_done = -1;
            // fresh barriers for every tile: chunk c of a tile uses slot c % 4 with parity (c / 4) & 1
            for (int c = 0; c < NCH; ++c) {
                if (!first_tile) mbar_inval(a_full + 8 * c);
                mbar_init(a_full + 8 * c, 1);
            }
            fence_mbar_init();
        }
        first_tile = false;
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
        const int ngroups16 = (nsteps + RING - 1) / RING;     // compute steps are padded to a multiple of RING
        const int nload = ngroups16 * RING + 4;               // row groups the loader delivers (a multiple of C)
        const int nchunks = nload / C;
        const bool has_u = U > 0, has_v = V > 0;
        const bool has_right = v0 + 32 < w.vhi;
        const int va_p = max(v0 - 32, w.vlo);
        const int nrows_p = (v0 - va_p) + w.nj - 1;
        const int dmf = m_first - (va_p - w.joff);
        const int chunk = p.chunk;
        const int nch = (nrows - 1) / chunk;
        const int ulast = w.nu - 1;
        const int nexist = min(NU, ulast - u0 + 1);           // rows of the tile that exist
        if (p.trace && threadIdx.x == 0) {
            p.trace[tile * 8 + 0] = gtime();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[tile * 8 + 7] = smid;
        }

        if (warp == NW) {
            // ================= loader (and poller) =================
            // lane j owns stream j:  [0, NU]  traveltime ring j (plane u0+j);  NU+1  plane u0-1;
            //                        [NU+2, 2NU+2)  slowness ring j-NU-2
            const int j = lane;
            const bool isT = j <= NU, isLo = j == NU + 1, isS = j >= NU + 2 && j < 2 * NU + 2;
            const int ring_i = isT ? j : (isS ? j - NU - 2 : 0);
            const int plane = isLo ? u0 - 1 : u0 + ring_i;
            bool active = (isT || isS) ? plane <= ulast : (isLo && has_u);
            // first lane of the copied segment in memory order: traveltime rows carry 4 halo lanes on each side
            const int halo = (isT || isLo) ? 4 : 0;
            const long long seg = w.rk ? -(long long)(v0 + 31 + halo) : (long long)(v0 - halo);
            const int row0 = (isLo ? m_first : m_first - ring_i) - 1;  // absolute row of group 0 (local index -1)
            const float* src = ((isS) ? slo : (const float*)tt) + w.base + (long long)plane * w.su + seg +
                               (long long)row0 * w.sm;
            const unsigned bytes = (isT || isLo) ? TROW : SROW;
            unsigned dst0;
            if (isT) dst0 = (unsigned)__cvta_generic_to_shared(&sm.T[ring_i][0][0]);
            else if (isLo) dst0 = (unsigned)__cvta_generic_to_shared(&sm.Tlo[0][0]);
            else dst0 = (unsigned)__cvta_generic_to_shared(&sm.S[ring_i][0][0]);
            if (!active) src = tt;   // never dereferenced
            // bytes landing per local row, all streams
            const unsigned group_bytes = (unsigned)(min(NU + 1, ulast - u0 + 1) * TROW + (has_u ? TROW : 0) + nexist * SROW);
            const int* fu = &p.flags[has_u ? tile - p.nV : tile];
            const int* fv = &p.flags[has_v ? tile - 1 : tile];
            int ku = has_u ? 0 : (1 << 30), kv = has_v ? 0 : (1 << 30);
            int dead = 0;
            for (int c = 0; c < nchunks && !dead; ++c) {
                const int g0 = c * C;
                if (lane == 0) {
                    long long spins = 0;
                    // (a) the slots of this chunk hold groups G-RING, last read at step G-RING (its v-1 halo)
                    // (b) upstream tiles must have finished what the halo lanes / row u0-1 of these rows hold:
                    //     group G is absolute row m_first-1+G of row u0-1, and at most row count G+dmf of tile V-1
                    const int need_done = g0 + C - 1 - RING;
                    const int need_u = min(g0 + C - 1, nrows);
                    const int need_v = min(g0 + C - 1 + dmf, nrows_p);
                    for (;;) {
                        bool ok = sm_step_done >= need_done;
                        if (ok && ku < need_u) {
                            const int k2 = ld_relaxed_gpu(fu);
                            if (k2 > ku) { ku = k2; fence_acq_rel_gpu(); fence_proxy_async(); }
                            ok = ku >= need_u;
                        }
                        if (ok && kv < need_v) {
                            const int k2 = ld_relaxed_gpu(fv);
                            if (k2 > kv) { kv = k2; fence_acq_rel_gpu(); fence_proxy_async(); }
                            ok = kv >= need_v;
                        }
                        if (ok) break;
                        if (sm_abort) { dead = 1; break; }
                        if (++spins > p.spin_limit) {
                            atomicExch(&p.ctrl[1], 1);
                            sm_abort = 1;
                            dead = 1;
                            break;
                        }
                    }
                    if (!dead) mbar_expect_tx(a_full + 8 * (c & 3), group_bytes * C);
                }
                dead = __shfl_sync(0xffffffffu, dead, 0);
                if (!dead && active) {
                    const unsigned mb = a_full + 8 * (c & 3);
#pragma unroll
                    for (int r = 0; r < C; ++r) {
                        const int g = g0 + r;
                        bulk_g2s(dst0 + (unsigned)((g & (RING - 1)) * bytes), src + (long long)g * w.sm, bytes, mb);
                    }
                }
            }
            // every copy must have landed before the barriers are re-initialised for the next tile
            if (lane == 0 && !dead) {
                for (int c = max(0, nchunks - NCH); c < nchunks; ++c) {
                    long long spins = 0;
                    while (!mbar_test(a_full + 8 * (c & 3), (c >> 2) & 1) && ++spins < (p.spin_limit >> 6)) {}
                }
            }
            __syncwarp();
        } else if (warp == NW + 1) {
            // ================= publisher =================
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
            // ================= compute warps: warp wq owns row u0 + wq =================
            const int wq = warp;
            const int u = u0 + wq;
            const int v = v0 + lane;
            const bool u_ok = u <= ulast;
            const bool v_ok = v >= w.vlo && v < w.vhi;
            const bool first_w = wq == 0;
            const bool has_um = !first_w || has_u;           // a (u-1) row exists
            const bool has_up = u < ulast;                   // a (u+1) row exists
            const bool lane_lo = lane == 0, lane_hi = lane == 31;
            const int sm32 = pin((int)w.sm);
            const int nj = pin(w.nj);
            const int chunk_mask = pin(chunk - 1);
            // column of this lane in a ring row (rows are in memory order: reversed sweeps run right to left)
            const int col = w.rk ? 35 - lane : 4 + lane;
            const int dcol = w.rk ? -1 : 1;                  // column of lane v+1 relative to col
            const unsigned aT = pin((int)__cvta_generic_to_shared(&sm.T[wq][0][col]));
            const unsigned aTvp = pin((int)__cvta_generic_to_shared(&sm.T[wq][0][col + dcol]));
            const unsigned aTvm = pin((int)__cvta_generic_to_shared(&sm.T[wq][0][col - dcol]));
            const unsigned aTup = pin((int)__cvta_generic_to_shared(&sm.T[wq + 1][0][col]));
            const unsigned aTlo = pin((int)__cvta_generic_to_shared(&sm.Tlo[0][col]));
            const unsigned aS = pin((int)__cvta_generic_to_shared(&sm.S[wq][0][w.rk ? 31 - lane : lane]));
            const unsigned aX = pin((int)__cvta_generic_to_shared(&sm.xnew[0][wq][lane]));
            constexpr int XS = NW * 32 * 4, XW = 32 * 4;
            // global store pointer: row of step 0 plus a running element offset
            const int row0 = m_first - wq;
            const long long e0 = w.base + (long long)min(u, ulast) * w.su + (long long)v * w.sv + (long long)row0 * w.sm;
            float* const stb = pin_ptr(tt + e0);
            int off = pin(0);
            const int jo0 = (u_ok && v_ok) ? row0 - v + w.joff : -(1 << 30);
            bool any_fz;
            {
                const int it = w.ri ? ulast - min(u, ulast) : min(u, ulast);
                const int ka0 = va - w.vlo, kb0 = vb - 1 - w.vlo;
                const int ka = w.rk ? p.d.nk - 1 - kb0 : ka0, kb = w.rk ? p.d.nk - 1 - ka0 : kb0;
                any_fz = it >= p.fb.ilo && it <= p.fb.ihi && !(kb < p.fb.klo || ka > p.fb.khi);
            }
            auto wait_chunk = [&](int c) {
                const unsigned a = a_full + 8 * (c & 3), par = (c >> 2) & 1;
                long long spins = 0;
                while (!mbar_test(a, par)) {   // each attempt suspends the thread for a bounded time
                    if (sm_abort) break;
                    if (++spins > (p.spin_limit >> 6)) {
                        atomicExch(&p.ctrl[1], 1);
                        sm_abort = 1;
                        break;
                    }
                }
            };

            // ---- prologue: groups 0..3 (local -1..2) have landed; old value of local row 0
            wait_chunk(0);
            float told = lds_at<TROW>(aT, 0.f);
            float t_prev = MAXV;
            float acc = 0.f;
            bar_compute<NC>();
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 1] = gtime();
            int dead = 0;
            int narrive = 0;

            for (int g = 0; g < ngroups16 && !dead; ++g) {
                if (p.trace && threadIdx.x == 0) {
                    if (g == ngroups16 / 4) p.trace[tile * 8 + 2] = gtime();
                    if (g == ngroups16 / 2) p.trace[tile * 8 + 3] = gtime();
                    if (g == (3 * ngroups16) / 4) p.trace[tile * 8 + 4] = gtime();
                }
#pragma unroll
                for (int q = 0; q < RING; ++q) {
                    const int s = g * RING + q;
                    // group s+2 (local s+1) is needed from here on: entering a new chunk?
                    if (((q + 2) & (C - 1)) == 0) wait_chunk((s + 2) >> 2);
                    const int s1 = (q + 2) & (RING - 1), s0 = (q + 1) & (RING - 1), sm1 = q;   // slots of local s+1, s, s-1
                    float jp, kp, up, sl, um, kmh;
                    // (switch on q is resolved at compile time: q is an unrolled loop index)
#define TTCR_LD(dstv, base, slot, rowb)                                                                       \
    switch (slot) {                                                                                            \
        case 0: dstv = lds_at<0 * rowb>(base, 0.f); break;   case 1: dstv = lds_at<1 * rowb>(base, 0.f); break;   \
        case 2: dstv = lds_at<2 * rowb>(base, 0.f); break;   case 3: dstv = lds_at<3 * rowb>(base, 0.f); break;   \
        case 4: dstv = lds_at<4 * rowb>(base, 0.f); break;   case 5: dstv = lds_at<5 * rowb>(base, 0.f); break;   \
        case 6: dstv = lds_at<6 * rowb>(base, 0.f); break;   case 7: dstv = lds_at<7 * rowb>(base, 0.f); break;   \
        case 8: dstv = lds_at<8 * rowb>(base, 0.f); break;   case 9: dstv = lds_at<9 * rowb>(base, 0.f); break;   \
        case 10: dstv = lds_at<10 * rowb>(base, 0.f); break; case 11: dstv = lds_at<11 * rowb>(base, 0.f); break; \
        case 12: dstv = lds_at<12 * rowb>(base, 0.f); break; case 13: dstv = lds_at<13 * rowb>(base, 0.f); break; \
        case 14: dstv = lds_at<14 * rowb>(base, 0.f); break; default: dstv = lds_at<15 * rowb>(base, 0.f); break; \
    }
                    TTCR_LD(jp, aT, s1, TROW)
                    TTCR_LD(kp, aTvp, s1, TROW)
                    TTCR_LD(up, aTup, s1, TROW)
                    TTCR_LD(sl, aS, s0, SROW)
                    TTCR_LD(kmh, aTvm, sm1, TROW)
                    if (first_w) { TTCR_LD(um, aTlo, s0, TROW) }
                    else um = (q & 1) ? lds_at<-XW>(aX, 0.f) : lds_at<XS - XW>(aX, 0.f);
#undef TTCR_LD
                    if (!has_um) um = MAXV;
                    if (!has_up) up = MAXV;
                    if (lane_hi && !has_right) kp = MAXV;
                    float km = __shfl_up_sync(0xffffffffu, t_prev, 1);
                    if (lane_lo) km = has_v ? kmh : MAXV;
                    const float t = godunov(tmin(km, kp), tmin(t_prev, jp), tmin(um, up), sl * dx);
                    bool valid = (unsigned)(jo0 + s) < (unsigned)nj;
                    if (any_fz) {
                        if (valid) {
                            const long long e = (long long)(stb + off - tt);
                            if ((frozen[e >> 5] >> (e & 31)) & 1u) valid = false;
                        }
                    }
                    float tnew = told;
                    if (valid && t < told) {
                        tnew = t;
                        st_stream(stb + off, t);
                        acc += told - t;
                    }
                    t_prev = tnew;
                    told = jp;
                    off = pin(off + sm32);
                    if (q & 1) sts_at<XS>(aX, tnew); else sts_at<0>(aX, tnew);
                    if (q == RING - 1)
                        dead = bar_compute_or<NC>(sm_abort);
                    else
                        bar_compute<NC>();
                    if (threadIdx.x == 0) sm_step_done = s;
                    {
                        const int rd = s - NU + 2;
                        if (rd > 0 && rd < nrows && (rd & chunk_mask) == 0) {
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
            for (; narrive <= nch; ++narrive) {
                if (narrive >= 2) {
                    long long spins = 0;
                    while (sm_pubseq < narrive - 1 && ++spins < p.spin_limit) {}
                }
                bar_pub_arrive<NP>(2 + (narrive & 1));
            }
            if (threadIdx.x == 0) sm_step_done = 1 << 29;   // let the loader run out its remaining chunks
            double dacc = (double)acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
            if (lane == 0) sred[wq] = dacc;
            bar_compute<NC>();
            if (threadIdx.x == 0) {
                double ssum = 0.0;
                for (int i = 0; i < NW; ++i) ssum += sred[i];
                p.partial[tile] = ssum;
                if (p.trace) p.trace[tile * 8 + 5] = gtime();
            }
        }
        __syncthreads();
    }
}

template <int NW>
inline int tile3_launch(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                        const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change,
                        cudaStream_t st) {
    constexpr int NU = NW;
    TileParams p;
    p.w = w; p.d = d; p.fb = fb;
    p.nV = d.kpad / 32;
    p.nU = (w.nu + NU - 1) / NU;
    p.ntiles = p.nU * p.nV;
    p.chunk = 1;
    while (p.chunk * 2 <= o.chunk) p.chunk *= 2;
    p.spin_limit = o.spin_limit;
    p.order = s.d_order; p.flags = s.d_flags; p.ctrl = s.d_ctrl; p.partial = s.d_partial;
    static long long* d_trace = nullptr;
    const char* trace_path = getenv("TTCR_B200_TRACE");
    if (trace_path && !d_trace) TCK(cudaMalloc(&d_trace, (size_t)s.cap_tiles * 8 * sizeof(long long)));
    p.trace = trace_path ? d_trace : nullptr;
    const int key = 3000000 + NU * 1000 + p.chunk;
    if (s.order_key != key || s.ntiles != p.ntiles) {
        std::vector<std::pair<long long, int>> k(p.ntiles);
        const long long lag_u = NU + 12 + p.chunk + 4;
        for (int U = 0; U < p.nU; ++U)
            for (int V = 0; V < p.nV; ++V) {
                const int va = std::max(V * 32, w.vlo);
                k[U * p.nV + V] = {U * lag_u + (long long)(va - w.joff), U * p.nV + V};
            }
        std::stable_sort(k.begin(), k.end());
        std::vector<int> order(p.ntiles);
        for (int i = 0; i < p.ntiles; ++i) order[i] = k[i].second;
        TCK(cudaMemcpyAsync(s.d_order, order.data(), p.ntiles * sizeof(int), cudaMemcpyHostToDevice, st));
        TCK(cudaStreamSynchronize(st));
        s.order_key = key;
        s.ntiles = p.ntiles;
    }
    TCK(cudaMemsetAsync(s.d_flags, 0, p.ntiles * sizeof(int), st));
    TCK(cudaMemsetAsync(s.d_ctrl, 0, sizeof(int), st));
    static int occ_cache = 0;
    if (!occ_cache) {
        TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, k_sweep_tile3<NW>, (NW + 2) * 32, 0));
        if (occ_cache < 1) throw std::runtime_error("tile3 kernel does not fit on an SM");
    }
    int occ = occ_cache;
    if (o.ctas_per_sm > 0) occ = std::min(occ, o.ctas_per_sm);
    const int grid = std::min(p.ntiles, occ * sm_count);
    k_sweep_tile3<NW><<<grid, (NW + 2) * 32, 0, st>>>(p, tt, slo, frozen, dx);
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

// fp32, first-order only
template <typename T> inline bool tile3_supported(bool weno_stage) { return false; }
template <> inline bool tile3_supported<float>(bool weno_stage) { return !weno_stage; }

template <typename T>
inline int tile3_sweep(TileState&, const TileOptions&, int, const SweepView&, const Dims&, T*, const T*, const uint32_t*,
                       const FrozenBox&, T, double*, cudaStream_t) {
    throw std::runtime_error("tile3 kernel: fp32 only");
}
template <>
inline int tile3_sweep<float>(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                              const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change,
                              cudaStream_t st) {
    if (o.warps == 4) return tile3_launch<4>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    return tile3_launch<8>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
}

}  // namespace ttcrb200
