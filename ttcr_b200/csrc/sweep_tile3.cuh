// Sweep kernel "TILE3": the tile-marching sweep of sweep_tile.cuh with TMA-fed, warp-specialised data
// movement.
//
// Same decomposition, dependency protocol and arithmetic as k_sweep_tile (read its header first); what
// changes is who moves the data.  In k_sweep_tile every compute thread streams its own rows through
// register queues, and a step of a compute warp is ~170 mostly serial instructions (address arithmetic,
// queue rotation, five global loads) of which ~40 are the Godunov update.  Here:
//
//   * a LOADER thread issues, per chunk of 8 rows, TWO tensor-map TMA copies (cp.async.bulk.tensor.3d,
//     SASS UTMALDG) into a shared-memory ring: a box of {40 lanes, 8 rows, NU+2 planes} of traveltimes
//     (the tile's NU rows of u, rows u0-1 and u0+NU, and the v0-1 / v0+32 halo lanes, so every halo the
//     tile needs arrives in the same box) and a box of {32 lanes, 8 rows, NU planes} of slowness; both
//     complete on one mbarrier.  The loader is also the poller of the two upstream flags: it issues a
//     chunk only when the tiles U-1 and V-1 have published the rows its halos hold.
//     (One 1-D bulk copy per row and stream was tried first: 18 streams x 128..160 B per step made the
//     TMA unit the bottleneck at ~70 cycles per request.)
//   * the boxes are rectangular, so all rows of u are loaded for the SAME row range.  The Gauss-Seidel
//     skew is kept by delaying the warps instead of the data: compute warp j (row u0+j) starts j steps
//     after warp 0, so at every moment consecutive u still lag one step.  A warp at its local step a
//     works on row m_first+a and reads from the ring: its own next row (jp) and the v+1 lane of it (kp),
//     the same row of plane u+1 (up), slowness, the v-1 halo of the previous row, and the (u-1) result
//     of the neighbouring warp from a small exchange buffer; one shuffle; godunov(); one predicated
//     global store; one named barrier per step.  No global loads, no register queues.
//   * the PUBLISHER warp is unchanged (bar.arrive hand-off, release fence, flag store).
//
// Ring geometry: NCH chunk slots of C = 8 rows.  Row group G (= local row + 1; group 0 is local row -1,
// needed by the v-1 halo of step 0) lives in chunk slot (G / 8) % NCH at box row G % 8 (7 - G % 8 when the
// sweep runs the row axis downwards, since a box arrives in memory order; likewise for planes and lanes).
//
// fp32 only; other types use k_sweep_tile.
#pragma once
#include <cuda.h>

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
// 3-D tensor-map TMA load: box at element coordinates (x, y, z) -> shared, completing on an mbarrier
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned mbar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(mbar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ float lds_f(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

struct Tile3Geom {
    static constexpr int C = 8;      // rows per chunk (one TMA box)
    static constexpr int TW = 40;    // floats per traveltime row: 4 + 32 + 4
};

template <int NW, int NCH>
struct Tile3Layout {
    static constexpr int C = Tile3Geom::C, TW = Tile3Geom::TW;
    static constexpr int TROW = TW * 4, SROW = 32 * 4;               // bytes per ring row
    static constexpr int TPL = C * TROW, SPL = C * SROW;             // bytes per plane of a box
    static constexpr int CHB_T = (NW + 1) * TPL, CHB_S = NW * SPL;   // bytes per chunk slot (one box)
    static constexpr int HC = 2, HN = 8;                             // halo plane u0-1: HN chunk slots of HC rows
    static constexpr int CHB_H = HC * SROW;                          // bytes per halo chunk (32 lanes x HC rows)
    static constexpr int OFF_T = 0;
    static constexpr int OFF_S = OFF_T + NCH * CHB_T;
    static constexpr int OFF_H = OFF_S + NCH * CHB_S;
    static constexpr int OFF_X = OFF_H + HN * CHB_H;                 // exchange buffer: 2 x NW x 32 floats
    static constexpr int OFF_BAR = OFF_X + 2 * NW * 32 * 4;          // NCH + HN mbarriers
    static constexpr int BYTES = OFF_BAR + (NCH + HN) * 8;
    static_assert(CHB_H % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(CHB_T % 128 == 0 && CHB_S % 128 == 0, "TMA destinations must stay 128-byte aligned");
};

template <int NW, int NCH>
__global__ void __launch_bounds__((NW + 2) * 32) k_sweep_tile3(const __grid_constant__ CUtensorMap tmT,
                                                              const __grid_constant__ CUtensorMap tmS,
                                                              const __grid_constant__ CUtensorMap tmH, TileParams p,
                                                              float* __restrict__ tt, const uint32_t* __restrict__ frozen,
                                                              float dx) {
    using L = Tile3Layout<NW, NCH>;
    constexpr int C = L::C;
    constexpr int NU = NW, NC = NW * 32, NP = (NW + 1) * 32;
    static_assert(NW + C + 8 <= GUARD, "guard rows too few");
    static_assert(C * NCH - 6 - NW - 2 >= 1, "ring too shallow: the loader could never run ahead of the compute warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double sred[NW];
    __shared__ int sm_tile;
    __shared__ volatile int sm_pubseq, sm_abort, sm_step_done;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const float MAXV = FLT_MAX;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned a_full = sbase + L::OFF_BAR;
    const unsigned a_fullh = a_full + 8 * NCH;
    constexpr int HC = L::HC, HN = L::HN;
    bool first_tile = true;

    for (;;) {
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sm_tile = (ab || t >= p.ntiles) ? -1 : t;
            sm_pubseq = 0; sm_abort = 0; sm_step_done = -1;
            // fresh barriers for every tile: chunk c uses slot c % NCH with parity (c / NCH) & 1
            for (int c = 0; c < NCH + HN; ++c) {
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
        const int nrows = (vb - va) + w.nj - 1;                // local rows 0 .. nrows-1
        const int nglobal = nrows + NU - 1;                    // steps of the tile (warp j runs steps j .. j+nrows-1)
        const int nchunks = (nrows + 2 + C - 1) / C;           // groups 0 .. nrows+1
        const bool has_u = U > 0, has_v = V > 0;
        const bool has_right = v0 + 32 < w.vhi;
        const int va_p = max(v0 - 32, w.vlo);
        const int nrows_p = (v0 - va_p) + w.nj - 1;
        const int dmf = m_first - (va_p - w.joff);
        const int chunk = p.chunk;
        const int nch = (nrows - 1) / chunk;
        const int ulast = w.nu - 1;
        if (p.trace && threadIdx.x == 0) {
            p.trace[tile * 8 + 0] = gtime();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[tile * 8 + 7] = smid;
        }

        if (warp == NW) {
            // ================= loader (and poller): one thread =================
            if (lane == 0) {
                // box origins in MEMORY coordinates (x: lane k, y: row within the padded plane, z: plane i)
                const int xT = w.rk ? p.d.kpad - 36 - v0 : v0 - 4;
                const int xS = w.rk ? p.d.kpad - 32 - v0 : v0;
                const int zT = w.ri ? p.d.ni - 1 - (u0 + NU) : u0;          // planes u0 .. u0+NU
                const int zS = w.ri ? p.d.ni - 1 - (u0 + NU - 1) : u0;
                const int zH = w.ri ? p.d.ni - 1 - (u0 - 1) : u0 - 1;       // plane u0-1
                const int* fu = &p.flags[has_u ? tile - p.nV : tile];
                const int* fv = &p.flags[has_v ? tile - 1 : tile];
                int ku = has_u ? 0 : (1 << 30), kv = has_v ? 0 : (1 << 30);
                const int nhch = has_u ? (nrows + HC - 1) / HC : 0;         // halo chunks: local rows 0 .. nrows-1
                int c = 0, hc = 0;
                bool dead = false;
                long long t0 = clock64();
                while ((c < nchunks || hc < nhch) && !dead) {
                    bool progress = false;
                    const int done = sm_step_done;
                    // ---- main chunk c: groups 8c .. 8c+7 = local rows 8c-1 .. 8c+6 of planes u0 .. u0+NU (old values,
                    //      except the v0-1 halo lanes, which hold new values of tile V-1: row a needs its count a+dmf+1)
                    if (c < nchunks) {
                        const int g0 = c * C;
                        // slot reuse: chunk c-NCH held local rows up to g0-C*NCH+6, last read (v-1 halo by the last warp)
                        // one step after that warp worked on the row: global step row + 1 + (NU-1)
                        bool ok = done >= g0 - C * NCH + 6 + NU;
                        const int need_v = min(g0 + C - 1 + dmf, nrows_p);
                        if (ok && kv < need_v) {
                            const int k2 = ld_relaxed_gpu(fv);
                            if (k2 > kv) { kv = k2; fence_acq_rel_gpu(); fence_proxy_async(); }
                            ok = kv >= need_v;
                        }
                        if (ok) {
                            const int mlo = m_first - 1 + g0;   // oriented rows mlo .. mlo+7 -> memory rows (ascending)
                            const int y = GUARD + (w.rj ? (w.nm - 1) - (mlo + C - 1) : mlo);
                            const unsigned mb = a_full + 8 * (c % NCH);
                            const unsigned slot = (unsigned)(c % NCH);
                            mbar_expect_tx(mb, L::CHB_T + L::CHB_S);
                            tma_load_3d(sbase + L::OFF_T + slot * L::CHB_T, &tmT, xT, y, zT, mb);
                            tma_load_3d(sbase + L::OFF_S + slot * L::CHB_S, &tmS, xS, y, zS, mb);
                            ++c;
                            progress = true;
                        }
                    }
                    // ---- halo chunk hc: local rows 2hc, 2hc+1 of plane u0-1 (new values of tile U-1: needs its count 2hc+2)
                    if (hc < nhch) {
                        const int a0 = hc * HC;
                        bool ok = done >= a0 + HC - 1 - HC * HN;            // row a is read at global step a (warp 0) only
                        const int need_u = min(a0 + HC, nrows);
                        if (ok && ku < need_u) {
                            const int k2 = ld_relaxed_gpu(fu);
                            if (k2 > ku) { ku = k2; fence_acq_rel_gpu(); fence_proxy_async(); }
                            ok = ku >= need_u;
                        }
                        if (ok) {
                            const int mlo = m_first + a0;
                            const int y = GUARD + (w.rj ? (w.nm - 1) - (mlo + HC - 1) : mlo);
                            const unsigned mb = a_fullh + 8 * (hc % HN);
                            mbar_expect_tx(mb, L::CHB_H);
                            tma_load_3d(sbase + L::OFF_H + (unsigned)(hc % HN) * L::CHB_H, &tmH, xS, y, zH, mb);
                            ++hc;
                            progress = true;
                        }
                    }
                    if (progress) {
                        t0 = clock64();
                    } else {
                        if (sm_abort) { dead = true; break; }
                        if (clock64() - t0 > (p.spin_limit << 9)) {
                            if (atomicCAS(&p.ctrl[1], 0, 10) == 0) {
                                p.ctrl[2] = tile; p.ctrl[3] = c; p.ctrl[4] = sm_step_done; p.ctrl[5] = ku; p.ctrl[6] = kv;
                            }
                            sm_abort = 1;
                            dead = true;
                            break;
                        }
                    }
                }
                // every copy must have landed before the barriers are re-initialised for the next tile
                if (!dead) {
                    for (int cc = max(0, nchunks - NCH); cc < nchunks; ++cc) {
                        const long long t1 = clock64();
                        while (!mbar_test(a_full + 8 * (cc % NCH), (cc / NCH) & 1) && clock64() - t1 < (p.spin_limit << 9)) {}
                    }
                    for (int cc = max(0, nhch - HN); cc < nhch; ++cc) {
                        const long long t1 = clock64();
                        while (!mbar_test(a_fullh + 8 * (cc % HN), (cc / HN) & 1) && clock64() - t1 < (p.spin_limit << 9)) {}
                    }
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
            // ================= compute warps: warp wq owns row u0 + wq and runs wq steps behind warp 0 =====
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
            // position of this thread inside a box (boxes arrive in memory order)
            const int col = w.rk ? 35 - lane : 4 + lane;     // lane column in a 40-float row
            const int dcol = w.rk ? -1 : 1;                  // column step towards lane v+1
            const int pT = w.ri ? NU - wq : wq;              // plane slot of row u0+wq in the traveltime box (planes u0 .. u0+NU)
            const int dpl = w.ri ? -1 : 1;                   // plane step towards u+1
            const int pS = w.ri ? NU - 1 - wq : wq;
            const unsigned bT = sbase + L::OFF_T + pT * L::TPL + col * 4;
            const unsigned bTvp = bT + dcol * 4, bTvm = bT - dcol * 4;
            const unsigned bTup = bT + dpl * L::TPL;         // plane u+1
            const unsigned bH = sbase + L::OFF_H + (w.rk ? 31 - lane : lane) * 4;   // plane u0-1 ring (32-lane rows)
            const int hflip = w.rj ? HC - 1 : 0;
            const unsigned bS = sbase + L::OFF_S + pS * L::SPL + (w.rk ? 31 - lane : lane) * 4;
            const unsigned aX = sbase + L::OFF_X + (wq * 32 + lane) * 4;
            constexpr int XS = NW * 32 * 4, XW = 32 * 4;
            const int rflip = w.rj ? C - 1 : 0;
            // ring offsets of row group G
            auto offT = [&](int G) -> unsigned { return (unsigned)(((G >> 3) % NCH) * L::CHB_T + ((G & 7) ^ rflip) * L::TROW); };
            auto offS = [&](int G) -> unsigned { return (unsigned)(((G >> 3) % NCH) * L::CHB_S + ((G & 7) ^ rflip) * L::SROW); };
            // global store pointer of local row 0 plus a running element offset
            const long long e0 = w.base + (long long)min(u, ulast) * w.su + (long long)v * w.sv + (long long)m_first * w.sm;
            float* const stb = pin_ptr(tt + e0);
            int off = pin(0);
            const int jo0 = (u_ok && v_ok) ? m_first - v + w.joff : -(1 << 30);
            bool any_fz;
            {
                const int it = w.ri ? ulast - min(u, ulast) : min(u, ulast);
                const int ka0 = va - w.vlo, kb0 = vb - 1 - w.vlo;
                const int ka = w.rk ? p.d.nk - 1 - kb0 : ka0, kb = w.rk ? p.d.nk - 1 - ka0 : kb0;
                any_fz = it >= p.fb.ilo && it <= p.fb.ihi && !(kb < p.fb.klo || ka > p.fb.khi);
            }
            int gstep = 0;                                   // global step counter of this thread
            auto wait_bar = [&](unsigned a, unsigned par, int why) {
                if (mbar_test(a, par)) return;
                const long long t0 = clock64();
                while (!mbar_test(a, par)) {   // each attempt suspends the thread for a bounded time
                    if (sm_abort) break;
                    if (clock64() - t0 > (p.spin_limit << 9)) {   // ~1 s at the default spin_limit
                        if (atomicCAS(&p.ctrl[1], 0, why) == 0) { p.ctrl[2] = tile; p.ctrl[3] = (int)a; p.ctrl[4] = gstep; p.ctrl[5] = wq; p.ctrl[6] = sm_step_done; }
                        sm_abort = 1;
                        break;
                    }
                }
            };
            auto wait_chunk = [&](int c) { wait_bar(a_full + 8 * (c % NCH), (c / NCH) & 1, 20); };
            int dead = 0;
            int narrive = 0;
            // one barrier per global step, then the hand-off bookkeeping of that step
            auto end_step = [&]() {
                dead |= bar_compute_or<NC>(sm_abort);
                if (threadIdx.x == 0) sm_step_done = gstep;
                const int rd = gstep - NU + 2;
                if (rd > 0 && rd < nrows && (rd & chunk_mask) == 0) {
                    if (narrive >= 2) {
                        long long spins = 0;
                        while (sm_pubseq < narrive - 1 && !sm_abort && ++spins < p.spin_limit) {}
                    }
                    bar_pub_arrive<NP>(2 + (narrive & 1));
                    ++narrive;
                }
                ++gstep;
            };

            // ---- warps start one step apart
            for (int k = 0; k < wq && !dead; ++k) end_step();
            wait_chunk(0);                                   // groups 0..7 = local rows -1..6
            float told = lds_f(bT + offT(1));                // old value of local row 0
            float t_prev = MAXV;
            float acc = 0.f;
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 1] = gtime();

            for (int a = 0; a < nrows && !dead; ++a) {
                if (p.trace && threadIdx.x == 0) {
                    if (a == nrows / 4) p.trace[tile * 8 + 2] = gtime();
                    if (a == nrows / 2) p.trace[tile * 8 + 3] = gtime();
                    if (a == (3 * nrows) / 4) p.trace[tile * 8 + 4] = gtime();
                }
                // group a+2 (local row a+1) is read from here on: entering a new chunk?
                if (((a + 2) & (C - 1)) == 0) wait_chunk((a + 2) >> 3);
                const unsigned o1 = offT(a + 2), o0 = offT(a + 1), om = offT(a), os = offS(a + 1);
                const float jp = lds_f(bT + o1);             // old (u, a+1, v)
                float kp = lds_f(bTvp + o1);                 // old (u, a+1, v+1)
                float up = lds_f(bTup + o0);                 // old (u+1, a, v)
                const float sl = lds_f(bS + os);
                float um;
                if (first_w) {                               // new (u0-1, a, v): halo of tile U-1
                    um = MAXV;
                    if (has_u) {
                        const int hcx = a / HC;
                        if ((a % HC) == 0) wait_bar(a_fullh + 8 * (hcx % HN), (hcx / HN) & 1, 21);
                        um = lds_f(bH + (unsigned)((hcx % HN) * L::CHB_H + (((a % HC) ^ hflip)) * L::SROW));
                    }
                }
                else um = lds_f(aX - XW + ((gstep & 1) ? 0 : XS));   // result of warp wq-1 at the previous step
                if (!has_um) um = MAXV;
                if (!has_up) up = MAXV;
                if (lane_hi && !has_right) kp = MAXV;
                float km = __shfl_up_sync(0xffffffffu, t_prev, 1);
                if (lane_lo) km = has_v ? lds_f(bTvm + om) : MAXV;   // new (u, a-1, v0-1): halo of tile V-1
                const float t = godunov(tmin(km, kp), tmin(t_prev, jp), tmin(um, up), sl * dx);
                bool valid = (unsigned)(jo0 + a) < (unsigned)nj;
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
                sts_f(aX + ((gstep & 1) ? XS : 0), tnew);
                end_step();
            }
            // ---- trailing steps of the warps that started earlier, then any arrivals skipped by an abort
            while (gstep < nglobal && !dead) end_step();
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

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qr;
        TCK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr));
        if (!ptr || qr != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled is not available");
        fn = (PFN_tmapEncodeTiled)ptr;
    }
    return fn;
}

// 3-D map over a sheared layout array: dims (kpad lanes, qs rows, ni planes), box (bw lanes, br rows, bp planes)
inline CUtensorMap make_tile3_map(const void* base, const Dims& d, int bw, int br, int bp) {
    CUtensorMap m;
    const cuuint64_t gdim[3] = {(cuuint64_t)d.kpad, (cuuint64_t)d.qs, (cuuint64_t)d.ni};
    const cuuint64_t gstr[2] = {(cuuint64_t)d.kpad * 4, (cuuint64_t)d.kpad * d.qs * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)br, (cuuint32_t)bp};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = tmap_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

template <int NW, int NCH>
inline int tile3_launch(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                        const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change,
                        cudaStream_t st) {
    using L = Tile3Layout<NW, NCH>;
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
        const long long lag_u = NU + Tile3Geom::C + p.chunk + 4;
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
    // tensor maps are cached per (array, box shape)
    struct MapKey { const void* a; int bw, br, bp, kpad, qs, ni; CUtensorMap m; };
    static thread_local std::vector<MapKey> cache;
    auto get_map = [&](const void* a, int bw, int br, int bp) -> CUtensorMap {
        for (auto& e : cache)
            if (e.a == a && e.bw == bw && e.br == br && e.bp == bp && e.kpad == d.kpad && e.qs == d.qs && e.ni == d.ni) return e.m;
        if (cache.size() > 96) cache.clear();
        cache.push_back({a, bw, br, bp, d.kpad, d.qs, d.ni, make_tile3_map(a, d, bw, br, bp)});
        return cache.back().m;
    };
    const CUtensorMap tmT = get_map(tt, Tile3Geom::TW, Tile3Geom::C, NU + 1);
    const CUtensorMap tmS = get_map(slo, 32, Tile3Geom::C, NU);
    const CUtensorMap tmH = get_map(tt, 32, L::HC, 1);
    TCK(cudaMemsetAsync(s.d_flags, 0, p.ntiles * sizeof(int), st));
    TCK(cudaMemsetAsync(s.d_ctrl, 0, sizeof(int), st));
    static int occ_cache = 0;
    if (!occ_cache) {
        TCK(cudaFuncSetAttribute(k_sweep_tile3<NW, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
        TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, k_sweep_tile3<NW, NCH>, (NW + 2) * 32, L::BYTES));
        if (occ_cache < 1) throw std::runtime_error("tile3 kernel does not fit on an SM");
    }
    int occ = occ_cache;
    if (o.ctas_per_sm > 0) occ = std::min(occ, o.ctas_per_sm);
    const int grid = std::min(p.ntiles, occ * sm_count);
    k_sweep_tile3<NW, NCH><<<grid, (NW + 2) * 32, L::BYTES, st>>>(tmT, tmS, tmH, p, tt, frozen, dx);
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
template <typename T> inline bool tile3_supported(bool) { return false; }
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
    if (o.warps == 4) return tile3_launch<4, 4>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    // ring depth: the loader must be able to run ahead, 8*NCH - 6 - NU - 2 >= 1 (see need_done in the kernel)
    if (o.warps >= 16) return tile3_launch<16, 4>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.depth <= 4) return tile3_launch<8, 3>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    return tile3_launch<8, 4>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
}

}  // namespace ttcrb200
