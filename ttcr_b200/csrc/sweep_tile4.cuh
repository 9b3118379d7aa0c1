// Sweep kernel "TILE4": the TMA-fed tile march of sweep_tile3.cuh with a flag-free, fence-free hand-off
// between tiles ("mailbox" protocol).
//
// In k_sweep_tile / k_sweep_tile3 a tile tells its U+1 and V+1 neighbours how many rows it has finished
// through a progress flag: store results, release fence, store flag | poll flag, acquire fence, load.
// Traces showed the hand-off (two fences at ~1 us each, chunked publication, a TMA round trip) costing
// 13-20 us per U hop against ~4 us of arithmetic; with 64 U hops and 15 V hops on the critical path of a
// 512^3 sweep that was two thirds of the sweep time.
//
// Here the only values a tile needs NEW from a neighbour -- the u0-1 row (32 lanes per step) and the v0-1
// lane (one value per warp and step) -- travel as 64-bit words {sweep serial : 32, float bits : 32}
// written with ONE 8-byte store into a mailbox slot that is private to (tile, row, lane).  An aligned
// 8-byte access is single-copy atomic, so the reader that finds the current serial in the upper half has
// the value in the lower half: no flag, no fence, no chunking, and the producer never waits (every slot
// is written once per sweep; the serial is the launch counter, so nothing is ever cleared).  Readers
// prefetch their slots three steps ahead and only re-poll when the tag is stale.
//
// Everything else a step reads is an OLD value (own rows, plane u0+NU, lane v0+32), which the neighbours
// cannot overwrite before this tile has published the row that depends on it -- so the TMA boxes need no
// gating at all, and the poller / publisher warps are gone: NW compute warps + one loader thread.
//
// fp32, first-order stage only.
#pragma once
#include "sweep_tile3.cuh"

namespace ttcrb200 {

__device__ __forceinline__ unsigned long long ld_mail(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_mail(unsigned long long* p, unsigned serial, float t) {
    const unsigned long long v = ((unsigned long long)serial << 32) | (unsigned long long)__float_as_uint(t);
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// predicated forms (no divergent branch around a one-lane access)
__device__ __forceinline__ void ld_mail_if(unsigned long long& v, const unsigned long long* p, int on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p ld.relaxed.gpu.global.b64 %0, [%1];\n\t}" : "+l"(v) : "l"(p), "r"(on) : "memory");
}
__device__ __forceinline__ void st_mail_if(unsigned long long* p, unsigned serial, float t, int on) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 x;\n\tsetp.ne.s32 p, %3, 0;\n\tmov.b64 x, {%1, %2};\n\t@p st.relaxed.gpu.global.b64 [%0], x;\n\t}" ::"l"(p),
        "f"(t), "r"(serial), "r"(on)
        : "memory");
}
__device__ __noinline__ bool frozen_bit(const uint32_t* frozen, long long e) { return (frozen[e >> 5] >> (e & 31)) & 1u; }

struct Tile4Mail {
    unsigned long long* u = nullptr;   // [tile][MARGIN + row][32 lanes]: results of the tile's last row of u
    unsigned long long* v = nullptr;   // [tile][MARGIN + row][NW warps]: results of lane 31
    int rows = 0;                      // row stride per tile (margins included: prefetches run past the ends)
    unsigned serial = 0;               // sweep launch counter = tag of the current sweep
    static constexpr int MARGIN = 2;
};

template <int NW, int NCH>
struct Tile4Layout {
    static constexpr int C = 8, TW = 36;                             // rows per chunk; floats per traveltime row: 32 + 4
    static constexpr int TROW = TW * 4, SROW = 32 * 4;
    static constexpr int TPL = C * TROW, SPL = C * SROW;             // bytes per plane of a box
    static constexpr int CHB_T = (NW + 1) * TPL, CHB_S = NW * SPL;   // bytes per chunk slot (one box)
    static constexpr int OFF_T = 0;
    static constexpr int OFF_S = OFF_T + NCH * CHB_T;
    static constexpr int OFF_X = OFF_S + NCH * CHB_S;                // exchange buffer: 2 x NW x 32 floats
    static constexpr int OFF_BAR = OFF_X + 2 * NW * 32 * 4;          // NCH mbarriers
    static constexpr int OFF_FLG = OFF_BAR + NCH * 8;                // abort, step_done, tile (32-bit each)
    static constexpr int OFF_RED = OFF_FLG + 16;                     // NW doubles
    static constexpr int BYTES = OFF_RED + NW * 8;
    static_assert(CHB_T % 128 == 0 && CHB_S % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(OFF_RED % 8 == 0, "alignment");
};

__device__ __forceinline__ int lds_i(unsigned a) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_i(unsigned a, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
template <int I> struct IntC { static constexpr int value = I; };

// RJ: the sweep runs the row axis downwards (rows of a box arrive in memory order, i.e. reversed)
template <int NW, int NCH, bool RJ>
__global__ void __launch_bounds__((NW + 1) * 32, (NW <= 8 ? 3 : 1)) k_sweep_tile4(const __grid_constant__ CUtensorMap tmT,
                                                              const __grid_constant__ CUtensorMap tmS, TileParams p,
                                                              Tile4Mail mail, float* __restrict__ tt,
                                                              const uint32_t* __restrict__ frozen, float dx) {
    using L = Tile4Layout<NW, NCH>;
    constexpr int C = L::C;
    constexpr int NU = NW, NC = NW * 32;
    static_assert(C * NCH - 6 - NW - 2 >= 1, "ring too shallow: the loader could never run ahead of the compute warps");
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const float MAXV = FLT_MAX;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned a_full = sbase + L::OFF_BAR;
    const unsigned a_abort = sbase + L::OFF_FLG, a_done = a_abort + 4, a_tile = a_abort + 8;
    const unsigned serial = mail.serial;
    bool first_tile = true;

    for (;;) {
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sts_i(a_tile, (ab || t >= p.ntiles) ? -1 : t);
            sts_i(a_abort, 0);
            sts_i(a_done, -1);
            for (int c = 0; c < NCH; ++c) {     // fresh barriers per tile: chunk c -> slot c % NCH, parity (c / NCH) & 1
                if (!first_tile) mbar_inval(a_full + 8 * c);
                mbar_init(a_full + 8 * c, 1);
            }
            fence_mbar_init();
        }
        first_tile = false;
        __syncthreads();
        const int ticket = lds_i(a_tile);
        if (ticket < 0) break;
        const int tile = p.order[ticket];
        const int U = tile / p.nV, V = tile - U * p.nV;
        const int u0 = U * NU, v0 = V * 32;
        const int va = max(v0, w.vlo), vb = min(v0 + 32, w.vhi);
        const int m_first = va - w.joff;
        const int nrows = (vb - va) + w.nj - 1;                // local rows 0 .. nrows-1
        const int nglobal = nrows + NU - 1;                    // steps of the tile (warp j runs steps j .. j+nrows-1)
        const int nchunks = nrows / C + 1;                     // chunk c = local rows 8c .. 8c+7; rows 0 .. nrows are read
        const bool has_u = U > 0, has_v = V > 0;
        const bool has_right = v0 + 32 < w.vhi;
        const bool has_down = U + 1 < p.nU;
        const int va_p = max(v0 - 32, w.vlo);
        const int nrows_p = (v0 - va_p) + w.nj - 1;            // rows of tile V-1
        const int dmf = m_first - (va_p - w.joff);             // its local row index of my local row 0
        const int ulast = w.nu - 1;
        if (p.trace && threadIdx.x == 0) {
            p.trace[tile * 8 + 0] = gtime();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[tile * 8 + 7] = smid;
        }

        if (warp == NW) {
            // ================= loader: one thread, never waits for another tile =================
            if (lane == 0) {
                // box origins in MEMORY coordinates (x: lane k, y: row within the padded plane, z: plane i)
                const int xT = w.rk ? p.d.kpad - 36 - v0 : v0;              // lanes v0 .. v0+35 (oriented)
                const int xS = w.rk ? p.d.kpad - 32 - v0 : v0;
                const int zT = w.ri ? p.d.ni - 1 - (u0 + NU) : u0;          // planes u0 .. u0+NU
                const int zS = w.ri ? p.d.ni - 1 - (u0 + NU - 1) : u0;
                bool dead = false;
                long long t0 = clock64();
                for (int c = 0; c < nchunks && !dead;) {
                    const int g0 = c * C;
                    // slot reuse: chunk c-NCH held rows up to g0-C*NCH+7, last read (row a of plane u+1, slowness) by
                    // the last warp at global step a + NU - 1
                    if (lds_i(a_done) >= g0 - C * NCH + 6 + NU) {
                        const int mlo = m_first + g0;       // oriented rows mlo .. mlo+7 -> memory rows (ascending)
                        const int y = GUARD + (RJ ? (w.nm - 1) - (mlo + C - 1) : mlo);
                        const unsigned mb = a_full + 8 * (c % NCH);
                        const unsigned slot = (unsigned)(c % NCH);
                        mbar_expect_tx(mb, L::CHB_T + L::CHB_S);
                        tma_load_3d(sbase + L::OFF_T + slot * L::CHB_T, &tmT, xT, y, zT, mb);
                        tma_load_3d(sbase + L::OFF_S + slot * L::CHB_S, &tmS, xS, y, zS, mb);
                        ++c;
                        t0 = clock64();
                    } else {
                        if (lds_i(a_abort)) { dead = true; break; }
                        if (clock64() - t0 > (p.spin_limit << 9)) {
                            if (atomicCAS(&p.ctrl[1], 0, 10) == 0) { p.ctrl[2] = tile; p.ctrl[3] = c; p.ctrl[4] = lds_i(a_done); }
                            sts_i(a_abort, 1);
                            dead = true;
                        }
                    }
                }
                // every copy must have landed before the barriers are re-initialised for the next tile
                if (!dead)
                    for (int cc = max(0, nchunks - NCH); cc < nchunks; ++cc) {
                        const long long t1 = clock64();
                        while (!mbar_test(a_full + 8 * (cc % NCH), (cc / NCH) & 1) && clock64() - t1 < (p.spin_limit << 9)) {}
                    }
            }
            __syncwarp();
        } else {
            // ================= compute warps: warp wq owns row u0 + wq and runs wq steps behind warp 0 =====
            const int wq = warp;
            const int u = u0 + wq;
            const int v = v0 + lane;
            const bool u_ok = u <= ulast;
            const bool v_ok = v >= w.vlo && v < w.vhi;
            const bool first_w = wq == 0;
            const bool has_up = u < ulast;                   // a (u+1) row exists
            const bool lane_lo = lane == 0;
            const bool kill_kp = lane == 31 && !has_right;
            const bool mail_u_in = first_w && has_u;         // this warp reads the U mailbox of tile U-1
            const int mail_v_in = pin((lane_lo && has_v) ? 1 : 0);          // this lane reads the V mailbox of tile V-1
            const int mail_u_out = pin((wq == NW - 1 && has_down) ? 1 : 0);
            const int mail_v_out = pin((lane == 31 && has_right) ? 1 : 0);
            const int nj = pin(w.nj);
            const int nrows_r = pin(nrows);
            const unsigned sb = (unsigned)pin((int)sbase);   // shared window base, kept in a register
            const unsigned a_ab = sb + L::OFF_FLG, a_dn = a_ab + 4;
            // position of this thread inside a box (boxes arrive in memory order)
            const int col = w.rk ? 35 - lane : lane;         // lane column in a 36-float row
            const int dvp = pin(w.rk ? -4 : 4);              // byte step towards lane v+1
            const int pT = w.ri ? NU - wq : wq;              // plane slot of row u0+wq in the traveltime box (planes u0 .. u0+NU)
            const int dup = pin(w.ri ? -L::TPL : L::TPL);    // byte step towards plane u+1
            const int pS = w.ri ? NU - 1 - wq : wq;
            const unsigned tT0 = (unsigned)pin((int)(sb + L::OFF_T + pT * L::TPL + col * 4));
            const unsigned tS0 = (unsigned)pin((int)(sb + L::OFF_S + pS * L::SPL + (w.rk ? 31 - lane : lane) * 4));
            // exchange buffer (u-1 results of the neighbouring warp), double-buffered by ROW parity: the reader of row a
            // is exactly one global step behind its writer
            constexpr int XS = NW * 32 * 4, XW = 32 * 4;
            const unsigned aX = (unsigned)pin((int)(sb + L::OFF_X + (wq * 32 + lane) * 4));
            const long long e0 = w.base + (long long)min(u, ulast) * w.su + (long long)v * w.sv + (long long)m_first * w.sm;
            char* const stb = reinterpret_cast<char*>(pin_ptr(tt + e0));
            const int smb = pin((int)w.sm * 4);              // bytes per row step
            const int jo0 = pin((u_ok && v_ok) ? m_first - v + w.joff : -(1 << 30));
            // rows of this thread that may hold a frozen node: oriented j = jo0 + a inside the source box
            int fz_lo = 0, fz_cnt = 0;
            if (u_ok && v_ok) {
                const int it = w.ri ? ulast - u : u;
                const int ko = v - w.vlo, kt = w.rk ? p.d.nk - 1 - ko : ko;
                if (it >= p.fb.ilo && it <= p.fb.ihi && kt >= p.fb.klo && kt <= p.fb.khi && p.fb.jhi >= p.fb.jlo) {
                    const int jol = RJ ? w.nj - 1 - p.fb.jhi : p.fb.jlo, joh = RJ ? w.nj - 1 - p.fb.jlo : p.fb.jhi;
                    fz_lo = jol - jo0;
                    fz_cnt = joh - jol + 1;
                }
            }
            fz_lo = pin(fz_lo); fz_cnt = pin(fz_cnt);
            // mailboxes: outgoing slots of this tile, incoming slots of the upstream tiles (row 0 of each)
            constexpr int MG = Tile4Mail::MARGIN;
            unsigned long long* const mu_out = pin_ptr(mail.u + ((size_t)tile * mail.rows + MG) * 32 + lane);
            unsigned long long* const mv_out = pin_ptr(mail.v + ((size_t)tile * mail.rows + MG) * NW + wq);
            const unsigned long long* const mu_in = pin_ptr(mail.u + ((size_t)(has_u ? tile - p.nV : tile) * mail.rows + MG) * 32 + lane);
            // V mailbox: my step a needs row a - 1 + dmf of tile V-1, a node row iff 0 <= a - 1 + dmf < nrows_p
            const unsigned long long* const mv_in =
                pin_ptr(mail.v + ((size_t)(has_v ? tile - 1 : tile) * mail.rows + MG + (dmf - 1)) * NW + wq);
            const int vin_lo = pin(1 - dmf);
            const int vin_cnt = pin(mail_v_in ? nrows_p : 0);
            const unsigned ser = (unsigned)pin((int)serial);
            int dead = 0;
            int a = 0;                                       // local row of this warp; its global step is a + wq
            auto give_up = [&](int why, int x) {
                if (atomicCAS(&p.ctrl[1], 0, why) == 0) { p.ctrl[2] = tile; p.ctrl[3] = x; p.ctrl[4] = a; p.ctrl[5] = wq; p.ctrl[6] = lane; }
                sts_i(a_ab, 1);
            };
            auto wait_slot = [&](unsigned slot, unsigned par) {
                const unsigned mb = sb + L::OFF_BAR + 8 * slot;
                if (mbar_test(mb, par)) return;
                const long long t0 = clock64();
                while (!mbar_test(mb, par)) {   // each attempt suspends the thread for a bounded time
                    if (lds_i(a_ab)) break;
                    if (clock64() - t0 > (p.spin_limit << 9)) { give_up(20, (int)slot); break; }
                }
            };
            // re-poll a mailbox slot until it carries this sweep's tag
            auto poll = [&](const unsigned long long* q, int why) -> unsigned long long {
                unsigned long long x;
                const long long t0 = clock64();
                for (;;) {
                    x = ld_mail(q);
                    if ((unsigned)(x >> 32) == ser) break;
                    if (lds_i(a_ab)) break;
                    if (clock64() - t0 > (p.spin_limit << 9)) { give_up(why, 0); break; }
                }
                return x;
            };
            auto idle_step = [&](int g) {
                dead |= bar_compute_or<NC>(lds_i(a_ab));
                if (threadIdx.x == 0) sts_i(a_dn, g);
            };

            // ---- warps start one step apart
            for (int k = 0; k < wq && !dead; ++k) idle_step(k);
            // prefetch queues of the incoming mailboxes: rows a and a+1
            unsigned long long qu0 = 0, qu1 = 0, qv0 = 0, qv1 = 0;
            if (mail_u_in) { qu0 = ld_mail(mu_in); qu1 = ld_mail(mu_in + 32); }
            if (mail_v_in) { qv0 = ld_mail(mv_in); qv1 = ld_mail(mv_in + NW); }
            unsigned slot = 0, par = 0;
            wait_slot(0, 0);
            unsigned tc = tT0, sc = tS0;                     // this thread's bases in the current chunk slot
            float told = lds_f(tc + (RJ ? 7 : 0) * L::TROW); // old value of local row 0
            float t_prev = MAXV;
            float acc = 0.f;
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 1] = gtime();
            const int tq1 = (nrows / 4) & ~7, tq2 = (nrows / 2) & ~7, tq3 = ((3 * nrows) / 4) & ~7;

            // one step; r = a % 8 is a compile-time constant so that every ring address is base + immediate
            auto step = [&](auto RC) -> bool {
                constexpr int r = decltype(RC)::value;
                constexpr int R0 = (RJ ? 7 - r : r), R1 = (RJ ? 7 - ((r + 1) & 7) : ((r + 1) & 7));
                constexpr int XO = (r & 1) * XS;
                if (a >= nrows_r) return false;
                unsigned tj = tc;                            // base of the chunk holding row a+1
                unsigned nslot = slot, npar = par;
                if constexpr (r == 7) {
                    nslot = slot + 1;
                    if (nslot == NCH) { nslot = 0; npar ^= 1; }
                    wait_slot(nslot, npar);
                    tj = tT0 + nslot * L::CHB_T;
                }
                const float jp = lds_f(tj + R1 * L::TROW);            // old (u, a+1, v)
                float kp = lds_f(tj + dvp + R1 * L::TROW);            // old (u, a+1, v+1)
                float up = lds_f(tc + dup + R0 * L::TROW);            // old (u+1, a, v)
                const float sl = lds_f(sc + R0 * L::SROW);
                float um;
                if (first_w) {                               // new (u0-1, a, v): U mailbox of tile U-1
                    um = MAXV;
                    if (mail_u_in) {
                        const unsigned long long* const q = mu_in + (size_t)(unsigned)a * 32;
                        unsigned long long x = (r & 1) ? qu1 : qu0;
                        if ((unsigned)(x >> 32) != ser) x = poll(q, 21);
                        um = __uint_as_float((unsigned)x);
                        if (r & 1) qu1 = ld_mail(q + 64); else qu0 = ld_mail(q + 64);
                    }
                } else um = lds_f(aX - XW + XO);             // result of warp wq-1 for row a (written one step ago)
                if (!has_up) up = MAXV;
                if (kill_kp) kp = MAXV;
                float km = __shfl_up_sync(0xffffffffu, t_prev, 1);
                {                                            // lane 0: new (u, a-1, v0-1) from the V mailbox of tile V-1
                    const unsigned long long* const q = mv_in + (size_t)(unsigned)a * NW;
                    unsigned long long x = (r & 1) ? qv1 : qv0;
                    const bool inr = (unsigned)(a - vin_lo) < (unsigned)vin_cnt;
                    if (inr && (unsigned)(x >> 32) != ser) x = poll(q, 22);
                    if (lane_lo) km = inr ? __uint_as_float((unsigned)x) : MAXV;
                    if (r & 1) ld_mail_if(qv1, q + 2 * NW, mail_v_in); else ld_mail_if(qv0, q + 2 * NW, mail_v_in);
                }
                const float t = godunov(tmin(km, kp), tmin(t_prev, jp), tmin(um, up), sl * dx);
                bool valid = (unsigned)(jo0 + a) < (unsigned)nj;
                float* const dst = reinterpret_cast<float*>(stb + (long long)a * smb);
                if ((unsigned)(a - fz_lo) < (unsigned)fz_cnt) {
                    if (frozen_bit(frozen, (long long)(dst - tt))) valid = false;
                }
                float tnew = told;
                if (valid && t < told) {
                    tnew = t;
                    st_stream(dst, t);
                    acc += told - t;
                }
                st_mail_if(mu_out + (size_t)(unsigned)a * 32, ser, tnew, mail_u_out);
                st_mail_if(mv_out + (size_t)(unsigned)a * NW, ser, tnew, mail_v_out);
                t_prev = tnew;
                told = jp;
                sts_f(aX + XO, tnew);
                dead |= bar_compute_or<NC>(lds_i(a_ab));
                if (threadIdx.x == 0) sts_i(a_dn, a);
                ++a;
                if constexpr (r == 7) {
                    slot = nslot; par = npar;
                    tc = tj;
                    sc = tS0 + nslot * L::CHB_S;
                    if (p.trace && threadIdx.x == 0) {
                        if (a == tq1) p.trace[tile * 8 + 2] = gtime();
                        if (a == tq2) p.trace[tile * 8 + 3] = gtime();
                        if (a == tq3) p.trace[tile * 8 + 4] = gtime();
                    }
                }
                return dead == 0;
            };
            if (!dead)
                for (;;) {
                    if (!step(IntC<0>{})) break;
                    if (!step(IntC<1>{})) break;
                    if (!step(IntC<2>{})) break;
                    if (!step(IntC<3>{})) break;
                    if (!step(IntC<4>{})) break;
                    if (!step(IntC<5>{})) break;
                    if (!step(IntC<6>{})) break;
                    if (!step(IntC<7>{})) break;
                }
            // ---- trailing steps of the warps that started earlier
            for (int g = a + wq; g < nglobal && !dead; ++g) idle_step(g);
            if (threadIdx.x == 0) sts_i(a_dn, 1 << 29);   // let the loader run out its remaining chunks
            double dacc = (double)acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
            double* const sred = reinterpret_cast<double*>(smem_raw + L::OFF_RED);
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
template <int NW, int NCH>
inline int tile4_launch(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                        const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change,
                        cudaStream_t st) {
    using L = Tile4Layout<NW, NCH>;
    constexpr int NU = NW;
    TileParams p;
    p.w = w; p.d = d; p.fb = fb;
    p.nV = d.kpad / 32;
    p.nU = (w.nu + NU - 1) / NU;
    p.ntiles = p.nU * p.nV;
    p.chunk = 1;
    p.spin_limit = o.spin_limit;
    p.order = s.d_order; p.flags = s.d_flags; p.ctrl = s.d_ctrl; p.partial = s.d_partial;
    static long long* d_trace = nullptr;
    const char* trace_path = getenv("TTCR_B200_TRACE");
    if (trace_path && !d_trace) TCK(cudaMalloc(&d_trace, (size_t)s.cap_tiles * 8 * sizeof(long long)));
    p.trace = trace_path ? d_trace : nullptr;
    // mailboxes: sized for the smallest tile height this kernel is instantiated with (8 rows of u)
    if (!s.d_mbu) {
        s.mb_rows = d.nj + 72;   // rows of a tile (<= nj + 31) + the reach of the prefetches (see the kernel)
        s.mb_tiles = ((d.ni + 7) / 8) * (d.kpad / 32);
        const size_t nu = (size_t)s.mb_tiles * s.mb_rows * 32, nv = (size_t)s.mb_tiles * s.mb_rows * 16;
        TCK(cudaMalloc(&s.d_mbu, nu * 8));
        TCK(cudaMalloc(&s.d_mbv, nv * 8));
        TCK(cudaMemsetAsync(s.d_mbu, 0, nu * 8, st));
        TCK(cudaMemsetAsync(s.d_mbv, 0, nv * 8, st));
        s.serial = 0;
    }
    if (p.ntiles > s.mb_tiles) throw std::runtime_error("tile4: mailbox capacity");
    Tile4Mail mail;
    mail.u = s.d_mbu; mail.v = s.d_mbv; mail.rows = s.mb_rows;
    mail.serial = ++s.serial;
    if (mail.serial == 0) {   // wrapped: clear the tags once every 2^32 sweeps
        TCK(cudaMemsetAsync(s.d_mbu, 0, (size_t)s.mb_tiles * s.mb_rows * 32 * 8, st));
        TCK(cudaMemsetAsync(s.d_mbv, 0, (size_t)s.mb_tiles * s.mb_rows * 16 * 8, st));
        mail.serial = s.serial = 1;
    }
    const int key = 4000000 + NU * 1000;
    if (s.order_key != key || s.ntiles != p.ntiles) {
        std::vector<std::pair<long long, int>> k(p.ntiles);
        const long long lag_u = NU + 4;
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
    struct MapKey { const void* a; int bw, br, bp, kpad, qs, ni; CUtensorMap m; };
    static thread_local std::vector<MapKey> cache;
    auto get_map = [&](const void* a, int bw, int br, int bp) -> CUtensorMap {
        for (auto& e : cache)
            if (e.a == a && e.bw == bw && e.br == br && e.bp == bp && e.kpad == d.kpad && e.qs == d.qs && e.ni == d.ni) return e.m;
        if (cache.size() > 96) cache.clear();
        cache.push_back({a, bw, br, bp, d.kpad, d.qs, d.ni, make_tile3_map(a, d, bw, br, bp)});
        return cache.back().m;
    };
    const CUtensorMap tmT = get_map(tt, L::TW, L::C, NU + 1);
    const CUtensorMap tmS = get_map(slo, 32, L::C, NU);
    TCK(cudaMemsetAsync(s.d_ctrl, 0, sizeof(int), st));
    static int occ_cache = 0;
    if (!occ_cache) {
        TCK(cudaFuncSetAttribute(k_sweep_tile4<NW, NCH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
        TCK(cudaFuncSetAttribute(k_sweep_tile4<NW, NCH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
        TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cache, k_sweep_tile4<NW, NCH, false>, (NW + 1) * 32, L::BYTES));
        if (occ_cache < 1) throw std::runtime_error("tile4 kernel does not fit on an SM");
    }
    int occ = occ_cache;
    if (o.ctas_per_sm > 0) occ = std::min(occ, o.ctas_per_sm);
    const int grid = std::min(p.ntiles, occ * sm_count);
    if (w.rj)
        k_sweep_tile4<NW, NCH, true><<<grid, (NW + 1) * 32, L::BYTES, st>>>(tmT, tmS, p, mail, tt, frozen, dx);
    else
        k_sweep_tile4<NW, NCH, false><<<grid, (NW + 1) * 32, L::BYTES, st>>>(tmT, tmS, p, mail, tt, frozen, dx);
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

template <typename T> inline bool tile4_supported(bool) { return false; }
template <> inline bool tile4_supported<float>(bool weno_stage) { return !weno_stage; }

template <typename T>
inline int tile4_sweep(TileState&, const TileOptions&, int, const SweepView&, const Dims&, T*, const T*, const uint32_t*,
                       const FrozenBox&, T, double*, cudaStream_t) {
    throw std::runtime_error("tile4 kernel: fp32 only");
}
template <>
inline int tile4_sweep<float>(TileState& s, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                              const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change,
                              cudaStream_t st) {
    if (o.warps >= 16) return tile4_launch<16, 4>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.depth <= 4) return tile4_launch<8, 3>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    return tile4_launch<8, 4>(s, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
}

}  // namespace ttcrb200
