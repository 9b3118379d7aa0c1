// Sweep kernel "MARCH" (k_sweep_march): one node per thread and step.
//
// A directional sweep is a chain of ni + nj + nk dependent steps, and what k_sweep_patch (sweep_tile5.cuh) pays per step
// is the time one thread needs for its FOUR nodes (775 cycles alone on a scheduler) -- 1535 x 0.49 us is 0.75 ms before
// any hand-off, three times the HBM time of the sweep.  This kernel keeps the decomposition of the earlier ones
// (sheared layouts, tiles marching along the row axis m, tickets in dependency order, tagged words between tiles,
// TMA boxes skewed by the tensor map) and changes the thread geometry so that a step is ONE Godunov update deep:
//
//   * a WARP owns a patch of 4 planes (u) x 8 lanes (v): thread (lu, lv) = lane (lu*8 + lv).  Plane lu runs lu rows
//     behind plane 0, so at step a the thread updates node (u0w + lu, m_first - 1 + a - pl, v0w + lv) and
//         (u-1, m, v)    = result of lane - 8 at the previous step      one SHFL   (lu == 0: ring word)
//         (u, m-1, v-1)  = result of lane - 1 at the previous step      one SHFL   (lv == 0: ring word)
//         (u, m-1, v)    = own previous result                          register
//         (u, m+1, v), (u, m+1, v+1), (u+1, m, v), slowness             four LDS.32 from the TMA boxes
//     Both upwind neighbours across the patch edges arrive as tagged 8-byte words {value, step tag} in a per-warp
//     shared-memory ring (8 U words + 4 V words per step), written by the warp above / to the left, or by the importer
//     warp from the global mailboxes of the neighbouring tiles.  No CTA barrier, no mbarrier in a compute warp.
//   * a CTA stacks WU x WV warps (default 4 x 4: a tile of 16 planes x 32 lanes, 16 compute warps + importer + loader),
//     two CTAs per SM.  All warps of a tile run the same step index; v-adjacent warps are NOT lagged (a lag would have
//     to be paid in ring depth of the TMA boxes).
//   * the loader warp issues one 3-D box of traveltimes {TW+4 lanes, 2 rows, PUT+1 planes} and one of slowness per chunk
//     of 2 steps into a ring of NCH slots and publishes "rows landed" as a plain shared word; compute warps publish
//     their progress every other step (ring reuse, back-pressure of the word rings).  The box is stored
//     [plane][row][lane] with 2*(TW+4) words per plane = 8 (mod 32): the 4 x 8 threads of a warp hit 32 different banks.
//   * slots that are no node hold +MAX (traveltime) and NaN (slowness), as for k_sweep_patch: no validity test.  Frozen
//     nodes and the grid's last plane / last lane take a warp-uniform slow branch.
//
// fp32, first-order stage.  Same DAG and same arithmetic (update.cuh) as every other sweep kernel: bit-identical field.
#pragma once
#include "sweep_tile5.cuh"

#ifndef TTCR_MARCH_STEP_TRACE
#define TTCR_MARCH_STEP_TRACE 0   // 1: per-step clocks of tile 0 (TTCR_B200_TRACE_STEPS), development builds only
#endif

namespace ttcrb200 {

struct MarchMail {
    unsigned long long* u = nullptr;   // [tile][rows][TW]:  last plane of the tile, one word per lane and row
    unsigned long long* v = nullptr;   // [tile][rows][PUT]: last lane of the tile, one word per plane and row
    int rows = 0;                      // row stride per tile (MARGIN rows before local row 0)
    unsigned serial = 0;               // tag of the current sweep (never cleared)
};

struct MarchParams {
    SweepView w;
    Dims d;
    FrozenBox fb;
    int nU, nV, ntiles;
    long long spin_cycles;
    const int* order;
    int* ctrl;
    double* partial;   // [tile][NW]
    long long* trace;  // [tile][8] or nullptr
    unsigned pause_ns; // sleep between two polls of a waiting compute warp (0 = spin)
    unsigned lpause_ns; // sleep of the loader when it can neither issue nor land a chunk
    int pf_chunks;     // L2 prefetch distance of the loader, in chunks beyond the box ring (0 = off)
    int trace_tile;    // tile whose steps a step-trace build records
};

template <int WU, int WV, int NCH, int D>
struct MarchLayout {
    static constexpr int NW = WU * WV, PUT = 4 * WU, TW = 8 * WV, BW = TW + 4, C = 2;
    static constexpr int NT = (NW + 2) * 32;
    static constexpr int ROWB = BW * 4, PSB = C * ROWB;                               // bytes per box row / per plane of a box
    static constexpr int TBYTES = (PUT + 1) * PSB, SBYTES = PUT * PSB;                // bytes a box delivers
    static constexpr int CHB_T = t5_round128(TBYTES), CHB_S = t5_round128(SBYTES), CHB = CHB_T + CHB_S;
    static constexpr int SLOTB = 128, RINGB = D * SLOTB;                              // ring slot: 8 U words | 4 V words | pad
    static constexpr int MARGIN = PUT + 4;                                            // mailbox rows before local row 0
    static constexpr int BIG = 1 << 29;
    static constexpr int OFF_BOX = 0;
    static constexpr int OFF_BAR = NCH * CHB;                                         // NCH mbarriers
    static constexpr int OFF_PROG = OFF_BAR + (NCH * 8 + 15) / 16 * 16;               // progress of the NW compute warps
    static constexpr int OFF_CTL = OFF_PROG + (NW * 4 + 15) / 16 * 16;                // landed, dead, tile, chunks issued
    static constexpr int OFF_RING = OFF_CTL + 16;                                     // word rings: aligned at run time to RINGB in the
    static constexpr int BYTES = OFF_RING + RINGB + NW * RINGB;                       // shared ADDRESS space (slot wrap by masking)
    static_assert((2 * BW) % 32 == 8 || (2 * BW) % 32 == 24, "plane stride of a box must spread the 4 planes of a warp over the banks");
    static_assert((D & (D - 1)) == 0 && D >= 8, "ring depth: a power of two, at least 8");
    // A warp at step a has seen chunk (a+2)/2 landed, which the loader issued only after EVERY warp had published step
    // a + 2 - 2 NCH: no warp can be more than 2 NCH steps ahead of another, so the word rings need no back-pressure of their own.
    static_assert(D >= 2 * NCH + 1, "word rings must outlast the box ring");
    static_assert(PUT + C + 6 <= GUARD, "guard rows too few");
    static_assert(PUT <= 32 && TW <= 32 && NW % 4 == 0, "one importer lane per plane and per lane of the tile");
};

// ---- geometry (host + device: tests/ emulate the TMA boxes with these) ---------------------------------------------
struct MarchTile {
    int U, V, u0, v0, va, vb, m_first, nrows, nA, nch;
    int has_u, has_v, has_down, has_right, nrows_p, dmf;
};

template <int PUT, int TW>
__host__ __device__ inline MarchTile march_tile(const SweepView& w, int nU, int nV, int tile) {
    MarchTile t;
    t.U = tile / nV; t.V = tile - t.U * nV;
    t.u0 = t.U * PUT; t.v0 = t.V * TW;
    t.va = t.v0 > w.vlo ? t.v0 : w.vlo;
    t.vb = t.v0 + TW < w.vhi ? t.v0 + TW : w.vhi;
    t.m_first = t.va - w.joff;                      // first row that holds a node of the tile (j = 0 on lane va)
    t.nrows = (t.vb - t.va) + w.nj - 1;             // local rows 0 .. nrows-1 hold all its nodes
    const int nsteps = t.nrows + PUT;               // at step a plane pl runs local row a - 1 - pl
    t.nA = (nsteps + 1) & ~1;                       // steps executed (pairs)
    t.nch = t.nA / 2 + 1;                           // chunks of 2 box rows: rows 0 .. nA (the last step still prefetches)
    t.has_u = t.U > 0; t.has_v = t.V > 0;
    t.has_down = t.U + 1 < nU; t.has_right = t.v0 + TW < w.vhi;
    const int va_p = (t.v0 - TW) > w.vlo ? (t.v0 - TW) : w.vlo;
    t.nrows_p = (t.v0 - va_p) + w.nj - 1;           // rows of tile V-1
    t.dmf = t.m_first - (va_p - w.joff);            // its local row index of my local row 0
    return t;
}

template <int WU, int WV, int NCH, int D, bool RI, bool RJ, bool RK>
struct MarchGeom {
    using L = MarchLayout<WU, WV, NCH, D>;
    // byte offsets inside a chunk slot, relative to the traveltime word (u, m+1, v) of the EVEN step of the chunk
    static constexpr int DR = RJ ? -L::ROWB : L::ROWB;                 // ... of the odd step
    static constexpr int DH = RK ? -4 : 4;                             // (u, m+1, v+1)
    static constexpr int DUP = RI ? -L::PSB : L::PSB;                  // (u+1, m, v)
    static constexpr int DS = L::CHB_T + (RI ? -L::PSB : 0);           // slowness (u, m, v)
    __host__ __device__ static int thread_off(int pl, int vl) {
        return (RI ? L::PUT - pl : pl) * L::PSB + (RK ? L::BW - 1 - vl : vl) * 4 + (RJ ? L::ROWB : 0);
    }
    // box origins in tensor-map coordinates (x lane, y skewed row, z plane) of chunk c; `s` = 1 for the slowness box.
    // Box row b (= step b) holds, for plane pl, row m_first + b - pl of the traveltimes and m_first - 1 + b - pl of slowness.
    __host__ __device__ static int box_x(const Dims& d, const MarchTile& t) { return RK ? d.kpad - L::BW - t.v0 : t.v0; }
    __host__ __device__ static int box_z(const Dims& d, const MarchTile& t, int s) {
        return RI ? d.ni - 1 - (t.u0 + L::PUT - s) : t.u0;
    }
    __host__ __device__ static int box_y(const SweepView& w, const Dims& d, const MarchTile& t, int c, int s) {
        const int mf = t.m_first - s;
        if (!RJ) return GUARD + mf + 2 * c + t.u0 + (RI ? 1 : 0);
        return GUARD + (w.nm - 1) - mf - (2 * c + 1) - t.u0 + d.ni - (RI ? 1 : 0);
    }
};

// ---- small PTX helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts_u2_if(unsigned a, unsigned x, unsigned y, int on) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t@q st.volatile.shared.v2.u32 [%0], {%1, %2};\n\t}" ::"r"(a), "r"(x), "r"(y), "r"(on) : "memory");
}
__device__ __forceinline__ void stg_f_stream_if(float* p, float x, int on) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q st.global.L1::no_allocate.f32 [%0], %1;\n\t}" ::"l"(p), "f"(x), "r"(on) : "memory");
}
__device__ __forceinline__ void sts_i_if(unsigned a, int v, unsigned on) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.volatile.shared.s32 [%0], %1;\n\t}" ::"r"(a), "r"(v), "r"(on) : "memory");
}
// next slot of a ring that is aligned to its size: (a & ~(SIZE-1)) | ((a + STEP) & (SIZE-1)), one add and one LOP3
template <unsigned SIZE, unsigned STEP>
__device__ __forceinline__ unsigned ring_next(unsigned a) {
    unsigned r;
    asm("{\n\t.reg .b32 t;\n\tadd.u32 t, %1, %2;\n\tlop3.b32 %0, %1, t, %3, 0xD8;\n\t}" : "=r"(r) : "r"(a), "n"(STEP), "n"(SIZE - 1));
    return r;
}
__device__ __forceinline__ int mbar_test_nb(unsigned a, unsigned parity) {
    int ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void sts_release_i(unsigned a, int v) { asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

template <int WU, int WV, int NCH, int D, bool RI, bool RJ, bool RK>
__global__ void __launch_bounds__((WU * WV + 2) * 32, 2)
k_sweep_march(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmS, MarchParams p, MarchMail mail,
              float* __restrict__ tt, const uint32_t* __restrict__ frozen, float dx) {
    using L = MarchLayout<WU, WV, NCH, D>;
    using G = MarchGeom<WU, WV, NCH, D, RI, RJ, RK>;
    constexpr int NW = L::NW, PUT = L::PUT, TW = L::TW, SLOTB = L::SLOTB, RINGB = L::RINGB, MG = L::MARGIN, BIG = L::BIG;
    constexpr int IG = 4;   // mailbox rows the importer keeps in flight
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const float MAXV = FLT_MAX;
    const unsigned sbase = (unsigned)pin((int)__cvta_generic_to_shared(smem_raw));
    const unsigned a_ctl = sbase + L::OFF_CTL, a_dead = a_ctl + 4, a_tile = a_ctl + 8;
    const unsigned a_prog = sbase + L::OFF_PROG;
    const unsigned a_ring = (sbase + L::OFF_RING + RINGB - 1) & ~(unsigned)(RINGB - 1);   // (the shared window does not start at 0)
    const unsigned serial = mail.serial;
    const long long spin_cycles = p.spin_cycles;
    const unsigned pause_ns = p.pause_ns;
    unsigned par = 0;   // loader: per chunk slot, parity of the mbarrier phase its next chunk completes

    if (threadIdx.x == 0) {
        for (int i = 0; i < NCH; ++i) mbar_init(sbase + L::OFF_BAR + 8 * i, 1);
        fence_mbar_init();
    }

    for (;;) {
        __syncthreads();   // everybody is done with the previous tile
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sts_i(a_tile, (ab || t >= p.ntiles) ? -1 : t);
            sts_i(a_dead, 0);
            sts_i(a_ctl, -3);
            sts_i(a_ctl + 12, 0);
        }
        if (threadIdx.x < NW) sts_i(a_prog + 4 * threadIdx.x, -1);   // progress a = even step a has issued its reads of box row a + 1
        __syncthreads();
        const int ticket = lds_i(a_tile);
        if (ticket < 0) break;
        const int tile = p.order[ticket];
        const MarchTile T = march_tile<PUT, TW>(w, p.nU, p.nV, tile);
        const int nA = T.nA;
        // Word rings.  A word is valid for step a when its tag is >= a + 1 (a slot only ever holds the tag of its step or an
        // older one).  Slot 0 holds the words of step 0 (+MAX: row -1 of a tile holds no node).  Words nobody will send --
        // U words of the first warp row of a tile without a tile above, V words of the first warp column of a tile without
        // a tile to the left -- are +MAX with tag BIG in every slot: those warps never wait.
        for (unsigned o = threadIdx.x * 16; o < (unsigned)(NW * RINGB); o += L::NT * 16) {
            const unsigned wr = o / RINGB, in = o & (SLOTB - 1);   // ring (= warp), byte inside the slot: U words 0..63, V words 64..95
            const bool open_end = in < 64 ? (wr < (unsigned)WV && !T.has_u) : (wr % WV == 0 && !T.has_v);
            const unsigned tag = open_end ? (unsigned)BIG : (((o & (RINGB - 1)) < (unsigned)SLOTB) ? 1u : 0u);
            sts_u4(a_ring + o, __float_as_uint(MAXV), tag, __float_as_uint(MAXV), tag);
        }
        __syncthreads();
        if (p.trace && threadIdx.x == 0) {
            p.trace[tile * 8 + 0] = gtime();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[tile * 8 + 7] = smid;
        }

        if (warp == NW) {
            // ================= importer ==================================================================================
            // step a (>= 1): ring of warp (0, l/8), slot a, U word l%8   <- row a-1 of the last plane of tile U-1, lane l
            //                ring of warp (pl/4, 0), slot a, V word pl%4 <- row a+dmf-2-pl of the last lane of tile V-1, plane pl
            // A tile writes one mailbox word per plane (lane) and STEP, also for the rows around its nodes (value +MAX), so
            // every word of steps 1 .. a_end exists sooner or later; the steps after a_end get +MAX.  Lane l runs one U
            // stream and (l < PUT) one V stream of IG independent slots each: a slot polls the word of ITS step until the
            // tag of this sweep shows up, hands it to the ring and moves IG steps on.  No slot ever waits for another one,
            // so a word is seen one poll (an L2 round trip) after it was written, whatever the words around it do.
            static_assert(TW <= 32, "one U word per importer lane");
            const bool ulane = T.has_u && lane < TW, vlane = T.has_v && lane < PUT;
            const int vpl = lane < PUT ? lane : 0;
            const int nA_left = (T.nrows_p + PUT + 1) & ~1;                                   // steps of tile V-1
            const int endU = min(nA - 1, nA - PUT), endV = min(nA - 1, nA_left - T.dmf);      // last step with a mailbox word
            const unsigned long long* const mu_p = mail.u + ((size_t)(T.has_u ? tile - p.nV : tile) * mail.rows + MG - 1) * TW + lane;                 // + a * TW
            const unsigned long long* const mv_p = mail.v + ((size_t)(T.has_v ? tile - 1 : tile) * mail.rows + MG + (T.dmf - 2 - vpl)) * PUT + vpl;   // + a * PUT
            const unsigned wU = a_ring + (unsigned)(lane >> 3) * RINGB + (lane & 7) * 8;
            const unsigned wV = a_ring + (unsigned)((vpl >> 2) * WV) * RINGB + 64 + (vpl & 3) * 8;
            const unsigned pU = a_prog + 4 * (lane >> 3), pV = a_prog + 4 * ((vpl >> 2) * WV);   // progress of the warp that reads the word
            unsigned long long qu[IG], qv[IG];
            int au[IG], av[IG];   // step each slot is working on
#pragma unroll
            for (int s = 0; s < IG; ++s) {
                au[s] = ulane ? 1 + s : nA; av[s] = vlane ? 1 + s : nA;
                qu[s] = qv[s] = 0;
                if (au[s] <= endU) qu[s] = ld_mail(mu_p + (size_t)au[s] * TW);
                if (av[s] <= endV) qv[s] = ld_mail(mv_p + (size_t)av[s] * PUT);
            }
            // one visit of a slot: poll again / hand over and move on / wait for the ring slot.  Returns 1 if it delivered.
            auto visit = [&](unsigned long long& q, int& a, const unsigned long long* base, int stride, int a_end, unsigned wring, unsigned pcons) {
                if (a >= nA) return 0;
                const bool real = a <= a_end;
                if (real && (unsigned)(q >> 32) != serial) {
                    q = ld_mail(base + (size_t)a * stride);
                    return 0;
                }
                if (lds_i(pcons) < a + 1 - D) return 0;   // ring slot a % D still holds the word of step a - D
                sts_u2(wring + (unsigned)(a & (D - 1)) * SLOTB, real ? (unsigned)q : __float_as_uint(MAXV), (unsigned)(a + 1));
                a += IG;
                if (a < nA && a <= a_end) q = ld_mail(base + (size_t)a * stride);
                return 1;
            };
            long long t0 = clock64();
            unsigned idle = 0;
            for (;;) {
                int done = 1, got = 0;
#pragma unroll
                for (int s = 0; s < IG; ++s) {
                    got += visit(qu[s], au[s], mu_p, TW, endU, wU, pU);
                    got += visit(qv[s], av[s], mv_p, PUT, endV, wV, pV);
                    done &= (au[s] >= nA) & (av[s] >= nA);
                }
                if (done) break;
                if (got) { idle = 0; continue; }
                if ((++idle & 63u) == 0) {   // nothing moved for a while: is the march still alive?
                    if (lds_i(a_dead)) break;
                    if (idle == 64) t0 = clock64();
                    else if (clock64() - t0 > spin_cycles) {
                        if (atomicCAS(&p.ctrl[1], 0, 31) == 0) { p.ctrl[2] = tile; p.ctrl[3] = au[0]; p.ctrl[4] = av[0]; p.ctrl[5] = NW; p.ctrl[6] = lane; }
                        sts_i(a_dead, 1);
                        break;
                    }
                }
            }
            __syncwarp();
        } else if (warp == NW + 1) {
            // ================= loader ====================================================================================
            // Chunk c of the tile lives in slot c % NCH.  Lane 0 issues (a chunk may go out once every warp has published the
            // last step that reads the chunk it replaces), lane 1 waits for the boxes to land (suspended in try_wait, no
            // polling) and publishes the rows that are there.  A slot's mbarrier is used once per chunk, in order: `par`
            // keeps, per slot, the parity of the phase its next landing chunk completes.
            const int nch = T.nch;
            const unsigned a_bar = sbase + L::OFF_BAR;
            if (lane == 0) {
                const int xb = G::box_x(p.d, T);
                const int zT = G::box_z(p.d, T, 0), zS = G::box_z(p.d, T, 1);
                const int yT0 = G::box_y(w, p.d, T, 0, 0), yS0 = G::box_y(w, p.d, T, 0, 1);
                const int dyc = RJ ? -2 : 2;
                int si = 0;
                for (int ci = 0; ci < nch; ++ci) {
                    if (ci >= NCH) {
                        // chunk ci - NCH (box rows 2(ci-NCH), +1) has been read by a warp once it has published step 2(ci-NCH)
                        const int need = 2 * (ci - NCH);
                        const long long t0 = clock64();
                        bool ok = true;
                        for (;;) {
                            int mp = BIG;
#pragma unroll
                            for (int i = 0; i < NW; i += 4) {
                                const uint4 x = lds_u4(a_prog + 4 * i);
                                mp = min(min(mp, (int)x.x), min(min((int)x.y, (int)x.z), (int)x.w));
                            }
                            if (mp >= need) break;
                            __nanosleep(100);   // (a chunk is two steps, and the ring is NCH chunks deep)
                            if (lds_i(a_dead)) { ok = false; break; }
                            if (clock64() - t0 > spin_cycles) {
                                if (atomicCAS(&p.ctrl[1], 0, 50) == 0) { p.ctrl[2] = tile; p.ctrl[3] = ci; p.ctrl[4] = mp; p.ctrl[5] = NW + 1; }
                                sts_i(a_dead, 1);
                                ok = false;
                                break;
                            }
                        }
                        if (!ok) break;
                    }
                    const unsigned mb = a_bar + 8 * si;
                    const unsigned dst = sbase + L::OFF_BOX + si * L::CHB;
                    mbar_expect_tx(mb, L::TBYTES + L::SBYTES);
                    tma_load_3d(dst, &tmT, xb, yT0 + ci * dyc, zT, mb);
                    tma_load_3d(dst + L::CHB_T, &tmS, xb, yS0 + ci * dyc, zS, mb);
                    sts_i(a_ctl + 12, ci + 1);   // chunks issued (what lane 1 may wait for)
                    si = si + 1 == NCH ? 0 : si + 1;
                }
            } else if (lane == 1) {
                int sl = 0;
                for (int cl = 0; cl < nch; ++cl) {
                    const long long t0 = clock64();
                    bool ok = true;
                    while (lds_i(a_ctl + 12) <= cl) {   // not issued yet
                        __nanosleep(100);
                        if (lds_i(a_dead) || clock64() - t0 > spin_cycles) { ok = false; break; }
                    }
                    while (ok && !mbar_test(a_bar + 8 * sl, (par >> sl) & 1u)) {   // try_wait: suspended until the phase completes or a time limit
                        if (lds_i(a_dead) || clock64() - t0 > spin_cycles) { ok = false; break; }
                    }
                    if (!ok) {   // give up; copies in flight still have to land before the CTA goes on
                        sts_i(a_dead, 1);
                        const int issued = lds_i(a_ctl + 12);
                        const long long t1 = clock64();
                        for (; cl < issued; ++cl) {
                            while (!mbar_test(a_bar + 8 * sl, (par >> sl) & 1u) && clock64() - t1 < spin_cycles) {}
                            par ^= 1u << sl;
                            sl = sl + 1 == NCH ? 0 : sl + 1;
                        }
                        break;
                    }
                    par ^= 1u << sl;
                    sl = sl + 1 == NCH ? 0 : sl + 1;
                    sts_release_i(a_ctl, 2 * (cl + 1) - 3);   // (rows landed) - 3: an even step with tag tg needs row tg + 2
                }
            }
            par = __shfl_sync(0xffffffffu, par, 1);   // (lane 1 owns the parities; every lane keeps a copy for the next tile)
        } else {
            // ================= compute warps =============================================================================
            const int lw = warp;
            const int wu = lw / WV, wv = lw - wu * WV, lu = lane >> 3, lv = lane & 7;
            const int pl = 4 * wu + lu, vl = 8 * wv + lv;
            const int u = T.u0 + pl, vt = T.v0 + vl;
            const int ulast = w.nu - 1;
            const bool ghost = vt >= p.d.kpad || u > ulast;   // TMA zero-fills these: `t < old` must never hold
            const bool out_u_in = wu < WU - 1, out_v_in = wv < WV - 1;
            // ring addresses: the output rings sit at compile-time distances from the input words (warp + WV, warp + 1)
            unsigned rU = a_ring + (unsigned)lw * RINGB + lv * 8;        // U word lv of the slot of the even step
            unsigned rV = a_ring + (unsigned)lw * RINGB + 64 + lu * 8;   // V word lu
            constexpr unsigned OUT_U = WV * RINGB, OUT_V = RINGB;
            const unsigned a_myprog = (unsigned)pin((int)(a_prog + 4 * lw));
            const unsigned aJ0 = sbase + L::OFF_BOX + (unsigned)G::thread_off(pl, vl);
            float* pg = tt + (w.base + (long long)min(u, ulast) * w.su + (long long)(T.m_first - 1 - pl) * w.sm + (long long)vt * w.sv);
            // ---- slow-path conditions: last plane / last lane of the grid, frozen nodes (source box)
            const bool edge_u = u >= ulast, edge_v = vt + 1 >= w.vhi;
            const bool edge_w = __any_sync(0xffffffffu, edge_u || edge_v);
            bool fzme = false;
            int wz_lo = 1 << 28, wz_hi = -(1 << 28);
            if (p.fb.jhi >= p.fb.jlo && u <= ulast && vt >= w.vlo && vt < w.vhi) {
                const int it = w.ri ? ulast - u : u;
                const int ko = vt - w.vlo, kt = w.rk ? p.d.nk - 1 - ko : ko;
                if (it >= p.fb.ilo && it <= p.fb.ihi && kt >= p.fb.klo && kt <= p.fb.khi) {
                    fzme = true;
                    const int jol = RJ ? w.nj - 1 - p.fb.jhi : p.fb.jlo, joh = RJ ? w.nj - 1 - p.fb.jlo : p.fb.jhi;
                    // oriented j = m_first + r - v + joff, step a = r + 1 + pl
                    wz_lo = jol - w.joff + vt - T.m_first + 1 + pl;
                    wz_hi = joh - w.joff + vt - T.m_first + 1 + pl;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                wz_lo = min(wz_lo, __shfl_xor_sync(0xffffffffu, wz_lo, o));
                wz_hi = max(wz_hi, __shfl_xor_sync(0xffffffffu, wz_hi, o));
            }
            // lane roles, one bit each, in a register the compiler cannot rematerialise from %tid
            enum : unsigned { F_L0 = 1, F_U0 = 2, F_V0 = 4, F_STU = 8, F_STV = 16, F_OMU = 32, F_OMV = 64, F_EU = 128, F_EV = 256, F_FZ = 512 };
            const unsigned fl = (unsigned)pin((int)((lane == 0 ? F_L0 : 0u) | (lu == 0 ? F_U0 : 0u) | (lv == 0 ? F_V0 : 0u) | ((lu == 3 && out_u_in) ? F_STU : 0u) |
                                                    ((lv == 7 && out_v_in) ? F_STV : 0u) | ((lu == 3 && !out_u_in && T.has_down) ? F_OMU : 0u) |
                                                    ((lv == 7 && !out_v_in && T.has_right) ? F_OMV : 0u) | (edge_u ? F_EU : 0u) | (edge_v ? F_EV : 0u) |
                                                    (fzme ? F_FZ : 0u)));
            const bool mail_warp = __any_sync(0xffffffffu, (fl & (F_OMU | F_OMV)) != 0);
            const float QNAN = __int_as_float(0x7fc00000);

            int dead = 0;
            auto give_up = [&](int why, int x, int a) {
                if (atomicCAS(&p.ctrl[1], 0, why) == 0) { p.ctrl[2] = tile; p.ctrl[3] = x; p.ctrl[4] = a; p.ctrl[5] = lw; p.ctrl[6] = lane; }
                sts_i(a_dead, 1);
            };
            // ---- prologue: chunk 0, operands of step 0
            {
                const long long t0 = clock64();
                while (lds_i(a_ctl) < 2 - 3) {
                    if (lds_i(a_dead)) { dead = 1; break; }
                    if (clock64() - t0 > spin_cycles) { give_up(40, 0, 0); dead = 1; break; }
                }
            }
            int deadw = __any_sync(0xffffffffu, dead);
            const float ini = ghost ? 0.f : MAXV;
            float nprev = ini, told = ini;
            float j = lds_f(aJ0), h = lds_f(aJ0 + G::DH), up = lds_f(aJ0 + G::DUP), s = lds_f(aJ0 + G::DS);
            float acc = 0.f;
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 1] = gtime();

            // The march runs in pairs of steps (tags tg, tg+1).  Pairs in [ts0, ts1) may touch the grid's last plane / lane
            // or a frozen node and run the SLOW body; warps that feed a global mailbox run the MAILW bodies.
            const unsigned tg_end = (unsigned)nA + 1u;
            unsigned ts0 = tg_end, ts1 = tg_end;
            if (edge_w) ts0 = 1;
            else if (wz_hi >= wz_lo) {
                ts0 = (unsigned)min(max((wz_lo & ~1) + 1, 1), (int)tg_end);
                ts1 = (unsigned)min(max((wz_hi & ~1) + 3, 1), (int)tg_end);
            }
            unsigned long long* mu = mail.u + ((size_t)tile * mail.rows + (MG - 1 - pl)) * TW + vl;
            unsigned long long* mv = mail.v + ((size_t)tile * mail.rows + (MG - 1 - pl)) * PUT + pl;
            unsigned rB = aJ0;       // thread's word in the chunk slot of the even step
            unsigned tg = 1;

            // One march step: ring words at rUi / rVi carry tag `tgs` (= step + 1), outputs go to rUo / rVo, next operands at `ao`.
            auto step = [&](const unsigned tgs, const unsigned rUi, const unsigned rVi, const unsigned rUo, const unsigned rVo, const unsigned ao,
                            auto odd_c, auto slow_c, auto mail_c) {
                constexpr bool ODD = decltype(odd_c)::value != 0, SLOW = decltype(slow_c)::value != 0, MAILW = decltype(mail_c)::value != 0;
#if TTCR_MARCH_STEP_TRACE
                long long* const tr = (p.trace && tile == p.trace_tile && lane == 0 && (unsigned)(tgs - 201u) < 64u) ? p.trace + (size_t)p.ntiles * 8 + ((lw * 64 + (tgs - 201u)) * 4) : nullptr;
                if (tr) tr[0] = clock64();
#endif
                // ---- (1) everything this step reads from shared memory
                uint2 xu = lds_u2(rUi), xv = lds_u2(rVi);
                int landed = BIG;
                if (!ODD) landed = lds_i(a_ctl);
                const float j2 = lds_f(ao), h2 = lds_f(ao + G::DH), up2 = lds_f(ao + G::DUP), s2 = lds_f(ao + G::DS);
                // ---- (2) the upwind neighbours inside the patch
                float um = __shfl_up_sync(0xffffffffu, nprev, 8);
                float km = __shfl_up_sync(0xffffffffu, nprev, 1);
                // (the shuffle is a convergence point: every lane has left the previous step, its ring slots and box rows are free)
                if (!ODD) sts_i_if(a_myprog, (int)tgs - 1, fl & F_L0);
                // ---- (3) the one branch: words not there yet / the chunk of the next pair has not landed
                bool bad = xu.y < tgs || xv.y < tgs;
                if (!ODD) bad = bad || landed < (int)tgs;
                if (bad) {
                    long long t0 = 0;
                    for (unsigned it = 1;; ++it) {
                        if (pause_ns) __nanosleep(pause_ns);
                        xu = lds_u2(rUi); xv = lds_u2(rVi);
                        bool b2 = xu.y < tgs || xv.y < tgs;
                        if (!ODD) b2 = b2 || lds_i(a_ctl) < (int)tgs;
                        if (!b2) break;
                        if ((it & 255u) == 0) {   // the expensive checks once in a while
                            if (lds_i(a_dead)) { dead = 1; break; }
                            if (t0 == 0) t0 = clock64();
                            else if (clock64() - t0 > spin_cycles) { give_up(41, (int)tgs, ODD); dead = 1; break; }
                        }
                    }
                }
#if TTCR_MARCH_STEP_TRACE
                if (tr) tr[1] = clock64();
#endif
                if (SLOW) {
                    if (fl & F_EU) up = MAXV;
                    if (fl & F_EV) h = MAXV;
                    if ((fl & F_FZ) && frozen_bit(frozen, (long long)(pg - tt))) s = QNAN;
                }
                if (fl & F_U0) um = __uint_as_float(xu.x);
                if (fl & F_V0) km = __uint_as_float(xv.x);
                // ---- (4) the update
                const float tn = godunov(tmin(km, h), tmin(nprev, j), tmin(um, up), s * dx);
                const float n = fminf(tn, told);   // (NaN -> told: slots that are no node, frozen nodes)
#if TTCR_MARCH_STEP_TRACE
                if (tr) tr[2] = clock64() + (long long)(n == 12345.f);   // (depends on n: stamped when the update is done)
#endif
                // ---- (5) hand-off: warp below / to the right (shared rings), tiles U+1 / V+1 (global mailboxes)
                sts_u2_if(rUo, __float_as_uint(n), tgs + 1, fl & F_STU);
                sts_u2_if(rVo, __float_as_uint(n), tgs + 1, fl & F_STV);
                if (MAILW) {
                    st_mail_if(mu, serial, n, fl & F_OMU);
                    st_mail_if(mv, serial, n, fl & F_OMV);
                    mu += TW; mv += PUT;
                }
                // ---- (6) result, change sum, rotate the operands
                stg_f_stream_if(pg, n, n < told ? 1 : 0);
                pg += RJ ? -(long long)p.d.kpad : (long long)p.d.kpad;
                acc += told - n;
                nprev = n; told = j;
                j = j2; h = h2; up = up2; s = s2;
#if TTCR_MARCH_STEP_TRACE
                if (tr) tr[3] = clock64();
#endif
            };
            const unsigned rBend = aJ0 + NCH * L::CHB;
            auto pair = [&](auto slow_c, auto mail_c) {
                step(tg, rU, rV, rU + OUT_U + SLOTB, rV + OUT_V + SLOTB, rB + (unsigned)G::DR, IntC<0>(), slow_c, mail_c);
                // slots of the next pair (rings are RINGB-aligned: wrap inside the low bits), chunk slot of the next pair
                const unsigned rU2 = ring_next<RINGB, 2 * SLOTB>(rU), rV2 = ring_next<RINGB, 2 * SLOTB>(rV);
                unsigned rB2 = rB + L::CHB;
                if (rB2 == rBend) rB2 = aJ0;
                step(tg + 1, rU + SLOTB, rV + SLOTB, rU2 + OUT_U, rV2 + OUT_V, rB2, IntC<1>(), slow_c, mail_c);
                rU = rU2; rV = rV2; rB = rB2;
                tg += 2;
                deadw = __any_sync(0xffffffffu, dead);
#if TTCR_MARCH_STEP_TRACE
                if (p.trace && threadIdx.x == 0) {   // quarter times of the march
                    const unsigned q = ((unsigned)nA / 8u) * 2u;
                    if (tg == q + 1) p.trace[tile * 8 + 2] = gtime();
                    if (tg == 2 * q + 1) p.trace[tile * 8 + 3] = gtime();
                    if (tg == 3 * q + 1) p.trace[tile * 8 + 4] = gtime();
                }
#endif
            };
            auto march = [&](auto mail_c) {
#pragma unroll 1
                for (int seg = 0; seg < 3; ++seg) {
                    const unsigned e = seg == 0 ? ts0 : (seg == 1 ? ts1 : tg_end);
                    if (seg == 1) {
                        while (tg < e && !deadw) pair(IntC<1>(), mail_c);
                    } else {
                        while (tg < e && !deadw) pair(IntC<0>(), mail_c);
                    }
                }
            };
            if (mail_warp) march(IntC<1>()); else march(IntC<0>());
            sts_i_if(a_myprog, BIG, fl & F_L0);   // release anybody still waiting for this warp
            double dacc = (double)acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
            if (lane == 0) p.partial[(size_t)tile * NW + lw] = dacc;
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 5] = gtime();
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct MarchState {
    unsigned long long* d_mbu = nullptr;
    unsigned long long* d_mbv = nullptr;
    double* d_partial = nullptr;
    int* d_order = nullptr;
    long long* d_trace = nullptr;
    int mb_rows = 0, mb_tiles = 0, mb_tw = 0, mb_put = 0;
    int order_key = -1, ntiles = 0, trace_cap = 0;
    unsigned serial = 0;
    std::vector<std::pair<const void*, int>> occ;   // per kernel instance, for the device this state (slot) belongs to
    std::vector<std::pair<std::vector<long long>, CUtensorMap>> maps;
};
inline void march_free(MarchState& s) {
    cudaFree(s.d_mbu); cudaFree(s.d_mbv); cudaFree(s.d_partial); cudaFree(s.d_order); cudaFree(s.d_trace);
    s = MarchState{};
}

template <int WU, int WV, int NCH, int D>
inline int march_launch(TileState& s, MarchState& ms, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                        const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change, cudaStream_t st) {
    using L = MarchLayout<WU, WV, NCH, D>;
    constexpr int PUT = L::PUT, TW = L::TW, NW = L::NW;
    MarchParams p;
    p.w = w; p.d = d; p.fb = fb;
    p.nV = (d.kpad + TW - 1) / TW;
    p.nU = (w.nu + PUT - 1) / PUT;
    p.ntiles = p.nU * p.nV;
    p.spin_cycles = o.spin_limit << 9;
    static const int pause_env = getenv("TTCR_B200_PAUSE") ? atoi(getenv("TTCR_B200_PAUSE")) : 0;
    p.pause_ns = (unsigned)pause_env;
    static const int lpause_env = getenv("TTCR_B200_LPAUSE") ? atoi(getenv("TTCR_B200_LPAUSE")) : 0;
    p.lpause_ns = (unsigned)lpause_env;
    static const int pf_env = getenv("TTCR_B200_PF") ? atoi(getenv("TTCR_B200_PF")) : 0;
    p.pf_chunks = pf_env;
    p.trace_tile = getenv("TTCR_B200_TRACE_TILE") ? atoi(getenv("TTCR_B200_TRACE_TILE")) : 0;
    p.ctrl = s.d_ctrl;
    const char* trace_path = getenv("TTCR_B200_TRACE");
    if (trace_path && ms.trace_cap < p.ntiles) {
        cudaFree(ms.d_trace);
        TCK(cudaMalloc(&ms.d_trace, ((size_t)p.ntiles * 8 + 32 * 64 * 4) * sizeof(long long)));
        ms.trace_cap = p.ntiles;
    }
    p.trace = trace_path ? ms.d_trace : nullptr;
    const int rows = d.nj + TW + 2 * L::MARGIN + 8;
    if (!ms.d_mbu || ms.mb_tiles < p.ntiles || ms.mb_rows != rows || ms.mb_tw != TW || ms.mb_put != PUT) {
        TCK(cudaStreamSynchronize(st));
        auto occ_keep = ms.occ;
        march_free(ms);
        ms.occ = occ_keep;
        ms.mb_rows = rows; ms.mb_tiles = p.ntiles; ms.mb_tw = TW; ms.mb_put = PUT;
        const size_t nu = (size_t)ms.mb_tiles * ms.mb_rows * TW, nv = (size_t)ms.mb_tiles * ms.mb_rows * PUT;
        TCK(cudaMalloc(&ms.d_mbu, nu * 8));
        TCK(cudaMalloc(&ms.d_mbv, nv * 8));
        TCK(cudaMalloc(&ms.d_partial, (size_t)p.ntiles * NW * sizeof(double)));
        TCK(cudaMalloc(&ms.d_order, (size_t)p.ntiles * sizeof(int)));
        TCK(cudaMemsetAsync(ms.d_mbu, 0, nu * 8, st));
        TCK(cudaMemsetAsync(ms.d_mbv, 0, nv * 8, st));
        ms.serial = 0;
        ms.order_key = -1;
        if (trace_path) {
            TCK(cudaMalloc(&ms.d_trace, ((size_t)p.ntiles * 8 + 32 * 64 * 4) * sizeof(long long)));
            ms.trace_cap = p.ntiles;
            p.trace = ms.d_trace;
        }
    }
    MarchMail mail;
    mail.u = ms.d_mbu; mail.v = ms.d_mbv; mail.rows = ms.mb_rows;
    mail.serial = ++ms.serial;
    if (mail.serial == 0) {   // wrapped: clear the tags once every 2^32 sweeps
        TCK(cudaMemsetAsync(ms.d_mbu, 0, (size_t)ms.mb_tiles * ms.mb_rows * TW * 8, st));
        TCK(cudaMemsetAsync(ms.d_mbv, 0, (size_t)ms.mb_tiles * ms.mb_rows * PUT * 8, st));
        mail.serial = ms.serial = 1;
    }
    p.order = ms.d_order; p.partial = ms.d_partial;
    const int key = 6000000 + PUT * 10000 + TW * 40 + w.vlo;
    if (ms.order_key != key || ms.ntiles != p.ntiles) {
        // ticket order: a linear extension of (U-1,V) < (U,V), (U,V-1) < (U,V), sorted by the step at which a tile can start
        std::vector<std::pair<long long, int>> k(p.ntiles);
        static const int lag_env = getenv("TTCR_B200_LAG_U") ? atoi(getenv("TTCR_B200_LAG_U")) : 0;
        const long long lag_u = lag_env > 0 ? lag_env : PUT + 2;
        for (int U = 0; U < p.nU; ++U)
            for (int V = 0; V < p.nV; ++V) {
                const int va = std::max(V * TW, w.vlo);
                k[U * p.nV + V] = {U * lag_u + (long long)(va - w.joff) + V, U * p.nV + V};
            }
        std::stable_sort(k.begin(), k.end());
        std::vector<int> order(p.ntiles);
        for (int i = 0; i < p.ntiles; ++i) order[i] = k[i].second;
        TCK(cudaMemcpyAsync(ms.d_order, order.data(), p.ntiles * sizeof(int), cudaMemcpyHostToDevice, st));
        TCK(cudaStreamSynchronize(st));
        ms.order_key = key;
        ms.ntiles = p.ntiles;
    }
    const bool minus = (w.ri != 0) == (w.rj != 0);
    auto get_map = [&](const void* a, int bp) -> CUtensorMap {
        const std::vector<long long> kk = {(long long)(size_t)a, minus, L::BW, bp, d.kpad, d.qs, d.ni};
        for (auto& e : ms.maps)
            if (e.first == kk) return e.second;
        if (ms.maps.size() > 64) ms.maps.clear();
        ms.maps.push_back({kk, make_tile5_map(a, d, minus, L::BW, 2, bp)});
        return ms.maps.back().second;
    };
    const CUtensorMap tmT = get_map(tt, PUT + 1);
    const CUtensorMap tmS = get_map(slo, PUT);
    TCK(cudaMemsetAsync(s.d_ctrl, 0, sizeof(int), st));
    const int variant = (w.ri ? 1 : 0) | (w.rj ? 2 : 0) | (w.rk ? 4 : 0);
    auto run = [&](auto kern) {
        int occ = 0;
        for (auto& e : ms.occ)
            if (e.first == (const void*)kern) occ = e.second;
        if (!occ) {   // once per kernel instance and state (= slot, hence device)
            TCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
            TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, L::NT, L::BYTES));
            if (occ < 1) throw std::runtime_error("march kernel does not fit on an SM");
            ms.occ.push_back({(const void*)kern, occ});
        }
        int per_sm = occ;
        if (o.ctas_per_sm > 0) per_sm = std::min(per_sm, o.ctas_per_sm);
        const int grid = std::min(p.ntiles, per_sm * sm_count);
        kern<<<grid, L::NT, L::BYTES, st>>>(tmT, tmS, p, mail, tt, frozen, dx);
    };
    switch (variant) {
        case 0: run(k_sweep_march<WU, WV, NCH, D, false, false, false>); break;
        case 1: run(k_sweep_march<WU, WV, NCH, D, true, false, false>); break;
        case 2: run(k_sweep_march<WU, WV, NCH, D, false, true, false>); break;
        case 3: run(k_sweep_march<WU, WV, NCH, D, true, true, false>); break;
        case 4: run(k_sweep_march<WU, WV, NCH, D, false, false, true>); break;
        case 5: run(k_sweep_march<WU, WV, NCH, D, true, false, true>); break;
        case 6: run(k_sweep_march<WU, WV, NCH, D, false, true, true>); break;
        default: run(k_sweep_march<WU, WV, NCH, D, true, true, true>); break;
    }
    k_sum_partials<<<1, 256, 0, st>>>(ms.d_partial, p.ntiles * NW, d_change);
    TCK(cudaMemcpyAsync(s.h_abort, s.d_ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    TCK(cudaGetLastError());
    if (trace_path) {
        std::vector<long long> h((size_t)p.ntiles * 8 + 32 * 64 * 4);
        TCK(cudaStreamSynchronize(st));
        TCK(cudaMemcpy(h.data(), ms.d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        FILE* f = fopen(trace_path, "ab");
        if (f) {
            const int hdr[4] = {p.ntiles, p.nU, p.nV, PUT};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h.data(), sizeof(long long), (size_t)p.ntiles * 8, f);
            fclose(f);
        }
        if (const char* sp = getenv("TTCR_B200_TRACE_STEPS")) {   // [warp][step 200..263][4 stamps] of tile 0 (step-trace builds)
            FILE* g = fopen(sp, "ab");
            if (g) { fwrite(h.data() + (size_t)p.ntiles * 8, sizeof(long long), 32 * 64 * 4, g); fclose(g); }
        }
    }
    return 2;
}

template <typename T> inline bool march_supported(bool) { return false; }
template <> inline bool march_supported<float>(bool weno_stage) { return !weno_stage; }

template <typename T>
inline int march_sweep(TileState&, MarchState&, const TileOptions&, int, const SweepView&, const Dims&, T*, const T*, const uint32_t*,
                       const FrozenBox&, T, double*, cudaStream_t) {
    throw std::runtime_error("march kernel: fp32 only");
}
template <>
inline int march_sweep<float>(TileState& s, MarchState& ms, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                              const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change, cudaStream_t st) {
    // <warps along u, warps along v, chunk slots of the box ring, depth of the word rings>
    if (o.depth == 4) return march_launch<4, 4, 4, 16>(s, ms, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.depth == 7) return march_launch<4, 4, 7, 16>(s, ms, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    return march_launch<4, 4, 6, 16>(s, ms, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);                                     // 16 planes x 32 lanes
}

}  // namespace ttcrb200
