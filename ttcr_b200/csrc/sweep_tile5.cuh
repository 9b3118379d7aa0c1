// Sweep kernel "TILE5" (k_sweep_patch): register-patch tile march.
//
// What bounds a directional sweep on a B200 is not HBM but the 3N-step dependency chain and the number of
// instructions issued per node: k_sweep_tile4 issues ~200 warp instructions per 32
// nodes and synchronises 8 warps with a CTA barrier at every step.  This kernel keeps the decomposition
// (sheared layouts, tiles marching along the row axis m, tickets in dependency order, tagged mailboxes) and
// changes the work assignment so that almost every instruction is arithmetic of the Godunov update:
//
//   * a THREAD owns a patch of 4 consecutive lanes (v) x R consecutive planes (u) and marches along m.
//     Plane r of the patch runs r rows behind plane 0, so at macro-step a the thread updates the 4R nodes
//     (u0+r, a-r, 4l..4l+3): all mutually independent (ILP 4R), and
//         (u-1, m, v)    = the thread's own result of plane r-1 at the previous step      (register)
//         (u, m-1, v)    = own previous result                                            (register)
//         (u, m-1, v-1)  = own previous result of the lane to the left / one shuffle for element 0
//         (u+1, m, v)    = the "next old row" of plane r+1, which the thread loads anyway (register)
//         (u, m+1, v), (u, m+1, v+1) = one LDS.128 (+ one LDS.32 for the lane beyond the patch)
//     i.e. per 4 nodes: 2 LDS.128 (traveltime row, slowness row), 1 LDS.32, 1 SHFL, <= 1 STG.128.
//   * a WARP (128 lanes x R planes) is self-sufficient: lane 0 issues its own TMA boxes (one 3-D box of
//     traveltimes {132 lanes, C rows, R+1 planes}, one of slowness {128, C, R} per chunk of C steps) into a
//     private ring and the warp waits on its own mbarriers.  The plane skew is put into the TENSOR MAP
//     (plane stride = (rows per plane -+ 1) rows), so a box arrives already skewed and every smem address of
//     a step is ring base + immediate.
//   * there is NO CTA barrier in the march.  A CTA stacks NU warps along u (tile = NU*R planes x 128
//     lanes).  Warp w hands the results of its last plane to warp w+1 through a shared-memory ring of
//     tagged 8-byte words {row tag, value} (an aligned 8-byte shared access is single-copy atomic: the
//     reader that sees the tag has the value; no fence, no barrier).  Between tiles the same words travel
//     through global mailboxes (as in k_sweep_tile4): the last warp writes them, and an IMPORTER warp of the
//     consuming CTA polls them with a rolling two-row prefetch and re-tags them into the shared rings, so
//     compute warps only ever read shared memory.  The importer also feeds the v0-1 lane results of tile
//     V-1 (one word per plane and row) and synthesises +MAX where no neighbour exists.
//   * slots that are no node hold +MAX (traveltime) and NaN (slowness): the update of such a slot is NaN
//     and `t < old` fails, so the march needs no validity test at all.  Frozen nodes (source box) and the
//     grid's last planes are handled by a slow variant of the step that only the few affected warps run.
//
// fp32, first-order stage only.  Results are bit-identical to the other sweep kernels (same DAG, same
// arithmetic: update.cuh).
#pragma once
#include "sweep_tile4.cuh"

#ifndef TTCR_T5_CLOCK
#define TTCR_T5_CLOCK clock64   // gtime: globaltimer (ns, comparable across SMs, coarse)
#endif
#ifndef TTCR_T5_STEP_TRACE
#define TTCR_T5_STEP_TRACE 0   // 1: per-step clocks of one tile (TTCR_B200_TRACE_STEPS); costs ~20 % of a step
#endif

namespace ttcrb200 {

struct Tile5Mail {
    unsigned long long* u = nullptr;   // [tile][row][2 halves][32 lanes][2 words]: last plane of the tile (1 KiB per row)
    unsigned long long* v = nullptr;   // [tile][MARGIN + row][PU words]: lane v0+127 of every plane
    int rows = 0;                      // row stride per tile
    unsigned serial = 0;               // sweep launch counter = tag of the current sweep
    static constexpr int MARGIN = 8;
};

struct Tile5Params {
    SweepView w;
    Dims d;
    FrozenBox fb;
    int nU, nV, ntiles;
    long long spin_limit;
    const int* order;
    int* ctrl;
    double* partial;   // [tile][NU]
    long long* trace;
    int trace_a0, trace_tile;   // per-step clocks: first step and tile of the window (debug)
    int* dbg;                   // [tile][NU + 2][16]: what every warp was waiting for when a march was given up
    int pf_chunks;              // L2 prefetch distance of the loader, in chunks (0 = off)
    int perm;                   // 1: the warps that share a scheduler with a helper warp come first in the chain
};

constexpr int t5_round128(int x) { return (x + 127) / 128 * 128; }

template <int NU, int R, int C, int NCH, int DU_, int NG>
struct Tile5Layout {
    static constexpr int PU = NU * R;
    static constexpr int TW = 132;                                   // floats per traveltime row: 128 + 4
    static constexpr int TROW = TW * 4, SROW = 128 * 4;
    static constexpr int TPL = C * TROW, SPL = C * SROW;             // bytes per plane of a box
    // The NU warps form NG ring groups of GW warps (GP planes): one TMA box pair per group and chunk (issuing a
    // UTMALDG costs the loader ~400 cycles, so one pair per WARP starves the march), planes skewed by the tensor map.
    static constexpr int GW = NU / NG, GP = GW * R;
    static constexpr int CHB_T = t5_round128((GP + 1) * TPL), CHB_S = GP * SPL;
    static constexpr int GROUP_BYTES = NCH * (CHB_T + CHB_S);
    static constexpr int DU = DU_;                                   // rows per U ring
    static constexpr int URING = DU * 1024;
    static constexpr int DV = 64;                                    // macro-steps per V ring
    static constexpr int VRING = DV * R * 8;
    static constexpr int OFF_RING = 0;                               // NG x [T slots][S slots]
    static constexpr int OFF_U = OFF_RING + NG * GROUP_BYTES;        // NU U rings (ring w = input of warp w)
    static constexpr int OFF_V = OFF_U + NU * URING;                 // NU V rings
    static constexpr int OFF_BAR = OFF_V + NU * VRING;               // NG x NCH mbarriers
    static constexpr int OFF_PROG = OFF_BAR + NG * NCH * 8;          // NU progress counters
    static constexpr int OFF_FLG = OFF_PROG + NU * 4;                // dead, tile
    static constexpr int OFF_RED = (OFF_FLG + 8 + 7) / 8 * 8;
    static constexpr int BYTES = OFF_RED + NU * 8;
    static_assert(NU % NG == 0 && CHB_S % 128 == 0 && GROUP_BYTES % 128 == 0 && URING % 128 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(DV >= NU * R + DU + 8, "V ring too shallow for the skew between the warps");
    static_assert(GP + C + 4 <= GUARD, "guard rows too few");
};

// ---- shared / global access helpers ---------------------------------------------------------------
__device__ __forceinline__ float4 lds_f4(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_u4(unsigned a) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds_u2(unsigned a) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u4(unsigned a, unsigned x, unsigned y, unsigned z, unsigned w) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts_u2(unsigned a, unsigned x, unsigned y) {
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
// two tagged words {value, tag} x 2 with one 16-byte store / load (each 8-byte half is single-copy atomic)
__device__ __forceinline__ void st_mail2(unsigned long long* p, unsigned tag, float a, float b) {
    asm volatile(
        "{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%1, %2};\n\tmov.b64 y, {%3, %2};\n\tst.relaxed.gpu.global.v2.b64 [%0], {x, y};\n\t}" ::"l"(p),
        "r"(__float_as_uint(a)), "r"(tag), "r"(__float_as_uint(b))
        : "memory");
}
__device__ __forceinline__ uint4 ld_mail2(const unsigned long long* p) {
    uint4 v;
    asm volatile(
        "{\n\t.reg .b64 x, y;\n\tld.relaxed.gpu.global.v2.b64 {x, y}, [%4];\n\tmov.b64 {%0, %1}, x;\n\tmov.b64 {%2, %3}, y;\n\t}"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
        : "l"(p)
        : "memory");
    return v;
}
__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool vt_ghost(int vt, int kpad) { return vt >= kpad; }
__device__ __forceinline__ void stg_f4_stream(float* p, float x, float y, float z, float w) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void stg_f4_stream_if(float* p, float x, float y, float z, float w, int on) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p),
                 "f"(x), "f"(y), "f"(z), "f"(w), "r"(on)
                 : "memory");
}
__device__ __forceinline__ void sts_u4_if(unsigned a, unsigned x, unsigned y, unsigned z, unsigned w, bool on) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t@q st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};\n\t}" ::"r"(a), "r"(x),
                 "r"(y), "r"(z), "r"(w), "r"((int)on)
                 : "memory");
}
__device__ __forceinline__ void st_mail2_if(unsigned long long* p, unsigned tag, float a, float b, int on) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b64 x, y;\n\tsetp.ne.s32 q, %4, 0;\n\tmov.b64 x, {%1, %2};\n\tmov.b64 y, {%3, %2};\n\t"
        "@q st.relaxed.gpu.global.v2.b64 [%0], {x, y};\n\t}" ::"l"(p),
        "r"(__float_as_uint(a)), "r"(tag), "r"(__float_as_uint(b)), "r"(on)
        : "memory");
}
// TMA prefetch of a box into L2 (no shared-memory slot needed: L2 is the deep buffer of the march)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}
template <bool REV> __device__ __forceinline__ float4 ord4(float4 v) { return REV ? make_float4(v.w, v.z, v.y, v.x) : v; }

// Cold path of a march step, out of line so that the hot loop stays small: poll the ring words of step `tag-1`
// until they are all there (U words of this lane, V words, progress of the next warp) or the march is given up.
// Returns 0 when ready, 1 when another warp gave up, 2 on timeout.
template <int R>
__device__ __noinline__ int t5_wait_words(unsigned ar, unsigned av, unsigned a_nxprog, unsigned a_dead, unsigned tag, int need_u,
                                          int nx_need, long long spin_cycles) {
    const long long t0 = clock64();
    for (;;) {
        const uint4 xa = lds_u4(ar), xb = lds_u4(ar + 512);
        unsigned bad = need_u ? ((xa.y ^ tag) | (xa.w ^ tag) | (xb.y ^ tag) | (xb.w ^ tag)) : 0u;
#pragma unroll
        for (int r = 0; r < R; ++r) bad |= lds_u2(av + r * 8).y ^ tag;
        if (lds_i(a_nxprog) < nx_need) bad |= 1u;
        if (!bad) return 0;
        if (lds_i(a_dead)) return 1;
        if (clock64() - t0 > spin_cycles) return 2;
    }
}

// RJ: the sweep runs the row axis downwards, RK: the lane axis (boxes arrive in memory order).
template <int NU, int R, int C, int NCH, int DU_, int NG, bool RJ, bool RK>
__global__ void __launch_bounds__((NU + 2) * 32, ((NU <= 4 || (C == 2 && R == 1)) ? 2 : 1)) k_sweep_patch(const __grid_constant__ CUtensorMap tmT,
                                                                  const __grid_constant__ CUtensorMap tmS, Tile5Params p,
                                                                  Tile5Mail mail, float* __restrict__ tt,
                                                                  const uint32_t* __restrict__ frozen, float dx) {
    using L = Tile5Layout<NU, R, C, NCH, DU_, NG>;
    constexpr int PU = L::PU, DU = L::DU, DV = L::DV, GW = L::GW, GP = L::GP;
    constexpr int MG = Tile5Mail::MARGIN;
    constexpr int G = 2;   // rows of mailbox loads the importer keeps in flight (3 and 4 measured slower: more stale polls)
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const float MAXV = FLT_MAX;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned a_dead = sbase + L::OFF_FLG, a_tile = a_dead + 4;
    const unsigned a_prog = sbase + L::OFF_PROG;
    const unsigned serial = mail.serial;
    const long long spin_cycles = p.spin_limit << 9;
    unsigned gc = 0;   // chunks this warp's ring group has been sent since the kernel started: slot gc % NCH, parity (gc / NCH) & 1

    if (threadIdx.x == 0) {
        for (int i = 0; i < NG * NCH; ++i) mbar_init(sbase + L::OFF_BAR + 8 * i, 1);
        fence_mbar_init();
    }

    for (;;) {
        __syncthreads();   // everybody is done with the previous tile (rings, flags)
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sts_i(a_tile, (ab || t >= p.ntiles) ? -1 : t);
            sts_i(a_dead, 0);
        }
        // clear the tags of all rings and the progress counters
        for (unsigned o = threadIdx.x * 16; o < (unsigned)(L::OFF_BAR - L::OFF_U); o += (NU + 2) * 32 * 16)
            sts_u4(sbase + L::OFF_U + o, 0u, 0u, 0u, 0u);
        if (threadIdx.x < NU) sts_i(a_prog + 4 * threadIdx.x, 0);
        __syncthreads();
        const int ticket = lds_i(a_tile);
        if (ticket < 0) break;
        const int tile = p.order[ticket];
        const int U = tile / p.nV, V = tile - U * p.nV;
        const int u0 = U * PU, v0 = V * 128;
        const int va = max(v0, w.vlo), vb = min(v0 + 128, w.vhi);
        const int m_first = va - w.joff;
        const int nrows = (vb - va) + w.nj - 1;                // local rows 0 .. nrows-1 hold all nodes of the tile
        const int nsteps = nrows + R - 1;                      // macro-steps (plane r runs r rows behind plane 0)
        const int nchunks = (nsteps + C - 1) / C;
        const int nA = nchunks * C;                            // macro-steps executed (the surplus ones touch no node)
        // A ring group's boxes are indexed by s = step + (plane of the group): warp wg reads box row a + wg*R at step a
        auto group_chunks = [&](int grp_) {   // boxes sent to ring group grp_: up to the last row its last marching warp reads
            const int first_plane = u0 + grp_ * GP;
            if (first_plane > w.nu - 1) return 0;
            const int wl_rel = min(GW - 1, (w.nu - 1 - first_plane) / R);
            return (nA + wl_rel * R + C - 1) / C;
        };
        const bool has_u = U > 0, has_v = V > 0;
        const bool has_right = v0 + 128 < w.vhi;
        const bool has_down = U + 1 < p.nU;
        const int va_p = max(v0 - 128, w.vlo);
        const int nrows_p = (v0 - va_p) + w.nj - 1;            // rows of tile V-1
        const int dmf = m_first - (va_p - w.joff);             // its local row index of my local row 0
        if (p.trace && threadIdx.x == 0) {
            p.trace[tile * 8 + 0] = gtime();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[tile * 8 + 7] = smid;
        }

        if (warp == NU) {
            // ================= importer =====================================================================
            // macro-step a: U ring 0 <- row a of the last plane of tile U-1; V ring w, slot a <- for every plane
            // (w, r) the lane v0-1 result of row a-r of tile V-1.  +MAX where there is no such node.
            const unsigned long long* const mu_in = mail.u + ((size_t)(has_u ? tile - p.nV : tile) * mail.rows) * 128 + lane * 2;
            const int pw = lane / R, pr = lane - pw * R;                      // lane < PU: plane (pw, pr)
            const bool vlane = lane < PU;
            const bool vlive = vlane && u0 + pw * R <= w.nu - 1;             // the warp that owns the plane marches (sends words)
            const unsigned long long* const mv_in =
                mail.v + ((size_t)(has_v ? tile - 1 : tile) * mail.rows + MG) * PU + lane;
            const unsigned a_vring = sbase + L::OFF_V + pw * L::VRING + pr * 8;
            const unsigned a_uring = sbase + L::OFF_U + lane * 16;
            const unsigned a_prog0 = a_prog, a_progl = a_prog + 4 * (NU - 1);
            bool dead = false;
            auto give_up = [&](int why, int x) {
                if (atomicCAS(&p.ctrl[1], 0, why) == 0) { p.ctrl[2] = tile; p.ctrl[3] = x; p.ctrl[4] = lane; p.ctrl[5] = NU; }
                sts_i(a_dead, 1);
            };
            // rolling prefetch of depth G
            uint4 qa[G], qb[G];
            unsigned long long qv[G];
            auto v_row = [&](int a) { return a - pr - 1 + dmf; };            // row of tile V-1 this lane needs at step a
            auto v_inr = [&](int a) {
                const int rr = a - pr, rp = rr - 1 + dmf;
                return has_v && vlive && rr >= 0 && rr < nrows && rp >= 0 && rp < nrows_p;
            };
            auto issue = [&](int a, int s) {
                if (has_u && a < nrows) {
                    qa[s] = ld_mail2(mu_in + (size_t)a * 128);
                    qb[s] = ld_mail2(mu_in + (size_t)a * 128 + 64);
                }
                if (v_inr(a)) qv[s] = ld_mail(mv_in + (long long)v_row(a) * PU);
            };
#pragma unroll
            for (int s = 0; s < G; ++s) {
                qa[s] = qb[s] = make_uint4(0, 0, 0, 0);
                qv[s] = 0;
                if (s < nA) issue(s, s);
            }
            for (int a0 = 0; a0 < nA && !dead; a0 += G) {
#pragma unroll
                for (int s = 0; s < G; ++s) {
                    const int a = a0 + s;
                    if (a >= nA || dead) break;
                    const unsigned tag = (unsigned)(a + 1);
                    // ---- back-pressure: the slots about to be overwritten have been read
                    {
                        const long long t0 = clock64();
                        while (lds_i(a_prog0) < a - DU + 1 || lds_i(a_progl) < a - DV + 1) {
                            __nanosleep(200);
                            if (lds_i(a_dead)) { dead = true; break; }
                            if (clock64() - t0 > spin_cycles) { give_up(30, a); dead = true; break; }
                        }
                        if (dead) break;
                    }
                    // ---- V words
                    if (vlane) {
                        float val = MAXV;
                        if (v_inr(a)) {
                            unsigned long long x = qv[s];
                            if ((unsigned)(x >> 32) != serial) {
                                const unsigned long long* const q = mv_in + (long long)v_row(a) * PU;
                                const long long t0 = clock64();
                                for (;;) {
                                    x = ld_mail(q);
                                    if ((unsigned)(x >> 32) == serial) break;
                                    if (lds_i(a_dead)) { dead = true; break; }
                                    if (clock64() - t0 > spin_cycles) { give_up(31, a); dead = true; break; }
                                }
                            }
                            val = __uint_as_float((unsigned)x);
                        }
                        sts_u2(a_vring + (unsigned)(a & (DV - 1)) * (R * 8), __float_as_uint(val), tag);
                    }
                    // ---- U row
                    if (a < nrows) {
                        uint4 xa, xb;
                        if (has_u) {
                            xa = qa[s]; xb = qb[s];
                            if (xa.y != serial || xa.w != serial || xb.y != serial || xb.w != serial) {
                                const unsigned long long* const q = mu_in + (size_t)a * 128;
                                const long long t0 = clock64();
                                for (;;) {
                                    xa = ld_mail2(q);
                                    xb = ld_mail2(q + 64);
                                    if (xa.y == serial && xa.w == serial && xb.y == serial && xb.w == serial) break;
                                    if (lds_i(a_dead)) { dead = true; break; }
                                    if (clock64() - t0 > spin_cycles) { give_up(32, a); dead = true; break; }
                                }
                            }
                        } else {
                            xa.x = xa.z = xb.x = xb.z = __float_as_uint(MAXV);
                        }
                        const unsigned ar = a_uring + (unsigned)(a & (DU - 1)) * 1024;
                        sts_u4(ar, xa.x, tag, xa.z, tag);
                        sts_u4(ar + 512, xb.x, tag, xb.z, tag);
                    }
                    dead = __any_sync(0xffffffffu, dead);
                    if (dead && lane == 0) {
                        int* const q = p.dbg + ((size_t)tile * (NU + 2) + NU) * 16;
                        q[0] = 30; q[1] = a; q[2] = lds_i(a_prog0); q[3] = lds_i(a_progl); q[4] = has_u; q[5] = has_v; q[6] = nA; q[7] = nrows;
                    }
                    if (!dead && a + G < nA) issue(a + G, s);
                }
            }
        } else if (warp == NU + 1) {
            // ================= loader: lane wu feeds the ring of compute warp wu ============================
            // Box origins in memory coordinates (x: lane, y: row, z: plane); plane p of a box sits at row Y0 -+ p (the
            // skew is in the tensor map, see make_tile5_map).  T box: local rows cC - r + 1 .. of plane r = 0 .. R
            // (the "next old row" of every plane at every step of the chunk), S box: local rows cC - r .. of plane
            // r = 0 .. R-1.  Chunk c may be issued once the warp has finished chunk c - NCH (same ring slot).
            const int wu = lane;                               // (here: the ring group this lane feeds)
            const int u0w = u0 + wu * GP;                      // first plane of the group
            const int nch = wu < NG ? group_chunks(wu) : 0;
            const unsigned a_ring = sbase + L::OFF_RING + (wu < NG ? wu : 0) * L::GROUP_BYTES;
            const unsigned a_bar = sbase + L::OFF_BAR + (wu < NG ? wu : 0) * NCH * 8;
            const bool minus_map = (w.ri != 0) == (w.rj != 0);
            const int xT = RK ? p.d.kpad - 132 - v0 : v0;
            const int xS = RK ? p.d.kpad - 128 - v0 : v0;
            const int zT = w.ri ? p.d.ni - 1 - (u0w + GP) : u0w;
            const int zS = w.ri ? p.d.ni - 1 - (u0w + GP - 1) : u0w;
            const int rT0 = w.ri ? GP : 0, rS0 = w.ri ? GP - 1 : 0;
            const int yT0 = RJ ? GUARD + (w.nm - 1) - (m_first - rT0 + C) : GUARD + m_first - rT0 + 1;
            const int yS0 = RJ ? GUARD + (w.nm - 1) - (m_first - rS0 + C - 1) : GUARD + m_first - rS0;
            const int yT = minus_map ? yT0 + zT : yT0 - zT + p.d.ni;   // the "plus" map is based ni rows below the array
            const int yS = minus_map ? yS0 + zS : yS0 - zS + p.d.ni;
            const int dyc = RJ ? -C : C;                       // rows per chunk in memory direction
            // readers of the group's ring: its warps that march (wfirst .. wlast)
            const int wfirst = (wu < NG ? wu : 0) * GW;
            const int wlast = wfirst + max(0, min(GW - 1, (w.nu - 1 - u0w) / R));
            const unsigned a_wprog = a_prog + 4 * wlast;
            auto ring_progress = [&]() {   // box rows below this index have been read by everybody
                int m = 1 << 30;
                for (int ww = wfirst; ww <= wlast; ++ww) m = min(m, lds_i(a_prog + 4 * ww) + (ww - wfirst) * R);
                return m;
            };
            long long t0 = clock64();
            bool dead = false;
            int c = 0;
            // One uniform loop for the whole warp (lanes spinning separately would issue 8 instruction streams and steal
            // the issue slots of the compute warps on this scheduler): look, issue where a slot is free, sleep.
            for (;;) {
                const bool todo = c < nch;
                if (!__any_sync(0xffffffffu, todo)) break;
                // chunk c - NCH is read for the last time while step (c-NCH+1)*C - 2 prefetches its successor's operands
                const bool go = todo && (c < NCH || ring_progress() >= (c - NCH + 1) * C - 1);
                if (go) {
                    const unsigned g = gc + (unsigned)c;
                    const unsigned slot = g % NCH;
                    const unsigned mb = a_bar + 8 * slot;
                    mbar_expect_tx(mb, (GP + 1) * L::TPL + GP * L::SPL);
                    tma_load_3d(a_ring + slot * L::CHB_T, &tmT, xT, yT + c * dyc, zT, mb);
                    tma_load_3d(a_ring + NCH * L::CHB_T + slot * L::CHB_S, &tmS, xS, yS + c * dyc, zS, mb);
                    if (c + p.pf_chunks < nch && p.pf_chunks > 0) {   // pull a later chunk from DRAM into L2 now
                        tma_prefetch_3d(&tmT, xT, yT + (c + p.pf_chunks) * dyc, zT);
                        tma_prefetch_3d(&tmS, xS, yS + (c + p.pf_chunks) * dyc, zS);
                    }
                    ++c;
                }
                if (__any_sync(0xffffffffu, go)) {
                    t0 = clock64();
                    continue;   // somebody made progress: look again at once (another slot may be free)
                }
                __nanosleep(250);   // a chunk is C steps of slack
                if (lds_i(a_dead) || clock64() - t0 > spin_cycles) {
                    if (todo) {
                        int* const q = p.dbg + ((size_t)tile * (NU + 2) + NU + 1) * 16 + (wu & 1) * 8;
                        const int pg = lds_i(a_wprog);
                        const unsigned ur = sbase + L::OFF_U + wlast * L::URING + (unsigned)(pg & (DU - 1)) * 1024;
                        q[0] = 50; q[1] = c; q[2] = pg; q[3] = wu; q[4] = lds_i(ur + 4);          // tag of lane 0, word 0
                        q[5] = lds_i(ur + 16 * 13 + 4); q[6] = lds_i(ur + 512 + 16 * 31 + 12);    // lane 13 word 0, lane 31 word 3
                        q[7] = lds_i(sbase + L::OFF_V + wlast * L::VRING + (unsigned)(pg & (DV - 1)) * (R * 8) + 4);
                        if (!lds_i(a_dead) && atomicCAS(&p.ctrl[1], 0, 50) == 0) { p.ctrl[2] = tile; p.ctrl[3] = c; p.ctrl[4] = pg; p.ctrl[5] = wu; }
                    }
                    sts_i(a_dead, 1);
                    dead = true;
                    break;
                }
            }
            // a march that was given up may leave copies in flight: they must land before the CTA goes on
            if (dead) {
                const long long t1 = clock64();
                for (int cc = max(0, c - NCH); cc < c; ++cc) {
                    const unsigned g = gc + (unsigned)cc;
                    while (!mbar_test(a_bar + 8 * (g % NCH), (g / NCH) & 1) && clock64() - t1 < spin_cycles) {}
                }
            }
            gc += (unsigned)nch;
            __syncwarp();
        } else {
            // ================= compute warps ==============================================================
            // Position in the tile's chain of warps.  Warps 0,4 and 1,5 share their schedulers with the importer and the
            // loader (warp id mod 4) and are the slower ones; a slow warp makes every warp ABOVE it run ahead to the end
            // of its ring (latency through the tile), while warps below a slow one simply follow one step behind.  With
            // `perm` the four slower warps take the first four positions.
            const int wu = (NU == 8 && p.perm) ? ((warp & 1) | ((warp & 2) << 1) | ((warp & 4) >> 1)) : warp;
            const int u0w = u0 + wu * R;                       // first plane of this warp
            const int ulast = w.nu - 1;
            const bool edge = u0w + R > ulast;                 // some (u+1) plane of the patch does not exist
            const int nch = u0w > ulast ? 0 : nchunks;         // warps wholly past the grid's last plane do nothing
            const bool ghost = vt_ghost(v0 + 4 * lane, p.d.kpad);   // lanes past the end of the row: TMA zero-fills them
            const bool last_w = wu == NU - 1;
            const int vt = v0 + 4 * lane;                      // oriented lane of element 0
            const bool kill = vt + 4 >= w.vhi;                 // element 3 has no v+1 neighbour
            const int mail_v_out = (lane == 31 && has_right) ? 1 : 0;
            const unsigned tag_g = serial;
            // ---- ring addresses
            const int grp = wu / GW, wg = wu - grp * GW;       // ring group and position in it
            const unsigned a_ring = sbase + L::OFF_RING + grp * L::GROUP_BYTES;
            const unsigned a_bar = sbase + L::OFF_BAR + grp * NCH * 8;
            const int colT = RK ? 128 - 4 * lane : 4 * lane;   // column of the float4 in a 132-float row (memory order)
            const int colH = RK ? colT - 1 : colT + 4;         // column of lane v+4
            const int colS = RK ? 124 - 4 * lane : 4 * lane;
            unsigned aT[R + 1], aS[R];
#pragma unroll
            for (int r = 0; r <= R; ++r) aT[r] = a_ring + (w.ri ? GP - (wg * R + r) : wg * R + r) * L::TPL + colT * 4;
#pragma unroll
            for (int r = 0; r < R; ++r) aS[r] = a_ring + NCH * L::CHB_T + (w.ri ? GP - 1 - (wg * R + r) : wg * R + r) * L::SPL + colS * 4;
            const int dH = (colH - colT) * 4;
            const unsigned a_uin = sbase + L::OFF_U + wu * L::URING + lane * 16;
            const unsigned a_uout = sbase + L::OFF_U + (wu + 1) * L::URING + lane * 16;   // not used by the last warp
            const unsigned a_vin = sbase + L::OFF_V + wu * L::VRING;
            const unsigned a_myprog = a_prog + 4 * wu, a_nxprog = a_prog + 4 * (last_w ? wu : wu + 1);
            // ---- global addresses
            // element offset of (u0w + r, local row 0, element 0 .. 3 in memory order)
            long long e0[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                e0[r] = w.base + (long long)min(u0w + r, ulast) * w.su + (long long)m_first * w.sm + (long long)vt * w.sv - (RK ? 3 : 0);
            unsigned long long* const mu_out = mail.u + ((size_t)tile * mail.rows) * 128 + lane * 2;
            unsigned long long* const mv_out = mail.v + ((size_t)tile * mail.rows + MG) * PU + wu * R;
            // ---- frozen nodes: macro-steps in which this warp may touch one, and which patch elements
            unsigned fzmask = 0;   // bit r*4+e: element may be frozen (plane and lane inside the source box)
            int wz_lo = 1 << 30, wz_hi = -(1 << 30);
            if (p.fb.jhi >= p.fb.jlo) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int uu = u0w + r;
                    const int it = w.ri ? ulast - uu : uu;
                    if (uu > ulast || it < p.fb.ilo || it > p.fb.ihi) continue;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int v = vt + e;
                        if (v < w.vlo || v >= w.vhi) continue;
                        const int ko = v - w.vlo, kt = w.rk ? p.d.nk - 1 - ko : ko;
                        if (kt < p.fb.klo || kt > p.fb.khi) continue;
                        fzmask |= 1u << (r * 4 + e);
                        const int jol = RJ ? w.nj - 1 - p.fb.jhi : p.fb.jlo, joh = RJ ? w.nj - 1 - p.fb.jlo : p.fb.jhi;
                        // oriented j = m_first + row - v + joff  ->  row = j - joff + v - m_first, macro-step = row + r
                        wz_lo = min(wz_lo, jol - w.joff + v - m_first + r);
                        wz_hi = max(wz_hi, joh - w.joff + v - m_first + r);
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                wz_lo = min(wz_lo, __shfl_xor_sync(0xffffffffu, wz_lo, o));
                wz_hi = max(wz_hi, __shfl_xor_sync(0xffffffffu, wz_hi, o));
            }
            const int wz_cnt = wz_hi >= wz_lo ? wz_hi - wz_lo + 1 : 0;

            const bool trace_on = p.trace && tile == p.trace_tile && lane == 0 && wu < 8;
            (void)trace_on;
            bool dead = false;
            auto give_up = [&](int why, int x, int a) {
                if (atomicCAS(&p.ctrl[1], 0, why) == 0) { p.ctrl[2] = tile; p.ctrl[3] = x; p.ctrl[4] = a; p.ctrl[5] = wu; p.ctrl[6] = lane; }
                sts_i(a_dead, 1);
            };
            auto wait_chunk = [&](unsigned g) {
                const unsigned mb = a_bar + 8 * (g % NCH), par = (g / NCH) & 1;
                if (mbar_test(mb, par)) return;
                const long long t0 = clock64();
                while (!mbar_test(mb, par)) {
                    if (lds_i(a_dead)) { dead = true; break; }
                    if (clock64() - t0 > spin_cycles) { give_up(40, (int)g, 0); dead = true; break; }
                }
                if (dead && lane == 0) {
                    int* const q = p.dbg + ((size_t)tile * (NU + 2) + wu) * 16;
                    q[0] = 40; q[1] = (int)g; q[2] = (int)gc;
                }
            };

            // ---- prologue: fill the ring, load the old values of local row 0
            const unsigned g0 = gc;
            float4 told[R], tprev[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                // rows before local row 0 hold no node of the tile; a ghost lane must never see `t < old`
                const float ini = ghost ? 0.f : MAXV;
                told[r] = make_float4(ini, ini, ini, ini);
                tprev[r] = told[r];
            }
            if (nch > 0 && !ghost) told[0] = ord4<RK>(ldg_f4_stream(tt + e0[0]));
            if (nch == 0 && lane == 0) sts_i(a_myprog, 1 << 29);
            float acc = 0.f;
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 1] = gtime();

            // old values of the step about to run, loaded one step ahead (software pipeline): the "next old row" of
            // every plane (jp) with its lane v+4 (h), slowness (sl), and the next old row of plane u0w+R (up)
            float4 jp[R], sl[R], up;
            float h[R];
            auto load_old = [&](unsigned oT, unsigned oS, float4 (&xj)[R], float (&xh)[R], float4 (&xs)[R], float4& xu) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    xj[r] = ord4<RK>(lds_f4(aT[r] + oT));
                    xh[r] = lds_f(aT[r] + oT + dH);
                    xs[r] = ord4<RK>(lds_f4(aS[r] + oS));
                }
                xu = ord4<RK>(lds_f4(aT[R] + oT));
            };
            // row i of a chunk slot sits at byte offset (RJ ? C-1-i : i) * row bytes: boxes arrive in memory order
            constexpr int RT0 = (RJ ? C - 1 : 0) * L::TROW, RS0 = (RJ ? C - 1 : 0) * L::SROW;
            constexpr int DRT = RJ ? -L::TROW : L::TROW, DRS = RJ ? -L::SROW : L::SROW;
            const int soff = wg * R;   // this warp reads box row (step + soff) of its group's ring
            const int nchunks_g = group_chunks(grp);
            unsigned oT = 0, oS = 0;   // byte offsets (slot + row) of the NEXT step's old values
            const float QNAN = __int_as_float(0x7fc00000);
            const int ulim = nrows;    // rows 0 .. nrows-1 of the last plane are handed on

            if (nch > 0) {
                const unsigned g = g0 + (unsigned)(soff / C);
                wait_chunk(g);
                oT = (g % NCH) * L::CHB_T + (unsigned)(RT0 + (soff % C) * DRT);
                oS = (g % NCH) * L::CHB_S + (unsigned)(RS0 + (soff % C) * DRS);
                load_old(oT, oS, jp, h, sl, up);
            }
            dead = __any_sync(0xffffffffu, dead);
            const int nsteps_w = dead ? 0 : nch * C;

            unsigned oTc = oT, oSc = oS;   // slot offsets (row 0) of the chunk after the current one, once it has landed
            if (nch > 0 && (soff & (C - 1)) == C - 1) {   // the first step is the last row of its chunk
                const unsigned g = g0 + (unsigned)(soff / C) + 1;
                wait_chunk(g);
                oTc = (g % NCH) * L::CHB_T + RT0;
                oSc = (g % NCH) * L::CHB_S + RS0;
            }
            bool dead_u = false;
            static_assert(C >= 2 && (C & (C - 1)) == 0, "steps per chunk: a power of two, at least 2");

            // One march step.  Exactly one (rarely or every C-th time taken) branch: everything that is not the plain update
            // -- words not there yet, a chunk boundary coming up, frozen nodes / last planes -- sits behind it.
            auto step = [&](const int a) {
#if TTCR_T5_STEP_TRACE
                long long* const tr = (trace_on && (unsigned)(a - p.trace_a0) < 128u)
                                          ? p.trace + (size_t)p.ntiles * 8 + ((wu * 128 + (a - p.trace_a0)) * 4) : nullptr;
                if (tr) tr[0] = TTCR_T5_CLOCK();
#endif
                // ---- (1) the words of the neighbours (rings): (u0w-1, a, .), lane v0-1 of every plane, progress of warp wu+1
                const unsigned tag = (unsigned)(a + 1);
                const unsigned ar = a_uin + (unsigned)(a & (DU - 1)) * 1024;
                const unsigned av = a_vin + (unsigned)(a & (DV - 1)) * (R * 8);
                const bool need_u = a < nrows;
                const int nx_need = last_w ? -(1 << 30) : a - (R - 1) - DU + 1;   // the slot this step overwrites has been read
                uint4 xa = lds_u4(ar), xb = lds_u4(ar + 512);
                uint2 xv[R];
#pragma unroll
                for (int r = 0; r < R; ++r) xv[r] = lds_u2(av + r * 8);
                const int nxp = lds_i(a_nxprog);
                // ---- (2) lane l-1's previous result
                float km0[R];
#pragma unroll
                for (int r = 0; r < R; ++r) km0[r] = __shfl_up_sync(0xffffffffu, tprev[r].w, 1);
                // The shuffle is a convergence point: every lane has left step a-1, so its ring slots may be reused.  (Lanes
                // can leave the wait below at different times; reporting at the end of a step would let a producer overwrite
                // a slot that a late lane is still polling.)
                if (lane == 0) sts_i(a_myprog, a);
                // ---- (3) the one branch
                unsigned bad = need_u ? ((xa.y ^ tag) | (xa.w ^ tag) | (xb.y ^ tag) | (xb.w ^ tag)) : 0u;
#pragma unroll
                for (int r = 0; r < R; ++r) bad |= xv[r].y ^ tag;
                if (nxp < nx_need) bad |= 1u;
                const int sb = a + soff;                                             // box row of this step
                const bool f_bound = (sb & (C - 1)) == C - 2;                       // time to make sure the next chunk has landed
                const bool f_slow = edge || (unsigned)(a - wz_lo) < (unsigned)wz_cnt;
                if (bad != 0 || f_bound || f_slow) {
                    if (bad) {   // the wait itself is out of line
                        const int rc = t5_wait_words<R>(ar, av, a_nxprog, a_dead, tag, need_u ? 1 : 0, nx_need, spin_cycles);
                        xa = lds_u4(ar); xb = lds_u4(ar + 512);
#pragma unroll
                        for (int r = 0; r < R; ++r) xv[r] = lds_u2(av + r * 8);
                        if (rc) {
                            if (rc == 2) give_up(41, lds_i(a_nxprog), a);
                            dead = true;
                            int* const q = p.dbg + ((size_t)tile * (NU + 2) + wu) * 16;
                            if (lane == 0 || lane == 31) {
                                int* const ql = q + (lane ? 8 : 0);
                                ql[0] = 41; ql[1] = a; ql[2] = need_u ? (int)xa.y : -1; ql[3] = need_u ? (int)xb.w : -1; ql[4] = (int)xv[0].y;
                                ql[5] = lds_i(a_nxprog);
                            }
                            atomicOr((unsigned*)&q[6], 1u << lane);          // lanes that were polling
                            atomicMax(&q[7], a + 1);                          // ... and their step
                        }
                    }
                    if (f_bound) {   // warp-uniform
                        const int cn = sb / C + 1;
                        if (cn < nchunks_g) {
                            const unsigned g = g0 + (unsigned)cn;
                            wait_chunk(g);
                            oTc = (g % NCH) * L::CHB_T + RT0;
                            oSc = (g % NCH) * L::CHB_S + RS0;
                        }   // (after the last chunk: a harmless re-read of the last slot)
                        dead_u = __any_sync(0xffffffffu, dead);   // (uniform: leaves the loop for the whole warp)
                        if (p.trace && threadIdx.x == 0) {
                            const int q = nrows / 4, a1 = (a / C + 1) * C;
                            if (a1 - C < q && a1 >= q) p.trace[tile * 8 + 2] = gtime();
                            if (a1 - C < 2 * q && a1 >= 2 * q) p.trace[tile * 8 + 3] = gtime();
                            if (a1 - C < 3 * q && a1 >= 3 * q) p.trace[tile * 8 + 4] = gtime();
                        }
                    }
                    if (f_slow) {
                        // last planes of the grid, rows that may hold frozen nodes.  A node that must not change gets NaN
                        // slowness: its update is NaN and fails `t < old` like every non-node slot.
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            unsigned fm = ((unsigned)(a - wz_lo) < (unsigned)wz_cnt) ? (fzmask >> (r * 4)) & 15u : 0u;
                            if (fm) {
                                const long long eb = e0[r] + (long long)(a - r) * w.sm;
                                unsigned fz = 0;
                                if ((fm & 1u) && frozen_bit(frozen, eb + (RK ? 3 : 0))) fz |= 1u;
                                if ((fm & 2u) && frozen_bit(frozen, eb + (RK ? 2 : 1))) fz |= 2u;
                                if ((fm & 4u) && frozen_bit(frozen, eb + (RK ? 1 : 2))) fz |= 4u;
                                if ((fm & 8u) && frozen_bit(frozen, eb + (RK ? 0 : 3))) fz |= 8u;
                                fm = fz;
                            }
                            if (u0w + r > ulast) fm = 15u;
                            if (fm & 1u) sl[r].x = QNAN;
                            if (fm & 2u) sl[r].y = QNAN;
                            if (fm & 4u) sl[r].z = QNAN;
                            if (fm & 8u) sl[r].w = QNAN;
                            if (r == R - 1 && u0w + r >= ulast) up = make_float4(MAXV, MAXV, MAXV, MAXV);   // no (u+1) plane
                        }
                    }
                }
#if TTCR_T5_STEP_TRACE
                if (tr) tr[1] = TTCR_T5_CLOCK();
#endif
                // What sits between the arrival of the neighbours' words and the hand-off of this step's last plane is the
                // sweep's critical path (the hop from warp to warp, 1535 times per sweep): only the update itself is left
                // there.  The loads of the NEXT step's old values, the stores of this step's results and the change sum
                // come after the hand-off, in the shadow of the next wait.
                float4 um = make_float4(__uint_as_float(xa.x), __uint_as_float(xa.z), __uint_as_float(xb.x), __uint_as_float(xb.z));
                if (!need_u) um = make_float4(MAXV, MAXV, MAXV, MAXV);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (lane == 0) km0[r] = __uint_as_float(xv[r].x);
                    if (kill) h[r] = MAXV;
                }
#if TTCR_T5_STEP_TRACE
                if (tr) tr[2] = TTCR_T5_CLOCK();
#endif
                // ---- (4) update, last plane first (plane r reads the previous-step result of plane r-1); branch free
                float4 nn[R], oo[R];
                int chg[R];
#pragma unroll
                for (int r = R - 1; r >= 0; --r) {
                    float4 upv = (r == R - 1) ? up : jp[r + 1 < R ? r + 1 : r];
                    if (R > 1 && r < R - 1 && u0w + r >= ulast) upv = make_float4(MAXV, MAXV, MAXV, MAXV);
                    const float4 umv = (r == 0) ? um : tprev[r > 0 ? r - 1 : 0];
                    const float4 tp = tprev[r], j = jp[r], s = sl[r], o = told[r];
                    const float t0 = godunov(tmin(km0[r], j.y), tmin(tp.x, j.x), tmin(umv.x, upv.x), s.x * dx);
                    const float t1 = godunov(tmin(tp.x, j.z), tmin(tp.y, j.y), tmin(umv.y, upv.y), s.y * dx);
                    const float t2 = godunov(tmin(tp.y, j.w), tmin(tp.z, j.z), tmin(umv.z, upv.z), s.z * dx);
                    const float t3 = godunov(tmin(tp.z, h[r]), tmin(tp.w, j.w), tmin(umv.w, upv.w), s.w * dx);
                    const bool c0 = t0 < o.x, c1 = t1 < o.y, c2 = t2 < o.z, c3 = t3 < o.w;
                    float4 n;
                    n.x = c0 ? t0 : o.x; n.y = c1 ? t1 : o.y; n.z = c2 ? t2 : o.z; n.w = c3 ? t3 : o.w;
                    chg[r] = (c0 || c1 || c2 || c3) ? 1 : 0;
                    nn[r] = n; oo[r] = o;
                    if (r == R - 1) {
                        // ---- (5) last plane of the patch -> warp wu+1 (shared ring) / tile U+1 (global mailbox); predicated
                        const int row = a - (R - 1);
                        const bool inr = (unsigned)row < (unsigned)ulim;
                        const unsigned ao = a_uout + (unsigned)(row & (DU - 1)) * 1024;
                        const unsigned tg = (unsigned)(row + 1);
                        sts_u4_if(ao, __float_as_uint(n.x), tg, __float_as_uint(n.y), tg, inr && !last_w);
                        sts_u4_if(ao + 512, __float_as_uint(n.z), tg, __float_as_uint(n.w), tg, inr && !last_w);
                        const int og = (inr && last_w && has_down) ? 1 : 0;
                        st_mail2_if(mu_out + (long long)row * 128, tag_g, n.x, n.y, og);
                        st_mail2_if(mu_out + (long long)row * 128 + 64, tag_g, n.z, n.w, og);
                    }
                    // lane v0+127 of this plane -> tile V+1
                    st_mail_if(mv_out + (long long)(a - r) * PU + r, tag_g, n.w, mail_v_out);
                    tprev[r] = n;
                    told[r] = j;
                }
#if TTCR_T5_STEP_TRACE
                if (tr) tr[3] = TTCR_T5_CLOCK();
#endif
                // ---- (6) old values of the NEXT step (they do not depend on anybody)
                {
                    const bool first = ((sb + 1) & (C - 1)) == 0;   // the next step opens a chunk
                    oT = first ? oTc : oT + (unsigned)DRT;
                    oS = first ? oSc : oS + (unsigned)DRS;
                    load_old(oT, oS, jp, h, sl, up);
                }
                // ---- (7) results to the field, change sum
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float* const dst = tt + (e0[r] + (long long)(a - r) * w.sm);
                    const float4 n = nn[r], o = oo[r];
                    if (RK) stg_f4_stream_if(dst, n.w, n.z, n.y, n.x, chg[r]); else stg_f4_stream_if(dst, n.x, n.y, n.z, n.w, chg[r]);
                    acc += ((o.x - n.x) + (o.y - n.y)) + ((o.z - n.z) + (o.w - n.w));
                }
            };
            for (int a = 0; a < nsteps_w && !dead_u; a += 2) {   // nsteps_w is a multiple of C, hence even
                step(a);
                step(a + 1);
            }
            gc = g0 + (unsigned)nchunks_g;   // what the loader sent to the group
            if (lane == 0) sts_i(a_myprog, 1 << 29);   // release anybody still waiting for this warp
            double dacc = (double)acc;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
            if (lane == 0) p.partial[(size_t)tile * NU + wu] = dacc;
            if (p.trace && threadIdx.x == 0) p.trace[tile * 8 + 5] = gtime();
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
// 3-D map with the plane skew built in: plane z starts (qs -+ 1) rows after plane z-1
inline CUtensorMap make_tile5_map(const void* base, const Dims& d, bool minus, int bw, int br, int bp, bool nan_fill = false) {
    CUtensorMap m;
    // "plus" map: row = y + z * (qs + 1) would need negative y for the rows of high planes; the base is moved ni rows
    // down instead and every y carries +ni (addresses of in-bounds coordinates the kernel uses stay inside the array)
    if (!minus) base = static_cast<const char*>(base) - (size_t)d.ni * d.kpad * 4;
    const cuuint64_t gdim[3] = {(cuuint64_t)d.kpad, (cuuint64_t)(d.qs + 2 * d.ni + 64), (cuuint64_t)d.ni};
    const cuuint64_t gstr[2] = {(cuuint64_t)d.kpad * 4, (cuuint64_t)d.kpad * (cuuint64_t)(minus ? d.qs - 1 : d.qs + 1) * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)br, (cuuint32_t)bp};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = tmap_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

struct Tile5State {
    unsigned long long* d_mbu = nullptr;
    unsigned long long* d_mbv = nullptr;
    double* d_partial = nullptr;
    int* d_order = nullptr;
    int* d_dbg = nullptr;
    int dbg_n = 0;
    int mb_rows = 0, mb_tiles = 0, mb_pu = 0;
    int order_key = -1, ntiles = 0;
    unsigned serial = 0;
};
inline void tile5_free(Tile5State& s) {
    cudaFree(s.d_mbu); cudaFree(s.d_mbv); cudaFree(s.d_partial); cudaFree(s.d_order); cudaFree(s.d_dbg);
    s = Tile5State{};
}

template <int NU, int R, int C, int NCH, int DU_, int NG>
inline int tile5_launch(TileState& s, Tile5State& s5, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d,
                        float* tt, const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change,
                        cudaStream_t st) {
    using L = Tile5Layout<NU, R, C, NCH, DU_, NG>;
    constexpr int PU = L::PU;
    Tile5Params p;
    p.w = w; p.d = d; p.fb = fb;
    p.nV = (d.kpad + 127) / 128;
    p.nU = (w.nu + PU - 1) / PU;
    p.ntiles = p.nU * p.nV;
    p.spin_limit = o.spin_limit;
    p.ctrl = s.d_ctrl;
    static long long* d_trace = nullptr;   // debug only, not thread safe
    static int trace_cap = 0;
    const char* trace_path = getenv("TTCR_B200_TRACE");
    if (trace_path && trace_cap < p.ntiles) {
        cudaFree(d_trace);
        TCK(cudaMalloc(&d_trace, ((size_t)p.ntiles * 8 + 8192) * sizeof(long long)));
        trace_cap = p.ntiles;
    }
    p.trace = trace_path ? d_trace : nullptr;
    p.trace_a0 = getenv("TTCR_B200_TRACE_A0") ? atoi(getenv("TTCR_B200_TRACE_A0")) : 200;
    p.trace_tile = getenv("TTCR_B200_TRACE_TILE") ? atoi(getenv("TTCR_B200_TRACE_TILE")) : 0;
    const int rows = d.nj + 128 + 2 * Tile5Mail::MARGIN;
    if (!s5.d_mbu || s5.mb_tiles < p.ntiles || s5.mb_rows != rows || s5.mb_pu != PU) {
        TCK(cudaStreamSynchronize(st));
        tile5_free(s5);
        s5.mb_rows = rows; s5.mb_tiles = p.ntiles; s5.mb_pu = PU;
        const size_t nu = (size_t)s5.mb_tiles * s5.mb_rows * 128, nv = (size_t)s5.mb_tiles * s5.mb_rows * PU;
        TCK(cudaMalloc(&s5.d_mbu, nu * 8));
        TCK(cudaMalloc(&s5.d_mbv, nv * 8));
        TCK(cudaMalloc(&s5.d_partial, (size_t)p.ntiles * NU * sizeof(double)));
        TCK(cudaMalloc(&s5.d_order, (size_t)p.ntiles * sizeof(int)));
        s5.dbg_n = p.ntiles * (NU + 2) * 16;
        TCK(cudaMalloc(&s5.d_dbg, (size_t)s5.dbg_n * sizeof(int)));
        TCK(cudaMemsetAsync(s5.d_dbg, 0, (size_t)s5.dbg_n * sizeof(int), st));
        TCK(cudaMemsetAsync(s5.d_mbu, 0, nu * 8, st));
        TCK(cudaMemsetAsync(s5.d_mbv, 0, nv * 8, st));
        s5.serial = 0;
        s5.order_key = -1;
    }
    Tile5Mail mail;
    mail.u = s5.d_mbu; mail.v = s5.d_mbv; mail.rows = s5.mb_rows;
    mail.serial = ++s5.serial;
    if (mail.serial == 0) {   // wrapped: clear the tags once every 2^32 sweeps
        TCK(cudaMemsetAsync(s5.d_mbu, 0, (size_t)s5.mb_tiles * s5.mb_rows * 128 * 8, st));
        TCK(cudaMemsetAsync(s5.d_mbv, 0, (size_t)s5.mb_tiles * s5.mb_rows * PU * 8, st));
        mail.serial = s5.serial = 1;
    }
    p.order = s5.d_order; p.partial = s5.d_partial; p.dbg = s5.d_dbg;
    p.pf_chunks = getenv("TTCR_B200_PF") ? atoi(getenv("TTCR_B200_PF")) : 0;
    static const int perm_env = getenv("TTCR_B200_PERM") ? atoi(getenv("TTCR_B200_PERM")) : 0;
    p.perm = perm_env;
    const int key = 5000000 + PU * 100000 + w.vlo;   // the order depends on the tile height and on where the lanes start
    if (s5.order_key != key || s5.ntiles != p.ntiles) {
        // ticket order: a linear extension of (U-1,V) < (U,V), (U,V-1) < (U,V), sorted by estimated start time
        std::vector<std::pair<long long, int>> k(p.ntiles);
        // start-time estimate of a tile: one hop along U costs PU steps plus hand-off; a larger value (TTCR_B200_LAG_U)
        // spreads the tiles in flight further apart along m, trading start-up time for slack between a tile and its
        // predecessor
        static const int lag_env = getenv("TTCR_B200_LAG_U") ? atoi(getenv("TTCR_B200_LAG_U")) : 0;
        const long long lag_u = lag_env > 0 ? lag_env : PU + 4;
        for (int U = 0; U < p.nU; ++U)
            for (int V = 0; V < p.nV; ++V) {
                const int va = std::max(V * 128, w.vlo);
                k[U * p.nV + V] = {U * lag_u + (long long)(va - w.joff), U * p.nV + V};
            }
        std::stable_sort(k.begin(), k.end());
        std::vector<int> order(p.ntiles);
        for (int i = 0; i < p.ntiles; ++i) order[i] = k[i].second;
        TCK(cudaMemcpyAsync(s5.d_order, order.data(), p.ntiles * sizeof(int), cudaMemcpyHostToDevice, st));
        TCK(cudaStreamSynchronize(st));
        s5.order_key = key;
        s5.ntiles = p.ntiles;
    }
    struct MapKey { const void* a; int minus, bw, br, bp, kpad, qs, ni; CUtensorMap m; };
    static thread_local std::vector<MapKey> cache;
    auto get_map = [&](const void* a, bool minus, int bw, int br, int bp) -> CUtensorMap {
        for (auto& e : cache)
            if (e.a == a && e.minus == (int)minus && e.bw == bw && e.br == br && e.bp == bp && e.kpad == d.kpad && e.qs == d.qs && e.ni == d.ni)
                return e.m;
        if (cache.size() > 96) cache.clear();
        cache.push_back({a, (int)minus, bw, br, bp, d.kpad, d.qs, d.ni, make_tile5_map(a, d, minus, bw, br, bp)});
        return cache.back().m;
    };
    const bool minus = (w.ri != 0) == (w.rj != 0);
    const CUtensorMap tmT = get_map(tt, minus, L::TW, C, L::GP + 1);
    const CUtensorMap tmS = get_map(slo, minus, 128, C, L::GP);
    TCK(cudaMemsetAsync(s.d_ctrl, 0, sizeof(int), st));
    // (the four variants share one function-pointer type, so they are set up together, once)
    static int occ_cache = 0;
    if (!occ_cache) {
        auto prep = [&](auto kern) {
            int occ = 0;
            TCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
            TCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (NU + 2) * 32, L::BYTES));
            return occ;
        };
        int occ = prep(k_sweep_patch<NU, R, C, NCH, DU_, NG, false, false>);
        occ = std::min(occ, prep(k_sweep_patch<NU, R, C, NCH, DU_, NG, false, true>));
        occ = std::min(occ, prep(k_sweep_patch<NU, R, C, NCH, DU_, NG, true, false>));
        occ = std::min(occ, prep(k_sweep_patch<NU, R, C, NCH, DU_, NG, true, true>));
        if (occ < 1) throw std::runtime_error("tile5 kernel does not fit on an SM");
        occ_cache = occ;
    }
    int occ = occ_cache;
    if (o.ctas_per_sm > 0) occ = std::min(occ, o.ctas_per_sm);
    const int grid = std::min(p.ntiles, occ * sm_count);
    auto launch = [&](auto kern) { kern<<<grid, (NU + 2) * 32, L::BYTES, st>>>(tmT, tmS, p, mail, tt, frozen, dx); };
    if (w.rj) {
        if (w.rk) launch(k_sweep_patch<NU, R, C, NCH, DU_, NG, true, true>); else launch(k_sweep_patch<NU, R, C, NCH, DU_, NG, true, false>);
    } else {
        if (w.rk) launch(k_sweep_patch<NU, R, C, NCH, DU_, NG, false, true>); else launch(k_sweep_patch<NU, R, C, NCH, DU_, NG, false, false>);
    }
    k_sum_partials<<<1, 256, 0, st>>>(s5.d_partial, p.ntiles * NU, d_change);
    TCK(cudaMemcpyAsync(s.h_abort, s.d_ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    TCK(cudaGetLastError());
    if (trace_path) {
        std::vector<long long> h((size_t)p.ntiles * 8 + 8192);
        TCK(cudaStreamSynchronize(st));
        TCK(cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        FILE* f = fopen(trace_path, "ab");
        if (f) {
            const int hdr[4] = {p.ntiles, p.nU, p.nV, PU};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h.data(), sizeof(long long), (size_t)p.ntiles * 8, f);
            fclose(f);
        }
        if (const char* sp = getenv("TTCR_B200_TRACE_STEPS")) {   // per-step clocks of one tile (steps 200..327), 4 stamps per step
            FILE* g = fopen(sp, "ab");
            if (g) { fwrite(h.data() + (size_t)p.ntiles * 8, sizeof(long long), 8192, g); fclose(g); }
        }
    }
    return 2;
}

// debugging aid (TTCR_B200_DEBUG=1): after an abort, print what every warp was waiting for
inline void tile5_dump(Tile5State& s5) {
    if (!s5.d_dbg || !getenv("TTCR_B200_DEBUG")) return;
    std::vector<int> h(s5.dbg_n);
    cudaMemcpy(h.data(), s5.d_dbg, h.size() * sizeof(int), cudaMemcpyDeviceToHost);
    for (int i = 0; i + 8 <= s5.dbg_n; i += 8)
        if (h[i])
            fprintf(stderr, "t5dbg slot %d (tile,warp = slot/16): why %d | %d %d %d %d %d %d %d\n", i / 8, h[i], h[i + 1], h[i + 2], h[i + 3],
                    h[i + 4], h[i + 5], h[i + 6], h[i + 7]);
}

template <typename T> inline bool tile5_supported(bool) { return false; }
template <> inline bool tile5_supported<float>(bool weno_stage) { return !weno_stage; }

template <typename T>
inline int tile5_sweep(TileState&, Tile5State&, const TileOptions&, int, const SweepView&, const Dims&, T*, const T*,
                       const uint32_t*, const FrozenBox&, T, double*, cudaStream_t) {
    throw std::runtime_error("tile5 kernel: fp32 only");
}
template <>
inline int tile5_sweep<float>(TileState& s, Tile5State& s5, const TileOptions& o, int sm_count, const SweepView& w,
                              const Dims& d, float* tt, const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx,
                              double* d_change, cudaStream_t st) {
    // <compute warps, planes per thread, steps per TMA chunk, ring slots, rows per U ring, ring groups>
    // 16-plane tiles (8 warps x 2 planes): half as many hops along u, 8 independent updates per thread and step
    if (o.rows >= 2 && o.warps >= 8 && o.depth >= 16) return tile5_launch<8, 2, 4, 2, 4, 4>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.rows >= 2 && o.warps >= 8) return tile5_launch<8, 2, 2, 3, 8, 4>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.rows >= 2) return tile5_launch<4, 2, 2, 3, 8, 4>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.warps <= 4) return tile5_launch<4, 1, 4, 4, 8, 1>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    // 6 compute warps + importer + loader = 8 warps: two per scheduler, the helpers paired with one compute warp each
    if (o.warps == 6) return tile5_launch<6, 1, 4, 5, 4, 2>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.depth >= 16) return tile5_launch<8, 1, 4, 4, 8, 2>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    // shallower U rings: a fast warp can run at most DU steps ahead of the slower warp below it, and every step of that
    // slack is latency on the way through the tile
    // two CTAs per SM with 8-plane tiles: 2-row chunks (93 KB of rings) and a register cap of 102 (launch bounds)
    if (o.depth == 3) return tile5_launch<8, 1, 2, 3, 4, 2>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.depth == 2) return tile5_launch<8, 1, 4, 5, 2, 2>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    if (o.depth == 1) return tile5_launch<8, 1, 4, 5, 1, 2>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
    return tile5_launch<8, 1, 4, 5, 4, 2>(s, s5, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
}

}  // namespace ttcrb200
