// Sweep kernel "MARCH", WENO stage (k_sweep_march_weno_weno): the marching kernel of sweep_march.cuh with the third-order stencil.
//
// Reference: Grid3Drn::sweep_weno3 / update_node_weno3 / weno3_upwind (ttcr/Grid3Drn.h:2962-3484).  Same decomposition, warp
// patches (4 planes x 16 lanes, two nodes per thread), TMA boxes skewed by the tensor map, tagged ring words, mailboxes and
// importers as k_sweep_march_weno; what changes is the reach of the stencil, two nodes in every direction:
//   (u-2, m, v)      = what lane - 8 used as ITS (u-1) neighbour at the previous step         one more SHFL per node
//                      (lu == 0: the second half of the U ring word, sent by the warp above)
//   (u, m-2, v)      = own result of two steps ago                                              registers
//   (u, m-2, v-2)    = A: result A of lane - 1 two steps ago (SHFL; lv == 0: second V word)     B: the thread's own (m-1, v-1) of the
//                      previous step
//   (u+1, m, v), (u+2, m, v), (u, m+1, v .. v+1), (u, m+2, v .. v+2)                             the boxes carry PUT + 2 planes
// A step runs one row later than in the first-order kernel (plane pl updates local row t - 3 - pl at step t) so that the
// row two ahead is the box row the step prefetches anyway.  The per-axis estimates are axis_weno() of update.cuh with the TRUE
// node index of every axis (the branch order of the CPU code at the grid's faces), so the field is bit-identical to the plane
// kernels' WENO stage.  fp32 only.
#pragma once
#include "sweep_march.cuh"


namespace ttcrb200 {

template <int WU, int WV, int NCH>
struct MarchWLayout {
    static constexpr int NW = WU * WV, PUT = 4 * WU, TW = 16 * WV, BW = TW + 4, C = 4, D = 32;
    static constexpr int NT = (NW + 3) * 32;
    static constexpr int ROWB = BW * 4, PSB = C * ROWB;                               // bytes per box row / per plane of a box
    static constexpr int TBYTES = (PUT + 2) * PSB, SBYTES = PUT * PSB;                // bytes a box delivers (two halo planes)
    // the two mbarriers of a chunk slot live in the padding behind its T box
    static constexpr int OFF_EMPTY = TBYTES, OFF_FULL = TBYTES + 8;
    static constexpr int CHB_T = round128(TBYTES + 16), CHB_S = round128(SBYTES), CHB = CHB_T + CHB_S;
    static constexpr int USLOT = 256, VSLOT = 64;   // 8 x {A, tag, B, tag} (u-1), 8 x {A', tag, B', tag} (u-2) | 4 x {B, tag} (m-1, v-1), 4 x {A', tag} (m-2, v-2)
    static constexpr int URING = D * USLOT, VRING = D * VSLOT;
    static constexpr int OFF_UR = NCH * CHB;
    static constexpr int OFF_VR = OFF_UR + NW * URING;
    static constexpr int OFF_PROG = OFF_VR + NW * VRING;                              // steps completed, per compute warp
    static constexpr int OFF_CTL = OFF_PROG + (NW * 4 + 15) / 16 * 16;                // dead, tile
    static constexpr int BYTES = OFF_CTL + 16;
    static constexpr int BIG = 1 << 29;
    static_assert(PSB % 128 == 64, "plane stride of a box must put the two planes of a half-warp on different banks");
    static_assert(NCH * C < D, "word rings must outlast the box ring");
    static_assert(PUT + 2 * C + 6 <= GUARD, "guard rows too few");
    static_assert(TW % 32 == 0, "tiles are cut on kpad's granularity: no tile without a valid lane");
    static_assert(BYTES <= 227 * 1024, "shared memory");
};

// ---- geometry: one more step than the first-order kernel (a row is updated a step later) -------------------------------
template <int PUT, int TW, int C>
__host__ __device__ inline MarchTile marchw_tile(const SweepView& w, int nU, int nV, int tile) {
    MarchTile t = march_tile<PUT, TW, C>(w, nU, nV, tile);
    t.nA = (t.nrows + PUT + 2 + C - 1) / C * C;
    t.nch = t.nA / C + 1;
    const int va_p = (t.v0 - TW) > w.vlo ? (t.v0 - TW) : w.vlo;
    const int nrows_p = (t.v0 - va_p) + w.nj - 1;
    t.nA_left = (nrows_p + PUT + 2 + C - 1) / C * C;
    return t;
}

template <int WU, int WV, int NCH, bool RI, bool RJ, bool RK>
struct MarchWGeom {
    using L = MarchWLayout<WU, WV, NCH>;
    // byte offsets inside a chunk slot, relative to the pair of traveltime words {vA, vB} of the thread's plane in box row 0
    static constexpr int DR = RJ ? -L::ROWB : L::ROWB;                 // next box row
    static constexpr int DB = RK ? -8 : 8;                             // the pair of lanes {vA + 2, vA + 3}
    static constexpr int DUP = RI ? -L::PSB : L::PSB;                  // plane u + 1
    static constexpr int DUP2 = 2 * DUP;                               // plane u + 2
    static constexpr int DS = L::CHB_T + (RI ? -2 * L::PSB : 0);       // slowness, same plane
    // vl: first (even) lane of the thread inside the tile.  With RK the pair sits mirrored in memory: vB below vA.
    __host__ __device__ static int thread_off(int pl, int vl) {
        return (RI ? L::PUT + 1 - pl : pl) * L::PSB + (RK ? L::BW - 2 - vl : vl) * 4 + (RJ ? (L::C - 1) * L::ROWB : 0);
    }
    // box origins in tensor-map coordinates (x lane, y skewed row, z plane) of chunk c; `s` = 1 for the slowness box.
    // Box row b holds, for plane pl, row m_first + b - pl of the traveltimes and m_first - 2 + b - pl of slowness (`s` = 2).
    __host__ __device__ static int box_x(const Dims& d, const MarchTile& t) { return RK ? d.kpad - L::BW - t.v0 : t.v0; }
    __host__ __device__ static int box_z(const Dims& d, const MarchTile& t, int s) {
        return RI ? d.ni - 1 - (t.u0 + (s ? L::PUT - 1 : L::PUT + 1)) : t.u0;
    }
    __host__ __device__ static int box_y(const SweepView& w, const Dims& d, const MarchTile& t, int c, int s) {
        const int mf = t.m_first - s;
        if (!RJ) return GUARD + mf + L::C * c + t.u0 + (RI ? 1 : 0);
        return GUARD + (w.nm - 1) - mf - (L::C * c + L::C - 1) - t.u0 + d.ni - (RI ? 1 : 0);
    }
};

// Cold path of a WENO march step: wait for BOTH halves of the ring words of step `tgs` (same return values as march_wait_words)
static __device__ __noinline__ int marchw_wait_words(unsigned rUi, unsigned rVi, unsigned tgs, unsigned a_dead, long long spin_cycles, unsigned spin_polls,
                                              unsigned sleep_ns) {
    long long t0 = 0;
    for (unsigned it = 1;; ++it) {
        const uint4 xu = lds_u4(rUi), xu2 = lds_u4(rUi + 128);
        const uint2 xv = lds_u2(rVi), xv2 = lds_u2(rVi + 32);
        if (min(min(min(xu.y, xu.w), min(xu2.y, xu2.w)), min(xv.y, xv2.y)) >= tgs) return 0;
        if (it > spin_polls) __nanosleep(sleep_ns);
        if ((it & 15u) == 0) {
            if (lds_i(a_dead)) return 1;
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > spin_cycles) return 2;
        }
    }
}

template <int WU, int WV, int NCH, bool RI, bool RJ, bool RK>
__global__ void __launch_bounds__((WU * WV + 3) * 32, (MarchWLayout<WU, WV, NCH>::BYTES <= 112 * 1024 ? 2 : 1))
k_sweep_march_weno(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmS, MarchParams p, MarchMail mail,
              float* __restrict__ tt, const uint32_t* __restrict__ frozen, float dx) {
    using L = MarchWLayout<WU, WV, NCH>;
    using G = MarchWGeom<WU, WV, NCH, RI, RJ, RK>;
    constexpr int NW = L::NW, PUT = L::PUT, TW = L::TW, C = L::C, D = L::D, BIG = L::BIG;
    constexpr int USLOT = L::USLOT, VSLOT = L::VSLOT, URING = L::URING, VRING = L::VRING;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SweepView& w = p.w;
    const float MAXV = FLT_MAX;
    const unsigned sbase = (unsigned)pin((int)__cvta_generic_to_shared(smem_raw));
    const unsigned a_ctl = sbase + L::OFF_CTL, a_dead = a_ctl, a_tile = a_ctl + 4;
    const unsigned a_prog = sbase + L::OFF_PROG;
    const unsigned a_ur = sbase + L::OFF_UR, a_vr = sbase + L::OFF_VR;
    const unsigned serial = mail.serial;
    const long long spin_cycles = p.spin_cycles;
    // Every warp sees the same sequence of chunks (nch per tile); fill k of the CTA goes to slot k % NCH.  Compute warps keep
    // the parity of the "full" phase of the fill they wait for (cpar), the loader the round of the ring it is in (cpar).
    unsigned cslot = 0, cpar = 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NCH; ++i) {
            mbar_init(sbase + i * L::CHB + L::OFF_FULL, 1);
            mbar_init(sbase + i * L::CHB + L::OFF_EMPTY, NW);
        }
        fence_mbar_init();
    }

    for (;;) {
        __syncthreads();   // everybody is done with the previous tile
        if (threadIdx.x == 0) {
            const int t = atomicAdd(&p.ctrl[0], 1);
            const int ab = *((volatile int*)&p.ctrl[1]);
            sts_i(a_tile, (ab || t >= p.ntiles) ? -1 : t);
            sts_i(a_dead, 0);
        }
        if (threadIdx.x < NW) sts_i(a_prog + 4 * threadIdx.x, 0);
        __syncthreads();
        const int ticket = lds_i(a_tile);
        if (ticket < 0) break;
        const int tile = p.order[ticket];
        const MarchTile T = marchw_tile<PUT, TW, C>(w, p.nU, p.nV, tile);
        const int nA = T.nA;
        // Word rings.  The words of step t live in slot (t - 1) % D and are valid when their tag is >= t (a slot only ever
        // holds the tag of its step or of an older one).  Step 1 reads +MAX (the row before the tile's first row holds
        // no node).  Words nobody will send -- U words of the first warp row of a tile without a tile above, V words of the
        // first warp column of a tile without a tile to the left -- are +MAX with tag BIG in every slot.
        for (unsigned o = threadIdx.x * 16; o < (unsigned)(NW * URING); o += L::NT * 16) {
            const unsigned wr = o / URING, slot = (o % URING) / USLOT;
            const unsigned tag = (wr < (unsigned)WV && !T.has_u) ? (unsigned)BIG : (slot == 0 ? 1u : 0u);
            sts_u4(a_ur + o, __float_as_uint(MAXV), tag, __float_as_uint(MAXV), tag);
        }
        for (unsigned o = threadIdx.x * 16; o < (unsigned)(NW * VRING); o += L::NT * 16) {
            const unsigned wr = o / VRING, slot = (o % VRING) / VSLOT;
            const unsigned tag = (wr % WV == 0 && !T.has_v) ? (unsigned)BIG : (slot == 0 ? 1u : 0u);
            sts_u4(a_vr + o, __float_as_uint(MAXV), tag, __float_as_uint(MAXV), tag);
        }
        __syncthreads();
        if (p.trace && threadIdx.x == 0) {
            p.trace[tile * 16 + 0] = gtime();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p.trace[tile * 16 + 7] = smid;
        }

        if (warp == NW || warp == NW + 1) {
            // ================= importers: warp NW feeds the U words, warp NW + 1 the V words ==================================
            // Mailboxes are indexed by the STEP of the tile that writes them (one word per lane / plane and step, also for
            // the rows around its nodes, value +MAX):
            //   U ring of warp (0, l/16), slot of step t, pair (l%16)/2  <- mail.u[tile U-1][step t + PUT - 1][lane l]
            //   V ring of warp (pl/4, 0), slot of step t, word pl%4      <- mail.v[tile V-1][step t + dmf - 1][plane pl]
            // and the steps after the last step of those tiles get +MAX.  A lane moves TWO words per access (16 bytes); the
            // lanes of a warp are spread over S consecutive steps and every lane works on K steps at a time.  A round issues
            // the polls of all K steps back to back and only then looks at what came back: a warp has six scoreboards, so
            // polls that are issued and consumed one by one wait for each other (measured: 0.3 us per poll, and the importer
            // paced every tile); a round costs ONE L2 round trip whatever K is.
            constexpr int NPU = TW / 2, LU = NPU >= 32 ? 32 : NPU, SU = 32 / LU;       // pairs per step, lanes per step, steps per round and slot
            constexpr int NPV = PUT / 2, LV = NPV > 8 ? 16 : 8, SV = 32 / LV;
            static_assert(NPU <= 32 && NPV <= 16, "importer lane mapping");
            const bool isU = warp == NW;
            const int L_ = isU ? LU : LV, S = isU ? SU : SV;
            const int h = lane / L_, j = lane % L_;
            const bool act = isU ? (T.has_u != 0) : (T.has_v != 0 && j < NPV);
            const int ja = act ? j : 0;
            // last step with mailbox words; first word of step 0 for this lane; ring address of this lane's words; consumer progress
            const int a_end = !act ? 0 : (isU ? nA - PUT + 1 : T.nA_left - T.dmf + 1);
            const int stride = isU ? 2 * TW : 2 * PUT, half = isU ? TW : PUT;   // a mailbox row: [2 halves][words]
            const unsigned long long* const base =
                isU ? mail.u + ((size_t)(T.has_u ? tile - p.nV : tile) * mail.rows + (PUT - 1)) * (2 * TW) + 2 * ja
                    : mail.v + ((size_t)(T.has_v ? tile - 1 : tile) * mail.rows + (T.dmf - 1)) * (2 * PUT) + 2 * ja;
            // first half of the lane's words in a ring slot; the second half (the u-2 / (m-2, v-2) values) sits half a slot further
            const unsigned wring = isU ? a_ur + (unsigned)(ja >> 3) * URING + (unsigned)(ja & 7) * 16 : a_vr + (unsigned)((ja >> 1) * WV) * VRING + (unsigned)(ja & 1) * 16;
            const unsigned slotb = isU ? USLOT : VSLOT;
            const unsigned pcons = a_prog + 4 * (isU ? (ja >> 3) : (ja >> 1) * WV);
            constexpr int K = 4;
            uint4 q[K], q2[K];
            int a[K];
            bool have[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { a[k] = act ? 2 + h + S * k : nA + 1; have[k] = false; q[k] = q2[k] = make_uint4(0, 0, 0, 0); }
            long long t0 = 0;
            unsigned idle = 0;
            for (;;) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (!have[k] && a[k] <= a_end) {
                        q[k] = ld_mail2(base + (size_t)a[k] * stride);
                        q2[k] = ld_mail2(base + (size_t)a[k] * stride + half);
                    }
                const int lim = lds_i(pcons) + D;   // ring slot (a - 1) % D holds the words of step a - D until the consumer has passed it
                int done = 1, got = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (a[k] <= nA) {
                        const bool real = a[k] <= a_end;
                        const bool ok = !real || (q[k].y == serial && q[k].w == serial && q2[k].y == serial && q2[k].w == serial);
                        if (ok && a[k] <= lim) {
                            const unsigned tag = (unsigned)a[k], mx = __float_as_uint(MAXV);
                            const unsigned dst = wring + (unsigned)((a[k] - 1) & (D - 1)) * slotb;
                            sts_u4(dst + slotb / 2, real ? q2[k].x : mx, tag, real ? q2[k].z : mx, tag);
                            sts_u4(dst, real ? q[k].x : mx, tag, real ? q[k].z : mx, tag);
                            a[k] += S * K;
                            have[k] = false;
                            got = 1;
                        } else {
                            have[k] = ok;   // (words that wait for their ring slot are kept)
                        }
                        done &= a[k] > nA;
                    }
                }
                if (done) break;
                if (got) { idle = 0; continue; }
                if (idle >= 3) __nanosleep(p.sleep_ns);   // (nothing to deliver for three round trips: the neighbours are far behind)
                if ((++idle & 15u) == 0) {   // nothing moved for a while: is the march still alive?
                    if (lds_i(a_dead)) break;
                    if (idle == 16) t0 = clock64();
                    else if (clock64() - t0 > spin_cycles) {
                        if (atomicCAS(&p.ctrl[1], 0, 31) == 0) { p.ctrl[2] = tile; p.ctrl[3] = a[0]; p.ctrl[4] = a[1]; p.ctrl[5] = warp; p.ctrl[6] = lane; }
                        sts_i(a_dead, 1);
                        break;
                    }
                }
            }
            __syncwarp();
        } else if (warp == NW + 2) {
            // ================= loader ====================================================================================
            // One thread.  Fill k of the CTA lives in slot k % NCH; it may be sent once every compute warp has arrived on the
            // slot's "empty" barrier for fill k - NCH (the thread sleeps in try_wait: no polling).  Here cpar counts the
            // rounds of the ring: round r >= 1 waits for the "empty" phase r - 1.
            if (lane == 0) {
                const int nch = T.nch;
                const int xb = G::box_x(p.d, T);
                const int zT = G::box_z(p.d, T, 0), zS = G::box_z(p.d, T, 1);
                const int yT0 = G::box_y(w, p.d, T, 0, 0), yS0 = G::box_y(w, p.d, T, 0, 2);
                const int dyc = RJ ? -C : C;
                for (int ci = 0; ci < nch; ++ci) {
                    const unsigned slot = sbase + cslot * L::CHB;
                    bool ok = true;
                    if (cpar >= 1) {
                        const long long t0 = clock64();
                        while (!mbar_test(slot + L::OFF_EMPTY, (cpar - 1u) & 1u)) {
                            if (lds_i(a_dead)) { ok = false; break; }
                            if (clock64() - t0 > spin_cycles) {
                                if (atomicCAS(&p.ctrl[1], 0, 50) == 0) { p.ctrl[2] = tile; p.ctrl[3] = ci; p.ctrl[4] = (int)cslot; p.ctrl[5] = NW + 2; }
                                sts_i(a_dead, 1);
                                ok = false;
                                break;
                            }
                        }
                    }
                    if (!ok) break;
                    mbar_expect_tx(slot + L::OFF_FULL, L::TBYTES + L::SBYTES);
                    tma_load_3d(slot, &tmT, xb, yT0 + ci * dyc, zT, slot + L::OFF_FULL);
                    tma_load_3d(slot + L::CHB_T, &tmS, xb, yS0 + ci * dyc, zS, slot + L::OFF_FULL);
                    if (p.pf_chunks > 0 && ci + p.pf_chunks < nch) {
                        tma_prefetch_3d(&tmT, xb, yT0 + (ci + p.pf_chunks) * dyc, zT);
                        tma_prefetch_3d(&tmS, xb, yS0 + (ci + p.pf_chunks) * dyc, zS);
                    }
                    if (++cslot == NCH) { cslot = 0; ++cpar; }
                }
            }
        } else {
            // ================= compute warps =============================================================================
            const int lw = warp;
            const int wu = lw / WV, wv = lw - wu * WV, lu = lane >> 3, lv = lane & 7;
            const int pl = 4 * wu + lu, vl = 16 * wv + 2 * lv;
            const int u = T.u0 + pl, vt = T.v0 + vl;
            const int ulast = w.nu - 1;
            // Lanes and planes beyond the arrays are filled with NaN by the TMA unit (the tensor maps ask for it): as an upwind
            // or downwind neighbour a NaN drops out of the minima, as an old value it never lets `t < old` hold.
            const bool ghost = vt >= p.d.kpad || u > ulast;
            const bool out_u_in = wu < WU - 1, out_v_in = wv < WV - 1;
            const unsigned aUin = a_ur + (unsigned)lw * URING + lv * 16;   // pair lv of a slot of my U ring
            const unsigned aVin = a_vr + (unsigned)lw * VRING + lu * 8;    // word lu of a slot of my V ring
            constexpr unsigned OUT_U = WV * URING, OUT_V = VRING;          // the rings of the warp below / to the right
            const unsigned a_myprog = (unsigned)pin((int)(a_prog + 4 * lw));
            const unsigned toff = (unsigned)pin((int)G::thread_off(pl, vl));
            // the thread's pair of step 1 (two rows before the tile's first one), and the byte offset of the current step's pair from it
            char* const pg0 = (char*)(tt + (w.base + (long long)min(u, ulast) * w.su + (long long)(T.m_first - 2 - pl) * w.sm + (long long)(RK ? vt + 1 : vt) * w.sv));
            const int rowbytes = (int)w.sm * 4;
            int poff = 0;
            // ---- slow-path condition: frozen nodes (source box)
            bool fzme = false;
            int wz_lo = 1 << 28, wz_hi = -(1 << 28);
            if (p.fb.jhi >= p.fb.jlo && u <= ulast) {
                const int it = w.ri ? ulast - u : u;
                bool kin = false;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int ve = vt + e;
                    const int ko = ve - w.vlo, kt = w.rk ? p.d.nk - 1 - ko : ko;
                    kin = kin || (ve >= w.vlo && ve < w.vhi && kt >= p.fb.klo && kt <= p.fb.khi);
                }
                if (it >= p.fb.ilo && it <= p.fb.ihi && kin) {
                    fzme = true;
                    const int jol = RJ ? w.nj - 1 - p.fb.jhi : p.fb.jlo, joh = RJ ? w.nj - 1 - p.fb.jlo : p.fb.jhi;
                    // oriented j = m_first + r - v + joff on local row r, which is updated at step t = r + 3 + pl
                    wz_lo = jol - w.joff + vt - T.m_first + 3 + pl;
                    wz_hi = joh - w.joff + vt + 1 - T.m_first + 3 + pl;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                wz_lo = min(wz_lo, __shfl_xor_sync(0xffffffffu, wz_lo, o));
                wz_hi = max(wz_hi, __shfl_xor_sync(0xffffffffu, wz_hi, o));
            }
            // lane roles, one bit each, in a register the compiler cannot rematerialise from %tid
            enum : unsigned { F_L0 = 1, F_U0 = 2, F_V0 = 4, F_STU = 8, F_STV = 16, F_OMU = 32, F_OMV = 64, F_FZ = 512 };
            const unsigned fl = (unsigned)pin((int)((lane == 0 ? F_L0 : 0u) | (lu == 0 ? F_U0 : 0u) | (lv == 0 ? F_V0 : 0u) | ((lu == 3 && out_u_in) ? F_STU : 0u) |
                                                    ((lv == 7 && out_v_in) ? F_STV : 0u) | ((lu == 3 && !out_u_in && T.has_down) ? F_OMU : 0u) |
                                                    ((lv == 7 && !out_v_in && T.has_right) ? F_OMV : 0u) |
                                                    (fzme ? F_FZ : 0u)));
            const float QNAN = __int_as_float(0x7fc00000);

            int dead = 0;
            auto give_up = [&](int why, int x, int a) {
                if (atomicCAS(&p.ctrl[1], 0, why) == 0) { p.ctrl[2] = tile; p.ctrl[3] = x; p.ctrl[4] = a; p.ctrl[5] = lw; p.ctrl[6] = lane; }
                sts_i(a_dead, 1);
            };
            auto wait_full = [&](unsigned a_full, unsigned par, int why) {
                if (__builtin_expect(mbar_test(a_full, par), 1)) return;
                const int st = march_wait_full(a_full, par, a_dead, spin_cycles);
                if (st == 2) give_up(why, (int)cslot, (int)par);
                if (st) dead = 1;
            };
            // ---- prologue: first chunk of the tile, operands of step 1
            unsigned sb = sbase + cslot * L::CHB;   // chunk slot the current group reads
            wait_full(sb + L::OFF_FULL, cpar, 40);
            int deadw = __any_sync(0xffffffffu, dead);
            unsigned rB = sb + toff;
            // true node indices of the three axes (axis_weno branches on them at the grid's faces): i and k are the thread's,
            // j changes by one per step; node B sits one lane further, i.e. one j lower on the same row
            const int it = RI ? ulast - u : u;
            const int koA = vt - w.vlo, ktA = RK ? p.d.nk - 1 - koA : koA, ktB = RK ? ktA - 1 : ktA + 1;
            int jtA;   // of the node of the current step
            {
                const int jo = T.m_first + (1 - 3 - pl) - vt + w.joff;   // step 1
                jtA = RJ ? w.nj - 1 - jo : jo;
            }
            const int nci = p.d.ni - 1, ncj = p.d.nj - 1, nck = p.d.nk - 1;
            const float ini = ghost ? 0.f : MAXV;
            // results of the last two steps (rows m-1, m-2), the (u-1) and (m-1, v-1) values the previous step used
            float p0 = ini, p1 = ini, pp0 = ini, pp1 = ini, q0 = ini, q1 = ini, kprevA = ini;
            // old values: O row m, A1 row m+1 (+ lane vA+2), C1 plane u+1 row m; A2 / B2 row m+2 (4 lanes), Cn plane u+1 row m+1, D0 plane u+2 row m
            float O0 = ini, O1 = ini, A10 = ini, A11 = ini, B1x = ini, C10 = ini, C11 = ini;
            float A20, A21, B20, B21, Cn0, Cn1, D00, D01, s0, s1;
            {
                const float2 a = lds_f2(rB), b = lds_f2(rB + (unsigned)G::DB), c = lds_f2(rB + (unsigned)G::DUP), d2 = lds_f2(rB + (unsigned)G::DUP2),
                             e = lds_f2(rB + (unsigned)G::DS);
                A20 = RK ? a.y : a.x; A21 = RK ? a.x : a.y;
                B20 = RK ? b.y : b.x; B21 = RK ? b.x : b.y;
                Cn0 = RK ? c.y : c.x; Cn1 = RK ? c.x : c.y;
                D00 = RK ? d2.y : d2.x; D01 = RK ? d2.x : d2.y;
                s0 = RK ? e.y : e.x; s1 = RK ? e.x : e.y;
            }
            float acc = 0.f;
            if (p.trace && threadIdx.x == 0) p.trace[tile * 16 + 1] = gtime();

            // Groups of C steps (tags t .. t+C-1) read chunk `sb`; groups in [ts0, ts1) may touch a frozen node.
            const unsigned t_end = (unsigned)nA + 1u;
            unsigned ts0 = t_end, ts1 = t_end;
            if (wz_hi >= wz_lo) {
                ts0 = (unsigned)min(max(((wz_lo - 1) / C) * C + 1, 1), (int)t_end);
                ts1 = (unsigned)min(max(((wz_hi - 1) / C + 1) * C + 1, 1), (int)t_end);
            }
            unsigned long long* mu = mail.u + ((size_t)tile * mail.rows + 1) * (2 * TW) + vl;    // [tile][step][2][lane]
            unsigned long long* mv = mail.v + ((size_t)tile * mail.rows + 1) * (2 * PUT) + pl;   // [tile][step][2][plane]
            unsigned t = 1;
            unsigned so = 0;                  // byte offset of the group's first slot in a U ring (V ring: so / 4)
            unsigned gU = aUin, gV = aVin;

            // values at oriented offsets -2 .. +2 of one axis -> axis_weno in TRUE axis order
            auto axis = [&](auto rev_c, float m2, float m1, float v0, float p1v, float p2v, int q, int nc) -> float {
                constexpr bool REV = decltype(rev_c)::value != 0;
                return REV ? axis_weno<float>(p2v, p1v, v0, m1, m2, q, nc, dx) : axis_weno<float>(m2, m1, v0, p1v, p2v, q, nc, dx);
            };

            // One march step: ring words at rUi / rVi must carry a tag >= tgs, outputs go to rUo / rVo with tag tgs + 1, the next
            // step's operands are at `ao`.  TWO copies of this body exist in the kernel (the group loop below is unrolled by two and
            // the frozen-node and mailbox cases are run-time predicates): a step is ~500 instructions, and every copy more is
            // 8 KB that eight warps at eight different places of the loop pull through a 32 KB instruction cache.
            auto step = [&](const unsigned tgs, const unsigned rUi, const unsigned rVi, const unsigned rUo, const unsigned rVo, const unsigned ao,
                            const bool slowg) {
                // ---- (1) everything this step reads from shared memory
                uint4 xu = lds_u4(rUi), xu2 = lds_u4(rUi + USLOT / 2);
                uint2 xv = lds_u2(rVi), xv2 = lds_u2(rVi + VSLOT / 2);
                const float2 na = lds_f2(ao), nb = lds_f2(ao + (unsigned)G::DB), nc = lds_f2(ao + (unsigned)G::DUP), nd = lds_f2(ao + (unsigned)G::DUP2),
                             ns = lds_f2(ao + (unsigned)G::DS);
                // ---- (2) the upwind neighbours inside the patch
                float um10 = __shfl_up_sync(0xffffffffu, p0, 8), um11 = __shfl_up_sync(0xffffffffu, p1, 8);
                float um20 = __shfl_up_sync(0xffffffffu, q0, 8), um21 = __shfl_up_sync(0xffffffffu, q1, 8);
                float km1A = __shfl_up_sync(0xffffffffu, p1, 1), km2A = __shfl_up_sync(0xffffffffu, pp0, 1);
                // ---- (3) the one branch: words not there yet
                if (__builtin_expect(__any_sync(0xffffffffu, min(min(min(xu.y, xu.w), min(xu2.y, xu2.w)), min(xv.y, xv2.y)) < tgs), 0)) {
                    const int st = marchw_wait_words(rUi, rVi, tgs, a_dead, spin_cycles, p.spin_polls, p.sleep_ns);
                    if (st == 2) give_up(41, (int)tgs, (int)min(xu.y, xv.y));
                    if (st) dead = 1;
                    xu = lds_u4(rUi); xu2 = lds_u4(rUi + USLOT / 2); xv = lds_u2(rVi); xv2 = lds_u2(rVi + VSLOT / 2);
                }
                if (slowg && (fl & F_FZ)) {   // (the pair shares a mask word: e is even)
                    const long long e = (long long)((float*)(pg0 + poff) - tt);
                    const unsigned bits = frozen[e >> 5] >> (e & 31);
                    if (bits & (RK ? 2u : 1u)) s0 = QNAN;
                    if (bits & (RK ? 1u : 2u)) s1 = QNAN;
                }
                if (fl & F_U0) {
                    um10 = __uint_as_float(xu.x); um11 = __uint_as_float(xu.z);
                    um20 = __uint_as_float(xu2.x); um21 = __uint_as_float(xu2.z);
                }
                if (fl & F_V0) { km1A = __uint_as_float(xv.x); km2A = __uint_as_float(xv2.x); }
                // ---- (4) the two updates: node A (lane vA), node B (lane vA + 1: its (m-1, v-1) is A's previous result, its
                //          (m-2, v-2) is what A took as (m-1, v-1) at the previous step)
                const int jtB = RJ ? jtA + 1 : jtA - 1;
                const float auA = axis(IntC<RI>(), um20, um10, O0, C10, D00, it, nci);
                const float ajA = axis(IntC<RJ>(), pp0, p0, O0, A10, A20, jtA, ncj);
                const float akA = axis(IntC<RK>(), km2A, km1A, O0, A11, B20, ktA, nck);
                const float auB = axis(IntC<RI>(), um21, um11, O1, C11, D01, it, nci);
                const float ajB = axis(IntC<RJ>(), pp1, p1, O1, A11, A21, jtB, ncj);
                const float akB = axis(IntC<RK>(), kprevA, p0, O1, B1x, B21, ktB, nck);
                const float tA = godunov(akA, ajA, auA, s0 * dx);
                const float tB = godunov(akB, ajB, auB, s1 * dx);
                const float n0 = fminf(tA, O0), n1 = fminf(tB, O1);   // (NaN -> old: slots that are no node, frozen nodes)
                // ---- (5) hand-off: warp below / to the right (shared rings), tiles U+1 / V+1 (global mailboxes)
                sts_u4_ifu(rUo + USLOT / 2, __float_as_uint(um10), tgs + 1, __float_as_uint(um11), tgs + 1, fl & F_STU);
                sts_u4_ifu(rUo, __float_as_uint(n0), tgs + 1, __float_as_uint(n1), tgs + 1, fl & F_STU);
                sts_u2_ifu(rVo + VSLOT / 2, __float_as_uint(p0), tgs + 1, fl & F_STV);
                sts_u2_ifu(rVo, __float_as_uint(n1), tgs + 1, fl & F_STV);
                st_mail2_if(mu + TW, serial, um10, um11, (int)(fl & F_OMU));   // (mu, mv: this step's mailbox rows)
                st_mail2_if(mu, serial, n0, n1, (int)(fl & F_OMU));
                st_mail_if(mv + PUT, serial, p0, (int)(fl & F_OMV));
                st_mail_if(mv, serial, n1, (int)(fl & F_OMV));
                mu += 2 * TW; mv += 2 * PUT;
                // ---- (6) result, change sum, rotate the operands
                stg_f2_stream_if((float*)(pg0 + poff), RK ? n1 : n0, RK ? n0 : n1, (n0 < O0 || n1 < O1) ? 1 : 0);
                poff += rowbytes;
                acc += (O0 - n0) + (O1 - n1);
                kprevA = km1A; q0 = um10; q1 = um11;
                pp0 = p0; pp1 = p1; p0 = n0; p1 = n1;
                O0 = A10; O1 = A11; A10 = A20; A11 = A21; B1x = B20; C10 = Cn0; C11 = Cn1;
                A20 = RK ? na.y : na.x; A21 = RK ? na.x : na.y;
                B20 = RK ? nb.y : nb.x; B21 = RK ? nb.x : nb.y;
                Cn0 = RK ? nc.y : nc.x; Cn1 = RK ? nc.x : nc.y;
                D00 = RK ? nd.y : nd.x; D01 = RK ? nd.x : nd.y;
                s0 = RK ? ns.y : ns.x; s1 = RK ? ns.x : ns.y;
                jtA += RJ ? -1 : 1;
            };
            // A group = the C steps that read chunk `sb`; its last step takes the next operands from the next chunk.
            while (t < t_end && !deadw) {
                const bool slowg = t >= ts0 && t < ts1;
                // chunk slot of the next group (and the parity of its "full" phase: it flips when the ring wraps)
                unsigned nslot = cslot + 1, npar = cpar;
                if (nslot == NCH) { nslot = 0; npar ^= 1u; }
                const unsigned sbn = sbase + nslot * L::CHB;
                const unsigned so_n = (so + C * USLOT) & (unsigned)(URING - 1);
                const unsigned gUn = aUin + so_n, gVn = aVin + (so_n >> 2);
                unsigned aU = gU, aV = gV, ao = rB;
#pragma unroll 2   // (two copies of the step: half of the register moves of the rolled loop, 18 KB of hot code; measured 933 -> 887 ms per 512^3 solve, four copies 876 ms)
                for (int r = 0; r < C; ++r) {
                    const bool last = r == C - 1;
                    if (last) wait_full(sbn + L::OFF_FULL, npar, 42);   // the chunk the last step takes the next operands from
                    ao = last ? sbn + toff : ao + (unsigned)G::DR;
                    step(t + (unsigned)r, aU, aV, last ? gUn + OUT_U : aU + OUT_U + USLOT, last ? gVn + OUT_V : aV + OUT_V + VSLOT, ao, slowg);
                    aU += USLOT; aV += VSLOT;
                }
                // every lane has read the group's words and the chunk's last row (its values were used by the update above)
                __syncwarp();
                mbar_arrive_ifu(sb + L::OFF_EMPTY, fl & F_L0);
                t += C;
                sts_i_ifu(a_myprog, (int)t - 1, fl & F_L0);
                sb = sbn; rB = sbn + toff; cslot = nslot; cpar = npar;
                so = so_n; gU = gUn; gV = gVn;
                deadw = __any_sync(0xffffffffu, dead);
            }
            // the last chunk (only its first row was read, by the prefetch of the last step) goes back to the loader, too
            if (!deadw) {
                __syncwarp();
                mbar_arrive_ifu(sb + L::OFF_EMPTY, fl & F_L0);
                if (++cslot == NCH) { cslot = 0; cpar ^= 1u; }
            }
            sts_i_ifu(a_myprog, BIG, fl & F_L0);   // release an importer still waiting for this warp
            double dacc = ghost ? 0.0 : (double)acc;   // (NaN - NaN in the ghosts)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
            if (lane == 0) p.partial[(size_t)tile * NW + lw] = dacc;
            if (p.trace && lane == 0) {
                if (lw == 0) p.trace[tile * 16 + 5] = gtime();
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct MarchWState {
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
inline void marchw_free(MarchWState& s) {
    cudaFree(s.d_mbu); cudaFree(s.d_mbv); cudaFree(s.d_partial); cudaFree(s.d_order); cudaFree(s.d_trace);
    s = MarchWState{};
}

template <int WU, int WV, int NCH>
inline int marchw_launch(TileState& s, MarchWState& ms, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                        const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change, cudaStream_t st) {
    using L = MarchWLayout<WU, WV, NCH>;
    constexpr int PUT = L::PUT, TW = L::TW, NW = L::NW;
    MarchParams p;
    p.w = w; p.d = d; p.fb = fb;
    p.nV = (d.kpad + TW - 1) / TW;
    p.nU = (w.nu + PUT - 1) / PUT;
    p.ntiles = p.nU * p.nV;
    p.spin_cycles = o.spin_limit << 9;
    static const int pf_env = getenv("TTCR_B200_PF") ? atoi(getenv("TTCR_B200_PF")) : 0;
    p.pf_chunks = pf_env;
    static const int spin_env = getenv("TTCR_B200_SPIN") ? atoi(getenv("TTCR_B200_SPIN")) : 24;
    static const int sleep_env = getenv("TTCR_B200_SLEEP") ? atoi(getenv("TTCR_B200_SLEEP")) : 400;
    p.spin_polls = (unsigned)spin_env; p.sleep_ns = (unsigned)sleep_env;
    p.ctrl = s.d_ctrl;
    const char* trace_path = getenv("TTCR_B200_TRACE");
    if (trace_path && ms.trace_cap < p.ntiles) {
        cudaFree(ms.d_trace);
        ms.d_trace = nullptr;
        TCK(cudaMalloc(&ms.d_trace, (size_t)p.ntiles * 16 * sizeof(long long)));
        ms.trace_cap = p.ntiles;
    }
    p.trace = trace_path ? ms.d_trace : nullptr;
    const int rows = d.nj + TW + PUT + 2 * L::C + 10;
    if (!ms.d_mbu || ms.mb_tiles < p.ntiles || ms.mb_rows != rows || ms.mb_tw != TW || ms.mb_put != PUT) {
        TCK(cudaStreamSynchronize(st));
        cudaFree(ms.d_mbu); cudaFree(ms.d_mbv); cudaFree(ms.d_partial); cudaFree(ms.d_order);
        ms.d_mbu = ms.d_mbv = nullptr; ms.d_partial = nullptr; ms.d_order = nullptr;
        ms.mb_rows = rows; ms.mb_tiles = p.ntiles; ms.mb_tw = TW; ms.mb_put = PUT;
        const size_t nu = (size_t)ms.mb_tiles * ms.mb_rows * TW * 2, nv = (size_t)ms.mb_tiles * ms.mb_rows * PUT * 2;   // two halves per step
        TCK(cudaMalloc(&ms.d_mbu, nu * 8));
        TCK(cudaMalloc(&ms.d_mbv, nv * 8));
        TCK(cudaMalloc(&ms.d_partial, (size_t)p.ntiles * NW * sizeof(double)));
        TCK(cudaMalloc(&ms.d_order, (size_t)p.ntiles * sizeof(int)));
        TCK(cudaMemsetAsync(ms.d_mbu, 0, nu * 8, st));
        TCK(cudaMemsetAsync(ms.d_mbv, 0, nv * 8, st));
        ms.serial = 0;
        ms.order_key = -1;
    }
    MarchMail mail;
    mail.u = ms.d_mbu; mail.v = ms.d_mbv; mail.rows = ms.mb_rows;
    mail.serial = ++ms.serial;
    if (mail.serial == 0) {   // wrapped: clear the tags once every 2^32 sweeps
        TCK(cudaMemsetAsync(ms.d_mbu, 0, (size_t)ms.mb_tiles * ms.mb_rows * TW * 16, st));
        TCK(cudaMemsetAsync(ms.d_mbv, 0, (size_t)ms.mb_tiles * ms.mb_rows * PUT * 16, st));
        mail.serial = ms.serial = 1;
    }
    p.order = ms.d_order; p.partial = ms.d_partial;
    const int key = 8000000 + PUT * 10000 + TW * 40 + w.vlo;
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
        ms.maps.push_back({kk, make_skew_map(a, d, minus, L::BW, L::C, bp, true)});
        return ms.maps.back().second;
    };
    const CUtensorMap tmT = get_map(tt, PUT + 2);
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
        int grid = std::min(p.ntiles, per_sm * sm_count);
        if (o.max_ctas > 0) grid = std::min(grid, o.max_ctas);
        kern<<<grid, L::NT, L::BYTES, st>>>(tmT, tmS, p, mail, tt, frozen, dx);
    };
    switch (variant) {
        case 0: run(k_sweep_march_weno<WU, WV, NCH, false, false, false>); break;
        case 1: run(k_sweep_march_weno<WU, WV, NCH, true, false, false>); break;
        case 2: run(k_sweep_march_weno<WU, WV, NCH, false, true, false>); break;
        case 3: run(k_sweep_march_weno<WU, WV, NCH, true, true, false>); break;
        case 4: run(k_sweep_march_weno<WU, WV, NCH, false, false, true>); break;
        case 5: run(k_sweep_march_weno<WU, WV, NCH, true, false, true>); break;
        case 6: run(k_sweep_march_weno<WU, WV, NCH, false, true, true>); break;
        default: run(k_sweep_march_weno<WU, WV, NCH, true, true, true>); break;
    }
    k_sum_partials<<<1, 256, 0, st>>>(ms.d_partial, p.ntiles * NW, d_change);
    TCK(cudaMemcpyAsync(s.h_abort, s.d_ctrl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    TCK(cudaGetLastError());
    if (trace_path) {
        std::vector<long long> h((size_t)p.ntiles * 16);
        TCK(cudaStreamSynchronize(st));
        TCK(cudaMemcpy(h.data(), ms.d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        FILE* f = fopen(trace_path, "ab");
        if (f) {
            const int hdr[4] = {p.ntiles, p.nU, p.nV, PUT + 1000 * 16};   // (record length in the thousands)
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h.data(), sizeof(long long), h.size(), f);
            fclose(f);
        }
    }
    return 2;
}

template <typename T> inline bool marchw_supported() { return false; }
template <> inline bool marchw_supported<float>() { return true; }

template <typename T>
inline int marchw_sweep(TileState&, MarchWState&, const TileOptions&, int, const SweepView&, const Dims&, T*, const T*, const uint32_t*,
                        const FrozenBox&, T, double*, cudaStream_t) {
    throw std::runtime_error("march kernel: fp32 only");
}
#if defined(TTCR_B200_SPLIT_BUILD) && !defined(TTCR_B200_MARCHW_DEFINE)
template <>
int marchw_sweep<float>(TileState& s, MarchWState& ms, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                               const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change, cudaStream_t st);   // defined in marchw_inst.cu
#else
#ifdef TTCR_B200_MARCHW_DEFINE
#define TTCR_B200_MARCHW_DEFINE_LINKAGE
#else
#define TTCR_B200_MARCHW_DEFINE_LINKAGE inline
#endif
template <>
TTCR_B200_MARCHW_DEFINE_LINKAGE int marchw_sweep<float>(TileState& s, MarchWState& ms, const TileOptions& o, int sm_count, const SweepView& w, const Dims& d, float* tt,
                               const float* slo, const uint32_t* frozen, const FrozenBox& fb, float dx, double* d_change, cudaStream_t st) {
    return marchw_launch<4, 2, 4>(s, ms, o, sm_count, w, d, tt, slo, frozen, fb, dx, d_change, st);
}
#endif

}  // namespace ttcrb200
