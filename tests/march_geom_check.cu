// Host-only check of the index arithmetic of k_sweep_march (ttcr_b200/csrc/sweep_march.cuh): the TMA boxes of every
// chunk of every tile are emulated from the tensor-map definition (skewed plane stride, out-of-bounds fill) and every
// shared-memory read of every thread and step is compared with the node the algorithm wants there.
// Built and run by tests/test_march_geometry.py (no GPU needed: nothing is launched).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ttcr_b200/csrc/sweep_march.cuh"

using namespace ttcrb200;

static long long fails = 0, checks = 0;

// element offset (relative to the array proper) a tensor-map coordinate resolves to; -1 = out of bounds (zero fill)
static long long tma_elem(const Dims& d, bool minus, int x, int y, int z, bool& oob_alloc) {
    if (x < 0 || x >= d.kpad || z < 0 || z >= d.ni || y < 0 || y >= d.qs + 2 * d.ni + 64) return -1;
    const long long row = minus ? (long long)z * (d.qs - 1) + y : (long long)z * (d.qs + 1) + y - d.ni;
    const long long e = row * d.kpad + x;
    if (e < -(long long)d.ni * d.kpad || e >= (long long)d.elems()) oob_alloc = true;
    return e;
}

// element offset of oriented node (u, m, v), -1 if the TMA unit must zero-fill it
static long long node_elem(const SweepView& w, const Dims& d, int u, int m, int v) {
    if (u < 0 || u >= w.nu || v < 0 || v >= d.kpad) return -1;
    return w.base + (long long)u * w.su + (long long)m * w.sm + (long long)v * w.sv;
}

template <int WU, int WV, int NCH, int D, bool RI, bool RJ, bool RK>
static void check_dir(const Dims& d, const SweepView& w) {
    using L = MarchLayout<WU, WV, NCH, D>;
    using G = MarchGeom<WU, WV, NCH, D, RI, RJ, RK>;
    constexpr int PUT = L::PUT, TW = L::TW, BW = L::BW;
    const int nV = (d.kpad + TW - 1) / TW, nU = (w.nu + PUT - 1) / PUT;
    const bool minus = (w.ri != 0) == (w.rj != 0);
    std::vector<long long> slot(L::CHB / 4);
    for (int tile = 0; tile < nU * nV; ++tile) {
        const MarchTile T = march_tile<PUT, TW>(w, nU, nV, tile);
        if (T.nrows <= 0 || T.vb <= T.va) { ++fails; printf("empty tile %d\n", tile); }
        for (int c = 0; c < T.nch; ++c) {
            std::fill(slot.begin(), slot.end(), -7);
            bool oob_alloc = false;
            for (int s = 0; s < 2; ++s) {
                const int x0 = G::box_x(d, T), y0 = G::box_y(w, d, T, c, s), z0 = G::box_z(d, T, s);
                const int np = s ? PUT : PUT + 1;
                const int base = s ? L::CHB_T / 4 : 0;
                for (int zi = 0; zi < np; ++zi)
                    for (int yi = 0; yi < 2; ++yi)
                        for (int xi = 0; xi < BW; ++xi)
                            slot[base + (zi * 2 + yi) * BW + xi] = tma_elem(d, minus, x0 + xi, y0 + yi, z0 + zi, oob_alloc);
            }
            if (oob_alloc) { ++fails; printf("box outside the allocation: tile %d chunk %d\n", tile, c); }
            for (int pl = 0; pl < PUT; ++pl)
                for (int vl = 0; vl < TW; ++vl)
                    for (int b = 2 * c; b < 2 * c + 2; ++b) {
                        const int off = G::thread_off(pl, vl) + ((b & 1) ? G::DR : 0);
                        const int u = T.u0 + pl, v = T.v0 + vl;
                        const int mt = T.m_first + b - pl;   // row of the "next old" traveltimes of step b
                        const long long want[4] = {node_elem(w, d, u, mt, v), node_elem(w, d, u, mt, v + 1), node_elem(w, d, u + 1, mt - 1, v),
                                                   node_elem(w, d, u, mt - 1, v)};
                        const int offs[4] = {off, off + G::DH, off + G::DUP, off + G::DS};
                        for (int q = 0; q < 4; ++q) {
                            ++checks;
                            if (offs[q] < 0 || offs[q] / 4 >= (int)slot.size() || (offs[q] & 3)) { ++fails; continue; }
                            const long long got = slot[offs[q] / 4];
                            // lanes past the row end: the halo lane v+1 of the tile's last lane may be a real slot the box does not
                            // need to hold a zero for (want == -1 requires a zero only where the update could otherwise store)
                            if (got != want[q]) {
                                if (++fails < 20)
                                    printf("dir ri%d rj%d rk%d tile %d chunk %d pl %d vl %d step %d operand %d: got %lld want %lld\n", RI, RJ, RK, tile, c,
                                           pl, vl, b, q, got, want[q]);
                            }
                        }
                        // global address of the node updated at step b: (u, m_first - 1 + b - pl, v)
                        if (u < w.nu && v < d.kpad) {
                            const long long e0 = w.base + (long long)u * w.su + (long long)(T.m_first - 1 - pl) * w.sm + (long long)v * w.sv;
                            ++checks;
                            if (e0 + (long long)b * w.sm != node_elem(w, d, u, mt - 1, v)) ++fails;
                        }
                    }
        }
        // every node of the tile is visited exactly once: rows 0 .. nrows-1 of lanes va .. vb-1 cover j = 0 .. nj-1
        for (int v = T.va; v < T.vb; ++v) {
            const int r0 = (0 - w.joff + v) - T.m_first, r1 = (w.nj - 1 - w.joff + v) - T.m_first;   // local rows of j = 0 and j = nj-1
            ++checks;
            if (r0 < 0 || r1 >= T.nrows) { ++fails; printf("rows of lane %d outside the tile's march\n", v); }
            // the last step of the last plane must be executed: step = r1 + 1 + (PUT-1) < nA
            if (r1 + PUT >= T.nA) { ++fails; printf("march too short\n"); }
        }
    }
}

template <int WU, int WV, int NCH, int D>
static void check_cfg(int ni, int nj, int nk) {
    const Dims d = make_dims(ni, nj, nk);
    for (int dir = 0; dir < 8; ++dir) {
        const SweepView w = make_view(d, dir);
        switch (dir) {
            case 0: check_dir<WU, WV, NCH, D, false, false, false>(d, w); break;
            case 1: check_dir<WU, WV, NCH, D, true, false, false>(d, w); break;
            case 2: check_dir<WU, WV, NCH, D, false, true, false>(d, w); break;
            case 3: check_dir<WU, WV, NCH, D, true, true, false>(d, w); break;
            case 4: check_dir<WU, WV, NCH, D, false, false, true>(d, w); break;
            case 5: check_dir<WU, WV, NCH, D, true, false, true>(d, w); break;
            case 6: check_dir<WU, WV, NCH, D, false, true, true>(d, w); break;
            default: check_dir<WU, WV, NCH, D, true, true, true>(d, w); break;
        }
    }
}

int main() {
    const int shapes[][3] = {{20, 13, 17}, {33, 40, 70}, {16, 32, 32}, {41, 41, 41}, {5, 3, 2}, {17, 9, 97}, {64, 64, 64}};
    for (auto& s : shapes) {
        check_cfg<4, 4, 6, 16>(s[0], s[1], s[2]);
        check_cfg<4, 4, 7, 16>(s[0], s[1], s[2]);
    }
    printf("march geometry: %lld checks, %lld failures\n", checks, fails);
    return fails ? 1 : 0;
}
