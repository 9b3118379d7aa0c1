// TEST INFRASTRUCTURE: the reference-side adapter (include/Grid3Drfs_B200.h), linked and EXECUTED.
//
// Builds the reference's own Grid3Drnfs / Grid3Drcfs and the adapter classes behind `Grid3D<T,uint32_t>*`, exactly as
// src/ttcrpy/rgrid.pyx:246-253 / ttcr/grids.h:575-599 would, and drives both through the multi-source overload
//     Grid3D::raytrace(vector<vector<sxyz>> Tx, vector<vector<T>> t0, vector<vector<sxyz>> Rx, vector<vector<T>>& tt)
// (ttcr/Grid3D.h:810-853) with nThreads = 2 and usePool off, so that the reference's own std::thread fan-out calls
// raytrace(..., threadNo = 0 / 1) of the adapter concurrently: two slots of one ttcr_b200 grid.  Receiver traveltimes and the
// fields of both thread slots must be bit-identical in double (node and cell slowness, first order and WENO), and within
// 1e-4 in float; the matrices M of the two m_data overloads bit-identical in double.  Compiled from /root/reference's headers by `make -C oracle adapter-run` into oracle/_ref/adapter_run (which
// travels to the GPU box); tests/test_gpu_parity.py::test_cxx_adapter_linked_and_run executes it.  No reference source is copied.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <memory>
#include <vector>

#include "Grid3Drcfs.h"
#include "Grid3Drnfs.h"
#include "Grid3Drfs_B200.h"

namespace ttcr {
int verbose = 0;       // ttcr/ttcr_t.h:35 (extern)
int gpu_profile = 0;   // ttcr/ttcr_t.h:40 (extern)
}

template <typename T>
static double max_rel(const std::vector<T>& a, const std::vector<T>& b, double floor) {
    double m = 0.0;
    for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::fabs((double)a[i] - (double)b[i]) / std::max((double)b[i], floor));
    return m;
}

template <typename T, bool CELL>
static int run(bool weno, const char* name) {
    const uint32_t ncx = 30, ncy = 26, ncz = 33;
    const T dx = T(0.25);
    const size_t nt = 2;
    using Ref = typename std::conditional<CELL, ttcr::Grid3Drcfs<T, uint32_t>, ttcr::Grid3Drnfs<T, uint32_t>>::type;
    std::unique_ptr<ttcr::Grid3D<T, uint32_t>> ref(new Ref(ncx, ncy, ncz, dx, T(0), T(0), T(0), T(1e-15), 50, weno, false, false, nt, false));
    std::unique_ptr<ttcr::Grid3D<T, uint32_t>> gpu(
        new ttcr::Grid3Drfs_B200<T, uint32_t, CELL>(ncx, ncy, ncz, dx, T(0), T(0), T(0), T(1e-15), 50, weno, false, false, nt, false));
    ref->setUsePool(false);   // the std::thread branch of Grid3D.h:833-852 passes threadNo = block index
    gpu->setUsePool(false);
    const size_t ns = CELL ? (size_t)ncx * ncy * ncz : (size_t)(ncx + 1) * (ncy + 1) * (ncz + 1);
    std::vector<T> s(ns);
    uint64_t st = 88172645463325252ull;
    for (auto& v : s) {   // xorshift: slowness in [0.3, 1.0)
        st ^= st << 13; st ^= st >> 7; st ^= st << 17;
        v = T(0.3 + 0.7 * (double)(st >> 11) / 9007199254740992.0);
    }
    ref->setSlowness(s);
    gpu->setSlowness(s);
    const double src[4][3] = {{1.0, 1.25, 2.0}, {6.3, 2.2, 7.1}, {0.0, 0.0, 0.0}, {7.5, 6.5, 8.25}};
    const double rcv[3][3] = {{7.0, 6.0, 8.0}, {0.1, 0.2, 0.3}, {3.75, 3.3, 4.9}};
    std::vector<std::vector<ttcr::sxyz<T>>> Tx(4), Rx(4);
    std::vector<std::vector<T>> t0(4), tt_ref(4), tt_gpu(4);
    for (int n = 0; n < 4; ++n) {
        Tx[n].push_back(ttcr::sxyz<T>(T(src[n][0]), T(src[n][1]), T(src[n][2])));
        t0[n].push_back(T(0.01 * n));
        for (auto& r : rcv) Rx[n].push_back(ttcr::sxyz<T>(T(r[0]), T(r[1]), T(r[2])));
        tt_ref[n].resize(3);
        tt_gpu[n].resize(3);
    }
    ref->raytrace(Tx, t0, Rx, tt_ref);
    gpu->raytrace(Tx, t0, Rx, tt_gpu);
    const bool exact = std::is_same<T, double>::value;
    int bad = 0;
    double worst = 0.0;
    for (int n = 0; n < 4; ++n) {
        if (exact) bad += std::memcmp(tt_ref[n].data(), tt_gpu[n].data(), 3 * sizeof(T)) != 0;
        worst = std::max(worst, max_rel(tt_gpu[n], tt_ref[n], (double)dx * 0.3));
    }
    for (size_t th = 0; th < nt; ++th) {   // slot th holds the field of the last source of block th (sources 1 and 3)
        std::vector<T> fr, fg;
        ref->getTT(fr, th);
        gpu->getTT(fg, th);
        if (fr.size() != fg.size()) { ++bad; continue; }
        if (exact) bad += std::memcmp(fr.data(), fg.data(), fr.size() * sizeof(T)) != 0;
        worst = std::max(worst, max_rel(fg, fr, (double)dx * 0.3));
    }
    std::vector<T> sr, sg;
    ref->getSlowness(sr);
    gpu->getSlowness(sg);
    bad += sr.size() != sg.size() || std::memcmp(sr.data(), sg.data(), sr.size() * sizeof(T)) != 0;   // cell -> node averaging: bit-exact
    if (exact && !weno) {   // Grid3Drn::saveTT formats 1 and 3: byte-identical files
        for (int fmt : {1, 3}) {
            ref->saveTT("/tmp/ttcr_adapter_ref", 0, 1, fmt);
            gpu->saveTT("/tmp/ttcr_adapter_gpu", 0, 1, fmt);
            const char* ext = fmt == 1 ? ".dat" : ".bin";
            std::ifstream a(std::string("/tmp/ttcr_adapter_ref") + ext, std::ios::binary), b(std::string("/tmp/ttcr_adapter_gpu") + ext, std::ios::binary);
            const std::string sa((std::istreambuf_iterator<char>(a)), std::istreambuf_iterator<char>()), sb((std::istreambuf_iterator<char>(b)), std::istreambuf_iterator<char>());
            if (sa.empty() || sa != sb) { ++bad; std::printf("  saveTT format %d differs (%zu vs %zu bytes)\n", fmt, sa.size(), sb.size()); }
        }
    }
    if (!exact && worst > (weno ? 2e-3 : 1e-4)) ++bad;
    std::printf("%-28s %s  max rel diff %.3g\n", name, bad ? "MISMATCH" : "ok", worst);
    return bad;
}

// The two M overloads (Grid3D.h:143-155; node slowness only): columns, values and their order, bit for bit in double.  A smooth
// model: the reference's raypath walk has no guard against cycling, and on the random model of run() it does not terminate.
static int run_m() {
    using T = double;
    const uint32_t nc = 24;
    const T dx = T(0.5);
    std::unique_ptr<ttcr::Grid3D<T, uint32_t>> ref(new ttcr::Grid3Drnfs<T, uint32_t>(nc, nc, nc, dx, T(0), T(0), T(0), T(1e-15), 20, true, true, false, 1, false));
    std::unique_ptr<ttcr::Grid3D<T, uint32_t>> gpu(
        new ttcr::Grid3Drfs_B200<T, uint32_t, false>(nc, nc, nc, dx, T(0), T(0), T(0), T(1e-15), 20, true, true, false, 1, false));
    std::vector<T> s((size_t)(nc + 1) * (nc + 1) * (nc + 1));
    for (uint32_t k = 0, n = 0; k <= nc; ++k)
        for (uint32_t j = 0; j <= nc; ++j)
            for (uint32_t i = 0; i <= nc; ++i, ++n)
                s[n] = (1.0 + 0.3 * std::sin(0.7 * i * dx) * std::cos(0.9 * j * dx)) / (1.0 + 0.1 * k * dx);
    ref->setSlowness(s);
    gpu->setSlowness(s);
    std::vector<ttcr::sxyz<T>> Tx{ttcr::sxyz<T>(3.3, 7.1, 8.9)}, Rx{ttcr::sxyz<T>(10.2, 1.7, 2.9), ttcr::sxyz<T>(2.5, 3.5, 1.0), ttcr::sxyz<T>(6.0, 6.0, 11.5)};
    std::vector<T> t0{0.25};
    int bad = 0;
    for (int with_rays = 0; with_rays < 2; ++with_rays) {
        std::vector<T> ta, tb;
        std::vector<std::vector<ttcr::sxyz<T>>> ra, rb;
        std::vector<std::vector<ttcr::sijv<T>>> ma, mb;
        if (with_rays) { ref->raytrace(Tx, t0, Rx, ta, ra, ma, 0); gpu->raytrace(Tx, t0, Rx, tb, rb, mb, 0); }
        else { ref->raytrace(Tx, t0, Rx, ta, ma, 0); gpu->raytrace(Tx, t0, Rx, tb, mb, 0); }
        int mbad = ma.size() != mb.size() || ta.size() != tb.size() || std::memcmp(ta.data(), tb.data(), ta.size() * sizeof(T)) != 0;
        size_t nnz = 0;
        for (size_t n = 0; !mbad && n < ma.size(); ++n) {
            mbad |= ma[n].size() != mb[n].size();
            for (size_t k = 0; !mbad && k < ma[n].size(); ++k)
                mbad |= ma[n][k].i != mb[n][k].i || ma[n][k].j != mb[n][k].j || std::memcmp(&ma[n][k].v, &mb[n][k].v, sizeof(T)) != 0;
            nnz += ma[n].size();
        }
        if (mbad || nnz == 0) ++bad;
        std::printf("%-28s %s  %zu entries\n", with_rays ? "M (r_data, m_data) <double>" : "M (m_data) <double>", (mbad || nnz == 0) ? "MISMATCH" : "ok", nnz);
    }
    return bad;
}

int main() {
    int bad = 0;
    try {
        bad += run<double, false>(false, "Grid3Drnfs<double> fo");
        bad += run<double, false>(true, "Grid3Drnfs<double> weno");
        bad += run<double, true>(false, "Grid3Drcfs<double> fo");
        bad += run<double, true>(true, "Grid3Drcfs<double> weno");
        bad += run<float, false>(false, "Grid3Drnfs<float> fo");
        bad += run<float, true>(false, "Grid3Drcfs<float> fo");
        bad += run_m();
    } catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 2;
    }
    std::printf(bad ? "adapter-run: FAILED\n" : "adapter-run: OK\n");
    return bad ? 1 : 0;
}
