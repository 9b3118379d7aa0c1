// Grid3Drfs_B200.h -- header-only adapter: a ttcr::Grid3D<T1,T2> subclass that forwards the FSM path
// to libttcr_b200.so through the C ABI of ttcr_b200.h.
//
// This is the reference-side half of the drop-in boundary.  The reference holds every 3-D grid as a
// `Grid3D<T,uint32_t>*` (src/ttcrpy/rgrid.pyx:153, ttcr/grids.h:549-599) and only ever calls the
// virtuals of ttcr/Grid3D.h:44-466, so a class with the constructor signature of Grid3Drnfs /
// Grid3Drcfs (ttcr/Grid3Drnfs.h:39-50, ttcr/Grid3Drcfs.h:40-52) can be selected exactly where
// Grid3Drnfs_OpenCL is selected today (rgrid.pyx:246-253, rgrid.pxd:122-125, grids.h:575-599).
//
// It is compiled only inside a ttcr build tree (it includes the reference's own "Grid3D.h"); nothing
// in this repository's product path includes it.  `make -C oracle adapter-check` syntax-checks it
// against /root/reference when that tree is present.
//
// Exceptions: status codes are mapped back onto the exception types the reference throws, so
// Cython's `except +` translation (rgrid.pxd:35-43) is unchanged.
#ifndef TTCR_GRID3DRFS_B200_H
#define TTCR_GRID3DRFS_B200_H

#include <fstream>
#include <iostream>   // ttcr/Grid3D.h uses std::cout without including it
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "Grid3D.h"      // the reference's abstract base (ttcr/Grid3D.h)
#include "ttcr_b200.h"

namespace ttcr {

template <typename T1, typename T2, bool CELL_SLOWNESS>
class Grid3Drfs_B200 : public Grid3D<T1, T2> {
    static_assert(std::is_same<T1, double>::value || std::is_same<T1, float>::value, "T1 must be float or double");

public:
    // same argument list as Grid3Drnfs / Grid3Drcfs (nx, ny, nz are CELL counts)
    Grid3Drfs_B200(const T2 nx, const T2 ny, const T2 nz, const T1 ddx, const T1 minx, const T1 miny, const T1 minz,
                   const T1 eps, const int maxit, const bool w, const bool ttrp = true, const bool intVel = false,
                   const size_t nt = 1, const bool _translateOrigin = false, const int device = -1)
        : Grid3D<T1, T2>(ttrp, static_cast<size_t>(nx) * ny * nz, nt, _translateOrigin), h(nullptr),
          nnodes(static_cast<size_t>(nx + 1) * (ny + 1) * (nz + 1)), ncx_(nx), ncy_(ny), ncz_(nz), dx_(ddx), xmin_(minx), ymin_(miny), zmin_(minz) {
        check(ttcr_b200_create(&h, nx, ny, nz, ddx, minx, miny, minz, eps, maxit, w, ttrp, intVel, nt, _translateOrigin,
                               CELL_SLOWNESS, std::is_same<T1, double>::value ? TTCR_B200_F64 : TTCR_B200_F32, device));
    }
    ~Grid3Drfs_B200() { ttcr_b200_destroy(h); }

    // Grid3Drn::setSlowness / Grid3Drcfs::setSlowness (Grid3Drn.h:82-89, Grid3Drcfs.h:88-171)
    void setSlowness(const std::vector<T1>& s) override {
        check(ttcr_b200_set_slowness(h, s.data(), s.size(), TTCR_B200_ORDER_X_FASTEST));
    }
    // Grid3Drn::getSlowness (Grid3Drn.h:90-97)
    void getSlowness(std::vector<T1>& s) const override {
        s.resize(nnodes);
        check(ttcr_b200_get_slowness(h, s.data(), TTCR_B200_ORDER_X_FASTEST));
    }
    // Grid3Drn::getTT (Grid3Drn.h:102-108)
    void getTT(std::vector<T1>& tt, const size_t threadNo = 0) const override {
        tt.resize(nnodes);
        check(ttcr_b200_get_tt(h, tt.data(), threadNo, TTCR_B200_ORDER_X_FASTEST));
    }
    size_t getNumberOfNodes() const override { return nnodes; }

    // Grid3D::raytrace(Tx, t0, Rx, traveltimes, threadNo) (Grid3D.h:115-119, :470-502)
    void raytrace(const std::vector<sxyz<T1>>& Tx, const std::vector<T1>& t0, const std::vector<sxyz<T1>>& Rx,
                  std::vector<T1>& traveltimes, const size_t threadNo = 0) const override {
        std::vector<T1> tx = flatten(Tx), rx = flatten(Rx);
        traveltimes.resize(Rx.size());
        check(ttcr_b200_raytrace(h, tx.data(), t0.data(), Tx.size(), rx.data(), Rx.size(), traveltimes.data(), threadNo));
    }

    // Grid3D::raytrace(Tx, t0, Rx, traveltimes, r_data, threadNo) (Grid3D.h:127-132, :545-586): traveltimes and raypaths
    void raytrace(const std::vector<sxyz<T1>>& Tx, const std::vector<T1>& t0, const std::vector<sxyz<T1>>& Rx,
                  std::vector<T1>& traveltimes, std::vector<std::vector<sxyz<T1>>>& r_data, const size_t threadNo = 0) const override {
        std::vector<T1> tx = flatten(Tx), rx = flatten(Rx);
        traveltimes.resize(Rx.size());
        std::vector<size_t> npts(Rx.size(), 0);
        check(ttcr_b200_raytrace_rays(h, tx.data(), t0.data(), Tx.size(), rx.data(), Rx.size(), traveltimes.data(), npts.data(),
                                      threadNo));
        size_t total = 0;
        for (size_t n : npts) total += n;
        std::vector<T1> xyz(3 * total);
        check(ttcr_b200_get_rays(h, threadNo, xyz.data()));
        r_data.resize(Rx.size());
        size_t o = 0;
        for (size_t n = 0; n < Rx.size(); ++n) {
            r_data[n].resize(npts[n]);
            for (size_t k = 0; k < npts[n]; ++k, o += 3) r_data[n][k] = sxyz<T1>(xyz[o], xyz[o + 1], xyz[o + 2]);
        }
    }

    // Grid3D::raytrace(Tx, t0, Rx, traveltimes, m_data, threadNo) (Grid3D.h:150-155, :743-780) and
    // (.., r_data, m_data, threadNo) (:143-148, :646-690): the sensitivity matrix M of Grid3Drn::getRaypath.  The library walks
    // the raw terms with the raypaths (the reference's two overloads differ, raypath.cuh: m_terms = 1 / 2); here they are
    // merged like the reference merges them (m_data[nm].v += m.v in order of appearance, Grid3Drn.h:1612-1623).
    void raytrace(const std::vector<sxyz<T1>>& Tx, const std::vector<T1>& t0, const std::vector<sxyz<T1>>& Rx,
                  std::vector<T1>& traveltimes, std::vector<std::vector<sijv<T1>>>& m_data, const size_t threadNo = 0) const override {
        std::vector<std::vector<sxyz<T1>>> r_data;
        raytrace_m(Tx, t0, Rx, traveltimes, r_data, m_data, threadNo, 1);
    }
    void raytrace(const std::vector<sxyz<T1>>& Tx, const std::vector<T1>& t0, const std::vector<sxyz<T1>>& Rx,
                  std::vector<T1>& traveltimes, std::vector<std::vector<sxyz<T1>>>& r_data,
                  std::vector<std::vector<sijv<T1>>>& m_data, const size_t threadNo = 0) const override {
        raytrace_m(Tx, t0, Rx, traveltimes, r_data, m_data, threadNo, 2);
    }
    void raytrace_m(const std::vector<sxyz<T1>>& Tx, const std::vector<T1>& t0, const std::vector<sxyz<T1>>& Rx,
                    std::vector<T1>& traveltimes, std::vector<std::vector<sxyz<T1>>>& r_data,
                    std::vector<std::vector<sijv<T1>>>& m_data, const size_t threadNo, const int mode) const {
        // (the option is per grid: like the reference's own M overloads, not meant to run concurrently with other calls)
        check(ttcr_b200_set_option(h, "m_terms", mode));
        try {
            raytrace(Tx, t0, Rx, traveltimes, r_data, threadNo);
        } catch (...) {
            ttcr_b200_set_option(h, "m_terms", 0);
            throw;
        }
        check(ttcr_b200_set_option(h, "m_terms", 0));
        size_t total = 0;
        for (const auto& r : r_data) total += r.size();
        std::vector<unsigned long long> node(8 * total);
        std::vector<T1> val(8 * total);
        check(ttcr_b200_get_m_terms(h, threadNo, node.data(), val.data()));
        m_data.assign(Rx.size(), std::vector<sijv<T1>>());
        size_t o = 0;
        for (size_t n = 0; n < Rx.size(); ++n) {
            for (size_t k = 8; k < 8 * r_data[n].size(); ++k) {   // (the receiver, point 0, closes no segment)
                const size_t j = node[o + k];
                bool found = false;
                for (auto& m : m_data[n])
                    if (m.j == j) { m.v += val[o + k]; found = true; break; }
                if (!found) m_data[n].push_back(sijv<T1>(n, j, val[o + k]));
            }
            o += 8 * r_data[n].size();
        }
    }

    // Grid3D::raytrace(vector<vector<sxyz>>&, ...) (Grid3D.h:172-175, :810-853) is NOT virtual: called
    // through a Grid3D*, the reference's own fan-out (ctpl pool / std::thread blocks) runs and calls the
    // single-source override above concurrently with distinct threadNo, which the library supports (one
    // slot = one field + one CUDA stream).  Called on the concrete type, this overload lets the library
    // deal the sources to its slots itself.
    void raytrace(const std::vector<std::vector<sxyz<T1>>>& Tx, const std::vector<std::vector<T1>>& t0,
                  const std::vector<std::vector<sxyz<T1>>>& Rx, std::vector<std::vector<T1>>& traveltimes) const {
        const size_t ns = Tx.size();
        std::vector<size_t> txo(ns + 1, 0), rxo(ns + 1, 0);
        std::vector<T1> tx, rx, tz;
        for (size_t n = 0; n < ns; ++n) {
            txo[n + 1] = txo[n] + Tx[n].size();
            rxo[n + 1] = rxo[n] + Rx[n].size();
            const std::vector<T1> a = flatten(Tx[n]), b = flatten(Rx[n]);
            tx.insert(tx.end(), a.begin(), a.end());
            rx.insert(rx.end(), b.begin(), b.end());
            tz.insert(tz.end(), t0[n].begin(), t0[n].end());
        }
        std::vector<T1> out(rxo[ns]);
        check(ttcr_b200_raytrace_multi(h, ns, txo.data(), tx.data(), tz.data(), rxo.data(), rx.data(), out.data(), nullptr,
                                       nullptr));
        traveltimes.resize(ns);
        for (size_t n = 0; n < ns; ++n) traveltimes[n].assign(out.begin() + rxo[n], out.begin() + rxo[n + 1]);
    }

    // Grid3Drnfs::get_niter / get_niterw (Grid3Drnfs.h:56-57; virtual in Grid3D.h:284-285); slot 0, plus
    // per-slot variants (the reference's single shared counter is racy with nt > 1)
    const int get_niter() const override { return get_niter_slot(0); }
    const int get_niterw() const override { return get_niterw_slot(0); }
    int get_niter_slot(const size_t threadNo) const { int a = 0, b = 0; check(ttcr_b200_get_niter(h, threadNo, &a, &b)); return a; }
    int get_niterw_slot(const size_t threadNo) const { int a = 0, b = 0; check(ttcr_b200_get_niter(h, threadNo, &a, &b)); return b; }

    void setTraveltimeFromRaypath(const bool ttrp) {
        Grid3D<T1, T2>::setTraveltimeFromRaypath(ttrp);
        check(ttcr_b200_set_option(h, "tt_from_rp", ttrp ? 1.0 : 0.0));
    }

    // Grid3Drn::saveTT (Grid3Drn.h:2678-2760): format 1 = text "x\ty\tz\ttt" (precision 12), format 3 = binary records of four
    // T1; every node of this grid is a primary node, so `all` makes no difference.  Format 2 (VTK) is written by the Python
    // surface (ttcr_b200/vtr.py); here it reports what the reference reports without VTK.
    void saveTT(const std::string& fname, const int, const size_t nt = 0, const int format = 1) const override {
        if (format == 2) { std::cerr << "VTK not included during compilation.\nNothing saved.\n"; return; }
        if (format != 1 && format != 3) throw std::runtime_error("Unsupported format for saving traveltimes");
        std::vector<T1> tt;
        getTT(tt, nt);
        std::ofstream fout;
        if (format == 1) { fout.open((fname + ".dat").c_str()); fout.precision(12); }
        else fout.open((fname + ".bin").c_str(), std::ios::out | std::ios::binary | std::ios::trunc);
        size_t n = 0;
        for (T2 k = 0; k <= ncz_; ++k)
            for (T2 j = 0; j <= ncy_; ++j)
                for (T2 i = 0; i <= ncx_; ++i, ++n) {   // node order and coordinates of Grid3Drn::buildGridNodes (Grid3Drn.h:2823)
                    const T1 rec[4] = {xmin_ + i * dx_, ymin_ + j * dx_, zmin_ + k * dx_, tt[n]};
                    if (format == 1) fout << rec[0] << '\t' << rec[1] << '\t' << rec[2] << '\t' << rec[3] << '\n';
                    else fout.write((const char*)rec, 4 * sizeof(T1));
                }
        fout.close();
    }

private:
    ttcr_b200_grid* h;
    size_t nnodes;
    T2 ncx_, ncy_, ncz_;
    T1 dx_, xmin_, ymin_, zmin_;

    static std::vector<T1> flatten(const std::vector<sxyz<T1>>& p) {
        std::vector<T1> v(3 * p.size());
        for (size_t n = 0; n < p.size(); ++n) { v[3 * n] = p[n].x; v[3 * n + 1] = p[n].y; v[3 * n + 2] = p[n].z; }
        return v;
    }
    void check(int rc) const {
        if (rc == TTCR_B200_OK) return;
        const std::string msg = ttcr_b200_last_error(h);
        switch (rc) {
            case TTCR_B200_ERR_LENGTH: throw std::length_error(msg);      // Grid3Drn.h:84
            case TTCR_B200_ERR_LOGIC: throw std::logic_error(msg);        // Grid3Drnfs.h:112
            case TTCR_B200_ERR_INVALID: throw std::invalid_argument(msg);
            default: throw std::runtime_error(msg);                       // Grid3Drn.h:785 and CUDA failures
        }
    }
};

template <typename T1, typename T2> using Grid3Drnfs_B200 = Grid3Drfs_B200<T1, T2, false>;   // node slowness
template <typename T1, typename T2> using Grid3Drcfs_B200 = Grid3Drfs_B200<T1, T2, true>;    // cell slowness

}  // namespace ttcr

#endif
