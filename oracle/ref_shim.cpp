// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// C-ABI shim around the UNMODIFIED reference headers (ttcr/Grid3Drnfs.h,
// ttcr/Grid3Drcfs.h under /root/reference).  It is compiled where the
// reference lies (see oracle/Makefile, target `ref`) into oracle/_ref/ and is
// used (a) to pin the plain-C restatement in fsm_oracle.c bit-for-bit,
// (b) to generate tests/golden/*.npz, and (c) as the CPU baseline
// (`bench.py --impl reference`, cpu_baseline.kind == "reference").
//
// No reference source is copied: this file only #includes the headers and
// forwards calls.  The public raytrace() overloads live on ttcr::Grid3D
// (ttcr/Grid3D.h:115-119, :172-175); the FSM classes re-declare private
// overloads that hide them, hence the calls through a Grid3D<T,uint32_t>&.

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "Grid3Drnfs.h"
#include "Grid3Drcfs.h"
#include "Grid2Drcfs.h"
#include "Grid2Drnfs.h"   // the 2-D twins: pins oracle/fsm2d_oracle.c

namespace ttcr {
int verbose = 0;       // ttcr/ttcr_t.h:35 (extern)
int gpu_profile = 0;   // ttcr/ttcr_t.h:40 (extern)
}

namespace {

thread_local std::string g_err;

struct Base {
    virtual ~Base() {}
    virtual void set_slowness(const double* s, size_t n) = 0;
    virtual void get_slowness(double* out) = 0;
    virtual void raytrace(const double* tx, const double* t0, size_t ntx,
                          const double* rx, size_t nrx, double* tt, size_t th) = 0;
    virtual void raytrace_multi(size_t nsrc, const double* tx, const double* t0,
                                const double* rx, size_t nrx, double* tt) = 0;
    virtual void raytrace_rays(const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx, double* tt,
                               size_t th, size_t* npts, double* xyz, size_t cap) = 0;
    virtual void raytrace_m(const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx, double* tt, size_t th,
                            size_t* nnz, unsigned long long* col, double* val, size_t cap, int with_rays) = 0;
    virtual void get_tt(double* out, size_t th) = 0;
    virtual void niter(int* a, int* b) = 0;
    virtual size_t nnodes() = 0;
};

template <typename T, typename G>
struct Impl : Base {
    std::unique_ptr<G> g;
    Impl(uint32_t nx, uint32_t ny, uint32_t nz, double dx, double x0, double y0, double z0,
         double eps, int maxit, int weno, int ttrp, int nt, int translate)
        : g(new G(nx, ny, nz, T(dx), T(x0), T(y0), T(z0), T(eps), maxit, weno != 0,
                  (ttrp & 1) != 0, (ttrp & 2) != 0, size_t(nt), translate != 0)) {}   // ttrp bit 1: intVel (processVel)
    ttcr::Grid3D<T, uint32_t>& base() { return *g; }
    void set_slowness(const double* s, size_t n) override {
        std::vector<T> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = T(s[i]);
        g->setSlowness(v);
    }
    void get_slowness(double* out) override {
        std::vector<T> v;
        g->getSlowness(v);
        for (size_t i = 0; i < v.size(); ++i) out[i] = double(v[i]);
    }
    static void pts(const double* p, size_t n, std::vector<ttcr::sxyz<T>>& v) {
        v.resize(n);
        for (size_t i = 0; i < n; ++i) v[i] = ttcr::sxyz<T>(T(p[3 * i]), T(p[3 * i + 1]), T(p[3 * i + 2]));
    }
    void raytrace(const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx,
                  double* tt, size_t th) override {
        std::vector<ttcr::sxyz<T>> Tx, Rx;
        pts(tx, ntx, Tx);
        pts(rx, nrx, Rx);
        std::vector<T> vt0(ntx), vtt(nrx);
        for (size_t i = 0; i < ntx; ++i) vt0[i] = T(t0[i]);
        base().raytrace(Tx, vt0, Rx, vtt, th);
        for (size_t i = 0; i < nrx; ++i) tt[i] = double(vtt[i]);
    }
    // Grid3D::raytrace(Tx,t0,Rx,traveltimes,r_data,threadNo) (Grid3D.h:545-586): ray n gets npts[n] points, the first
    // `cap` of them stored at xyz + 3 * cap * n
    void raytrace_rays(const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx, double* tt, size_t th,
                       size_t* npts, double* xyz, size_t cap) override {
        std::vector<ttcr::sxyz<T>> Tx, Rx;
        pts(tx, ntx, Tx);
        pts(rx, nrx, Rx);
        std::vector<T> vt0(ntx), vtt(nrx);
        for (size_t i = 0; i < ntx; ++i) vt0[i] = T(t0[i]);
        std::vector<std::vector<ttcr::sxyz<T>>> r_data;
        base().raytrace(Tx, vt0, Rx, vtt, r_data, th);
        for (size_t i = 0; i < nrx; ++i) {
            tt[i] = double(vtt[i]);
            npts[i] = r_data[i].size();
            for (size_t k = 0; k < r_data[i].size() && k < cap; ++k) {
                xyz[3 * (cap * i + k)] = double(r_data[i][k].x);
                xyz[3 * (cap * i + k) + 1] = double(r_data[i][k].y);
                xyz[3 * (cap * i + k) + 2] = double(r_data[i][k].z);
            }
        }
    }
    // Grid3D::raytrace(Tx,t0,Rx,traveltimes,m_data,threadNo) (Grid3D.h:743-780) or -- with_rays -- the overload with r_data and
    // m_data (:646-690; the two walk the M terms differently): receiver n gets nnz[n] (column, value) pairs of the matrix M in
    // the order the reference leaves them in m_data[n], the first `cap` of them stored at cap * n
    void raytrace_m(const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx, double* tt, size_t th,
                    size_t* nnz, unsigned long long* col, double* val, size_t cap, int with_rays) override {
        std::vector<ttcr::sxyz<T>> Tx, Rx;
        pts(tx, ntx, Tx);
        pts(rx, nrx, Rx);
        std::vector<T> vt0(ntx), vtt(nrx);
        for (size_t i = 0; i < ntx; ++i) vt0[i] = T(t0[i]);
        std::vector<std::vector<ttcr::sijv<T>>> m_data;
        std::vector<std::vector<ttcr::sxyz<T>>> r_data;
        if (with_rays) base().raytrace(Tx, vt0, Rx, vtt, r_data, m_data, th);
        else base().raytrace(Tx, vt0, Rx, vtt, m_data, th);
        for (size_t i = 0; i < nrx; ++i) {
            tt[i] = double(vtt[i]);
            nnz[i] = m_data[i].size();
            for (size_t k = 0; k < m_data[i].size() && k < cap; ++k) {
                col[cap * i + k] = m_data[i][k].j;
                val[cap * i + k] = double(m_data[i][k].v);
            }
        }
    }
    // one Tx point per source, same receivers for every source; fan-out is the
    // reference's own (ttcr/Grid3D.h:810-853: thread pool / std::thread blocks)
    void raytrace_multi(size_t nsrc, const double* tx, const double* t0, const double* rx,
                        size_t nrx, double* tt) override {
        std::vector<std::vector<ttcr::sxyz<T>>> Tx(nsrc), Rx(nsrc);
        std::vector<std::vector<T>> vt0(nsrc), vtt(nsrc);
        for (size_t s = 0; s < nsrc; ++s) {
            pts(tx + 3 * s, 1, Tx[s]);
            pts(rx, nrx, Rx[s]);
            vt0[s].assign(1, T(t0[s]));
            vtt[s].resize(nrx);
        }
        base().raytrace(Tx, vt0, Rx, vtt);
        for (size_t s = 0; s < nsrc; ++s)
            for (size_t i = 0; i < nrx; ++i) tt[s * nrx + i] = double(vtt[s][i]);
    }
    void get_tt(double* out, size_t th) override {
        std::vector<T> v;
        g->getTT(v, th);
        for (size_t i = 0; i < v.size(); ++i) out[i] = double(v[i]);
    }
    void niter(int* a, int* b) override {
        *a = g->get_niter();
        *b = g->get_niterw();
    }
    size_t nnodes() override { return g->getNumberOfNodes(); }
};

template <typename F>
int guard(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::length_error& e) {
        g_err = e.what();
        return 2;
    } catch (const std::logic_error& e) {
        g_err = e.what();
        return 3;
    } catch (const std::runtime_error& e) {
        g_err = e.what();
        return 1;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 4;
    }
}

}  // namespace

// One solve of the reference's 2-D Grid2Drnfs / Grid2Drcfs <T, uint32_t, sxz<T>> (Grid2Drnfs.h:83-95, Grid2Drcfs.h:36-50):
// slowness s at the nodes ((ncx+1)(ncz+1) values) or in the cells (ncx ncz values), z fastest; Tx points tx (ntx x 2:
// x, z) with origin times t0; receivers rx (nrx x 2) get Grid2Drn::getTraveltime (tt_from_rp = false).  Returns the
// traveltime field, the receiver times and the iteration counts.
template <typename T, typename G>
static void solve2d(uint32_t ncx, uint32_t ncz, double dx, double dz, double xmin, double zmin, double eps, int maxit, int weno,
                    int rotated, const double* s, size_t ns, const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx,
                    double* tt, double* tt_rx, int* niter, int* niterw) {
    typedef ttcr::sxz<T> S;
    G g(ncx, ncz, T(dx), T(dz), T(xmin), T(zmin), T(eps), maxit, weno != 0, rotated != 0, false, 1);
    std::vector<T> v(ns);
    for (size_t i = 0; i < ns; ++i) v[i] = T(s[i]);
    g.setSlowness(v);
    std::vector<S> Tx(ntx), Rx(nrx);
    std::vector<T> vt0(ntx), tr;
    for (size_t i = 0; i < ntx; ++i) { Tx[i].x = T(tx[2 * i]); Tx[i].z = T(tx[2 * i + 1]); vt0[i] = T(t0[i]); }
    for (size_t i = 0; i < nrx; ++i) { Rx[i].x = T(rx[2 * i]); Rx[i].z = T(rx[2 * i + 1]); }
    static_cast<ttcr::Grid2D<T, uint32_t, S>&>(g).raytrace(Tx, vt0, Rx, tr, 0);
    for (size_t i = 0; i < nrx; ++i) tt_rx[i] = double(tr[i]);
    std::vector<T> f;
    g.getTT(f, 0);
    for (size_t i = 0; i < f.size(); ++i) tt[i] = double(f[i]);
    *niter = g.get_niter();
    *niterw = g.get_niterw();
}

extern "C" {

const char* ttcr_ref_last_error() { return g_err.c_str(); }

// dtype: 0 = double, 1 = float.  cell: 0 = Grid3Drnfs (node slowness), 1 = Grid3Drcfs.
// nx,ny,nz are CELL counts, as in the reference constructors (Grid3Drnfs.h:39-50).
void* ttcr_ref_create(int dtype, int cell, uint32_t nx, uint32_t ny, uint32_t nz, double dx,
                      double x0, double y0, double z0, double eps, int maxit, int weno, int ttrp,
                      int nthreads, int translate) {
    Base* b = nullptr;
    int rc = guard([&] {
        if (dtype == 0 && !cell)
            b = new Impl<double, ttcr::Grid3Drnfs<double, uint32_t>>(nx, ny, nz, dx, x0, y0, z0, eps, maxit, weno, ttrp, nthreads, translate);
        else if (dtype == 0)
            b = new Impl<double, ttcr::Grid3Drcfs<double, uint32_t>>(nx, ny, nz, dx, x0, y0, z0, eps, maxit, weno, ttrp, nthreads, translate);
        else if (!cell)
            b = new Impl<float, ttcr::Grid3Drnfs<float, uint32_t>>(nx, ny, nz, dx, x0, y0, z0, eps, maxit, weno, ttrp, nthreads, translate);
        else
            b = new Impl<float, ttcr::Grid3Drcfs<float, uint32_t>>(nx, ny, nz, dx, x0, y0, z0, eps, maxit, weno, ttrp, nthreads, translate);
    });
    return rc == 0 ? b : nullptr;
}
void ttcr_ref_destroy(void* h) { delete static_cast<Base*>(h); }
size_t ttcr_ref_nnodes(void* h) { return static_cast<Base*>(h)->nnodes(); }
int ttcr_ref_set_slowness(void* h, const double* s, size_t n) {
    return guard([&] { static_cast<Base*>(h)->set_slowness(s, n); });
}
int ttcr_ref_get_slowness(void* h, double* out) {
    return guard([&] { static_cast<Base*>(h)->get_slowness(out); });
}
int ttcr_ref_raytrace(void* h, const double* tx, const double* t0, size_t ntx, const double* rx,
                      size_t nrx, double* tt, size_t thread_no, double* seconds) {
    auto t_0 = std::chrono::high_resolution_clock::now();
    int rc = guard([&] { static_cast<Base*>(h)->raytrace(tx, t0, ntx, rx, nrx, tt, thread_no); });
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t_0).count();
    return rc;
}
int ttcr_ref_raytrace_rays(void* h, const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx, double* tt,
                           size_t thread_no, size_t* npts, double* xyz, size_t cap) {
    return guard([&] { static_cast<Base*>(h)->raytrace_rays(tx, t0, ntx, rx, nrx, tt, thread_no, npts, xyz, cap); });
}
int ttcr_ref_raytrace_m(void* h, const double* tx, const double* t0, size_t ntx, const double* rx, size_t nrx, double* tt,
                        size_t thread_no, size_t* nnz, unsigned long long* col, double* val, size_t cap, int with_rays) {
    return guard([&] { static_cast<Base*>(h)->raytrace_m(tx, t0, ntx, rx, nrx, tt, thread_no, nnz, col, val, cap, with_rays); });
}
int ttcr_ref_raytrace_multi(void* h, size_t nsrc, const double* tx, const double* t0,
                            const double* rx, size_t nrx, double* tt, double* seconds) {
    auto t_0 = std::chrono::high_resolution_clock::now();
    int rc = guard([&] { static_cast<Base*>(h)->raytrace_multi(nsrc, tx, t0, rx, nrx, tt); });
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t_0).count();
    return rc;
}
int ttcr_ref_get_tt(void* h, double* out, size_t thread_no) {
    return guard([&] { static_cast<Base*>(h)->get_tt(out, thread_no); });
}
int ttcr_ref_get_niter(void* h, int* niter, int* niterw) {
    return guard([&] { static_cast<Base*>(h)->niter(niter, niterw); });
}

int ttcr_ref2d_solve(int dtype, int cell, uint32_t ncx, uint32_t ncz, double dx, double dz, double xmin, double zmin, double eps,
                     int maxit, int weno, int rotated, const double* s, size_t ns, const double* tx, const double* t0, size_t ntx,
                     const double* rx, size_t nrx, double* tt, double* tt_rx, int* niter, int* niterw) {
    return guard([&] {
#define TTCR_SOLVE2D(T, G) solve2d<T, ttcr::G<T, uint32_t, ttcr::sxz<T>>>(ncx, ncz, dx, dz, xmin, zmin, eps, maxit, weno, rotated, s, ns, \
                                                                         tx, t0, ntx, rx, nrx, tt, tt_rx, niter, niterw)
        if (dtype == 0 && !cell) TTCR_SOLVE2D(double, Grid2Drnfs);
        else if (dtype == 0) TTCR_SOLVE2D(double, Grid2Drcfs);
        else if (!cell) TTCR_SOLVE2D(float, Grid2Drnfs);
        else TTCR_SOLVE2D(float, Grid2Drcfs);
#undef TTCR_SOLVE2D
    });
}

}  // extern "C"
