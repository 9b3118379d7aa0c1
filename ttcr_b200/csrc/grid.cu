// Host side of libttcr_b200.so: the grid object behind the C ABI of include/ttcr_b200.h.
//
// Mirrors, for the FSM path only, what the reference spreads over Grid3D (Grid3D.h:470-502,
// :810-853: origin translation, receiver extraction, source fan-out), Grid3Drn (setSlowness /
// getTT / checkPts, Grid3Drn.h:82-108, :771-790) and the Grid3Drnfs / Grid3Drcfs drivers
// (Grid3Drnfs.h:84-155, Grid3Drcfs.h:88-247).  There is no CPU solver in this library: if a CUDA
// call fails the entry point returns TTCR_B200_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ttcr_b200.h"
#include "kernels.cuh"
#include "sweep_tile.cuh"
#include "raypath.cuh"
#include "sweep_march.cuh"
#include "sweep_march4.cuh"
#include "sweep_march_weno.cuh"
#include "grid2d.cuh"

namespace ttcrb200 {

struct Err : std::exception {
    int code;
    std::string msg;
    Err(int c, std::string m) : code(c), msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            std::ostringstream os_;                                                                   \
            os_ << "CUDA error: " << cudaGetErrorString(e_) << " at " << __FILE__ << ":" << __LINE__  \
                << " (" #call ")";                                                                    \
            throw Err(TTCR_B200_ERR_CUDA, os_.str());                                                 \
        }                                                                                             \
    } while (0)

static inline unsigned nblocks(size_t n, unsigned threads = 256, unsigned cap = 148 * 16) {
    size_t b = (n + threads - 1) / threads;
    return (unsigned)std::min<size_t>(std::max<size_t>(b, 1), cap);
}

// stream-ordered scratch allocation that is released on every exit path
struct ScratchBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ScratchBuf() = default;
    ScratchBuf(const ScratchBuf&) = delete;
    ScratchBuf& operator=(const ScratchBuf&) = delete;
    ~ScratchBuf() { if (p) cudaFreeAsync(p, st); }
    void alloc(size_t bytes, cudaStream_t s) {
        st = s;
        const cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 1, s);
        if (e != cudaSuccess) { p = nullptr; throw Err(TTCR_B200_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e)); }
    }
    template <typename U> U* as() const { return static_cast<U*>(p); }
};

struct GridBase {
    virtual ~GridBase() {}
    virtual void set_slowness(const void* s, size_t n, int order) = 0;
    virtual void set_slowness_device(const void* s, size_t n, int order) = 0;
    virtual void set_slowness_device_planes(const void* s, size_t n, int i0, int cnt) = 0;
    virtual void get_tt_device(void* out, size_t slot, int order) = 0;
    virtual void get_slowness(void* out, int order) = 0;
    virtual void solve(const void* tx, const void* t0, size_t ntx, size_t slot) = 0;
    virtual void raytrace(const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt,
                          size_t slot) = 0;
    virtual void raytrace_multi(size_t nsrc, const size_t* tx_off, const void* tx, const void* t0,
                                const size_t* rx_off, const void* rx, void* tt, int* niter, int* niterw) = 0;
    virtual void raytrace_rays(const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt, size_t* npts,
                               size_t slot) = 0;
    virtual void get_rays(size_t slot, void* xyz) = 0;
    virtual void get_m_terms(size_t slot, unsigned long long* node, void* value) = 0;
    virtual void get_tt(void* out, size_t slot, int order) = 0;
    virtual void stats(size_t slot, ttcr_b200_stats* out) = 0;
    virtual void set_option(const std::string& key, double v) = 0;
    virtual size_t n_slots() const = 0;
    virtual size_t device_bytes() const = 0;
};

template <typename T>
class Grid final : public GridBase {
   public:
    Grid(uint32_t ncx, uint32_t ncy, uint32_t ncz, double dx, double xmin, double ymin, double zmin, double eps,
         int maxit, bool weno, bool ttrp, bool interp_vel, size_t nslots, bool translate, bool cell, int device)
        : maxit_(maxit), weno_(weno), ttrp_(ttrp), cell_(cell), translate_(translate), intvel_(interp_vel) {
        if (ncx < 1 || ncy < 1 || ncz < 1) throw Err(TTCR_B200_ERR_INVALID, "grid must have at least one cell per axis");
        if (nslots < 1) throw Err(TTCR_B200_ERR_INVALID, "n_slots must be >= 1");
        if ((double)(ncx + 1) * (ncy + 1) * (ncz + 1) > 4294967295.0)
            throw Err(TTCR_B200_ERR_INVALID, "grid has more nodes than a uint32_t index can address");
        if (device < 0) CK(cudaGetDevice(&device));
        dev_ = device;
        CK(cudaSetDevice(dev_));
        d_ = make_dims((int)ncx + 1, (int)ncy + 1, (int)ncz + 1);
        // reference constructor arithmetic, in T (Grid3Drn.h:68-77, buildGridNodes :362-372)
        g_.dx = T(dx);
        g_.xmin = T(xmin); g_.ymin = T(ymin); g_.zmin = T(zmin);
        g_.xmax = g_.xmin + T(ncx) * g_.dx;
        g_.ymax = g_.ymin + T(ncy) * g_.dx;
        g_.zmax = g_.zmin + T(ncz) * g_.dx;
        g_.ncx = (int)ncx; g_.ncy = (int)ncy; g_.ncz = (int)ncz;
        origin_[0] = origin_[1] = origin_[2] = T(0);
        if (translate_) {
            origin_[0] = g_.xmin; origin_[1] = g_.ymin; origin_[2] = g_.zmin;
            g_.xmax -= g_.xmin; g_.ymax -= g_.ymin; g_.zmax -= g_.zmin;
            g_.xmin = g_.ymin = g_.zmin = T(0);
        }
        // per-node tolerance -> L1 threshold, in T (Grid3Drnfs.h:49)
        epsilon_ = T(eps);
        epsilon_ *= static_cast<T>(d_.nodes());

        const size_t ne = d_.elems();
        for (int l = 0; l < 2; ++l) {
            slo_[l] = alloc_field(ne);
            // slots that are no node hold NaN: their update is NaN and never passes `t < old` (sweep_march.cuh)
            CK(cudaMemset(slo_[l], 0xFF, ne * sizeof(T)));
        }
        slots_.resize(nslots);
        for (auto& s : slots_) {
            CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            CK(cudaEventCreate(&s.e0));
            CK(cudaEventCreate(&s.e1));
            for (int l = 0; l < 2; ++l) {
                s.tt[l] = alloc_field(ne);
                CK(cudaMalloc(&s.mask[l], ne / 32 * sizeof(uint32_t)));
                bytes_ += ne / 8;
                k_fill<T><<<nblocks(ne), 256, 0, s.stream>>>(s.tt[l], ne, Lim<T>::max());
                CK(cudaMemsetAsync(s.mask[l], 0, ne / 8, s.stream));
            }
            CK(cudaMalloc(&s.d_change, sizeof(double)));
            CK(cudaMalloc(&s.d_fb, sizeof(FrozenBox)));
            CK(cudaMalloc(&s.d_bar, sizeof(unsigned)));
            CK(cudaMallocHost(&s.h_change, sizeof(double)));
            tile_alloc(s.tile, d_, bytes_);
            CK(cudaStreamSynchronize(s.stream));
        }
        int sms = 0;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_));
        sm_count_ = sms;
    }

    ~Grid() override {
        cudaSetDevice(dev_);
        for (auto& s : slots_) {
            cudaStreamSynchronize(s.stream);
            for (int l = 0; l < 2; ++l) { free_field(s.tt[l]); cudaFree(s.mask[l]); }
            for (auto e : s.sweep_ev) cudaEventDestroy(e);
            cudaFree(s.d_change); cudaFreeHost(s.h_change);
            cudaFree(s.d_fb); cudaFree(s.d_bar); cudaFree(s.d_ppart);
            for (auto& a : s.pgraph)
                for (auto& b : a)
                    if (b.exec) cudaGraphExecDestroy(b.exec);
            cudaFree(s.d_pts); cudaFreeHost(s.h_pts);
            tile_free(s.tile);
            march_free(s.march);
            march_free(s.march4);
            marchw_free(s.marchw);
            cudaEventDestroy(s.e0); cudaEventDestroy(s.e1);
            cudaStreamDestroy(s.stream);
        }
        for (int l = 0; l < 2; ++l) free_field(slo_[l]);
        cudaFree(lin_);
        if (copy_st_) {
            cudaStreamDestroy(copy_st_);
            for (auto& e : copy_ev_) cudaEventDestroy(e);
        }
    }

    size_t n_slots() const override { return slots_.size(); }
    size_t device_bytes() const override { return bytes_; }

    // ---- model -------------------------------------------------------------------------------
    void set_slowness(const void* s, size_t n, int order) override { set_slowness_any(s, n, order, cudaMemcpyHostToDevice); }
    void set_slowness_device(const void* s, size_t n, int order) override { set_slowness_any(s, n, order, cudaMemcpyDeviceToDevice); }

    void set_slowness_device_planes(const void* s, size_t n, int i0, int cnt) override {
        CK(cudaSetDevice(dev_));
        if (cell_) throw Err(TTCR_B200_ERR_INVALID, "set_slowness_device_planes: node models only");
        if (n != d_.nodes()) throw Err(TTCR_B200_ERR_LENGTH, "Error: slowness vectors of incompatible size.");
        if (i0 < 0 || cnt <= 0 || i0 + cnt > d_.ni) throw Err(TTCR_B200_ERR_INVALID, "bad plane range");
        std::lock_guard<std::mutex> lk(lin_mu_);
        cudaStream_t st = slots_[0].stream;
        if (i0 == 0) have_slowness_ = false;
        const size_t ne = (size_t)cnt * d_.qs * d_.kpad;
        for (int l = 0; l < 2; ++l) k_import<T><<<nblocks(ne), 256, 0, st>>>((const T*)s, 1, slo_[l], l, d_, i0, i0 + cnt);
        CK(cudaGetLastError());
        if (i0 + cnt == d_.ni) {
            CK(cudaStreamSynchronize(st));
            have_slowness_ = true;
        }
    }

    void set_slowness_any(const void* s, size_t n, int order, cudaMemcpyKind kind) {
        CK(cudaSetDevice(dev_));
        const size_t want = cell_ ? (size_t)g_.ncx * g_.ncy * g_.ncz : d_.nodes();
        if (n != want) throw Err(TTCR_B200_ERR_LENGTH, "Error: slowness vectors of incompatible size.");
        if (order != 0 && order != 1) throw Err(TTCR_B200_ERR_INVALID, "bad order");
        std::lock_guard<std::mutex> lk(lin_mu_);
        ensure_lin();
        cudaStream_t st = slots_[0].stream;
        T* const stage = staging(slots_[0]);
        // A large node model in numpy order coming from the host: the x planes are contiguous in both the source and the
        // sheared layouts, so the model is copied in chunks of planes on a copy stream and every chunk is imported (into
        // both layouts) while the next one is still on the bus.  The copy is the longer leg; only the last import shows.
        constexpr int NCHUNK = 8;
        if (kind == cudaMemcpyHostToDevice && !cell_ && order == 1 && d_.ni >= 4 * NCHUNK && n * sizeof(T) >= (size_t(64) << 20) &&
            !getenv("TTCR_B200_NO_PIPELINED_IMPORT")) {
            if (!copy_st_) {
                CK(cudaStreamCreateWithFlags(&copy_st_, cudaStreamNonBlocking));
                for (auto& e : copy_ev_) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
            CK(cudaEventRecord(copy_ev_[NCHUNK], st));               // earlier users of the staging buffer
            CK(cudaStreamWaitEvent(copy_st_, copy_ev_[NCHUNK], 0));
            const size_t plane = (size_t)d_.nj * d_.nk;
            for (int c = 0; c < NCHUNK; ++c) {
                const int i0 = (int)((long long)d_.ni * c / NCHUNK), i1 = (int)((long long)d_.ni * (c + 1) / NCHUNK);
                CK(cudaMemcpyAsync(stage + i0 * plane, (const T*)s + i0 * plane, (size_t)(i1 - i0) * plane * sizeof(T), kind, copy_st_));
                CK(cudaEventRecord(copy_ev_[c], copy_st_));
                CK(cudaStreamWaitEvent(st, copy_ev_[c], 0));
                const size_t ne = (size_t)(i1 - i0) * d_.qs * d_.kpad;
                for (int l = 0; l < 2; ++l) k_import<T><<<nblocks(ne), 256, 0, st>>>(stage, order, slo_[l], l, d_, i0, i1);
            }
            release_staging(slots_[0], st);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
            have_slowness_ = true;
            return;
        }
        CK(cudaMemcpyAsync(stage, s, n * sizeof(T), kind, st));
        const T* nodes = stage;
        if (cell_) {
            k_cell_to_node<T><<<nblocks(d_.nodes()), 256, 0, st>>>(stage, lin_, order, g_.ncx, g_.ncy, g_.ncz);
            nodes = lin_;
        }
        for (int l = 0; l < 2; ++l) k_import<T><<<nblocks(d_.elems()), 256, 0, st>>>(nodes, order, slo_[l], l, d_);
        release_staging(slots_[0], st);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        have_slowness_ = true;
    }

    void get_slowness(void* out, int order) override {
        CK(cudaSetDevice(dev_));
        if (order != 0 && order != 1) throw Err(TTCR_B200_ERR_INVALID, "bad order");
        std::lock_guard<std::mutex> lk(lin_mu_);
        cudaStream_t st = slots_[0].stream;
        T* const stage = staging(slots_[0]);
        k_export<T><<<nblocks(d_.nodes()), 256, 0, st>>>(slo_[0], 0, stage, order, d_);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, stage, d_.nodes() * sizeof(T), cudaMemcpyDeviceToHost, st));
        release_staging(slots_[0], st);
        CK(cudaStreamSynchronize(st));
    }

    void get_tt(void* out, size_t slot, int order) override {
        CK(cudaSetDevice(dev_));
        Slot& s = slot_at(slot);
        if (order != 0 && order != 1) throw Err(TTCR_B200_ERR_INVALID, "bad order");
        T* const stage = staging(s);   // (the slot's own scratch: no lock, slots export concurrently)
        k_export<T><<<nblocks(d_.nodes()), 256, 0, s.stream>>>(s.tt[0], 0, stage, order, d_);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, stage, d_.nodes() * sizeof(T), cudaMemcpyDeviceToHost, s.stream));
        release_staging(s, s.stream);
        CK(cudaStreamSynchronize(s.stream));
    }

    void get_tt_device(void* out, size_t slot, int order) override {
        CK(cudaSetDevice(dev_));
        Slot& s = slot_at(slot);
        if (order != 0 && order != 1) throw Err(TTCR_B200_ERR_INVALID, "bad order");
        k_export<T><<<nblocks(d_.nodes()), 256, 0, s.stream>>>(s.tt[0], 0, (T*)out, order, d_);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(s.stream));
    }

    // ---- solve -------------------------------------------------------------------------------
    void solve(const void* tx, const void* t0, size_t ntx, size_t slot) override {
        CK(cudaSetDevice(dev_));
        Slot& s = slot_at(slot);
        std::vector<T> vtx, vt0;
        prepare_tx(tx, t0, ntx, vtx, vt0);
        solve_device(s, vtx, vt0);
    }

    void raytrace(const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt, size_t slot) override {
        raytrace_impl(tx, t0, ntx, rx, nrx, tt, slot, nullptr);
    }

    // Grid3D::raytrace(Tx,t0,Rx,traveltimes,r_data,threadNo) (Grid3D.h:545-586): traveltimes AND raypaths, both from
    // Grid3Drn::getRaypath (Grid3Drn.h:1339-1500), whatever tt_from_rp says.  The points stay on the slot until get_rays.
    void raytrace_rays(const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt, size_t* npts,
                       size_t slot) override {
        raytrace_impl(tx, t0, ntx, rx, nrx, tt, slot, npts);
    }

    void get_rays(size_t slot, void* xyz) override {
        Slot& s = slot_at(slot);
        if (!s.rays.empty()) std::memcpy(xyz, s.rays.data(), s.rays.size() * sizeof(T));
    }

    // The raw terms of the matrix M of the last raytrace_rays call made with option "m_terms" = 1 (Grid3Drn::getRaypath with
    // m_data, Grid3Drn.h:1500-1801): 8 (column, value) pairs per ray point, ray after ray, in the order the reference produces
    // them; the 8 pairs of every ray's FIRST point (the receiver closes no segment) are zero.  8 * sum(ray_npts) elements each.
    void get_m_terms(size_t slot, unsigned long long* node, void* value) override {
        Slot& s = slot_at(slot);
        if (s.m_node.size() != 8 * (s.rays.size() / 3)) throw Err(TTCR_B200_ERR_LOGIC, "no M terms: set option m_terms = 1 before raytrace_rays");
        if (!s.m_node.empty()) {
            std::memcpy(node, s.m_node.data(), s.m_node.size() * sizeof(unsigned long long));
            std::memcpy(value, s.m_val.data(), s.m_val.size() * sizeof(T));
        }
    }

    void raytrace_impl(const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt, size_t slot, size_t* npts) {
        CK(cudaSetDevice(dev_));
        Slot& s = slot_at(slot);
        const bool want_rays = npts != nullptr;
        const bool walk = ttrp_ || want_rays;
        std::vector<T> vtx, vt0, vrx;
        prepare_tx(tx, t0, ntx, vtx, vt0);
        // checkPts(Rx) before any work, as Grid3Drnfs.h:89-90
        vrx.assign((const T*)rx, (const T*)rx + 3 * nrx);
        translate_pts(vrx);
        check_pts(vrx);
        ensure_pts(s, 4 * ntx + 5 * nrx);   // Tx, t0 | Rx, traveltimes, status (before the solve: it uploads Tx into this buffer)
        solve_device(s, vtx, vt0);
        s.rays.clear();
        s.m_node.clear(); s.m_val.clear();
        if (nrx) {
            T* const h_rx = s.h_pts + 4 * ntx;
            T* const d_rx = s.d_pts + 4 * ntx;
            std::memcpy(h_rx, vrx.data(), 3 * nrx * sizeof(T));
            CK(cudaMemcpyAsync(d_rx, h_rx, 3 * nrx * sizeof(T), cudaMemcpyHostToDevice, s.stream));
            T* d_out = d_rx + 3 * nrx;
            const unsigned rp_blocks = (unsigned)((nrx + 63) / 64);
            ScratchBuf b_rn, b_off, b_xyz, b_mn, b_mv;
            if (want_rays) {   // counts and offsets of the ray points
                b_rn.alloc(nrx * sizeof(int), s.stream);
                b_off.alloc(nrx * sizeof(unsigned long long), s.stream);
            }
            int* const d_rn = b_rn.as<int>();
            unsigned long long* const d_off = b_off.as<unsigned long long>();
            if (walk) {
                // Grid3D.h:493-501 with tt_from_rp: traveltimes integrated along the raypaths (raypath.cuh)
                k_tt_from_rp<T><<<rp_blocks, 64, 0, s.stream>>>(g_, d_, s.tt[0], slo_[0], s.d_pts, s.d_pts + 3 * ntx, (int)ntx, d_rx, (int)nrx,
                                                                d_out, d_out + nrx, d_rn, nullptr, nullptr, intvel_, want_rays ? m_terms_ : 0);
            } else {
                k_interp<T><<<(unsigned)((nrx + 127) / 128), 128, 0, s.stream>>>(g_, d_, s.tt[0], d_rx, (int)nrx, d_out);
            }
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(h_rx + 3 * nrx, d_out, (walk ? 2 : 1) * nrx * sizeof(T), cudaMemcpyDeviceToHost, s.stream));
            std::vector<int> rn(want_rays ? nrx : 0);
            if (want_rays) CK(cudaMemcpyAsync(rn.data(), d_rn, nrx * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            std::memcpy(tt, h_rx + 3 * nrx, nrx * sizeof(T));
            if (walk) {
                const T* st = h_rx + 4 * nrx;
                for (size_t n = 0; n < nrx; ++n) {
                    if (st[n] == T(0)) continue;
                    std::ostringstream msg;
                    if (st[n] == T(1))
                        msg << "Error while computing raypaths: going outside grid \n                Rx: " << vrx[3 * n] << " " << vrx[3 * n + 1] << " "
                            << vrx[3 * n + 2] << "\n                Tx: " << vtx[0] << " " << vtx[1] << " " << vtx[2] << "\n";
                    else
                        msg << "Error while computing raypaths: the ray from Rx " << vrx[3 * n] << " " << vrx[3 * n + 1] << " " << vrx[3 * n + 2]
                            << " did not reach a source point";
                    throw Err(TTCR_B200_ERR_RUNTIME, msg.str());
                }
            }
            if (want_rays) {
                // second pass of the same walk, now storing the points at their offsets
                std::vector<unsigned long long> off(nrx);
                unsigned long long total = 0;
                for (size_t n = 0; n < nrx; ++n) { off[n] = total; total += (unsigned long long)rn[n]; npts[n] = (size_t)rn[n]; }
                b_xyz.alloc(3 * total * sizeof(T), s.stream);
                T* const d_xyz = b_xyz.as<T>();
                CK(cudaMemcpyAsync(d_off, off.data(), nrx * sizeof(unsigned long long), cudaMemcpyHostToDevice, s.stream));
                unsigned long long* d_mn = nullptr;
                T* d_mv = nullptr;
                if (m_terms_) {   // (cell grids: the reference defines M for node slowness only, rgrid.pyx:910-911)
                    if (cell_) throw Err(TTCR_B200_ERR_INVALID, "M terms are defined for node slowness only");
                    b_mn.alloc(8 * total * sizeof(unsigned long long), s.stream);
                    b_mv.alloc(8 * total * sizeof(T), s.stream);
                    d_mn = b_mn.as<unsigned long long>(); d_mv = b_mv.as<T>();
                    CK(cudaMemsetAsync(d_mn, 0, 8 * total * sizeof(unsigned long long), s.stream));
                    CK(cudaMemsetAsync(d_mv, 0, 8 * total * sizeof(T), s.stream));
                }
                k_tt_from_rp<T><<<rp_blocks, 64, 0, s.stream>>>(g_, d_, s.tt[0], slo_[0], s.d_pts, s.d_pts + 3 * ntx, (int)ntx, d_rx, (int)nrx,
                                                                d_out, d_out + nrx, d_rn, d_off, d_xyz, intvel_, m_terms_, d_mn, d_mv);
                CK(cudaGetLastError());
                s.rays.resize(3 * total);
                CK(cudaMemcpyAsync(s.rays.data(), d_xyz, 3 * total * sizeof(T), cudaMemcpyDeviceToHost, s.stream));
                if (m_terms_) {
                    s.m_node.resize(8 * total); s.m_val.resize(8 * total);
                    CK(cudaMemcpyAsync(s.m_node.data(), d_mn, 8 * total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
                    CK(cudaMemcpyAsync(s.m_val.data(), d_mv, 8 * total * sizeof(T), cudaMemcpyDeviceToHost, s.stream));
                }
                CK(cudaStreamSynchronize(s.stream));
                if (translate_)   // Grid3D.h:578-584: r_data += origin
                    for (size_t n = 0; n < s.rays.size(); n += 3) { s.rays[n] += origin_[0]; s.rays[n + 1] += origin_[1]; s.rays[n + 2] += origin_[2]; }
            }
        }
    }

    // Grid3D.h:810-853: sources dealt to the slots; one host thread per slot drives its stream.
    void raytrace_multi(size_t nsrc, const size_t* tx_off, const void* tx, const void* t0, const size_t* rx_off,
                        const void* rx, void* tt, int* niter, int* niterw) override {
        const size_t nt = std::min(slots_.size(), nsrc);
        if (nt == 0) return;
        std::vector<std::string> errs(nt);
        std::vector<int> codes(nt, 0);
        auto work = [&](size_t t) {
            try {
                for (size_t sidx = t; sidx < nsrc; sidx += nt) {
                    const size_t a = tx_off[sidx], b = tx_off[sidx + 1], ra = rx_off[sidx], rb = rx_off[sidx + 1];
                    raytrace((const T*)tx + 3 * a, (const T*)t0 + a, b - a, (const T*)rx + 3 * ra, rb - ra,
                             (T*)tt + ra, t);
                    if (niter) niter[sidx] = slots_[t].st.niter;
                    if (niterw) niterw[sidx] = slots_[t].st.niterw;
                }
            } catch (const Err& e) {
                codes[t] = e.code; errs[t] = e.msg;
            } catch (const std::exception& e) {
                codes[t] = TTCR_B200_ERR_RUNTIME; errs[t] = e.what();
            }
        };
        if (nt == 1) {
            work(0);
        } else {
            std::vector<std::thread> th;
            for (size_t t = 0; t < nt; ++t) th.emplace_back(work, t);
            for (auto& x : th) x.join();
        }
        for (size_t t = 0; t < nt; ++t)
            if (codes[t]) throw Err(codes[t], errs[t]);
    }

    void stats(size_t slot, ttcr_b200_stats* out) override { *out = slot_at(slot).st; }

    void set_option(const std::string& key, double v) override {
        if (key == "tt_from_rp") ttrp_ = v != 0;
        else if (key == "kernel") {
            if (v != TTCR_B200_KERNEL_AUTO && v != TTCR_B200_KERNEL_PLANE && v != TTCR_B200_KERNEL_TILE &&
                v != TTCR_B200_KERNEL_COOP && v != TTCR_B200_KERNEL_MARCH)
                throw Err(TTCR_B200_ERR_INVALID, "unknown kernel id");
            kernel_ = (int)v;
        } else if (key == "tile_rows") tile_opt_.chunk = std::max(1, (int)v);
        else if (key == "ctas_per_sm") tile_opt_.ctas_per_sm = std::max(0, (int)v);
        else if (key == "tile_warps") {   // 0: back to the choice by grid size
            warps_set_ = v != 0;
            tile_opt_.warps = warps_set_ ? std::max(1, std::min(16, (int)v)) : 8;
        }
        else if (key == "tile_urows") tile_opt_.rows = std::max(1, std::min(4, (int)v));
        else if (key == "tile_depth") tile_opt_.depth = (int)v;
        else if (key == "spin_limit") tile_opt_.spin_limit = (long long)v;
        else if (key == "max_ctas") tile_opt_.max_ctas = std::max(0, (int)v);
        else if (key == "march_nodes") {
            if (v != 0 && v != 2 && v != 4) throw Err(TTCR_B200_ERR_INVALID, "march_nodes: 0 (by grid size), 2 or 4");
            tile_opt_.nodes = (int)v;
        }
        else if (key == "plane_graph") plane_graph_ = v != 0;
        else if (key == "coop_ctas") coop_ctas_ = std::max(1, std::min(8, (int)v));
        else if (key == "weno_kernel") {
            if (v != TTCR_B200_KERNEL_AUTO && v != TTCR_B200_KERNEL_PLANE && v != TTCR_B200_KERNEL_COOP && v != TTCR_B200_KERNEL_MARCH)
                throw Err(TTCR_B200_ERR_INVALID, "weno_kernel: AUTO, PLANE, COOP or MARCH");
            weno_kernel_ = (int)v;
        }
        else if (key == "plane_pdl") plane_pdl_ = v != 0;
        else if (key == "m_terms") {
            if (v != 0 && v != 1 && v != 2) throw Err(TTCR_B200_ERR_INVALID, "m_terms: 0, 1 (as the m_data overload) or 2 (as the r_data + m_data overload)");
            m_terms_ = (int)v;
        }
        else if (key == "use_pool") {}
        else if (key == "maxit") maxit_ = (int)v;
        else throw Err(TTCR_B200_ERR_INVALID, "unknown option '" + key + "'");
    }

   private:
    struct Slot {
        T* tt[2] = {nullptr, nullptr};          // traveltime field in layouts L1, L2 (padding = MAX)
        uint32_t* mask[2] = {nullptr, nullptr}; // frozen bit per slot, both layouts
        FrozenBox prev_fb{};                    // bounding box of the bits that may be set (the previous source's)
        bool have_prev_fb = false;
        cudaStream_t stream = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        double* d_change = nullptr;
        double* h_change = nullptr;
        T* d_pts = nullptr;   // small point buffer (Tx, t0, Rx, out)
        T* h_pts = nullptr;
        size_t pts_cap = 0;
        TileState tile;
        MarchState march;
        MarchState march4;
        MarchWState marchw;
        ttcr_b200_stats st{};
        unsigned* d_bar = nullptr;           // arrival counter of k_sweep_planes_coop's grid barrier
        double* d_ppart = nullptr;           // per-block sums of the plane kernels' decreases (one slot per block and plane)
        size_t ppart_cap = 0;
        FrozenBox* d_fb = nullptr;           // the source's frozen box, for k_sweep_plane (launch arguments stay source independent)
        struct PlaneGraph { cudaGraphExec_t exec = nullptr; bool pdl = false; };
        PlaneGraph pgraph[8][2];             // captured plane launches of a direction: [dir][first order | WENO]
        std::vector<T> rays;                 // points of the last raytrace_rays call, ray after ray (x,y,z)
        std::vector<unsigned long long> m_node;   // raw M terms of that call (option m_terms): 8 per ray point
        std::vector<T> m_val;
        std::vector<cudaEvent_t> sweep_ev;   // pairs of events around the directional sweeps
        size_t sweep_ev_used = 0;
    };

    // Field arrays (traveltime, slowness) carry ni rows of front padding: the skewed tensor map of k_sweep_patch
    // (make_skew_map, "plus" variant) is based that far below the array and the TMA unit wants a mapped base.
    size_t front_pad() const { return (size_t)d_.ni * d_.kpad; }
    T* alloc_field(size_t ne) {
        T* raw = nullptr;
        CK(cudaMalloc(&raw, (ne + front_pad()) * sizeof(T)));
        CK(cudaMemset(raw, 0, front_pad() * sizeof(T)));
        bytes_ += (ne + front_pad()) * sizeof(T);
        return raw + front_pad();
    }
    void free_field(T* p) {
        if (p) cudaFree(p - front_pad());
    }

    Slot& slot_at(size_t slot) {
        if (slot >= slots_.size()) throw Err(TTCR_B200_ERR_INVALID, "Thread number is larger than number of threads");
        return slots_[slot];
    }

    // Linear staging (a model or a field in host order, on its way in or out).  Between solves the L2-layout traveltime array
    // of a slot is scratch -- a solve starts in L1 and k_relayout2 rewrites every node of L2 before the first sweep that reads
    // it -- so its first nodes() elements serve as the staging buffer (it has elems() > 2 nodes() of them) instead of a
    // separate allocation (0.5 GB at 512^3); release_staging() restores the +MAX the non-node slots of that range must hold.
    T* staging(Slot& s) { return s.tt[1]; }
    void release_staging(Slot& s, cudaStream_t st) {
        const size_t n = (d_.nodes() + 3) / 4 * 4;   // (elems() is a multiple of 32)
        k_fill16<T><<<nblocks(n / (16 / sizeof(T)), 256, 148 * 32), 256, 0, st>>>(s.tt[1], n, Lim<T>::max());
    }
    void ensure_lin() {   // the averaged node model of a cell grid; node grids never need it
        if (!cell_ || lin_) return;
        CK(cudaMalloc(&lin_, d_.nodes() * sizeof(T)));
        bytes_ += d_.nodes() * sizeof(T);
    }

    void ensure_ppart(Slot& s, size_t n) {
        if (n <= s.ppart_cap) return;
        CK(cudaStreamSynchronize(s.stream));
        for (auto& a : s.pgraph)          // (captured graphs hold the old pointer)
            for (auto& b : a)
                if (b.exec) { cudaGraphExecDestroy(b.exec); b.exec = nullptr; }
        cudaFree(s.d_ppart);
        s.d_ppart = nullptr;
        CK(cudaMalloc(&s.d_ppart, n * sizeof(double)));
        s.ppart_cap = n;
    }

    void ensure_pts(Slot& s, size_t n) {
        if (n <= s.pts_cap) return;
        CK(cudaStreamSynchronize(s.stream));
        cudaFree(s.d_pts); cudaFreeHost(s.h_pts);
        s.pts_cap = std::max<size_t>(n, 1024);
        CK(cudaMalloc(&s.d_pts, s.pts_cap * sizeof(T)));
        CK(cudaMallocHost(&s.h_pts, s.pts_cap * sizeof(T)));
    }

    void translate_pts(std::vector<T>& p) const {   // Grid3D.h:478-485
        if (!translate_) return;
        for (size_t n = 0; n < p.size(); n += 3) { p[n] -= origin_[0]; p[n + 1] -= origin_[1]; p[n + 2] -= origin_[2]; }
    }

    void check_pts(const std::vector<T>& p) const {   // Grid3Drn.h:771-790
        for (size_t n = 0; n < p.size(); n += 3) {
            if (p[n] < g_.xmin || p[n] > g_.xmax || p[n + 1] < g_.ymin || p[n + 1] > g_.ymax || p[n + 2] < g_.zmin ||
                p[n + 2] > g_.zmax) {
                std::ostringstream msg;
                msg << "Error: Point (" << p[n] << " " << p[n + 1] << " " << p[n + 2] << ") outside grid.";
                throw Err(TTCR_B200_ERR_RUNTIME, msg.str());
            }
        }
    }

    void prepare_tx(const void* tx, const void* t0, size_t ntx, std::vector<T>& vtx, std::vector<T>& vt0) const {
        if (!have_slowness_) throw Err(TTCR_B200_ERR_LOGIC, "slowness has not been set");
        if (ntx == 0) throw Err(TTCR_B200_ERR_INVALID, "source has no Tx point");
        vtx.assign((const T*)tx, (const T*)tx + 3 * ntx);
        vt0.assign((const T*)t0, (const T*)t0 + ntx);
        translate_pts(vtx);
        check_pts(vtx);
    }

    FrozenBox frozen_box(const std::vector<T>& vtx, int npts) const {
        FrozenBox fb{INT32_MAX, -1, INT32_MAX, -1, INT32_MAX, -1};
        auto upd = [&](double p, double mn, int nc, int& lo, int& hi) {
            const int c = (int)std::floor((p - mn) / (double)g_.dx);
            lo = std::min(lo, std::max(0, c - npts - 1));
            hi = std::max(hi, std::min(nc, c + npts + 2));
        };
        for (size_t n = 0; n < vtx.size(); n += 3) {
            upd(vtx[n], g_.xmin, g_.ncx, fb.ilo, fb.ihi);
            upd(vtx[n + 1], g_.ymin, g_.ncy, fb.jlo, fb.jhi);
            upd(vtx[n + 2], g_.zmin, g_.ncz, fb.klo, fb.khi);
        }
        return fb;
    }

    // one directional sweep (dir 0..7 in the reference order) on the layout the field is in
    void launch_sweep(Slot& s, int dir, bool weno_stage, const FrozenBox& fb, int kernel) {
        const SweepView w = make_view(d_, dir);
        T* tt = s.tt[w.layout];
        if (kernel == TTCR_B200_KERNEL_MARCH && weno_stage) {
            const int nl = marchw_sweep<T>(s.tile, s.marchw, tile_opt_, sm_count_, w, d_, tt, slo_[w.layout], s.mask[w.layout], fb,
                                          g_.dx, s.d_change, s.stream);
            s.st.launches += nl; s.st.sweep_launches += nl;
            return;
        }
        // Two or four nodes per thread: the four-node step is 1.7x as long and does twice the work.  Where a sweep is bound by
        // its chain of dependent steps (512^3: 3.5 tiles per SM) the two-node kernel wins (0.69 against 0.80 ms), where it is
        // bound by the SMs' throughput the four-node kernel does (measured cross-over between 640^3 and 768^3; 1024^3: 3.39
        // against 3.77 ms).
        // In between (768^3: 1152 tiles) the two-node kernel with 12 compute warps (tiles of 24 planes x 32 lanes) is the fastest:
        // 1.60 ms against 1.70 (four nodes) and 1.75 (two nodes, 8 warps); at 512^3 it ties with 8 warps, at 896^3 (1568 tiles) and
        // 1024^3 it is 2 % behind the four-node kernel.
        const long long tiles2 = (long long)((w.nu + 15) / 16) * (d_.kpad / 32);
        const int nodes = tile_opt_.nodes ? tile_opt_.nodes : (tiles2 >= 1400 ? 4 : 2);
        TileOptions topt = tile_opt_;
        if (!warps_set_ && !tile_opt_.nodes && tiles2 >= 1000 && tiles2 < 1400) topt.warps = 12;
        if (kernel == TTCR_B200_KERNEL_MARCH && nodes == 4) {
            const int nl = march4_sweep<T>(s.tile, s.march4, topt, sm_count_, w, d_, tt, slo_[w.layout], s.mask[w.layout], fb,
                                          g_.dx, s.d_change, s.stream);
            s.st.launches += nl; s.st.sweep_launches += nl;
            return;
        }
        if (kernel == TTCR_B200_KERNEL_MARCH) {
            const int nl = march_sweep<T>(s.tile, s.march, topt, sm_count_, w, d_, tt, slo_[w.layout], s.mask[w.layout], fb,
                                         g_.dx, s.d_change, s.stream);
            s.st.launches += nl; s.st.sweep_launches += nl;
            return;
        }
        if (kernel == TTCR_B200_KERNEL_TILE) {
            const int nl = tile_sweep<T>(s.tile, tile_opt_, sm_count_, w, d_, tt, slo_[w.layout], s.mask[w.layout], fb,
                                        g_.dx, weno_stage, s.d_change, s.stream);
            s.st.launches += nl; s.st.sweep_launches += nl;
            return;
        }
        if (kernel == TTCR_B200_KERNEL_COOP) {
            // all planes of the direction in one cooperative launch with a grid-wide barrier per plane (kernels.cuh)
            const int np = w.nu + w.nm - 1;
            const int max_blocks = (d_.kpad / 32) * ((std::min(w.nu, w.nm) + 7) / 8);
            int& occ = coop_occ_[weno_stage ? 1 : 0];
            if (!occ) {
                if (weno_stage) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_planes_coop<T, true>, 256, 0));
                else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_planes_coop<T, false>, 256, 0));
                if (occ < 1) throw Err(TTCR_B200_ERR_CUDA, "k_sweep_planes_coop does not fit on an SM");
            }
            // About one CTA per block of the widest plane (half of a plane's blocks hold no node: the sheared rows), at most
            // coop_ctas_ per SM: the barrier costs one atomic per CTA.  CTA c takes blocks c, c + G, ...: G is made odd so
            // that a CTA does not always land on the same lane chunk (chunks near the row ends are mostly empty).
            const int per_sm = std::max(1, std::min(std::min(occ, coop_ctas_), (max_blocks + sm_count_ - 1) / sm_count_));
            int grid = std::max(1, std::min(max_blocks, per_sm * sm_count_));
            if (grid > 1 && grid < max_blocks && grid % 2 == 0) --grid;
            CK(cudaMemsetAsync(s.d_bar, 0, sizeof(unsigned), s.stream));
            ensure_ppart(s, (size_t)grid);
            SweepView wv = w;
            Dims dv = d_;
            const T* slo = slo_[w.layout];
            const uint32_t* mask = s.mask[w.layout];
            const FrozenBox* fbp = s.d_fb;
            T dx = g_.dx;
            double* chg = s.d_ppart;
            unsigned* bar = s.d_bar;
            void* args[] = {&wv, &dv, &tt, &slo, &mask, &fbp, &dx, &chg, &bar};
            const void* fn = weno_stage ? (const void*)k_sweep_planes_coop<T, true> : (const void*)k_sweep_planes_coop<T, false>;
            CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(32, 8), args, 0, s.stream));
            (void)np;
            k_sum_partials<<<1, 256, 0, s.stream>>>(s.d_ppart, grid, s.d_change);
            s.st.launches += 2; s.st.sweep_launches += 1;
            return;
        }
        // One launch per wavefront plane (the OpenCL design, Grid3Drn_OpenCL.h:839-848).  Launch bound: the np launches of
        // a direction are captured once per slot into a CUDA graph (their arguments do not depend on the source) and
        // replayed, optionally chained by programmatic dependent launch so that plane p+1 is resident when p drains.
        const int np = w.nu + w.nm - 1;
        // one slot of s.d_ppart per block and plane (offsets do not depend on the source or the direction: they are part of the
        // captured graph); the buffer is sized once per slot, before any capture
        size_t nparts = 0;
        for (int p = 0; p < np; ++p) {
            const int u_lo = std::max(0, p - w.nm + 1), u_hi = std::min(w.nu - 1, p);
            nparts += (size_t)(d_.kpad / 32) * ((u_hi - u_lo + 1 + 7) / 8);
        }
        ensure_ppart(s, nparts);
        auto launch_all = [&]() {
            const dim3 block(32, 8);
            const T* slo = slo_[w.layout];
            const uint32_t* mask = s.mask[w.layout];
            const FrozenBox* fbp = s.d_fb;
            T dx = g_.dx;
            double* chg = s.d_ppart;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            for (int p = 0; p < np; ++p) {
                int u_lo = std::max(0, p - w.nm + 1), u_hi = std::min(w.nu - 1, p);
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(d_.kpad / 32, (u_hi - u_lo + 1 + 7) / 8);
                cfg.blockDim = block;
                cfg.dynamicSmemBytes = 0;
                cfg.stream = s.stream;
                cfg.attrs = at;
                cfg.numAttrs = (plane_pdl_ && p > 0) ? 1 : 0;
                if (weno_stage)
                    CK(cudaLaunchKernelEx(&cfg, k_sweep_plane<T, true>, w, d_, tt, slo, mask, fbp, p, u_lo, u_hi, dx, chg));
                else
                    CK(cudaLaunchKernelEx(&cfg, k_sweep_plane<T, false>, w, d_, tt, slo, mask, fbp, p, u_lo, u_hi, dx, chg));
                chg += (size_t)cfg.gridDim.x * cfg.gridDim.y;
            }
            k_sum_partials<<<1, 256, 0, s.stream>>>(s.d_ppart, (int)nparts, s.d_change);
        };
        if (plane_graph_) {
            auto& pg = s.pgraph[dir][weno_stage ? 1 : 0];
            if (pg.exec && pg.pdl != plane_pdl_) { cudaGraphExecDestroy(pg.exec); pg.exec = nullptr; }
            if (!pg.exec) {
                cudaGraph_t gr = nullptr;
                CK(cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
                try {
                    launch_all();
                } catch (...) {
                    cudaStreamEndCapture(s.stream, &gr);
                    if (gr) cudaGraphDestroy(gr);
                    throw;
                }
                CK(cudaStreamEndCapture(s.stream, &gr));
                const cudaError_t e = cudaGraphInstantiate(&pg.exec, gr, 0);
                cudaGraphDestroy(gr);
                CK(e);
                pg.pdl = plane_pdl_;
            }
            CK(cudaGraphLaunch(pg.exec, s.stream));
        } else {
            launch_all();
        }
        s.st.launches += np + 1; s.st.sweep_launches += np;
    }

    int plane_kernel() const {
        // fp32: the marching kernel's WENO variant (sweep_march_weno.cuh), measured faster than the plane kernels at every
        // size from 7 x 6 x 5 to 512^3 (3.4x at 512^3); fp64: plane kernels
        if ((weno_kernel_ == TTCR_B200_KERNEL_MARCH || weno_kernel_ == TTCR_B200_KERNEL_AUTO) && marchw_supported<T>())
            return TTCR_B200_KERNEL_MARCH;
        if (weno_kernel_ != TTCR_B200_KERNEL_AUTO && weno_kernel_ != TTCR_B200_KERNEL_MARCH) return weno_kernel_;
        const int widest = (d_.kpad / 32) * ((std::min(d_.ni, d_.q) + 7) / 8);
        return widest >= 4 * sm_count_ ? TTCR_B200_KERNEL_COOP : TTCR_B200_KERNEL_PLANE;
    }

    int pick_kernel(bool weno_stage) const {
        const int weno_kernel_ = plane_kernel();
        if (kernel_ != TTCR_B200_KERNEL_AUTO) {
            if (kernel_ == TTCR_B200_KERNEL_MARCH && !march_supported<T>(weno_stage))
                return tile_supported<T>(weno_stage) ? TTCR_B200_KERNEL_TILE : weno_kernel_;
            if (kernel_ == TTCR_B200_KERNEL_TILE && !tile_supported<T>(weno_stage)) return weno_kernel_;
            return kernel_;
        }
        if (march_supported<T>(weno_stage)) return TTCR_B200_KERNEL_MARCH;   // fp32, first order: the marching kernel (sweep_march.cuh)
        if (!tile_supported<T>(weno_stage)) return weno_kernel_;                // WENO stage: plane kernels
        return TTCR_B200_KERNEL_TILE;                                           // fp64, first order
    }

    // Grid3Drnfs::raytrace body (Grid3Drnfs.h:92-154) on the device
    void solve_device(Slot& s, const std::vector<T>& vtx, const std::vector<T>& vt0) {
        const size_t ntx = vt0.size();
        const int npts = weno_ ? 2 : 1;
        const size_t ne = d_.elems();
        s.st = ttcr_b200_stats{};
        ensure_pts(s, 4 * ntx);
        std::memcpy(s.h_pts, vtx.data(), 3 * ntx * sizeof(T));
        std::memcpy(s.h_pts + 3 * ntx, vt0.data(), ntx * sizeof(T));
        const FrozenBox fb = frozen_box(vtx, npts);

        s.sweep_ev_used = 0;
        CK(cudaEventRecord(s.e0, s.stream));
        CK(cudaMemcpyAsync(s.d_fb, &fb, sizeof(FrozenBox), cudaMemcpyHostToDevice, s.stream));   // (pageable: staged before the call returns)
        CK(cudaMemcpyAsync(s.d_pts, s.h_pts, 4 * ntx * sizeof(T), cudaMemcpyHostToDevice, s.stream));
        // Node3Dn::reinit: every node +MAX; the slots that are no node hold +MAX anyway, so this is a plain fill
        k_fill16<T><<<nblocks(ne / (16 / sizeof(T)), 256, 148 * 32), 256, 0, s.stream>>>(s.tt[0], ne, Lim<T>::max());
        // frozen bits: only the previous source's box can hold any (the masks were zeroed when the slot was made)
        if (s.have_prev_fb) {
            const FrozenBox& pb = s.prev_fb;
            const long long vol = (long long)(pb.ihi - pb.ilo + 1) * (pb.jhi - pb.jlo + 1) * (pb.khi - pb.klo + 1);
            if (vol > 0 && vol <= (long long)d_.nodes() / 64) {
                k_clear_frozen_box<<<(unsigned)std::min<long long>((vol + 255) / 256, 1184), 256, 0, s.stream>>>(d_, pb, s.mask[0], s.mask[1]);
            } else if (vol > 0) {
                CK(cudaMemsetAsync(s.mask[0], 0, ne / 8, s.stream));
                CK(cudaMemsetAsync(s.mask[1], 0, ne / 8, s.stream));
            }
        }
        s.prev_fb = fb; s.have_prev_fb = true;
        k_init_fsm<T><<<1, 32, 0, s.stream>>>(g_, d_, s.d_pts, s.d_pts + 3 * ntx, (int)ntx, npts, s.tt[0], slo_[0], s.mask[0],
                                              s.mask[1]);
        CK(cudaGetLastError());
        s.st.launches += 2;   // k_reinit_l1, k_init_fsm

        int cur = 0;   // layout the field currently lives in
        float sweep_ms = 0.f;
        for (int stage = 0; stage < (weno_ ? 2 : 1); ++stage) {
            const bool wstage = stage == 1;
            const int kernel = pick_kernel(wstage);
            s.st.kernel = kernel;
            int it = 0;
            T change = Lim<T>::max();
            while (change >= epsilon_ && it < maxit_) {
                CK(cudaMemsetAsync(s.d_change, 0, sizeof(double), s.stream));
                static const int dbg_dirs = getenv("TTCR_B200_DEBUG_DIRS") ? atoi(getenv("TTCR_B200_DEBUG_DIRS")) : 255;
                for (int dir = 0; dir < 8; ++dir) {
                    if (!((dbg_dirs >> dir) & 1)) continue;   // debugging aid: run a subset of the sweep directions
                    const int want = make_view(d_, dir).layout;
                    if (want != cur) {
                        // (a shared-memory version that reads whole row segments was measured 50 % slower than these gathers
                        // through the L1: 0.43 ms instead of 0.28 ms per relayout at 512^3)
                        static const bool old_relayout = getenv("TTCR_B200_OLD_RELAYOUT") != nullptr;
                        if (old_relayout) {
                            const dim3 grid((d_.nk + 31) / 32, (d_.nj + 31) / 32, d_.ni);
                            k_relayout<T><<<grid, dim3(32, 8), 0, s.stream>>>(s.tt[cur], cur, s.tt[want], d_);
                        } else {
                            constexpr int RB = 64;   // (measured at 512^3: 64 rows per block 12.35 ms per solve, 32: 12.49, 128: 12.45-12.51, 256: 12.59)
                            const dim3 grid(d_.kpad / 32, (d_.q + RB - 1) / RB, d_.ni);
                            k_relayout2<T, RB><<<grid, dim3(32, 8), 0, s.stream>>>(s.tt[cur], cur, s.tt[want], d_);
                        }
                        s.st.launches += 1;
                        cur = want;
                    }
                    cudaEvent_t ea = nullptr, eb = nullptr;
                    if (s.sweep_ev_used + 2 <= kMaxSweepEvents) {
                        while (s.sweep_ev.size() < s.sweep_ev_used + 2) {
                            cudaEvent_t e;
                            CK(cudaEventCreate(&e));
                            s.sweep_ev.push_back(e);
                        }
                        ea = s.sweep_ev[s.sweep_ev_used]; eb = s.sweep_ev[s.sweep_ev_used + 1];
                        s.sweep_ev_used += 2;
                        CK(cudaEventRecord(ea, s.stream));
                    }
                    launch_sweep(s, dir, wstage, fb, kernel);
                    if (eb) CK(cudaEventRecord(eb, s.stream));
                    s.st.sweeps += 1;
                }
                CK(cudaGetLastError());
                CK(cudaMemcpyAsync(s.h_change, s.d_change, sizeof(double), cudaMemcpyDeviceToHost, s.stream));
                CK(cudaStreamSynchronize(s.stream));
                tile_check(s.tile);
                const double c = *s.h_change;
                s.st.last_change = c;
                // the reference accumulates `change` in T; an overflowing double sum maps to T's inf
                change = c > (double)Lim<T>::max() ? std::numeric_limits<T>::infinity() : (T)c;
                ++it;
            }
            if (wstage) s.st.niterw = it; else s.st.niter = it;
        }
        if (cur != 0) {   // only reachable with TTCR_B200_DEBUG_DIRS
            const dim3 grid((d_.nk + 31) / 32, (d_.nj + 31) / 32, d_.ni);
            k_relayout<T><<<grid, dim3(32, 8), 0, s.stream>>>(s.tt[cur], cur, s.tt[0], d_);
            cur = 0;
        }
        CK(cudaEventRecord(s.e1, s.stream));
        CK(cudaEventSynchronize(s.e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s.e0, s.e1));
        s.st.solve_ms = ms;
        for (size_t i = 0; i + 1 < s.sweep_ev_used; i += 2) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, s.sweep_ev[i], s.sweep_ev[i + 1]));
            sweep_ms += t;
        }
        if (s.sweep_ev_used < (size_t)2 * s.st.sweeps && s.sweep_ev_used > 0)   // more sweeps than events: extrapolate
            sweep_ms *= (float)(2.0 * s.st.sweeps / (double)s.sweep_ev_used);
        s.st.sweep_ms = sweep_ms;
    }

    static constexpr size_t kMaxSweepEvents = 2 * 8 * 64;
    Dims d_{};
    Geom<T> g_{};
    T origin_[3];
    T epsilon_;
    int maxit_;
    bool weno_, ttrp_, cell_, translate_, intvel_;
    bool have_slowness_ = false;
    int m_terms_ = 0;   // raytrace_rays also produces the raw terms of the matrix M (get_m_terms): 1 / 2 = which overload
    cudaStream_t copy_st_ = nullptr;   // H2D leg of the pipelined model import
    cudaEvent_t copy_ev_[9] = {};
    // plane-per-launch sweeps (WENO stage, small grids): replay a captured graph / chain the planes by PDL
    bool plane_graph_ = getenv("TTCR_B200_PLANE_GRAPH") ? atoi(getenv("TTCR_B200_PLANE_GRAPH")) != 0 : true;
    bool plane_pdl_ = getenv("TTCR_B200_PLANE_PDL") ? atoi(getenv("TTCR_B200_PLANE_PDL")) != 0 : true;
    // plane kernel of the WENO stage: AUTO = one cooperative launch per direction where the planes are wide (>= 4 blocks per
    // SM at the widest: 512^3 and up, measured 13 % faster), graph-replayed plane launches otherwise (measured 15-25 % faster)
    int weno_kernel_ = getenv("TTCR_B200_WENO_KERNEL") ? atoi(getenv("TTCR_B200_WENO_KERNEL")) : TTCR_B200_KERNEL_AUTO;
    int coop_ctas_ = 4;          // CTAs per SM of the cooperative plane kernel, at most
    int coop_occ_[2] = {0, 0};
    int dev_ = 0, sm_count_ = 148;
    int kernel_ = TTCR_B200_KERNEL_AUTO;
    TileOptions tile_opt_{};
    bool warps_set_ = false;   // "tile_warps" given: no choice by grid size
    T* slo_[2] = {nullptr, nullptr};   // node slowness in layouts L1, L2
    T* lin_ = nullptr;   // cell grids: the averaged node model in host order (the other staging buffer is a slot's idle array)
    std::mutex lin_mu_;
    std::vector<Slot> slots_;
    size_t bytes_ = 0;
};

// ---- 2-D twins (grid2d.cuh) ------------------------------------------------------------------------
struct Grid2Base {
    virtual ~Grid2Base() {}
    virtual void set_slowness(const void* s, size_t n) = 0;
    virtual void get_slowness(void* out) = 0;
    virtual void raytrace_multi(size_t ns, const size_t* txo, const void* tx, const void* t0, const size_t* rxo, const void* rx, void* out,
                                int* niter, size_t first_slot) = 0;
    virtual void get_tt(void* out, size_t slot) = 0;
    virtual void get_niter(size_t slot, int* a, int* b) = 0;
    virtual size_t n_slots() const = 0;
    virtual double last_ms() const = 0;
};

template <typename T>
class Grid2 final : public Grid2Base {
public:
    Grid2(uint32_t ncx, uint32_t ncz, double dx, double dz, double xmin, double zmin, double eps, int maxit, bool weno, bool rotated,
          size_t nslots, bool cell, int dev)
        : ncx_((int)ncx), ncz_((int)ncz), cell_(cell), dev_(dev), nslots_(std::max<size_t>(nslots, 1)) {
        if (ncx == 0 || ncz == 0) throw Err(TTCR_B200_ERR_INVALID, "grid needs at least one cell per axis");
        CK(cudaSetDevice(dev_));
        N_ = (size_t)(ncx + 1) * (ncz + 1);
        p_.ncx = ncx_; p_.ncz = ncz_;
        p_.dx = (T)dx; p_.dz = (T)dz; p_.xmin = (T)xmin; p_.zmin = (T)zmin;
        p_.eps_total = (T)eps;
        p_.eps_total *= static_cast<T>(N_);          // Grid2Drnfs.h:92
        p_.maxit = maxit; p_.weno = weno; p_.rotated = rotated;
        CK(cudaMalloc(&d_s_, N_ * sizeof(T)));
        CK(cudaMalloc(&d_tt_, nslots_ * N_ * sizeof(T)));
        CK(cudaMalloc(&d_frozen_, nslots_ * N_));
        CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0_)); CK(cudaEventCreate(&e1_));
        niter_.assign(2 * nslots_, 0);
    }
    ~Grid2() override {
        cudaSetDevice(dev_);
        cudaFree(d_s_); cudaFree(d_tt_); cudaFree(d_frozen_);
        cudaEventDestroy(e0_); cudaEventDestroy(e1_);
        cudaStreamDestroy(st_);
    }
    size_t n_slots() const override { return nslots_; }
    double last_ms() const override { return last_ms_; }

    void set_slowness(const void* s, size_t n) override {
        CK(cudaSetDevice(dev_));
        const size_t want = cell_ ? (size_t)ncx_ * ncz_ : N_;
        if (n != want) throw Err(TTCR_B200_ERR_LENGTH, "Error: slowness vectors of incompatible size.");
        if (cell_) {
            ScratchBuf c;
            c.alloc(n * sizeof(T), st_);
            CK(cudaMemcpyAsync(c.p, s, n * sizeof(T), cudaMemcpyHostToDevice, st_));
            k2d_cell_to_node<T><<<nblocks(N_), 256, 0, st_>>>(c.as<T>(), ncx_, ncz_, d_s_);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st_));
        } else {
            CK(cudaMemcpyAsync(d_s_, s, n * sizeof(T), cudaMemcpyHostToDevice, st_));
            CK(cudaStreamSynchronize(st_));
        }
        have_s_ = true;
    }
    void get_slowness(void* out) override {
        CK(cudaSetDevice(dev_));
        CK(cudaMemcpy(out, d_s_, N_ * sizeof(T), cudaMemcpyDeviceToHost));
    }
    void get_tt(void* out, size_t slot) override {
        if (slot >= nslots_) throw Err(TTCR_B200_ERR_INVALID, "Thread number is larger than number of threads");
        CK(cudaSetDevice(dev_));
        CK(cudaMemcpy(out, d_tt_ + slot * N_, N_ * sizeof(T), cudaMemcpyDeviceToHost));
    }
    void get_niter(size_t slot, int* a, int* b) override {
        if (slot >= nslots_) throw Err(TTCR_B200_ERR_INVALID, "Thread number is larger than number of threads");
        *a = niter_[2 * slot]; *b = niter_[2 * slot + 1];
    }

    // sources n = 0 .. ns-1 (Tx points txo[n] .. txo[n+1], receivers rxo[n] .. rxo[n+1]) in rounds of up to n_slots - first_slot
    // sources per launch, source n of a round on slot first_slot + (n - round start): the fan-out of Grid2D::raytrace over a
    // vector of sources (one slot = one of the reference's threadNo)
    void raytrace_multi(size_t ns, const size_t* txo, const void* tx, const void* t0, const size_t* rxo, const void* rx, void* out, int* niter,
                        size_t first_slot) override {
        CK(cudaSetDevice(dev_));
        if (!have_s_) throw Err(TTCR_B200_ERR_LOGIC, "slowness has not been set");
        if (first_slot >= nslots_) throw Err(TTCR_B200_ERR_INVALID, "Thread number is larger than number of threads");
        if (ns == 0) return;
        int ntx_max = 0;
        for (size_t n = 0; n < ns; ++n) {
            if (txo[n + 1] <= txo[n]) throw Err(TTCR_B200_ERR_INVALID, "source has no Tx point");
            ntx_max = std::max(ntx_max, (int)(txo[n + 1] - txo[n]));
        }
        // pack the Tx points per source
        std::vector<T> htx((size_t)ns * ntx_max * 2, T(0)), ht0((size_t)ns * ntx_max, T(0));
        std::vector<int> hn(ns);
        for (size_t n = 0; n < ns; ++n) {
            hn[n] = (int)(txo[n + 1] - txo[n]);
            std::memcpy(&htx[n * ntx_max * 2], (const T*)tx + 2 * txo[n], (size_t)hn[n] * 2 * sizeof(T));
            std::memcpy(&ht0[n * ntx_max], (const T*)t0 + txo[n], (size_t)hn[n] * sizeof(T));
        }
        const size_t nrx = rxo[ns];
        ScratchBuf btx, bt0, bn, bit, berr, brx, bout;
        btx.alloc(htx.size() * sizeof(T), st_); bt0.alloc(ht0.size() * sizeof(T), st_); bn.alloc(ns * sizeof(int), st_);
        bit.alloc(2 * ns * sizeof(int), st_); berr.alloc(ns * sizeof(int), st_);
        brx.alloc(std::max<size_t>(nrx, 1) * 2 * sizeof(T), st_); bout.alloc(std::max<size_t>(nrx, 1) * sizeof(T), st_);
        CK(cudaMemcpyAsync(btx.p, htx.data(), htx.size() * sizeof(T), cudaMemcpyHostToDevice, st_));
        CK(cudaMemcpyAsync(bt0.p, ht0.data(), ht0.size() * sizeof(T), cudaMemcpyHostToDevice, st_));
        CK(cudaMemcpyAsync(bn.p, hn.data(), ns * sizeof(int), cudaMemcpyHostToDevice, st_));
        if (nrx) CK(cudaMemcpyAsync(brx.p, rx, nrx * 2 * sizeof(T), cudaMemcpyHostToDevice, st_));
        CK(cudaMemsetAsync(bit.p, 0, 2 * ns * sizeof(int), st_));
        P2<T> p = p_;
        p.s = d_s_; p.tx = btx.as<T>(); p.t0 = bt0.as<T>(); p.ntx = bn.as<int>(); p.ntx_max = ntx_max;
        p.niter = bit.as<int>(); p.err = berr.as<int>();
        const size_t per = nslots_ - first_slot;
        CK(cudaEventRecord(e0_, st_));
        for (size_t n0 = 0; n0 < ns; n0 += per) {
            const size_t nb = std::min(per, ns - n0);
            p.tt = d_tt_ + first_slot * N_;
            p.frozen = d_frozen_ + first_slot * N_;
            // wide grids: a cluster of CTAs per source (grid2d.cuh); narrow ones: one CTA (a wavefront fits its threads and
            // its L1, and more sources run at a time)
            const int cl_env = getenv("TTCR_B200_2D_CLUSTER") ? atoi(getenv("TTCR_B200_2D_CLUSTER")) : -1;   // (read per call: the tests run both kernels)
            const bool use_cluster = cl_env >= 0 ? cl_env != 0 : std::min(ncx_, ncz_) + 1 > 1024;   // (measured: 1001^2 0.74 s with one CTA, 0.84 s with a cluster; 2001^2 3.13 s / 1.99 s)
            if (use_cluster) {
                cudaLaunchConfig_t cfg{};
                static const int cs_env = getenv("TTCR_B200_2D_CLUSTER_SIZE") ? atoi(getenv("TTCR_B200_2D_CLUSTER_SIZE")) : K2D_CLUSTER;
                static const int bs_env = getenv("TTCR_B200_2D_BLOCK") ? atoi(getenv("TTCR_B200_2D_BLOCK")) : 256;
                const int cs = std::max(1, std::min(cs_env, K2D_CLUSTER));
                cfg.gridDim = dim3((unsigned)nb * cs);
                cfg.blockDim = dim3(std::max(32, std::min(bs_env, 1024)));
                cfg.stream = st_;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                CK(cudaLaunchKernelEx(&cfg, k2d_solve<T, true>, p, (int)n0));
            } else {
                k2d_solve<T, false><<<(unsigned)nb, 1024, 0, st_>>>(p, (int)n0);
            }
            for (size_t b = 0; b < nb; ++b) {
                const size_t n = n0 + b, m = rxo[n + 1] - rxo[n];
                if (m) k2d_interp<T><<<(unsigned)((m + 127) / 128), 128, 0, st_>>>(p, d_tt_ + (first_slot + b) * N_, brx.as<T>() + 2 * rxo[n], (int)m,
                                                                                    bout.as<T>() + rxo[n]);
            }
        }
        CK(cudaEventRecord(e1_, st_));
        CK(cudaGetLastError());
        std::vector<int> hit(2 * ns), herr(ns);
        CK(cudaMemcpyAsync(hit.data(), bit.p, 2 * ns * sizeof(int), cudaMemcpyDeviceToHost, st_));
        CK(cudaMemcpyAsync(herr.data(), berr.p, ns * sizeof(int), cudaMemcpyDeviceToHost, st_));
        if (nrx) CK(cudaMemcpyAsync(out, bout.p, nrx * sizeof(T), cudaMemcpyDeviceToHost, st_));
        CK(cudaStreamSynchronize(st_));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0_, e1_));
        last_ms_ = ms;
        for (size_t n = 0; n < ns; ++n)
            if (herr[n]) throw Err(TTCR_B200_ERR_RUNTIME, "Error: Point outside grid.");   // checkPts, Grid2Drn.h:333-342
        for (size_t n = 0; n < ns; ++n) {
            const size_t slot = first_slot + n % per;
            niter_[2 * slot] = hit[2 * n]; niter_[2 * slot + 1] = hit[2 * n + 1];
        }
        if (niter) std::memcpy(niter, hit.data(), 2 * ns * sizeof(int));
    }

private:
    int ncx_, ncz_;
    bool cell_, have_s_ = false;
    int dev_;
    size_t nslots_, N_ = 0;
    P2<T> p_{};
    T* d_s_ = nullptr;
    T* d_tt_ = nullptr;
    unsigned char* d_frozen_ = nullptr;
    cudaStream_t st_ = nullptr;
    cudaEvent_t e0_ = nullptr, e1_ = nullptr;
    std::vector<int> niter_;
    double last_ms_ = 0.0;
};

}  // namespace ttcrb200

// =============================================================================================
// C ABI
// =============================================================================================
using namespace ttcrb200;

struct ttcr_b200_grid {
    GridBase* impl = nullptr;
};
struct ttcr_b200_grid2d {
    ttcrb200::Grid2Base* impl = nullptr;
};

static thread_local std::string g_last_error;

template <typename F>
static int guard(F&& f) {
    try {
        f();
        return TTCR_B200_OK;
    } catch (const Err& e) {
        g_last_error = e.msg;
        return e.code;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return TTCR_B200_ERR_RUNTIME;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return TTCR_B200_ERR_RUNTIME;
    }
}

#define NEED(g)                                                    \
    if (!(g) || !(g)->impl) {                                      \
        g_last_error = "null ttcr_b200_grid handle";               \
        return TTCR_B200_ERR_INVALID;                              \
    }

extern "C" {

const char* ttcr_b200_version(void) { return "ttcr_b200 0.1 (sm_100a)"; }

const char* ttcr_b200_last_error(const ttcr_b200_grid*) { return g_last_error.c_str(); }

int ttcr_b200_create(ttcr_b200_grid** out, uint32_t nx, uint32_t ny, uint32_t nz, double dx, double xmin, double ymin,
                     double zmin, double eps, int maxit, int weno, int tt_from_rp, int interp_vel, size_t n_slots,
                     int translate_origin, int cell_slowness, int dtype, int device) {
    if (!out) { g_last_error = "null output pointer"; return TTCR_B200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Err(TTCR_B200_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e) +
                                              " (ttcr_b200 has no CPU path)");
        GridBase* impl;
        if (dtype == TTCR_B200_F64)
            impl = new Grid<double>(nx, ny, nz, dx, xmin, ymin, zmin, eps, maxit, weno != 0, tt_from_rp != 0, interp_vel != 0,
                                    n_slots, translate_origin != 0, cell_slowness != 0, device);
        else if (dtype == TTCR_B200_F32)
            impl = new Grid<float>(nx, ny, nz, dx, xmin, ymin, zmin, eps, maxit, weno != 0, tt_from_rp != 0, interp_vel != 0,
                                   n_slots, translate_origin != 0, cell_slowness != 0, device);
        else
            throw Err(TTCR_B200_ERR_INVALID, "dtype must be TTCR_B200_F64 or TTCR_B200_F32");
        *out = new ttcr_b200_grid{impl};
    });
}

void ttcr_b200_destroy(ttcr_b200_grid* g) {
    if (!g) return;
    delete g->impl;
    delete g;
}

int ttcr_b200_set_slowness(ttcr_b200_grid* g, const void* s, size_t n, int order) {
    NEED(g);
    return guard([&] { g->impl->set_slowness(s, n, order); });
}
int ttcr_b200_set_slowness_device_planes(ttcr_b200_grid* g, const void* s, size_t n, int i_first, int i_count) {
    NEED(g);
    return guard([&] { g->impl->set_slowness_device_planes(s, n, i_first, i_count); });
}

int ttcr_b200_set_slowness_device(ttcr_b200_grid* g, const void* s, size_t n, int order) {
    NEED(g);
    return guard([&] { g->impl->set_slowness_device(s, n, order); });
}
int ttcr_b200_get_tt_device(ttcr_b200_grid* g, void* out, size_t slot, int order) {
    NEED(g);
    return guard([&] { g->impl->get_tt_device(out, slot, order); });
}
int ttcr_b200_get_slowness(ttcr_b200_grid* g, void* out, int order) {
    NEED(g);
    return guard([&] { g->impl->get_slowness(out, order); });
}
int ttcr_b200_raytrace(ttcr_b200_grid* g, const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt,
                       size_t slot) {
    NEED(g);
    return guard([&] { g->impl->raytrace(tx, t0, ntx, rx, nrx, tt, slot); });
}

int ttcr_b200_raytrace_rays(ttcr_b200_grid* g, const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt,
                            size_t* ray_npts, size_t slot) {
    NEED(g);
    return guard([&] {
        if (nrx && !ray_npts) throw Err(TTCR_B200_ERR_INVALID, "ray_npts must not be NULL");
        size_t dummy = 0;
        g->impl->raytrace_rays(tx, t0, ntx, rx, nrx, tt, ray_npts ? ray_npts : &dummy, slot);
    });
}

int ttcr_b200_get_m_terms(ttcr_b200_grid* g, size_t slot, unsigned long long* node_out, void* value_out) {
    if (!g) return TTCR_B200_ERR_INVALID;
    return guard([&] { g->impl->get_m_terms(slot, node_out, value_out); });
}

int ttcr_b200_get_rays(ttcr_b200_grid* g, size_t slot, void* xyz_out) {
    NEED(g);
    return guard([&] { g->impl->get_rays(slot, xyz_out); });
}
int ttcr_b200_raytrace_multi(ttcr_b200_grid* g, size_t nsrc, const size_t* tx_off, const void* tx, const void* t0,
                             const size_t* rx_off, const void* rx, void* tt, int* niter, int* niterw) {
    NEED(g);
    return guard([&] { g->impl->raytrace_multi(nsrc, tx_off, tx, t0, rx_off, rx, tt, niter, niterw); });
}
int ttcr_b200_get_tt(ttcr_b200_grid* g, void* out, size_t slot, int order) {
    NEED(g);
    return guard([&] { g->impl->get_tt(out, slot, order); });
}
int ttcr_b200_get_niter(ttcr_b200_grid* g, size_t slot, int* niter, int* niterw) {
    NEED(g);
    return guard([&] {
        ttcr_b200_stats st;
        g->impl->stats(slot, &st);
        if (niter) *niter = st.niter;
        if (niterw) *niterw = st.niterw;
    });
}
int ttcr_b200_set_option(ttcr_b200_grid* g, const char* key, double value) {
    NEED(g);
    return guard([&] { g->impl->set_option(key ? key : "", value); });
}
size_t ttcr_b200_n_slots(const ttcr_b200_grid* g) { return g && g->impl ? g->impl->n_slots() : 0; }
int ttcr_b200_solve(ttcr_b200_grid* g, const void* tx, const void* t0, size_t ntx, size_t slot) {
    NEED(g);
    return guard([&] { g->impl->solve(tx, t0, ntx, slot); });
}
int ttcr_b200_get_stats(ttcr_b200_grid* g, size_t slot, ttcr_b200_stats* out) {
    NEED(g);
    if (!out) { g_last_error = "null output pointer"; return TTCR_B200_ERR_INVALID; }
    return guard([&] { g->impl->stats(slot, out); });
}
size_t ttcr_b200_device_bytes(const ttcr_b200_grid* g) { return g && g->impl ? g->impl->device_bytes() : 0; }


/* ---- 2-D twins ---------------------------------------------------------------------------------- */
int ttcr_b200_create2d(ttcr_b200_grid2d** out, uint32_t nx, uint32_t nz, double dx, double dz, double xmin, double zmin, double eps,
                       int maxit, int weno, int rotated_template, size_t n_slots, int cell_slowness, int dtype, int device) {
    if (!out) { g_last_error = "null output pointer"; return TTCR_B200_ERR_INVALID; }
    *out = nullptr;
    return guard([&] {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw Err(TTCR_B200_ERR_CUDA, "no CUDA device: ttcr_b200 has no CPU path");
        int dev = device;
        if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
        if (dev >= ndev) throw Err(TTCR_B200_ERR_INVALID, "device index out of range");
        std::unique_ptr<ttcr_b200_grid2d> g(new ttcr_b200_grid2d);
        if (dtype == TTCR_B200_F64) g->impl = new ttcrb200::Grid2<double>(nx, nz, dx, dz, xmin, zmin, eps, maxit, weno != 0, rotated_template != 0, n_slots, cell_slowness != 0, dev);
        else if (dtype == TTCR_B200_F32) g->impl = new ttcrb200::Grid2<float>(nx, nz, dx, dz, xmin, zmin, eps, maxit, weno != 0, rotated_template != 0, n_slots, cell_slowness != 0, dev);
        else throw Err(TTCR_B200_ERR_INVALID, "dtype must be TTCR_B200_F32 or TTCR_B200_F64");
        *out = g.release();
    });
}
void ttcr_b200_destroy2d(ttcr_b200_grid2d* g) {
    if (!g) return;
    delete g->impl;
    delete g;
}
int ttcr_b200_set_slowness2d(ttcr_b200_grid2d* g, const void* s, size_t n) {
    NEED(g);
    return guard([&] { g->impl->set_slowness(s, n); });
}
int ttcr_b200_get_slowness2d(ttcr_b200_grid2d* g, void* out) {
    NEED(g);
    return guard([&] { g->impl->get_slowness(out); });
}
int ttcr_b200_raytrace2d(ttcr_b200_grid2d* g, const void* tx, const void* t0, size_t ntx, const void* rx, size_t nrx, void* tt_out, size_t slot) {
    NEED(g);
    return guard([&] {
        const size_t txo[2] = {0, ntx}, rxo[2] = {0, nrx};
        g->impl->raytrace_multi(1, txo, tx, t0, rxo, rx, tt_out, nullptr, slot);
    });
}
int ttcr_b200_raytrace2d_multi(ttcr_b200_grid2d* g, size_t n_sources, const size_t* tx_off, const void* tx, const void* t0, const size_t* rx_off,
                               const void* rx, void* tt_out, int* niter_out) {
    NEED(g);
    return guard([&] { g->impl->raytrace_multi(n_sources, tx_off, tx, t0, rx_off, rx, tt_out, niter_out, 0); });
}
int ttcr_b200_get_tt2d(ttcr_b200_grid2d* g, void* out, size_t slot) {
    NEED(g);
    return guard([&] { g->impl->get_tt(out, slot); });
}
int ttcr_b200_get_niter2d(ttcr_b200_grid2d* g, size_t slot, int* niter, int* niterw) {
    NEED(g);
    return guard([&] { g->impl->get_niter(slot, niter, niterw); });
}
double ttcr_b200_last_solve_ms2d(const ttcr_b200_grid2d* g) { return (g && g->impl) ? g->impl->last_ms() : 0.0; }
}  // extern "C"
