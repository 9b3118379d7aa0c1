// Shared-memory / mailbox / mbarrier / TMA helpers of the marching sweep kernel (sweep_march.cuh), and the host-side
// tensor-map encoder.  PTX only: no library calls on the device side.
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
static __device__ __noinline__ bool frozen_bit(const uint32_t* frozen, long long e) { return (frozen[e >> 5] >> (e & 31)) & 1u; }

__device__ __forceinline__ int lds_i(unsigned a) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_i(unsigned a, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
template <int I> struct IntC { static constexpr int value = I; };

constexpr int round128(int x) { return (x + 127) / 128 * 128; }

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

// 3-D map over a sheared layout array whose plane stride is (rows per plane -+ 1) rows: a box arrives with its planes skewed by one
// row each, which is the lag between consecutive planes of a marching tile.
inline CUtensorMap make_skew_map(const void* base, const Dims& d, bool minus, int bw, int br, int bp, bool nan_fill = false) {
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


}  // namespace ttcrb200
