// Persistent tile-marching sweep kernel (placeholder: the PLANE kernel is the only sweep path
// until this file is filled in).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace ttcrb200 {

struct TileOptions {
    int chunk = 8;         // rows between progress-flag publications
    int ctas_per_sm = 0;   // 0 = occupancy query
    int warps = 8;         // warps (u rows) per tile
    long long spin_limit = 1ll << 26;
};
struct TileState {};

inline void tile_alloc(TileState&, const Dims&, size_t&) {}
inline void tile_free(TileState&) {}
inline void tile_check(TileState&) {}
template <typename T> inline bool tile_supported(bool) { return false; }
template <typename T>
inline int tile_sweep(TileState&, const TileOptions&, int, const SweepView&, const Dims&, T*, const T*, const uint32_t*,
                      const FrozenBox&, T, bool, double*, cudaStream_t) {
    return 0;
}

}  // namespace ttcrb200
