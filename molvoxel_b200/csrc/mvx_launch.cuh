// mvx_launch.cuh — host-side launch helpers shared by the API translation unit (mvx_api.cu) and the per-(mode, channel
// chunk, density) instantiation units of the voxelize kernels (mvx_vox_inst.cu, compiled once per combination so that the
// library builds in parallel).
#pragma once
#include <atomic>
#include <cuda_runtime.h>

#include "mvx_common.cuh"

namespace mvx {

// Voxelize kernel forms.  All give identical results; they differ in how a tile's atoms reach the warps.
//   CELLS  column entries (48 B) with cell masks, features gathered at staging: sparse / medium batches
//   TILES  entries regrouped per 16-voxel z layer with their feature rows, one flat staging copy,
//          per-layer warp filter: packed complexes (hundreds of atoms per column)
//   ROWS   generic lock-step form (any D, scalar stores when D % 4 != 0)
//   PIPE   the tile form made persistent: one CTA per SM walks the tiles, staging is a bulk copy (cp.async.bulk +
//          mbarrier) issued several tiles ahead, no CTA-wide barrier in the steady state: dense batches
enum Form { FORM_ROWS = 0, FORM_CELLS = 1, FORM_TILES = 3, FORM_PIPE = 4 };

// The dynamic shared-memory limit is a per-device function attribute: remember which devices have it (the kernels are
// template instantiations, so one DeviceSet per call site).  Lock-free: concurrent callers at worst set it twice.
struct DeviceSet {
    std::atomic<unsigned long long> bits[4] = {};   // devices 0..255
    bool test_and_set(int dev) {
        if (dev < 0 || dev > 255) return false;
        const unsigned long long bit = 1ull << (dev & 63);
        return (bits[dev >> 6].fetch_or(bit, std::memory_order_acq_rel) & bit) != 0ull;
    }
};

template <typename K>
cudaError_t set_smem(K kernel, size_t smem, DeviceSet* done = nullptr) {
    if (done != nullptr) {
        int dev = -1;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (done->test_and_set(dev)) return cudaSuccess;
    }
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// Launches the voxelize kernel(s) of `form` for one (mode, channel chunk, density) combination; the output element
// type (vp.out_kind) is dispatched inside.  Defined in mvx_vox_inst.cu, one explicit instantiation per object file.
template <int MODE, int CH, bool BINARY>
cudaError_t launch_form(const VoxParams& vp, int form, int nv, unsigned grid, cudaStream_t st);

}  // namespace mvx
