// mvx_vox_inst.cu — one explicit instantiation of the voxelize launchers, chosen by
//   -DMVX_INST_MODE={0,1,2} -DMVX_INST_CH={1,4,8,12,16} -DMVX_INST_BINARY={0,1}
// (molvoxel_b200/_lib.py compiles this file once per combination, in parallel, and links the objects with mvx_api.o).
#include "mvx_launch.cuh"
#include "mvx_vox_kernels.cuh"
#ifdef MVX_WITH_WS
#include "mvx_vox_ws.cuh"
#endif

namespace mvx {
namespace {

// CTAs of the persistent form: one per SM of the current device
static int pipe_grid(unsigned ntiles, unsigned* grid) {
    static std::atomic<int> sms[256] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 255) return -1;
    int n = sms[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) return -1;
        sms[dev].store(n, std::memory_order_relaxed);
    }
    const unsigned want = (unsigned)n;
    *grid = ntiles < want ? ntiles : want;
    return 0;
}

template <int MODE, int CH, bool BINARY, bool O16, bool CL>
cudaError_t launch_form_out(const VoxParams& vp, int form, int nv, unsigned grid, cudaStream_t st) {
    if (form == FORM_PIPE) {
        constexpr size_t smem = kPipeSmemBytes;
        static DeviceSet cfg, cfg_m, cfg_t;
        unsigned pg = 0;
        if (pipe_grid(grid, &pg) != 0) return cudaErrorInvalidDevice;
        bool ws_done = false;
#ifdef MVX_WITH_WS
        if constexpr (MODE == 2 && CH == 16) {   // warp-specialised instance (list-builder + accumulator warps)
            if (vp.ws_nb == 4 || vp.ws_nb == 8 || vp.ws_nb == 20) {
                static DeviceSet cfg_w[6];
                const bool multi = vp.pipe_q == ws_ring_q(ws_builders(vp.ws_nb), ws_walkers(vp.ws_nb), true);
                cudaError_t ew = cudaSuccess;
#define MVX_WS_LAUNCH(SLOT, MULTI_, NB_, NWK_, RB_, RW_)                                                                              \
    {                                                                                                                                \
        ew = set_smem(mvx_voxelize_ws_kernel<MODE, CH, BINARY, O16, MULTI_, NB_, NWK_, RB_, RW_>, smem, &cfg_w[SLOT]);                \
        if (ew == cudaSuccess) mvx_voxelize_ws_kernel<MODE, CH, BINARY, O16, MULTI_, NB_, NWK_, RB_, RW_><<<pg, (NB_ + NWK_) * 32, smem, st>>>(vp, grid); \
    }
                if (vp.ws_nb == 4) { if (multi) MVX_WS_LAUNCH(0, true, 4, 12, 56, 152) else MVX_WS_LAUNCH(1, false, 4, 12, 56, 152) }
                else if (vp.ws_nb == 8) { if (multi) MVX_WS_LAUNCH(2, true, 8, 8, 64, 192) else MVX_WS_LAUNCH(3, false, 8, 8, 64, 192) }
                else { if (multi) MVX_WS_LAUNCH(4, true, 8, 12, 48, 128) else MVX_WS_LAUNCH(5, false, 8, 12, 48, 128) }
#undef MVX_WS_LAUNCH
                if (ew != cudaSuccess) return ew;
                ws_done = true;
            }
        }
#endif
        if constexpr (CH == 16) {   // several channel chunks per cell (C > 16): hit-weight cache
            if (!ws_done && vp.pipe_q == pipe_ring_q(true)) {
                cudaError_t em = set_smem(mvx_voxelize_pipe_kernel<MODE, CH, BINARY, O16, true, CL>, smem, &cfg_m);
                if (em != cudaSuccess) return em;
                mvx_voxelize_pipe_kernel<MODE, CH, BINARY, O16, true, CL><<<pg, kPipeThreads, smem, st>>>(vp, grid);
            }
        }
        if (!ws_done && (CH != 16 || vp.pipe_q != pipe_ring_q(true))) {
            cudaError_t em = set_smem(mvx_voxelize_pipe_kernel<MODE, CH, BINARY, O16, false, CL>, smem, &cfg);
            if (em != cudaSuccess) return em;
            mvx_voxelize_pipe_kernel<MODE, CH, BINARY, O16, false, CL><<<pg, kPipeThreads, smem, st>>>(vp, grid);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        // tiles with more entries than the pipelined form takes (usually none): a small scanning grid
        constexpr size_t smem_t = tiles_smem_bytes<MODE>();
        e = set_smem(mvx_voxelize_sweep_kernel<MODE, CH, BINARY, O16, CL>, smem_t, &cfg_t);
        if (e != cudaSuccess) return e;
        mvx_voxelize_sweep_kernel<MODE, CH, BINARY, O16, CL><<<2 * pg < grid ? 2 * pg : grid, kThreads, smem_t, st>>>(vp, grid);
    } else if (form == FORM_TILES) {
        constexpr size_t smem = tiles_smem_bytes<MODE>();
        static DeviceSet cfg;
        { cudaError_t e = set_smem(mvx_voxelize_tiles_kernel<MODE, CH, BINARY, O16, CL>, smem, &cfg); if (e != cudaSuccess) return e; }
        mvx_voxelize_tiles_kernel<MODE, CH, BINARY, O16, CL><<<grid, kThreads, smem, st>>>(vp);
    } else if (form == FORM_CELLS) {
        constexpr size_t smem = cells_smem_bytes<MODE, CH>();
        static DeviceSet cfg;
        if constexpr (MODE == 1 && CH >= 12) {
            if (vp.lean) {
                static DeviceSet cfg_l;
                { cudaError_t e = set_smem(mvx_voxelize_cells_lean_kernel<MODE, CH, BINARY, O16, CL>, smem, &cfg_l); if (e != cudaSuccess) return e; }
                mvx_voxelize_cells_lean_kernel<MODE, CH, BINARY, O16, CL><<<grid, kThreads, smem, st>>>(vp);
                return cudaGetLastError();
            }
        }
        { cudaError_t e = set_smem(mvx_voxelize_cells_kernel<MODE, CH, BINARY, O16, CL>, smem, &cfg); if (e != cudaSuccess) return e; }
        mvx_voxelize_cells_kernel<MODE, CH, BINARY, O16, CL><<<grid, kThreads, smem, st>>>(vp);
    } else if (nv == 4) {
        mvx_voxelize_kernel<MODE, CH, BINARY, 4, O16, CL><<<grid, kThreads, 0, st>>>(vp);
    } else {
        mvx_voxelize_kernel<MODE, CH, BINARY, 1, O16, CL><<<grid, kThreads, 0, st>>>(vp);
    }
    return cudaGetLastError();
}

}  // namespace

template <int MODE, int CH, bool BINARY>
cudaError_t launch_form(const VoxParams& vp, int form, int nv, unsigned grid, cudaStream_t st) {
    if (vp.clast)   // channels-last instances: the same kernels with the other store / zero-fill code
        return vp.out_kind == 0 ? launch_form_out<MODE, CH, BINARY, false, true>(vp, form, nv, grid, st)
                                : launch_form_out<MODE, CH, BINARY, true, true>(vp, form, nv, grid, st);
    return vp.out_kind == 0 ? launch_form_out<MODE, CH, BINARY, false, false>(vp, form, nv, grid, st)
                            : launch_form_out<MODE, CH, BINARY, true, false>(vp, form, nv, grid, st);
}

template cudaError_t launch_form<MVX_INST_MODE, MVX_INST_CH, (MVX_INST_BINARY != 0)>(const VoxParams&, int, int, unsigned, cudaStream_t);

}  // namespace mvx
