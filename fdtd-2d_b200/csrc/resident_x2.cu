// The twelve instantiations of the packed cluster-resident kernel (grid_resident_x2.cuh: uniform / mapped dt/(mu*dx) x the row
// at which the first band ends) and their launcher, as a translation unit of their own: the build compiles it next to
// api.cu (these kernels take longer than everything else together -- ten role-specialised copies of the time loop each).
// api.cu calls resident_x2_launch and never names the kernel template.
#include "grid_resident_x2.cuh"

namespace fdtd2d {

template <bool UCH, int RL>
static cudaError_t launch_t(const cudaLaunchConfig_t* cfg, const PassParams<float>& p, float ch_uniform, bool set_smem_attr, size_t smem,
                            int* max_active_clusters) {
    if (set_smem_attr) {
        const cudaError_t e = cudaFuncSetAttribute(grid_resident_x2_kernel<UCH, RL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (max_active_clusters) cudaOccupancyMaxActiveClusters(max_active_clusters, grid_resident_x2_kernel<UCH, RL>, cfg);
    const unsigned long long negzero = 0x8000000080000000ull;  // the -0 pair of the packed products (strip_wave.cuh)
    return cudaLaunchKernelEx(cfg, grid_resident_x2_kernel<UCH, RL>, p, ch_uniform, negzero);
}

// rl: the row inside a row block at which the first band ends, 0..5 (5 = with the block)
cudaError_t resident_x2_launch(bool uch, int rl, const cudaLaunchConfig_t* cfg, const PassParams<float>& p, float ch_uniform, bool set_smem_attr,
                               size_t smem, int* max_active_clusters) {
#define FDTD2D_RX_CASE(RL) \
    case RL: return uch ? launch_t<true, RL>(cfg, p, ch_uniform, set_smem_attr, smem, max_active_clusters) \
                        : launch_t<false, RL>(cfg, p, ch_uniform, set_smem_attr, smem, max_active_clusters);
    switch (rl) {
        FDTD2D_RX_CASE(0) FDTD2D_RX_CASE(1) FDTD2D_RX_CASE(2) FDTD2D_RX_CASE(3) FDTD2D_RX_CASE(4) FDTD2D_RX_CASE(5)
        default: return cudaErrorInvalidValue;
    }
#undef FDTD2D_RX_CASE
}

}  // namespace fdtd2d
