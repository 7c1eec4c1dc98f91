// fused env-step kernel, block shape: 2 arms, 14 envs per block, 12 warps (see fm_kernel.cuh)
#define FM_KERNEL_IMPL
#include "fm_kernel.cuh"
template cudaError_t fm_launch_step<2, FM_BIG2_E, FM_BIG2_TW, true>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
