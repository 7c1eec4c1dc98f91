// fused env-step kernel, block shape: 4 arms, 7 envs per block, 7 warps (see fm_kernel.cuh)
#define FM_KERNEL_IMPL
#include "fm_kernel.cuh"
template cudaError_t fm_launch_step<4, FM_BIG4_E, FM_BIG4_TW, true>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
