// fused env-step kernel, block shape: 6 arms, 6 envs per block, 8 warps (see fm_kernel.cuh)
#define FM_KERNEL_IMPL
#include "fm_kernel.cuh"
template cudaError_t fm_launch_step<6, FM_BIG6_E, FM_BIG6_TW>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
