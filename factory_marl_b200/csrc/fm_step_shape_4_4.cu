// fused env-step kernel, block shape: 4 arms, 4 envs per block, 4 warps (see fm_kernel.cuh)
#define FM_KERNEL_IMPL
#include "fm_kernel.cuh"
template cudaError_t fm_launch_step<4, 4, 4>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
