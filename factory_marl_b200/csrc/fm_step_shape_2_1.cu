// fused env-step kernel, block shape: 2 arms, 1 envs per block, 1 warps (see fm_kernel.cuh)
#define FM_KERNEL_IMPL
#include "fm_kernel.cuh"
template cudaError_t fm_launch_step<2, 1, 1>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
