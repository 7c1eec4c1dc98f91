// fused env-step kernel, block shape: 6 arms, 1 env per block, 1 warp (see fm_kernel.cuh)
#define FM_KERNEL_IMPL
#include "fm_kernel.cuh"
template cudaError_t fm_launch_step<6, 1, 1>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
