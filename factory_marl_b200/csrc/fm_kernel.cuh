// fm_kernel.cuh — the fused env-step kernel and its launcher, one explicit instantiation per block shape (fm_step_shape_*.cu).
//   NA = arms, E = envs per block (state of E envs lives in shared memory for the 100 sub-steps), TW = warps per block.
// Work items of a phase are spread over all 32*TW lanes of the block (item w -> env w % E); phases are separated by
// __syncthreads().  fm_lib.cu only sees the declaration of fm_launch_step and the extern template lines below.
#pragma once
#include <cuda_runtime.h>
#include "fm_driver.cuh"

// the one-block-per-SM shape of the two-arm / four-arm scene (overridable for shape experiments, tools/ab_variant.py)
#ifndef FM_BIG2_E
#define FM_BIG2_E 14
#define FM_BIG2_TW 12
#endif
#ifndef FM_BIG4_E
#define FM_BIG4_E 7
#define FM_BIG4_TW 7
#endif
#ifndef FM_BIG6_E
#define FM_BIG6_E 6
#define FM_BIG6_TW 8
#endif
#ifndef FM_MINBLK
#define FM_MINBLK 1
#endif
struct FmLaunch { const FmDevModel* model; const unsigned int* cpairs; const int* geom_class; FmEnvCfg cfg; FmStatePtrs st; };

template <int NA, int E, int TW, bool BIGN = false>
cudaError_t fm_launch_step(const FmLaunch& L, const FmStepArgs& ar, cudaStream_t s);

#ifdef FM_KERNEL_IMPL

// barrier among W consecutive warps of the block, starting at warp BASE (named barrier ID; barrier 0 is __syncthreads)
template <int ID, int BASE, int W>
struct DevGroupRun {
  template <class F> __device__ __forceinline__ void operator()(F f) {
    f((int)threadIdx.x - 32 * BASE);
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(32 * W) : "memory");
  }
};
// the 32 lanes of a warp working through a sequence of stages (island_newton_warp): a stage is the same code on every lane followed
// by __syncwarp(); data that crosses stages lives in shared / global memory
struct DevLanes {
  int lane;
  template <class G> __device__ __forceinline__ void stage(G g) { g(lane); __syncwarp(); }
  __device__ __forceinline__ void sync() { __syncwarp(); }
  template <class G> __device__ __forceinline__ void run(G g) { g(lane); }   // no synchronisation after it
  template <class T, int N> struct Arr { T a[N]; __device__ __forceinline__ T* operator()(int) { return a; } };   // per-lane registers
};
struct DevRun {
  template <class F> __device__ __forceinline__ void operator()(F f) { f((int)threadIdx.x); __syncthreads(); }
  // one barrier interval whose work items are taken by whole warps: f(warp index, lanes)
  template <class F> __device__ __forceinline__ void warps(F f) { DevLanes L{(int)threadIdx.x & 31}; f((int)threadIdx.x >> 5, L); __syncthreads(); }
  // warps [0, WA) run the intervals of progA, warps [WA, WA + WB) those of progB meanwhile; the whole block joins at the end
  template <int WA, int WB, class PA, class PB> __device__ __forceinline__ void split(PA progA, PB progB) {
    if constexpr (WB == 0) {
      progA(*this, 32 * WA); progB(*this, 32 * WA);
    } else {
      if (threadIdx.x < 32 * WA) { DevGroupRun<1, 0, WA> sub; progA(sub, 32 * WA); }
      else { DevGroupRun<2, WA, WB> sub; progB(sub, 32 * WB); }
      __syncthreads();
    }
  }
};

template <int NA, int E, int TW, bool BIGN>
__global__ void __launch_bounds__(32 * TW, FM_MINBLK) fm_step_kernel(const FmDevModel* __restrict__ mp, const unsigned int* __restrict__ cpairs,
                                                             const int* __restrict__ geom_class, FmEnvCfg cfg, FmStatePtrs st, FmStepArgs ar) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EnvS<NA>* S = reinterpret_cast<EnvS<NA>*>(smem_raw);
  TaskS<NA>* TS = reinterpret_cast<TaskS<NA>*>(S + E);
  // zero the block's shared state (counters, lists)
  {
    int* z = reinterpret_cast<int*>(smem_raw);
    const int nw = (int)((E * (sizeof(EnvS<NA>) + sizeof(TaskS<NA>))) / 4);
    for (int i = threadIdx.x; i < nw; i += 32 * TW) z[i] = 0;
    __syncthreads();
  }
  __builtin_assume(__isShared(S)); __builtin_assume(__isShared(TS));
  DevRun run;
#ifdef FM_BLOCK_CYCLES
  const long long t0 = clock64();
#endif
  fm_block_step<NA, E, TW, BIGN>(run, S, TS, *mp, cpairs, geom_class, cfg, st, ar, (int)blockIdx.x * E);
#ifdef FM_BLOCK_CYCLES
  // profiling variant (tools/block_cycles.py): the block's cycle count goes into ep_info[env][0] of its envs that did NOT terminate
  // (the monitor reads ep_info only for terminated envs)
  __syncthreads();
  if (ar.mode == 0 && ar.ep_info && threadIdx.x < E) {
    const int env = slot_env(ar.order, (int)blockIdx.x * E + (int)threadIdx.x, st.n);
    if (env < st.n && !ar.terminated[env]) ar.ep_info[4 * env] = (float)(clock64() - t0);
  }
#endif
}

template <int NA, int E, int TW, bool BIGN>
cudaError_t fm_launch_step(const FmLaunch& L, const FmStepArgs& ar, cudaStream_t s) {
  const size_t smem = E * (sizeof(EnvS<NA>) + sizeof(TaskS<NA>));
  {   // per device, not per process: set on every launch (cheap) so that a second GPU in the same process gets it too
    cudaError_t e = cudaFuncSetAttribute(fm_step_kernel<NA, E, TW, BIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int nblk = (L.st.n + E - 1) / E;
  fm_step_kernel<NA, E, TW, BIGN><<<nblk, 32 * TW, smem, s>>>(L.model, L.cpairs, L.geom_class, L.cfg, L.st, ar);
  return cudaGetLastError();
}
#else
extern template cudaError_t fm_launch_step<2, 1, 1>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<2, 4, 4>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<2, FM_BIG2_E, FM_BIG2_TW>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<2, FM_BIG2_E, FM_BIG2_TW, true>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<4, 1, 1>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<4, 4, 4>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<4, FM_BIG4_E, FM_BIG4_TW>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<4, FM_BIG4_E, FM_BIG4_TW, true>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<6, 1, 1>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<6, FM_BIG6_E, FM_BIG6_TW>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
extern template cudaError_t fm_launch_step<6, FM_BIG6_E, FM_BIG6_TW, true>(const FmLaunch&, const FmStepArgs&, cudaStream_t);
#endif
