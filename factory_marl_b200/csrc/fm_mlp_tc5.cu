// fm_mlp_tc5.cu — fused policy/value MLP forward on the 5th-generation tensor cores (sm_100a):
//   OUT = (tanh(tanh(X W1 + b1) W2 + b2)) W3 + b3      for a tile of 128 observations per CTA,
// bf16 operands staged in shared memory in the canonical K-major core-matrix layout, `tcgen05.mma` (cta_group::1,
// kind::f16, M = 128) issued by one thread, fp32 accumulators in TMEM, `tcgen05.ld` epilogues that apply bias + tanh and
// write the next layer's A operand straight back to shared memory (the hidden activations never touch HBM unless asked).
// Replaces ActorCriticPolicy.forward of stable-baselines3 at the reference's call site src/learning.py:112 (rollout).
// Descriptor formats follow cute/arch/mma_sm100_desc.hpp (vendored CUTLASS headers, read for the bit layouts only).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

typedef __nv_bfloat16 bf16;
extern thread_local std::string g_tc5_err;
thread_local std::string g_tc5_err;

// ---- shared-memory operand layout (K-major, no swizzle): 16-byte row chunks (8 bf16 along K), 8 rows per core matrix stored
//      contiguously (128 B), core matrices ordered row-group-major inside a K-chunk: byte address of (row r, chunk kc) =
//      ((kc * (ROWS/8) + r/8) * 8 + r%8) * 16.   => SBO (between 8-row groups) = 128 B, LBO (between K-chunks) = ROWS*16 B.
__device__ __forceinline__ uint32_t chunk_off(int r, int kc, int rows) { return (uint32_t)(((kc * (rows >> 3) + (r >> 3)) * 8 + (r & 7)) * 16); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address, 16-byte units
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading-dimension byte offset (K direction)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // stride-dimension byte offset (M/N direction)
  d |= (uint64_t)1 << 46;                                // descriptor version 1 (Blackwell)
  return d;                                              // layout type 0 = no swizzle, base offset 0
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, N and M encoded as N>>3, M>>4
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc5_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc5_commit(uint32_t bar_smem) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_smem) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar_smem, uint32_t parity) {
  // bounded spin: a descriptor mistake must not hang the GPU
  for (int it = 0; it < (1 << 22); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar_smem), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

struct Tc5Args {
  const float* x; int ld_x; int d_in;       // fp32 observations [n, d_in]
  const int* idx;                            // optional row gather
  const bf16* w1; const bf16* w2; const bf16* w3;      // [Kp,128], [128,128], [128,Op] in the packed B-operand image (load_b)
  const float* b1; const float* b2; const float* b3;
  float* out; int op;                        // fp32 [n, op]
  bf16* h1; bf16* h2;                        // optional activations [n,128] (training forward), may be null
  int n, kp;
  int* status;                               // set to 1 on an mbarrier timeout
};

// B operand: the caller passes the weights already in the shared-memory image (16-byte chunks [K / 8][rows][8 consecutive k],
// fm_ppo.cu pack_weights_kernel), so staging is a straight coalesced 16-byte copy
__device__ __forceinline__ void load_b(unsigned char* sB, const bf16* __restrict__ Wp, int K, int rows) {
  const int nchunk = (K >> 3) * rows;
  const uint4* src = reinterpret_cast<const uint4*>(Wp);
  uint4* dst = reinterpret_cast<uint4*>(sB);
  for (int i = threadIdx.x; i < nchunk; i += blockDim.x) dst[i] = src[i];
}

__global__ void __launch_bounds__(128, 1) mlp_tc5_kernel(Tc5Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int kp = a.kp;
  unsigned char* sA = smem;                                   // 128 x max(kp,128) bf16
  unsigned char* sB = smem + 128 * (kp > 128 ? kp : 128) * 2; // 128 x max(kp,128) bf16
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);

  if (warp == 0) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_addr), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // A operand of layer 1: observations -> bf16, zero padded to kp, gathered through idx if given
  {
    const int nch = kp >> 3;
    for (int i = tid; i < 128 * nch; i += 128) {
      int r = i / nch, kc = i % nch;
      int gr = row0 + r;
      __align__(16) bf16 v[8];
      const float* src = nullptr;
      if (gr < a.n) src = a.x + (size_t)(a.idx ? a.idx[gr] : gr) * a.ld_x;
#pragma unroll
      for (int q = 0; q < 8; ++q) { int c = 8 * kc + q; v[q] = __float2bfloat16((src && c < a.d_in) ? src[c] : 0.f); }
      *reinterpret_cast<uint4*>(sA + chunk_off(r, kc, 128)) = *reinterpret_cast<const uint4*>(v);
    }
  }
  load_b(sB, a.w1, kp, 128);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t sA_addr = (uint32_t)__cvta_generic_to_shared(sA), sB_addr = (uint32_t)__cvta_generic_to_shared(sB);
  uint32_t parity = 0;
  bool ok = true;

  for (int layer = 0; layer < 3; ++layer) {
    const int K = layer == 0 ? kp : 128;
    const int N = layer == 2 ? a.op : 128;
    const int brows = layer == 2 ? a.op : 128;
    if (tid == 0) {
      const uint32_t idesc = make_idesc(128, N);
      for (int j = 0; j < K / 16; ++j) {
        // one MMA consumes K = 16 = two 16-byte chunks; consecutive chunks are ROWS*16 bytes apart (LBO)
        uint64_t ad = make_smem_desc(sA_addr + (uint32_t)j * 2u * 128u * 16u, 128u * 16u, 128u);
        uint64_t bd = make_smem_desc(sB_addr + (uint32_t)j * 2u * (uint32_t)brows * 16u, (uint32_t)brows * 16u, 128u);
        tc5_mma(tmem, ad, bd, idesc, j > 0 ? 1u : 0u);
      }
      tc5_commit(bar_addr);
    }
    ok = mbar_wait(bar_addr, parity) && ok;
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w..32w+31 = tile rows; 8 columns per tcgen05.ld
    const int r = warp * 32 + lane;
    const int gr = row0 + r;
    const float* bias = layer == 0 ? a.b1 : (layer == 1 ? a.b2 : a.b3);
    bf16* hsave = layer == 0 ? a.h1 : (layer == 1 ? a.h2 : nullptr);
    for (int c = 0; c < N / 8; ++c) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(8 * c), v);
      if (layer < 2) {
        __align__(16) bf16 hv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) hv[q] = __float2bfloat16(tanhf(v[q] + bias[8 * c + q]));
        *reinterpret_cast<uint4*>(sA + chunk_off(r, c, 128)) = *reinterpret_cast<const uint4*>(hv);   // next layer's A operand
        if (hsave && gr < a.n) *reinterpret_cast<uint4*>(hsave + (size_t)gr * 128 + 8 * c) = *reinterpret_cast<const uint4*>(hv);
      } else if (gr < a.n) {
#pragma unroll
        for (int q = 0; q < 8; ++q) a.out[(size_t)gr * a.op + 8 * c + q] = v[q] + bias[8 * c + q];
      }
    }
    if (layer < 2) {
      // the accumulator has been drained and sA rewritten: stage the next weights, then hand over to the tensor core again
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (layer == 0) load_b(sB, a.w2, 128, 128); else load_b(sB, a.w3, 128, a.op);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
  if (!ok && tid == 0 && a.status) *a.status = 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u) : "memory");
}

extern "C" {
const char* fm_tc5_last_error(void) { return g_tc5_err.c_str(); }
/* Fused 3-layer MLP forward on tcgen05 for one net.  All pointers are device pointers; weights bf16 row-major [K, N] as in
 * fm_ppo_net (w1 [kp,128], w2 [128,128], w3 [128,op]); out fp32 [n, op]; h1/h2 optional bf16 [n,128]; status: int on device. */
int fm_mlp_forward_tc5(const float* x, int ld_x, int d_in, const int* idx, const void* w1, const void* w2, const void* w3,
                       const float* b1, const float* b2, const float* b3, float* out, int op, void* h1, void* h2, int n, int kp,
                       int* status, void* stream) {
  if (kp % 16 || kp > 448 || op % 16 || op < 16 || op > 64) { g_tc5_err = "fm_mlp_forward_tc5: kp % 16 == 0, kp <= 448, op in {16, 32, 48, 64}"; return 1; }
  Tc5Args a;
  a.x = x; a.ld_x = ld_x; a.d_in = d_in; a.idx = idx; a.w1 = (const bf16*)w1; a.w2 = (const bf16*)w2; a.w3 = (const bf16*)w3;
  a.b1 = b1; a.b2 = b2; a.b3 = b3; a.out = out; a.op = op; a.h1 = (bf16*)h1; a.h2 = (bf16*)h2; a.n = n; a.kp = kp; a.status = status;
  size_t smem = (size_t)2 * 128 * (kp > 128 ? kp : 128) * 2;
  {  // the opt-in is per device and this process may drive several: set it on every launch (a host-side table look-up)
    cudaError_t e = cudaFuncSetAttribute(mlp_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_tc5_err = cudaGetErrorString(e); return 2; }
  }
  mlp_tc5_kernel<<<(n + 127) / 128, 128, smem, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_tc5_err = cudaGetErrorString(e); return 2; }
  return 0;
}
}
