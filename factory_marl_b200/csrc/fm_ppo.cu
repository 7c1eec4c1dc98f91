// fm_ppo.cu — on-GPU PPO for the batched TaskEnv: policy/value MLP as bf16 tensor-core
// GEMMs (fp32 accumulate), Gaussian / categorical heads, GAE, clipped-surrogate loss + backward, global-norm clip + Adam.
// Replaces what stable-baselines3 computes at the reference's call site  src/learning.py:106-118
//   PPO("MlpPolicy", net_arch=[128,128], gamma=.99) .learn()  — rollout forward, RolloutBuffer.compute_returns_and_advantage,
//   PPO.train  (third party; hyper-parameters recorded in runs/r666unuv.json:11-73; restated [3P]).
// Network (runs/*.json `policy`): separate policy and value MLPs  D -> 128 -> 128 (tanh)  + heads 128 -> A and 128 -> 1;
// DiagGaussian with state-independent log_std (Box actions) or independent 2-way categoricals (MultiDiscrete).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include "../../include/fm_api.h"

static thread_local std::string g_ppo_err;
#define PPO_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_ppo_err = std::string(#x) + ": " + cudaGetErrorString(e_); return 2; } } while (0)
extern "C" const char* fm_ppo_last_error(void) { return g_ppo_err.c_str(); }

typedef __nv_bfloat16 bf16;
extern "C" int fm_mlp_forward_tc5(const float* x, int ld_x, int d_in, const int* idx, const void* w1, const void* w2, const void* w3,
                                  const float* b1, const float* b2, const float* b3, float* out, int op, void* h1, void* h2, int n, int kp,
                                  int* status, void* stream);
extern "C" const char* fm_tc5_last_error(void);
static int g_use_tc5 = 3;     // bit 0: rollout forward on the tcgen05 path (fm_mlp_tc5.cu), bit 1: the update's forward too; 0 = mma.sync GEMM chain
extern "C" void fm_ppo_use_tcgen05(int on) { g_use_tc5 = on; }

// ----------------------------------------------------------------------------------------------------------------------
// Generic small GEMM on the tensor cores: C[M,N] = epi(A[M,K] * B[K,N]); bf16 operands, fp32 accumulate.
// CTA = 128 threads, tile 64 x BN x 16; operands staged K-contiguous in shared memory so every fragment register is one
// 32-bit LDS.  M % 64 == 0, N % BN == 0, K % 16 == 0 (callers pad).  grid = (M/64, N/BN, splitK).
enum { EPI_BIAS_TANH_BF16 = 0, EPI_BIAS_F32 = 1, EPI_DTANH_BF16 = 2, EPI_ATOMIC_F32 = 3 };

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BN, bool A_KM, bool B_KN, int EPI>
__global__ void __launch_bounds__(128) gemm_bf16_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ B, int ldb,
                                                        void* __restrict__ Cv, int ldc, int M, int N, int K, int k_chunk,
                                                        const float* __restrict__ bias, const bf16* __restrict__ aux, int ldaux) {
  constexpr int BM = 64, BK = 16, LD = BK + 8;
  constexpr int WN = BN >= 64 ? 2 : 1;            // warps along N
  constexpr int WM = 4 / WN;                       // warps along M
  constexpr int TM = BM / WM / 16, TN = BN / WN / 8;
  __shared__ __align__(16) bf16 As[BM][LD];
  __shared__ __align__(16) bf16 Bs[BN][LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
  float acc[TM][TN][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  for (int k0 = kb; k0 < ke; k0 += BK) {
    for (int i = tid; i < BM * BK; i += 128) {
      int m, k;
      if (A_KM) { m = i % BM; k = i / BM; } else { k = i % BK; m = i / BK; }
      As[m][k] = A_KM ? A[(size_t)(k0 + k) * lda + m0 + m] : A[(size_t)(m0 + m) * lda + k0 + k];
    }
    for (int i = tid; i < BN * BK; i += 128) {
      int n, k;
      if (B_KN) { n = i % BN; k = i / BN; } else { k = i % BK; n = i / BK; }
      Bs[n][k] = B_KN ? B[(size_t)(k0 + k) * ldb + n0 + n] : B[(size_t)(n0 + n) * ldb + k0 + k];
    }
    __syncthreads();
    uint32_t af[TM][4], bfr[TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int r = wm * (BM / WM) + i * 16 + g;
      af[i][0] = *reinterpret_cast<const uint32_t*>(&As[r][2 * t]);
      af[i][1] = *reinterpret_cast<const uint32_t*>(&As[r + 8][2 * t]);
      af[i][2] = *reinterpret_cast<const uint32_t*>(&As[r][2 * t + 8]);
      af[i][3] = *reinterpret_cast<const uint32_t*>(&As[r + 8][2 * t + 8]);
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int c = wn * (BN / WN) + j * 8 + g;
      bfr[j][0] = *reinterpret_cast<const uint32_t*>(&Bs[c][2 * t]);
      bfr[j][1] = *reinterpret_cast<const uint32_t*>(&Bs[c][2 * t + 8]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) mma16816(acc[i][j], af[i], bfr[j]);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int r = m0 + wm * (BM / WM) + i * 16 + g + (q >= 2 ? 8 : 0);
        int c = n0 + wn * (BN / WN) + j * 8 + 2 * t + (q & 1);
        float v = acc[i][j][q];
        if (EPI == EPI_BIAS_TANH_BF16) reinterpret_cast<bf16*>(Cv)[(size_t)r * ldc + c] = __float2bfloat16(tanhf(v + bias[c]));
        else if (EPI == EPI_BIAS_F32) reinterpret_cast<float*>(Cv)[(size_t)r * ldc + c] = v + bias[c];
        else if (EPI == EPI_DTANH_BF16) { float h = __bfloat162float(aux[(size_t)r * ldaux + c]); reinterpret_cast<bf16*>(Cv)[(size_t)r * ldc + c] = __float2bfloat16(v * (1.f - h * h)); }
        else atomicAdd(&reinterpret_cast<float*>(Cv)[(size_t)r * ldc + c], v);
      }
}

template <int BN, bool A_KM, bool B_KN, int EPI>
static int gemm(const bf16* A, int lda, const bf16* B, int ldb, void* C, int ldc, int M, int N, int K, int splitk, const float* bias,
                const bf16* aux, int ldaux, cudaStream_t s) {
  if (M % 64 || N % BN || K % 16) { g_ppo_err = "gemm: dimensions must be padded (M%64, N%BN, K%16)"; return 1; }
  int chunk = ((K / 16 + splitk - 1) / splitk) * 16;
  dim3 grid(M / 64, N / BN, (K + chunk - 1) / chunk);
  gemm_bf16_kernel<BN, A_KM, B_KN, EPI><<<grid, 128, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, chunk, bias, aux, ldaux);
  PPO_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------------
// network description shared with Python (ctypes Structure in factory_marl_b200/ppo.py)
struct fm_ppo_net {
  int32_t obs_dim, obs_pad, act_dim, out_pad, discrete, n_params, reserved0, reserved1;
  // offsets (in elements) into the flat parameter / gradient buffers; policy net then value net
  int32_t w1[2], b1[2], w2[2], b2[2], w3[2], b3[2], log_std, reserved2;
  float* params;        // fp32 master copy [n_params]
  bf16* params_bf16;    // bf16 copy used by the GEMMs
  float* grads;         // fp32 [n_params]
  float* adam_m; float* adam_v;
  bf16* params_pk;      // the weight matrices once more, in the shared-memory image of the tcgen05 B operand (pack_weights_kernel)
};

struct fm_ppo_ws {       // caller-allocated workspace for up to `cap` rows
  int32_t cap, reserved;
  bf16* x;               // [cap, obs_pad]
  bf16* h1[2]; bf16* h2[2];      // [cap, 128]
  float* out[2];         // [cap, out_pad] policy head, [cap, 16] value head
  bf16* dout[2];         // same shapes, bf16
  bf16* dh2[2]; bf16* dh1[2];
  float* scal;           // [16]: adv mean, adv rstd, stats...
};

__global__ void cast_pad_kernel(const float* __restrict__ src, int n, int d, bf16* __restrict__ dst, int dp, const int* __restrict__ idx) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * dp) return;
  int r = (int)(i / dp), c = (int)(i % dp);
  int sr = idx ? idx[r] : r;
  dst[i] = __float2bfloat16(c < d ? src[(size_t)sr * d + c] : 0.f);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }

// DiagGaussian head (SB3 DiagGaussianDistribution [3P]): a = mean + exp(log_std) * eps; log-prob of the UNclipped sample;
// the env receives clip(a, -1, 1) (SB3 clips Box actions before env.step) and then applies its own tanh (environments.py:98).
__global__ void sample_gaussian_kernel(const float* __restrict__ mean, int ldm, const float* __restrict__ log_std, int n, int A, uint64_t seed,
                                       float* __restrict__ act_raw, float* __restrict__ act_env, float* __restrict__ logp, int deterministic) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float lp = 0.f;
  for (int k = 0; k < A; ++k) {
    uint64_t h = splitmix64(seed ^ ((uint64_t)e * 0x100000001B3ULL + k));
    float u1 = u01(h), u2 = u01(splitmix64(h));
    float eps = deterministic ? 0.f : sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
    float ls = log_std[k], sd = expf(ls);
    float a = mean[(size_t)e * ldm + k] + sd * eps;
    act_raw[(size_t)e * A + k] = a;
    act_env[(size_t)e * A + k] = fminf(fmaxf(a, -1.f), 1.f);
    lp += -0.5f * eps * eps - ls - 0.9189385332046727f;
  }
  logp[e] = lp;
}
// MultiDiscrete([2]*n_arms) head: independent 2-way categoricals from 2*n_arms logits (SB3 MultiCategoricalDistribution [3P])
__global__ void sample_categorical_kernel(const float* __restrict__ logits, int ldm, int n, int A, uint64_t seed, int* __restrict__ act,
                                          float* __restrict__ act_raw, float* __restrict__ logp, int deterministic) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float lp = 0.f;
  for (int k = 0; k < A; ++k) {
    float l0 = logits[(size_t)e * ldm + 2 * k], l1 = logits[(size_t)e * ldm + 2 * k + 1];
    float mx = fmaxf(l0, l1), lse = mx + logf(expf(l0 - mx) + expf(l1 - mx));
    float p1 = expf(l1 - lse);
    int a = deterministic ? (l1 > l0) : (u01(splitmix64(seed ^ ((uint64_t)e * 0x100000001B3ULL + k))) < p1);
    act[(size_t)e * A + k] = a; act_raw[(size_t)e * A + k] = (float)a;
    lp += (a ? l1 : l0) - lse;
  }
  logp[e] = lp;
}
__global__ void copy_col0_kernel(const float* __restrict__ src, int ld, int n, float* __restrict__ dst) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) dst[e] = src[(size_t)e * ld];
}

// GAE (SB3 RolloutBuffer.compute_returns_and_advantage [3P]); thread per env, reverse scan over T; arrays are [T, N]
__global__ void gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const unsigned char* __restrict__ done,
                           const float* __restrict__ last_val, int T, int N, float gamma, float lam, float* __restrict__ adv, float* __restrict__ ret) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  float last = 0.f, nv = last_val[e];
  for (int t = T - 1; t >= 0; --t) {
    size_t i = (size_t)t * N + e;
    float nnt = done[i] ? 0.f : 1.f;       // truncated is always False in this env (environments.py:202)
    float v = val[i];
    float delta = rew[i] + gamma * nv * nnt - v;
    last = delta + gamma * lam * nnt * last;
    adv[i] = last; ret[i] = last + v;
    nv = v;
  }
}

// per-minibatch advantage statistics (mean, 1/(std+1e-8) with the unbiased std torch uses), single CTA
__global__ void adv_stats_kernel(const float* __restrict__ adv, const int* __restrict__ idx, int n, float* __restrict__ scal) {
  __shared__ double s1[256], s2[256];
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < n; i += 256) { double v = adv[idx ? idx[i] : i]; a += v; b += v * v; }
  s1[threadIdx.x] = a; s2[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) { s1[threadIdx.x] += s1[threadIdx.x + s]; s2[threadIdx.x] += s2[threadIdx.x + s]; } __syncthreads(); }
  if (threadIdx.x == 0) {
    double mean = s1[0] / n, var = (s2[0] - n * mean * mean) / (n > 1 ? n - 1 : 1);
    scal[0] = (float)mean; scal[1] = (float)(1.0 / (sqrt(var > 0 ? var : 0) + 1e-8));
  }
}

// PPO clipped-surrogate + value loss, gradients w.r.t. the head outputs (SB3 PPO.train [3P], runs/r666unuv.json:15-23,50-52,73)
__global__ void ppo_loss_kernel(const float* __restrict__ out_pi, int ldp, const float* __restrict__ out_v, int ldv, const float* __restrict__ log_std,
                                const float* __restrict__ act, const float* __restrict__ old_logp, const float* __restrict__ adv,
                                const float* __restrict__ ret, const int* __restrict__ idx, int n, int A, int discrete, float clip, float vf_coef,
                                float ent_coef, int normalize_adv, const float* __restrict__ scal, bf16* __restrict__ dpi, bf16* __restrict__ dv,
                                float* __restrict__ g_log_std, float* __restrict__ stats) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  float pl = 0.f, vl = 0.f, kl = 0.f, cf = 0.f, ent = 0.f;
  float gls[64];                           // act_dim <= 64 (six arms: 48)
  for (int k = 0; k < 64; ++k) gls[k] = 0.f;
  if (r < n) {
    int s = idx ? idx[r] : r;
    float a_hat = adv[s];
    if (normalize_adv) a_hat = (a_hat - scal[0]) * scal[1];
    float lp = 0.f;
    if (!discrete) {
      for (int k = 0; k < A; ++k) {
        float ls = log_std[k], isd = expf(-ls);
        float z = (act[(size_t)s * A + k] - out_pi[(size_t)r * ldp + k]) * isd;
        lp += -0.5f * z * z - ls - 0.9189385332046727f;
        ent += 1.4189385332046727f + ls;
      }
    } else {
      for (int k = 0; k < A; ++k) {
        float l0 = out_pi[(size_t)r * ldp + 2 * k], l1 = out_pi[(size_t)r * ldp + 2 * k + 1];
        float mx = fmaxf(l0, l1), lse = mx + logf(expf(l0 - mx) + expf(l1 - mx));
        int a = act[(size_t)s * A + k] > 0.5f;
        lp += (a ? l1 : l0) - lse;
        float p0 = expf(l0 - lse), p1 = expf(l1 - lse);
        ent -= p0 * (l0 - lse) + p1 * (l1 - lse);
      }
    }
    float lr_ = lp - old_logp[s];
    float ratio = expf(lr_);
    float s1 = ratio * a_hat, s2 = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip) * a_hat;
    pl = -fminf(s1, s2);
    bool unclipped = s1 <= s2;              // gradient flows through `ratio` only when the unclipped term is the min
    float dlp = unclipped ? -a_hat * ratio / n : 0.f;
    cf = fabsf(ratio - 1.f) > clip ? 1.f : 0.f;
    kl = (ratio - 1.f) - lr_;
    float v = out_v[(size_t)r * ldv];
    float dvv = v - ret[s];
    vl = dvv * dvv;
    float dvalue = vf_coef * 2.f * dvv / n;
    for (int c = 0; c < 16; ++c) dv[(size_t)r * 16 + c] = __float2bfloat16(c == 0 ? dvalue : 0.f);
    if (!discrete) {
      for (int k = 0; k < ldp; ++k) {
        float gmean = 0.f;
        if (k < A) {
          float ls = log_std[k], isd = expf(-ls);
          float z = (act[(size_t)s * A + k] - out_pi[(size_t)r * ldp + k]) * isd;
          gmean = dlp * z * isd;                                // dlogp/dmean = (a - mean)/sigma^2
          gls[k] = dlp * (z * z - 1.f) - ent_coef / n;           // dlogp/dlog_std = z^2 - 1 ; entropy bonus d/dlog_std = 1
        }
        dpi[(size_t)r * ldp + k] = __float2bfloat16(gmean);
      }
    } else {
      for (int k = 0; k < ldp / 2; ++k) {
        float g0 = 0.f, g1 = 0.f;
        if (k < A) {
          float l0 = out_pi[(size_t)r * ldp + 2 * k], l1 = out_pi[(size_t)r * ldp + 2 * k + 1];
          float mx = fmaxf(l0, l1), lse = mx + logf(expf(l0 - mx) + expf(l1 - mx));
          float p0 = expf(l0 - lse), p1 = expf(l1 - lse);
          int a = act[(size_t)s * A + k] > 0.5f;
          g0 = dlp * ((a ? 0.f : 1.f) - p0); g1 = dlp * ((a ? 1.f : 0.f) - p1);
          // entropy gradient (ent_coef is 0 in the reference runs): dH/dl_i = -p_i (log p_i + H)
          float H = -(p0 * (l0 - lse) + p1 * (l1 - lse));
          g0 += ent_coef / n * p0 * ((l0 - lse) + H); g1 += ent_coef / n * p1 * ((l1 - lse) + H);
        }
        dpi[(size_t)r * ldp + 2 * k] = __float2bfloat16(g0); dpi[(size_t)r * ldp + 2 * k + 1] = __float2bfloat16(g1);
      }
    }
  }
  // warp-reduce the statistics and the log_std gradient, one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    pl += __shfl_xor_sync(0xffffffff, pl, o); vl += __shfl_xor_sync(0xffffffff, vl, o); kl += __shfl_xor_sync(0xffffffff, kl, o);
    cf += __shfl_xor_sync(0xffffffff, cf, o); ent += __shfl_xor_sync(0xffffffff, ent, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&stats[0], pl / n); atomicAdd(&stats[1], vl / n); atomicAdd(&stats[2], kl / n); atomicAdd(&stats[3], cf / n); atomicAdd(&stats[4], ent / n); }
  if (!discrete)
    for (int k = 0; k < A && k < 64; ++k) {
      float v = gls[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(&g_log_std[k], v);
    }
}

// bias gradients: column sums of a bf16 [n, ld] matrix
__global__ void colsum_kernel(const bf16* __restrict__ d, int n, int ld, int cols, float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  int r0 = blockIdx.y * 256, r1 = min(n, r0 + 256);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += __bfloat162float(d[(size_t)r * ld + c]);
  atomicAdd(&out[c], s);
}

// global-norm clip + Adam (torch.optim.Adam semantics, eps 1e-5 from policy.optimizer.pth) + refresh of the bf16 copy.
// Two launches over the ~1e5 parameters: the squared norm into norm2[0], then the update.
#define FM_ADAM_BLOCKS 64
// ONE block, fixed summation order: with several ranks every replica must compute bit-identical clip factors from the all-reduced
// gradient (an atomicAdd of per-block partial sums differs in the last bit from run to run and lets the replicas drift apart)
__global__ void __launch_bounds__(1024) grad_norm_kernel(const float* __restrict__ g, int n, float grad_scale, float* __restrict__ norm2) {
  __shared__ float red[1024];
  // 16-byte loads, eight in flight per thread: ~9e4 parameters are 22 float4 per thread (a scalar dependent loop took ~150 us)
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const int n4 = n >> 2;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
  for (int i = threadIdx.x; i < n4; i += 1024) {
    const float4 v = g4[i];
    const float a = v.x * grad_scale, b = v.y * grad_scale, c = v.z * grad_scale, d = v.w * grad_scale;
    s0 += a * a; s1 += b * b; s2 += c * c; s3 += d * d;
  }
  float s = (s0 + s1) + (s2 + s3);
  for (int i = 4 * n4 + threadIdx.x; i < n; i += 1024) { const float x = g[i] * grad_scale; s += x * x; }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 512; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) norm2[0] = red[0];
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, bf16* __restrict__ pb, float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int n, float lr, float b1, float b2, float eps, float max_norm,
                                                   float grad_scale, int step, const float* __restrict__ norm2, float* __restrict__ stats) {
  float norm = sqrtf(norm2[0]);
  float cs = fminf(1.f, max_norm / (norm + 1e-6f)) * grad_scale;   // torch.nn.utils.clip_grad_norm_
  // step <= 0: the step count lives on the device in stats[7] (a captured CUDA graph replays this launch with fixed arguments);
  // it is advanced by adam_count_kernel after this launch
  const int st = step > 0 ? step : (int)stats[7] + 1;
  float bc1 = 1.f - powf(b1, (float)st), bc2 = 1.f - powf(b2, (float)st);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    float x = g[i] * cs;
    float mi = b1 * m[i] + (1.f - b1) * x, vi = b2 * v[i] + (1.f - b2) * x * x;
    m[i] = mi; v[i] = vi;
    float upd = (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    float np_ = p[i] - upd;
    p[i] = np_; pb[i] = __float2bfloat16(np_);
    g[i] = 0.f;
  }
}
__global__ void adam_count_kernel(int step, const float* __restrict__ norm2, float* __restrict__ stats) {
  stats[5] = sqrtf(norm2[0]);
  if (step <= 0) stats[7] = stats[7] + 1.f;
}
// tcgen05 B operand of a layer (fm_mlp_tc5.cu): 16-byte chunks [K / 8][N rows][8 consecutive k] — the fused MLP kernel copies this
// image into shared memory with plain 16-byte loads instead of gathering W[k][n] element by element in every CTA.  One thread per
// chunk over the six weight matrices; runs after every parameter change (Adam step, state-dict load).
__global__ void pack_weights_kernel(fm_ppo_net net) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int which = 0; which < 2; ++which) {
    const int op = which == 0 ? net.out_pad : 16;
    const int K[3] = {net.obs_pad, 128, 128}, N[3] = {128, 128, op}, off[3] = {net.w1[which], net.w2[which], net.w3[which]};
    for (int l = 0; l < 3; ++l) {
      const int nchunk = (K[l] >> 3) * N[l];
      if (t < nchunk) {
        const int kc = t / N[l], nr = t % N[l];
        __align__(16) bf16 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = net.params_bf16[off[l] + (size_t)(8 * kc + q) * N[l] + nr];
        *reinterpret_cast<uint4*>(net.params_pk + off[l] + (size_t)t * 8) = *reinterpret_cast<const uint4*>(v);
        return;
      }
      t -= nchunk;
    }
  }
}
static int pack_weights(const fm_ppo_net* net, cudaStream_t s) {
  if (!net->params_pk) return 0;
  const int total = 2 * ((net->obs_pad >> 3) * 128 + 16 * 128) + 16 * (net->out_pad + 16);
  pack_weights_kernel<<<(total + 255) / 256, 256, 0, s>>>(*net);
  return 0;
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ s, bf16* __restrict__ d, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __float2bfloat16(s[i]);
}

// ----------------------------------------------------------------------------------------------------------------------
static int mlp_forward(const fm_ppo_net* net, const fm_ppo_ws* ws, int which, int n, cudaStream_t s) {
  const bf16* P = net->params_bf16; const float* Pf = net->params;
  int Dp = net->obs_pad, Op = which == 0 ? net->out_pad : 16;
  if (int rc = gemm<64, false, true, EPI_BIAS_TANH_BF16>(ws->x, Dp, P + net->w1[which], 128, ws->h1[which], 128, n, 128, Dp, 1, Pf + net->b1[which], nullptr, 0, s)) return rc;
  if (int rc = gemm<64, false, true, EPI_BIAS_TANH_BF16>(ws->h1[which], 128, P + net->w2[which], 128, ws->h2[which], 128, n, 128, 128, 1, Pf + net->b2[which], nullptr, 0, s)) return rc;
  return gemm<16, false, true, EPI_BIAS_F32>(ws->h2[which], 128, P + net->w3[which], Op, ws->out[which], Op, n, Op, 128, 1, Pf + net->b3[which], nullptr, 0, s);
}

static int mlp_backward(const fm_ppo_net* net, const fm_ppo_ws* ws, int which, int n, cudaStream_t s) {
  const bf16* P = net->params_bf16; float* G = net->grads;
  int Dp = net->obs_pad, Op = which == 0 ? net->out_pad : 16;
  int sk = n >= 4096 ? 64 : (n >= 1024 ? 16 : (n >= 256 ? 4 : 1));   // split-K of the weight-gradient GEMMs (K = rows): >= 148 CTAs
  dim3 cb(64);
  colsum_kernel<<<dim3((Op + 63) / 64, (n + 255) / 256), cb, 0, s>>>(ws->dout[which], n, Op, Op, G + net->b3[which]);
  // dW3[128,Op] += H2^T dOUT ; dH2 = (dOUT W3^T) * (1 - H2^2)
  if (int rc = gemm<16, true, true, EPI_ATOMIC_F32>(ws->h2[which], 128, ws->dout[which], Op, G + net->w3[which], Op, 128, Op, n, sk, nullptr, nullptr, 0, s)) return rc;
  if (int rc = gemm<64, false, false, EPI_DTANH_BF16>(ws->dout[which], Op, P + net->w3[which], Op, ws->dh2[which], 128, n, 128, Op, 1, nullptr, ws->h2[which], 128, s)) return rc;
  colsum_kernel<<<dim3(2, (n + 255) / 256), cb, 0, s>>>(ws->dh2[which], n, 128, 128, G + net->b2[which]);
  if (int rc = gemm<64, true, true, EPI_ATOMIC_F32>(ws->h1[which], 128, ws->dh2[which], 128, G + net->w2[which], 128, 128, 128, n, sk, nullptr, nullptr, 0, s)) return rc;
  if (int rc = gemm<64, false, false, EPI_DTANH_BF16>(ws->dh2[which], 128, P + net->w2[which], 128, ws->dh1[which], 128, n, 128, 128, 1, nullptr, ws->h1[which], 128, s)) return rc;
  colsum_kernel<<<dim3(2, (n + 255) / 256), cb, 0, s>>>(ws->dh1[which], n, 128, 128, G + net->b1[which]);
  // dW1[Dp,128] += X^T dH1
  return gemm<64, true, true, EPI_ATOMIC_F32>(ws->x, Dp, ws->dh1[which], 128, G + net->w1[which], 128, Dp, 128, n, sk, nullptr, nullptr, 0, s);
}

#define FM_MON_WINDOW 100
__global__ void monitor_kernel(const uint8_t* __restrict__ term, const float* __restrict__ ep, int n, int16_t* hist, int16_t* first,
                               int32_t* count, float* ring, int32_t* ring_head) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !term[e]) return;
  const int c = count[e];
  const int16_t s0 = (int16_t)ep[4 * e + 2], s1 = (int16_t)ep[4 * e + 3];
  if (c == 0) { first[2 * e] = s0; first[2 * e + 1] = s1; }
  int16_t* h = hist + ((size_t)e * FM_MON_WINDOW + c % FM_MON_WINDOW) * 2;
  h[0] = s0; h[1] = s1;
  count[e] = c + 1;
  const int slot = atomicAdd(ring_head, 1) % FM_MON_WINDOW;
  ring[2 * slot] = ep[4 * e]; ring[2 * slot + 1] = ep[4 * e + 1];
}

extern "C" {

/* SB3 ActorCriticPolicy.forward over a batch of observations (collect_rollouts, learning.py:112): actions for the env
 * (clipped for Box), raw actions + log-probs + values for the rollout buffer.  obs f32 [n, obs_dim] (device). */
int fm_policy_forward(const fm_ppo_net* net, const fm_ppo_ws* ws, const float* obs, int n, uint64_t seed, int deterministic,
                      void* act_env, float* act_raw, float* logp, float* value, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n > ws->cap || n % 64) { g_ppo_err = "fm_policy_forward: n must be a multiple of 64 and <= workspace capacity"; return 1; }
  if (g_use_tc5 & 1) {
    // fused tcgen05 path: the three layers of a net in one kernel, hidden activations stay in shared memory / TMEM
    const bf16* P = net->params_pk; const float* Pf = net->params;
    int* status = reinterpret_cast<int*>(ws->scal + 8);
    if (int rc = fm_mlp_forward_tc5(obs, net->obs_dim, net->obs_dim, nullptr, P + net->w1[0], P + net->w2[0], P + net->w3[0], Pf + net->b1[0],
                                    Pf + net->b2[0], Pf + net->b3[0], ws->out[0], net->out_pad, nullptr, nullptr, n, net->obs_pad, status, stream)) { g_ppo_err = fm_tc5_last_error(); return rc; }
    if (value) {
      if (int rc = fm_mlp_forward_tc5(obs, net->obs_dim, net->obs_dim, nullptr, P + net->w1[1], P + net->w2[1], P + net->w3[1], Pf + net->b1[1],
                                      Pf + net->b2[1], Pf + net->b3[1], ws->out[1], 16, nullptr, nullptr, n, net->obs_pad, status, stream)) { g_ppo_err = fm_tc5_last_error(); return rc; }
      copy_col0_kernel<<<(n + 255) / 256, 256, 0, s>>>(ws->out[1], 16, n, value);
    }
  } else {
    size_t tot = (size_t)n * net->obs_pad;
    cast_pad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(obs, n, net->obs_dim, ws->x, net->obs_pad, nullptr);
    if (int rc = mlp_forward(net, ws, 0, n, s)) return rc;
    if (value) { if (int rc = mlp_forward(net, ws, 1, n, s)) return rc; copy_col0_kernel<<<(n + 255) / 256, 256, 0, s>>>(ws->out[1], 16, n, value); }
  }
  if (net->discrete) sample_categorical_kernel<<<(n + 127) / 128, 128, 0, s>>>(ws->out[0], net->out_pad, n, net->act_dim, seed, (int*)act_env, act_raw, logp, deterministic);
  else sample_gaussian_kernel<<<(n + 127) / 128, 128, 0, s>>>(ws->out[0], net->out_pad, net->params + net->log_std, n, net->act_dim, seed, act_raw, (float*)act_env, logp, deterministic);
  PPO_CUDA(cudaGetLastError());
  return 0;
}

/* RolloutBuffer.compute_returns_and_advantage (gamma learning.py:47, gae_lambda runs/r666unuv.json:42); arrays [T, N] */
int fm_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, int T, int N, float gamma,
           float gae_lambda, float* advantages, float* returns, void* stream) {
  gae_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, values, dones, last_values, T, N, gamma, gae_lambda, advantages, returns);
  PPO_CUDA(cudaGetLastError());
  return 0;
}

/* One minibatch of PPO.train: forward, clipped-surrogate + value loss, backward; gradients are ACCUMULATED into net->grads
 * (caller all-reduces them across ranks, then calls fm_adam_step).  idx: row indices of the minibatch inside the rollout
 * arrays (obs [*, obs_dim], actions [*, act_dim] as f32, old_logp, adv, ret [*]); n % 64 == 0.  stats[0..4] += policy loss,
 * value loss, approx_kl, clip fraction, entropy. */
int fm_ppo_update(const fm_ppo_net* net, const fm_ppo_ws* ws, const float* obs, const float* actions, const float* old_logp,
                  const float* adv, const float* ret, const int32_t* idx, int n, float clip_range, float vf_coef, float ent_coef,
                  int normalize_advantage, float* stats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n > ws->cap || n % 64) { g_ppo_err = "fm_ppo_update: n must be a multiple of 64 and <= workspace capacity"; return 1; }
  size_t tot = (size_t)n * net->obs_pad;
  cast_pad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(obs, n, net->obs_dim, ws->x, net->obs_pad, idx);   // x in bf16: the dW1 operand
  if (g_use_tc5 & 2) {
    // the SAME fused tcgen05 forward as the rollout (gathers the minibatch rows itself, keeps h1 / h2 for the backward pass): the
    // first epoch's probability ratio is exactly 1 and the forward is 2 launches instead of 6 (80 us per optimiser step less at
    // 8192-row minibatches); fm_ppo_use_tcgen05(1) keeps the mma.sync chain here
    const bf16* P = net->params_pk; const float* Pf = net->params;
    int* status = reinterpret_cast<int*>(ws->scal + 8);
    for (int which = 0; which < 2; ++which)
      if (int rc = fm_mlp_forward_tc5(obs, net->obs_dim, net->obs_dim, idx, P + net->w1[which], P + net->w2[which], P + net->w3[which],
                                      Pf + net->b1[which], Pf + net->b2[which], Pf + net->b3[which], ws->out[which], which == 0 ? net->out_pad : 16,
                                      ws->h1[which], ws->h2[which], n, net->obs_pad, status, stream)) { g_ppo_err = fm_tc5_last_error(); return rc; }
  } else {
    if (int rc = mlp_forward(net, ws, 0, n, s)) return rc;
    if (int rc = mlp_forward(net, ws, 1, n, s)) return rc;
  }
  if (normalize_advantage) adv_stats_kernel<<<1, 256, 0, s>>>(adv, idx, n, ws->scal);
  ppo_loss_kernel<<<(n + 127) / 128, 128, 0, s>>>(ws->out[0], net->out_pad, ws->out[1], 16, net->params + net->log_std, actions, old_logp, adv, ret, idx,
                                                 n, net->act_dim, net->discrete, clip_range, vf_coef, ent_coef, normalize_advantage, ws->scal,
                                                 ws->dout[0], ws->dout[1], net->grads + net->log_std, stats);
  PPO_CUDA(cudaGetLastError());
  if (int rc = mlp_backward(net, ws, 0, n, s)) return rc;
  if (int rc = mlp_backward(net, ws, 1, n, s)) return rc;
  PPO_CUDA(cudaGetLastError());
  return 0;
}

/* clip_grad_norm_(max_grad_norm) + Adam step (lr 3e-4, eps 1e-5, runs/r666unuv.json:51-52); grads *= grad_scale first
 * (1/world after the NCCL sum all-reduce); zeroes the gradients and refreshes the bf16 parameter copy.  step = Adam's t
 * (1-based); step <= 0 takes t from stats[7] on the device and increments it there (for CUDA-graph replay). */
int fm_adam_step(const fm_ppo_net* net, float lr, float beta1, float beta2, float eps, float max_grad_norm, float grad_scale, int step,
                 float* stats, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!stats) { g_ppo_err = "fm_adam_step: stats (8 floats on the device) is required"; return 1; }
  float* norm2 = stats + 6;
  PPO_CUDA(cudaMemsetAsync(norm2, 0, sizeof(float), s));
  grad_norm_kernel<<<1, 1024, 0, s>>>(net->grads, net->n_params, grad_scale, norm2);
  adam_kernel<<<FM_ADAM_BLOCKS, 256, 0, s>>>(net->params, net->params_bf16, net->grads, net->adam_m, net->adam_v, net->n_params, lr, beta1, beta2, eps,
                                             max_grad_norm, grad_scale, step, norm2, stats);
  adam_count_kernel<<<1, 1, 0, s>>>(step, norm2, stats);
  pack_weights(net, s);
  PPO_CUDA(cudaGetLastError());
  return 0;
}
/* Monitor + ep_score_history + AdditionalMetricsCallback state (learning.py:53-72,98; environments.py:53,200-201), one launch per
 * env step instead of a get_attr round trip over pipes: for every env that terminated in this step, its episode's scores go into
 * the env's history ring (the callback looks at the last <= 100 episodes per env) and (return, length) into the ring of the last
 * 100 episodes over all envs (SB3's ep_info_buffer, stats_window_size 100). */
int fm_monitor_update(const uint8_t* terminated, const float* ep_info, int n, int16_t* hist, int16_t* first, int32_t* count,
                      float* ring, int32_t* ring_head, void* stream) {
  monitor_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(terminated, ep_info, n, hist, first, count, ring, ring_head);
  PPO_CUDA(cudaGetLastError());
  return 0;
}
int fm_ppo_sync_bf16(const fm_ppo_net* net, void* stream) {
  f32_to_bf16_kernel<<<(net->n_params + 255) / 256, 256, 0, (cudaStream_t)stream>>>(net->params, net->params_bf16, net->n_params);
  pack_weights(net, (cudaStream_t)stream);
  PPO_CUDA(cudaGetLastError());
  return 0;
}
}  // extern "C"
