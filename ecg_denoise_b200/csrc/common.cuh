// common.cuh -- shared device helpers for the RA-LENet sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "ralenet_b200.h"

#define RL_NT 256          // threads per CTA for all window kernels
#define RL_HD 4            // head dim (model/transformer.py:277: dim // num_heads == 4 at every stage)
#define RL_LOG2E 1.4426950408889634f
#define RL_LN_EPS 1e-5f

void rl_set_error(const char* fmt, ...);
void rl_count_launch();
int rl_check_launch(const char* what, int tag0 = -1, int tag1 = -1);

#define RL_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      rl_set_error(__VA_ARGS__);               \
      return (code);                           \
    }                                          \
  } while (0)

template <typename K>
static inline int rl_set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    rl_set_error("cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e));
    return RL_ERR_CUDA;
  }
  return RL_OK;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float x) {          // nn.GELU() exact erf form
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * __expf(-0.5f * x * x) * 0.3989422804014327f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int GS>
__device__ __forceinline__ float group_sum(float v) {        // sum over aligned groups of GS lanes
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of `v` (all RL_NT threads must call); result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float* s_red /* >= 32 floats */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[w] = v;
  __syncthreads();
  float r = (lane < (RL_NT >> 5)) ? s_red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Register-tiled SIMT micro GEMM on shared-memory operands:
//   acc[i][j] += sum_k A[(m0+i)*sam + k*sak] * B[k*sbk + n0 + j*nst]
template <int RM, int RN>
__device__ __forceinline__ void micro_gemm(float (&acc)[RM][RN], const float* __restrict__ A, int sam, int sak,
                                           const float* __restrict__ B, int sbk, int m0, int n0, int nst, int K) {
  const float* a_base = A + m0 * sam;
  const float* b_base = B + n0;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[RM], b[RN];
#pragma unroll
    for (int i = 0; i < RM; ++i) a[i] = a_base[i * sam + k * sak];
#pragma unroll
    for (int j = 0; j < RN; ++j) b[j] = b_base[k * sbk + j * nst];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// stage W^T chunk: dst[k*ldd + n] = W[(n0+n)*ldw + k0 + k]   for n < NC, k < KC   (coalesced along k)
__device__ __forceinline__ void stage_wT(float* dst, int ldd, const float* __restrict__ W, int ldw, int n0, int NC,
                                         int k0, int KC) {
  for (int i = threadIdx.x; i < NC * KC; i += RL_NT) {
    const int k = i % KC, n = i / KC;
    dst[k * ldd + n] = __ldg(W + (size_t)(n0 + n) * ldw + k0 + k);
  }
}
// stage W chunk as is: dst[r*ldd + c] = W[(r0+r)*ldw + c0 + c]   for r < R, c < Cc   (coalesced along c)
__device__ __forceinline__ void stage_w(float* dst, int ldd, const float* __restrict__ W, int ldw, int r0, int R,
                                        int c0, int Cc) {
  for (int i = threadIdx.x; i < R * Cc; i += RL_NT) {
    const int c = i % Cc, r = i / Cc;
    dst[r * ldd + c] = __ldg(W + (size_t)(r0 + r) * ldw + c0 + c);
  }
}

// copy a contiguous [n] float block global <-> shared with float4 (n % 4 == 0, 16B aligned)
__device__ __forceinline__ void copy_g2s(float* dst, const float* __restrict__ src, int n) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < n / 4; i += RL_NT) d4[i] = __ldg(s4 + i);
}
__device__ __forceinline__ void copy_s2g(float* __restrict__ dst, const float* src, int n) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < n / 4; i += RL_NT) d4[i] = s4[i];
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over C channels of `rows` tokens (eps 1e-5, biased variance, like nn.LayerNorm).
// A group of GS = min(C,32) lanes owns one token; each lane holds CPL = C/GS channels.
// load(t, c) returns the pre-norm value; store(t, c, zhat, rstd_is_unused) receives zhat.
template <int C>
struct LnGeom {
  static constexpr int GS = (C < 32) ? C : 32;
  static constexpr int CPL = C / GS;
  static constexpr int TPW = 32 / GS;                 // tokens per warp per iteration
  static constexpr int TPI = TPW * (RL_NT / 32);      // tokens per CTA iteration
};

template <int C, class Load, class Store>
__device__ __forceinline__ void ln_forward_rows(int rows, Load load, Store store) {
  using G = LnGeom<C>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % G::GS, gi = lane / G::GS;
  for (int t0 = 0; t0 < rows; t0 += G::TPI) {
    const int t = t0 + warp * G::TPW + gi;
    const bool ok = t < rows;
    float z[G::CPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < G::CPL; ++i) {
      z[i] = ok ? load(t, gl + i * G::GS) : 0.f;
      s += z[i];
    }
    const float mu = group_sum<G::GS>(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < G::CPL; ++i) {
      z[i] -= mu;
      q += z[i] * z[i];
    }
    const float rstd = rsqrtf(group_sum<G::GS>(q) * (1.0f / C) + RL_LN_EPS);
    if (ok) {
#pragma unroll
      for (int i = 0; i < G::CPL; ++i) store(t, gl + i * G::GS, z[i] * rstd);
    }
  }
}

// LayerNorm backward for `rows` tokens.
//   loadz(t,c): pre-norm input;  loaddu(t,c): gradient w.r.t. LN output;  gamma: [C]
//   emit(t, c, dz, zhat): receives the gradient w.r.t. the pre-norm input and zhat
//   dgam/dbet partial sums are accumulated into s_gb[0:C] / s_gb[C:2C] (shared, pre-zeroed) with atomics.
template <int C, class LoadZ, class LoadDu, class Emit>
__device__ __forceinline__ void ln_backward_rows(int rows, const float* __restrict__ gamma, float* s_gb, LoadZ loadz,
                                                 LoadDu loaddu, Emit emit) {
  using G = LnGeom<C>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane % G::GS, gi = lane / G::GS;
  float gam[G::CPL], ag[G::CPL], ab[G::CPL];
#pragma unroll
  for (int i = 0; i < G::CPL; ++i) {
    gam[i] = __ldg(gamma + gl + i * G::GS);
    ag[i] = 0.f;
    ab[i] = 0.f;
  }
  for (int t0 = 0; t0 < rows; t0 += G::TPI) {
    const int t = t0 + warp * G::TPW + gi;
    const bool ok = t < rows;
    float z[G::CPL], du[G::CPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < G::CPL; ++i) {
      z[i] = ok ? loadz(t, gl + i * G::GS) : 0.f;
      du[i] = ok ? loaddu(t, gl + i * G::GS) : 0.f;
      s += z[i];
    }
    const float mu = group_sum<G::GS>(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < G::CPL; ++i) {
      z[i] -= mu;
      q += z[i] * z[i];
    }
    const float rstd = rsqrtf(group_sum<G::GS>(q) * (1.0f / C) + RL_LN_EPS);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < G::CPL; ++i) {
      z[i] *= rstd;                       // zhat
      ag[i] += du[i] * z[i];
      ab[i] += du[i];
      du[i] *= gam[i];                    // dzhat
      s1 += du[i];
      s2 += du[i] * z[i];
    }
    s1 = group_sum<G::GS>(s1) * (1.0f / C);
    s2 = group_sum<G::GS>(s2) * (1.0f / C);
    if (ok) {
#pragma unroll
      for (int i = 0; i < G::CPL; ++i) emit(t, gl + i * G::GS, rstd * (du[i] - s1 - z[i] * s2), z[i]);
    }
  }
  if (s_gb != nullptr) {
#pragma unroll
    for (int i = 0; i < G::CPL; ++i) {
      atomicAdd(&s_gb[gl + i * G::GS], ag[i]);
      atomicAdd(&s_gb[C + gl + i * G::GS], ab[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
constexpr int SW_FLOATS = 4608;   // weight staging area (18 KB)

// CTA-wide GEMM accumulator: out[M x N] spread over RL_NT threads, RM x RN micro tile per thread at
// rows m0+i, cols n0 + j*nst.  Requires (M/RM)*(N/RN) <= RL_NT.
template <int RM, int RN>
struct TileAcc {
  float acc[RM][RN];
  int m0, n0, nst;
  bool active;
  __device__ __forceinline__ void init(int M, int N) {
    const int ntn = N / RN;
    const int t = threadIdx.x;
    active = t < (M / RM) * ntn;
    m0 = (t / ntn) * RM;
    n0 = t % ntn;
    nst = ntn;
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
  }
  __device__ __forceinline__ void mac(const float* A, int sam, int sak, const float* B, int sbk, int K) {
    if (active) micro_gemm<RM, RN>(acc, A, sam, sak, B, sbk, m0, n0, nst, K);
  }
  template <class F>
  __device__ __forceinline__ void epilogue(F f) {
    if (active) {
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) f(m0 + i, n0 + j * nst, acc[i][j]);
    }
  }
};

__device__ __forceinline__ int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

template <int C>
__host__ __device__ constexpr int lda_of() { return C + 1; }   // odd row stride: conflict-free A operand

// generic weight-gradient GEMM launcher (wgrad.cu):
//   dW[n*K + k] += sum_m dY[m*ldy + n] * X[m*ldx + k]      n < N, k < K, m < M
//   db[n]       += sum_m dY[m*ldy + n]                     (db may be NULL)
int rl_launch_wgrad(const float* dY, int ldy, const float* X, int ldx, int M, int N, int K, float* dW, float* db,
                    cudaStream_t st);
