// small_ops.cu -- the stand-alone forwards of the helper modules that the network kernels fuse away:
// LinearProjection.forward (model/transformer.py:226-247: q = to_q(x), kv = to_kv(x)),
// AbsPositionalEncoding.forward (:179-181: X + P[:, :L]) and PartialConv_1d.forward_split_cat (:54-59: Conv1d(1,1,3)
// on channel 0 of a channels-first tensor, the other channels pass through).  They are NOT on the hot path (inside
// ralenet.forward all three are fused into the attention / feed-forward kernels); they exist so that a user who calls
// these sub-modules directly, as the reference allows, runs on the B200 kernels too instead of hitting an error.
#include "common.cuh"

namespace {

// y[m][n] = sum_k x[m][k] w[n][k] (+ b[n]);  one warp per output row block of 32 columns, k-loop over shared tiles
constexpr int LT = 32;
__global__ void __launch_bounds__(LT * 8) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, float* __restrict__ y, int M,
                                                            int K, int N) {
  __shared__ float sx[8][LT + 1], sw[LT][LT + 1];
  const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
  const int m = blockIdx.y * 8 + ty, n = blockIdx.x * LT + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += LT) {
    sx[ty][tx] = (m < M && k0 + tx < K) ? x[(size_t)m * K + k0 + tx] : 0.f;
    for (int r = ty; r < LT; r += 8) {
      const int nn = blockIdx.x * LT + r;
      sw[r][tx] = (nn < N && k0 + tx < K) ? w[(size_t)nn * K + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < LT; ++k) acc = fmaf(sx[ty][k], sw[tx][k], acc);
    __syncthreads();
  }
  if (m < M && n < N) y[(size_t)m * N + n] = acc + (b ? b[n] : 0.f);
}

// dx[m][k] = sum_n dy[m][n] w[n][k]
__global__ void __launch_bounds__(LT * 8) linear_bwd_data_kernel(const float* __restrict__ dy,
                                                                 const float* __restrict__ w, float* __restrict__ dx,
                                                                 int M, int K, int N) {
  __shared__ float sd[8][LT + 1], sw[LT][LT + 1];
  const int tx = threadIdx.x % LT, ty = threadIdx.x / LT;
  const int m = blockIdx.y * 8 + ty, k = blockIdx.x * LT + tx;
  float acc = 0.f;
  for (int n0 = 0; n0 < N; n0 += LT) {
    sd[ty][tx] = (m < M && n0 + tx < N) ? dy[(size_t)m * N + n0 + tx] : 0.f;
    for (int r = ty; r < LT; r += 8) {
      const int kk = blockIdx.x * LT + tx;
      sw[r][tx] = (n0 + r < N && kk < K) ? w[(size_t)(n0 + r) * K + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < LT; ++n) acc = fmaf(sd[ty][n], sw[n][tx], acc);
    __syncthreads();
  }
  if (m < M && k < K) dx[(size_t)m * K + k] = acc;
}

__global__ void pe_add_kernel(const float* __restrict__ x, const float* __restrict__ pe, float* __restrict__ y,
                              long long total, int per) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] + __ldg(pe + (int)(i % per));
}

// channels-first [B][C][L]; channel 0: y[t] = w0 x[t-1] + w1 x[t] + w2 x[t+1] (zero padded), others: copy.
// transpose = 1 computes the adjoint (data gradient): y[t] = w2 x[t-1] + w1 x[t] + w0 x[t+1].
__global__ void pconv1_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int B,
                              int C, int L, int transpose) {
  const long long total = (long long)B * C * L;
  const float w0 = __ldg(w + (transpose ? 2 : 0)), w1 = __ldg(w + 1), w2 = __ldg(w + (transpose ? 0 : 2));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % L);
    const int c = (int)((i / L) % C);
    float v = x[i];
    if (c == 0) v = (t > 0 ? w0 * x[i - 1] : 0.f) + w1 * v + (t + 1 < L ? w2 * x[i + 1] : 0.f);
    y[i] = v;
  }
}

// dw[k] += sum_{b,t} dy[b][0][t] x[b][0][t+k-1]
__global__ void __launch_bounds__(256) pconv1_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           float* __restrict__ dw, int B, int C, int L) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  const long long total = (long long)B * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % L);
    const long long o = (i / L) * (long long)C * L + t;
    const float d = dy[o];
    a0 += t > 0 ? d * x[o - 1] : 0.f;
    a1 += d * x[o];
    a2 += t + 1 < L ? d * x[o + 1] : 0.f;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, s);
    a1 += __shfl_xor_sync(0xffffffffu, a1, s);
    a2 += __shfl_xor_sync(0xffffffffu, a2, s);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(dw + 0, a0);
    atomicAdd(dw + 1, a1);
    atomicAdd(dw + 2, a2);
  }
}

int grid1d(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

extern "C" int ralenet_linear_fwd(const float* x, const float* w, const float* b, float* y, int32_t M, int32_t K,
                                  int32_t N, void* stream) {
  RL_REQUIRE(x && w && y, RL_ERR_NULL, "linear_fwd: NULL tensor");
  RL_REQUIRE(M > 0 && K > 0 && N > 0, RL_ERR_SHAPE, "linear_fwd: M=%d K=%d N=%d", M, K, N);
  rl_prof_pre((cudaStream_t)stream);
  linear_fwd_kernel<<<dim3((N + LT - 1) / LT, (M + 7) / 8), LT * 8, 0, (cudaStream_t)stream>>>(x, w, b, y, M, K, N);
  return rl_check_launch("linear_fwd_kernel");
}

extern "C" int ralenet_linear_bwd_data(const float* dy, const float* w, float* dx, int32_t M, int32_t K, int32_t N,
                                       void* stream) {
  RL_REQUIRE(dy && w && dx, RL_ERR_NULL, "linear_bwd_data: NULL tensor");
  RL_REQUIRE(M > 0 && K > 0 && N > 0, RL_ERR_SHAPE, "linear_bwd_data: M=%d K=%d N=%d", M, K, N);
  rl_prof_pre((cudaStream_t)stream);
  linear_bwd_data_kernel<<<dim3((K + LT - 1) / LT, (M + 7) / 8), LT * 8, 0, (cudaStream_t)stream>>>(dy, w, dx, M, K, N);
  return rl_check_launch("linear_bwd_data_kernel");
}

extern "C" int ralenet_pe_add(const float* x, const float* pe, float* y, int32_t B, int32_t per, void* stream) {
  RL_REQUIRE(x && pe && y, RL_ERR_NULL, "pe_add: NULL tensor");
  RL_REQUIRE(B > 0 && per > 0, RL_ERR_SHAPE, "pe_add: B=%d per=%d", B, per);
  const long long total = (long long)B * per;
  rl_prof_pre((cudaStream_t)stream);
  pe_add_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(x, pe, y, total, per);
  return rl_check_launch("pe_add_kernel");
}

extern "C" int ralenet_pconv1(const float* x, const float* w, float* y, int32_t B, int32_t C, int32_t L,
                              int32_t transpose, void* stream) {
  RL_REQUIRE(x && w && y, RL_ERR_NULL, "pconv1: NULL tensor");
  RL_REQUIRE(B > 0 && C > 0 && L > 0, RL_ERR_SHAPE, "pconv1: B=%d C=%d L=%d", B, C, L);
  rl_prof_pre((cudaStream_t)stream);
  pconv1_kernel<<<grid1d((long long)B * C * L), 256, 0, (cudaStream_t)stream>>>(x, w, y, B, C, L, transpose);
  return rl_check_launch("pconv1_kernel");
}

extern "C" int ralenet_pconv1_wgrad(const float* dy, const float* x, float* dw, int32_t B, int32_t C, int32_t L,
                                    void* stream) {
  RL_REQUIRE(dy && x && dw, RL_ERR_NULL, "pconv1_wgrad: NULL tensor");
  RL_REQUIRE(B > 0 && C > 0 && L > 0, RL_ERR_SHAPE, "pconv1_wgrad: B=%d C=%d L=%d", B, C, L);
  rl_prof_pre((cudaStream_t)stream);
  pconv1_wgrad_kernel<<<grid1d((long long)B * L), 256, 0, (cudaStream_t)stream>>>(dy, x, dw, B, C, L);
  return rl_check_launch("pconv1_wgrad_kernel");
}
