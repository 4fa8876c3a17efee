// wgrad.cu -- weight-gradient GEMM over the token dimension, split-K across CTAs.
//   dW[n*K + k] += sum_m dY[m*ldy + n] * X[m*ldx + k]        db[n] += sum_m dY[m*ldy + n]
// M = all tokens of the batch (B*L: 65536 ... 4096 at B=256), N,K in {4..512}.  The reduction dim is
// the long one, so each CTA owns a (TN x TK) output tile and a slice of MC tokens, accumulates in
// registers and finishes with fp32 red.global (atomicAdd).  HBM-bound: reads (N+K)*4 B per token.
#include "common.cuh"

namespace {

constexpr int MS = 32;   // tokens staged per inner step

template <int TN, int TK>
__global__ void __launch_bounds__(RL_NT) wgrad_kernel(const float* __restrict__ dY, int ldy,
                                                      const float* __restrict__ X, int ldx, int M, int N, int K,
                                                      float* __restrict__ dW, float* __restrict__ db, int MC) {
  constexpr int RM = (TN >= 16) ? TN / 16 : 1;
  constexpr int RN = (TK >= 16) ? TK / 16 : 1;
  constexpr int NTN = TK / RN;                 // threads along k
  constexpr int NTM = TN / RM;                 // threads along n
  constexpr int LDA = TN + 1, LDB = TK + 1;
  __shared__ float sA[MS * LDA];
  __shared__ float sB[MS * LDB];

  const int tiles_k = K / TK;
  const int tile_n = blockIdx.x / tiles_k, tile_k = blockIdx.x % tiles_k;
  const int n_base = tile_n * TN, k_base = tile_k * TK;
  const int m_begin = blockIdx.y * MC;
  const int m_end = min(M, m_begin + MC);

  const int tid = threadIdx.x;
  const bool active = tid < NTN * NTM;
  const int tn = tid % NTN, tm = tid / NTN;
  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;
  float bacc = 0.f;
  const bool do_bias = (db != nullptr) && (tile_k == 0);

  for (int m0 = m_begin; m0 < m_end; m0 += MS) {
    const int rows = min(MS, m_end - m0);
    for (int i = tid; i < MS * TN; i += RL_NT) {
      const int c = i % TN, r = i / TN;
      sA[r * LDA + c] = (r < rows) ? __ldg(dY + (size_t)(m0 + r) * ldy + n_base + c) : 0.f;
    }
    for (int i = tid; i < MS * TK; i += RL_NT) {
      const int c = i % TK, r = i / TK;
      sB[r * LDB + c] = (r < rows) ? __ldg(X + (size_t)(m0 + r) * ldx + k_base + c) : 0.f;
    }
    __syncthreads();
    if (active) micro_gemm<RM, RN>(acc, sA, 1, LDA, sB, LDB, tm * RM, tn, NTN, MS);
    if (do_bias && tid < TN) {
#pragma unroll 8
      for (int r = 0; r < MS; ++r) bacc += sA[r * LDA + tid];
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j)
        atomicAdd(dW + (size_t)(n_base + tm * RM + i) * K + k_base + tn + j * NTN, acc[i][j]);
  }
  if (do_bias && tid < TN) atomicAdd(db + n_base + tid, bacc);
}

template <int TN, int TK>
int launch(const float* dY, int ldy, const float* X, int ldx, int M, int N, int K, float* dW, float* db,
           cudaStream_t st) {
  const int tiles = (N / TN) * (K / TK);
  // aim for ~4 CTAs per SM overall, but never slices shorter than 128 tokens
  int splits = (592 + tiles - 1) / tiles;
  const int max_splits = (M + 127) / 128;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int MC = (M + splits - 1) / splits;
  MC = ((MC + MS - 1) / MS) * MS;
  splits = (M + MC - 1) / MC;
  dim3 grid(tiles, splits);
  wgrad_kernel<TN, TK><<<grid, RL_NT, 0, st>>>(dY, ldy, X, ldx, M, N, K, dW, db, MC);
  return rl_check_launch("wgrad_kernel", N, K);
}

template <int TN>
int dispatch_k(int TK, const float* dY, int ldy, const float* X, int ldx, int M, int N, int K, float* dW, float* db,
               cudaStream_t st) {
  switch (TK) {
    case 4: return launch<TN, 4>(dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 8: return launch<TN, 8>(dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 16: return launch<TN, 16>(dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 32: return launch<TN, 32>(dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 64: return launch<TN, 64>(dY, ldy, X, ldx, M, N, K, dW, db, st);
  }
  rl_set_error("wgrad: unsupported K tile %d", TK);
  return RL_ERR_SHAPE;
}

inline int pick_tile(int n) {
  if (n >= 64 && n % 64 == 0) return 64;
  if (n == 32 || n == 16 || n == 8 || n == 4) return n;
  return 0;
}

}  // namespace

int rl_launch_wgrad(const float* dY, int ldy, const float* X, int ldx, int M, int N, int K, float* dW, float* db,
                    cudaStream_t st) {
  if (dW == nullptr) return RL_OK;
  const int TN = pick_tile(N), TK = pick_tile(K);
  RL_REQUIRE(TN && TK && M > 0, RL_ERR_SHAPE, "wgrad: unsupported shape M=%d N=%d K=%d", M, N, K);
  switch (TN) {
    case 4: return dispatch_k<4>(TK, dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 8: return dispatch_k<8>(TK, dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 16: return dispatch_k<16>(TK, dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 32: return dispatch_k<32>(TK, dY, ldy, X, ldx, M, N, K, dW, db, st);
    case 64: return dispatch_k<64>(TK, dY, ldy, X, ldx, M, N, K, dW, db, st);
  }
  return RL_ERR_SHAPE;
}

extern "C" int ralenet_wgrad(const float* dY, int32_t ldy, const float* X, int32_t ldx, int32_t M, int32_t N, int32_t K,
                             float* dW, float* db, void* stream) {
  RL_REQUIRE(dY && X && dW, RL_ERR_NULL, "wgrad: NULL tensor");
  return rl_launch_wgrad(dY, ldy, X, ldx, M, N, K, dW, db, (cudaStream_t)stream);
}
