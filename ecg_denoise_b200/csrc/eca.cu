// eca.cu -- the two small per-window ops either side of the hot path that SURVEY.md section 8f lists:
//   * eca_layer_1d (model/transformer.py:100-113): channel gate on the feed-forward output, forward + backward
//   * single_snr_noise_add (local_utils/local_utils.py:176-192): SNR-targeted noise mixing (data synthesis prologue)
// One CTA per window; a window is 8 KB, so both are latency/HBM-bound streaming kernels (2-3 tensor crossings).
#include "common.cuh"

namespace {

constexpr int ECA_MAX_K = 15, ECA_MAX_C = 256;

// column means of x[L][C] over t into sm[C] (C divides RL_NT): thread = (row group, column)
__device__ __forceinline__ void column_sums(const float* __restrict__ x, const float* __restrict__ g, int L, int C,
                                            float* sm /* [C], zeroed */) {
  const int c = threadIdx.x % C, r0 = threadIdx.x / C, RG = RL_NT / C;
  float acc = 0.f;
  for (int t = r0; t < L; t += RG) {
    const float v = __ldg(x + (size_t)t * C + c);
    acc += g ? v * __ldg(g + (size_t)t * C + c) : v;
  }
  atomicAdd(&sm[c], acc);
}

__global__ void __launch_bounds__(RL_NT) eca_fwd_kernel(const rl_eca_fwd_args a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sm[ECA_MAX_C], ss[ECA_MAX_C], swk[ECA_MAX_K];
  const int L = a.L, C = a.C, K = a.K, P = (K - 1) / 2, tid = threadIdx.x;
  const size_t off = (size_t)blockIdx.x * L * C;
  const float* xw = a.x + off;
  for (int i = tid; i < C; i += RL_NT) sm[i] = 0.f;
  if (tid < K) swk[tid] = __ldg(a.w + tid);
  __syncthreads();
  column_sums(xw, nullptr, L, C, sm);
  __syncthreads();
  for (int c = tid; c < C; c += RL_NT) {
    float z = 0.f;
    for (int k = 0; k < K; ++k) {
      const int j = c + k - P;
      if (j >= 0 && j < C) z = fmaf(swk[k], sm[j] * (1.0f / L), z);
    }
    const float s = 1.0f / (1.0f + expf(-z));
    ss[c] = s;
    if (a.s) a.s[(size_t)blockIdx.x * C + c] = s;
  }
  __syncthreads();
  const float* rw = a.res ? a.res + off : nullptr;
  float* yw = a.y + off;
  for (int i = tid; i < L * C; i += RL_NT) {
    const float v = __ldg(xw + i) * ss[i % C];
    yw[i] = rw ? v + __ldg(rw + i) : v;
  }
}

__global__ void __launch_bounds__(RL_NT) eca_bwd_kernel(const rl_eca_bwd_args a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sm[ECA_MAX_C], sds[ECA_MAX_C], sdz[ECA_MAX_C], sdm[ECA_MAX_C], ss[ECA_MAX_C], swk[ECA_MAX_K];
  const int L = a.L, C = a.C, K = a.K, P = (K - 1) / 2, tid = threadIdx.x;
  const size_t off = (size_t)blockIdx.x * L * C;
  const float* xw = a.x + off;
  const float* gw = a.g + off;
  for (int i = tid; i < C; i += RL_NT) {
    sm[i] = 0.f;
    sds[i] = 0.f;
    ss[i] = __ldg(a.s + (size_t)blockIdx.x * C + i);
  }
  if (tid < K) swk[tid] = __ldg(a.w + tid);
  __syncthreads();
  column_sums(xw, nullptr, L, C, sm);     // L * mean
  column_sums(xw, gw, L, C, sds);         // dL/ds = sum_t g x
  __syncthreads();
  for (int c = tid; c < C; c += RL_NT) sdz[c] = sds[c] * ss[c] * (1.0f - ss[c]);
  __syncthreads();
  for (int c = tid; c < C; c += RL_NT) {  // dm[c] = sum_k w[k] dz[c - k + P]
    float d = 0.f;
    for (int k = 0; k < K; ++k) {
      const int j = c - k + P;
      if (j >= 0 && j < C) d = fmaf(swk[k], sdz[j], d);
    }
    sdm[c] = d * (1.0f / L);
  }
  if (a.d_w && tid < K) {                 // dw[k] = sum_c dz[c] m[c + k - P]
    float d = 0.f;
    for (int c = 0; c < C; ++c) {
      const int j = c + tid - P;
      if (j >= 0 && j < C) d = fmaf(sdz[c], sm[j] * (1.0f / L), d);
    }
    atomicAdd(a.d_w + tid, d);
  }
  __syncthreads();
  float* dxw = a.dx + off;
  for (int i = tid; i < L * C; i += RL_NT) dxw[i] = fmaf(__ldg(gw + i), ss[i % C], sdm[i % C]);
}

__global__ void __launch_bounds__(RL_NT) snr_mix_kernel(const float* __restrict__ data, const float* __restrict__ noise,
                                                        const float* __restrict__ snr_db, float* __restrict__ out,
                                                        int per) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_red[32];
  const size_t off = (size_t)blockIdx.x * per;
  float ps = 0.f, pn = 0.f;
  for (int i = threadIdx.x; i < per; i += RL_NT) {
    const float d = __ldg(data + off + i), n = __ldg(noise + off + i);
    ps = fmaf(d, d, ps);
    pn = fmaf(n, n, pn);
  }
  ps = block_sum(ps, s_red);
  pn = block_sum(pn, s_red);
  // scale = sqrt(target_noise_energy / noise_energy), target = signal_energy / 10^(snr/10)   (:184-189)
  const float scale = sqrtf(ps / (exp10f(__ldg(snr_db + blockIdx.x) * 0.1f) * pn));
  for (int i = threadIdx.x; i < per; i += RL_NT) out[off + i] = fmaf(__ldg(noise + off + i), scale, __ldg(data + off + i));
}

int eca_check(int B, int L, int C, int K) {
  RL_REQUIRE(B > 0 && L > 0, RL_ERR_SHAPE, "eca: B=%d L=%d", B, L);
  RL_REQUIRE(C > 0 && C <= ECA_MAX_C && RL_NT % C == 0, RL_ERR_SHAPE, "eca: C=%d must divide %d", C, RL_NT);
  RL_REQUIRE(K >= 1 && K <= ECA_MAX_K && (K & 1), RL_ERR_SHAPE, "eca: kernel size %d must be odd and <= %d", K, ECA_MAX_K);
  return RL_OK;
}

}  // namespace

extern "C" int ralenet_eca_fwd(const rl_eca_fwd_args* a, void* stream) {
  RL_REQUIRE(a, RL_ERR_NULL, "eca_fwd: args is NULL");
  if (int rc = eca_check(a->B, a->L, a->C, a->K)) return rc;
  RL_REQUIRE(a->x && a->w && a->y, RL_ERR_NULL, "eca_fwd: NULL tensor");
  rl_launch_pdl(eca_fwd_kernel, dim3(a->B), dim3(RL_NT), 0, (cudaStream_t)stream, *a);
  return rl_check_launch("eca_fwd_kernel", a->C);
}

extern "C" int ralenet_eca_bwd(const rl_eca_bwd_args* a, void* stream) {
  RL_REQUIRE(a, RL_ERR_NULL, "eca_bwd: args is NULL");
  if (int rc = eca_check(a->B, a->L, a->C, a->K)) return rc;
  RL_REQUIRE(a->g && a->x && a->w && a->s && a->dx, RL_ERR_NULL, "eca_bwd: NULL tensor");
  rl_launch_pdl(eca_bwd_kernel, dim3(a->B), dim3(RL_NT), 0, (cudaStream_t)stream, *a);
  return rl_check_launch("eca_bwd_kernel", a->C);
}

extern "C" int ralenet_snr_mix(const float* data, const float* noise, const float* snr_db, float* out, int32_t B,
                               int32_t per, void* stream) {
  RL_REQUIRE(data && noise && snr_db && out, RL_ERR_NULL, "snr_mix: NULL tensor");
  RL_REQUIRE(B > 0 && per > 0, RL_ERR_SHAPE, "snr_mix: B=%d per=%d", B, per);
  rl_launch_pdl(snr_mix_kernel, dim3(B), dim3(RL_NT), 0, (cudaStream_t)stream, data, noise, snr_db, out, (int)per);
  return rl_check_launch("snr_mix_kernel");
}
