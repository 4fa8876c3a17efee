// ffn_cluster.cu -- feed-forward half for the WIDE stages (C = 64, 128) on thread-block clusters.
//
// At the wide stages a window has only L = 32 / 16 tokens but the weights are large (fc1 + fc2 = 512 KB at
// C = 128): one CTA per window would stream all of them through shared memory for 16 rows of work.  Here a
// cluster of CL = 4 CTAs owns 4 consecutive windows (TM = 4L tokens) and SPLITS THE HIDDEN DIMENSION: CTA r
// computes hidden units [r*4C/CL, (r+1)*4C/CL) for all TM tokens (fc1 slice -> GELU -> local enhancement ->
// GELU), multiplies by its K-slice of fc2, and the four partial outputs are reduced through distributed shared
// memory (DSMEM): CTA r sums window r's rows from all four CTAs and finishes that window (residual / LayerNorm
// backward).  Per CTA the weight stream is 4x shorter, the MMA tiles are 4x taller, L2 weight traffic drops 4x,
// and no global atomics or extra kernels are needed.
//
// Same arithmetic as ffn.cu (reference model/transformer.py:392-395, 149-161, 54-59); modes RL_LE_NONE and
// RL_LE_PARTIAL (the convolved hidden channel 0 lives in slice 0).  Depthwise mode and 512-sample windows use ffn.cu.
#define RL_NT 512
#define RL_MINB 2
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int CL = 4;   // cluster size = windows per cluster = hidden-dimension split

template <int C>
struct Geo {
  static constexpr int L = 2048 / C;          // tokens per window
  static constexpr int TM = CL * L;           // tokens per cluster
  static constexpr int HC = 4 * C;
  static constexpr int TH = HC / CL;          // hidden units per CTA
  static constexpr int LDC = ld_mk(C);        // row stride of [TM][C] tiles
  static constexpr int LDT = ld_mk(TH);       // row stride of [TM][TH] tiles
};

template <int C>
__host__ __device__ constexpr int fwd_swf() {
  return cmax(WStream<Geo<C>::TH, C, B_NK>::FLOATS, WStream<C, Geo<C>::TH, B_NK>::FLOATS);
}
template <int C>
constexpr size_t fwd_smem() {
  return sizeof(float) * ((size_t)Geo<C>::TM * Geo<C>::LDC + (size_t)Geo<C>::TM * Geo<C>::LDT + fwd_swf<C>() +
                          Geo<C>::TM + 64);
}

template <int C>
__global__ void __launch_bounds__(RL_NT, RL_MINB) ffn_fwd_cluster_kernel(const rl_ffn_fwd_args a) {
  pdl_wait();
  pdl_trigger();
  using G = Geo<C>;
  constexpr int L = G::L, TM = G::TM, HC = G::HC, TH = G::TH, LDC = G::LDC, LDT = G::LDT;
  extern __shared__ __align__(16) float smem[];
  float* su = smem;                 // LN2 output [TM][LDC]; later the partial fc2 output of this CTA
  float* sh = su + TM * LDC;        // hidden slice [TM][LDT]
  float* sw = sh + TM * LDT;
  float* sfir = sw + fwd_swf<C>();
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int w0 = (blockIdx.x / CL) * CL;                 // first window of this cluster
  const int nvalid = min(CL, a.B - w0) * L;              // valid token rows in this cluster
  const int tid = threadIdx.x;
  const size_t tok0 = (size_t)w0 * L;
  const float* xw = a.x + tok0 * C;

  // 1. LN2 over all TM tokens (each CTA of the cluster needs the full rows)
  if (a.flags & RL_F_PRENORM) {
    const float* lw = a.ln_w;
    const float* lb = a.ln_b;
    ln_forward_rows<C>(
        TM, [&](int t, int c) { return t < nvalid ? __ldg(xw + t * C + c) : 0.f; },
        [&](int t, int c, float zh) { su[t * LDC + c] = fmaf(zh, __ldg(lw + c), __ldg(lb + c)); });
  } else {
    for (int i = tid; i < TM * C; i += RL_NT) {
      const int t = i / C, c = i % C;
      su[t * LDC + c] = t < nvalid ? __ldg(xw + i) : 0.f;
    }
  }
  __syncthreads();

  // 2. hidden slice: h = u W1[r*TH : (r+1)*TH]^T + b1;  g1 = GELU(h) -> sh
  {
    MmaTile<TM, TH> acc;
    acc.init();
    WStream<TH, C, B_NK>::run(acc, su, LDC, sw, a.w1 + (size_t)r * TH * C, TH, nullptr, C);
    const float* b1 = a.b1 ? a.b1 + r * TH : nullptr;
    float* hs = a.h ? a.h + tok0 * HC + r * TH : nullptr;
    acc.epilogue([&](int t, int n, float v) {
      v += b1 ? __ldg(b1 + n) : 0.f;
      if (hs && t < nvalid) hs[(size_t)t * HC + n] = v;
      sh[t * LDT + n] = gelu_f(v);
    });
  }
  __syncthreads();

  // 3. local enhancement (3-tap FIR along the tokens of each window on hidden channel 0 = column 0 of slice 0)
  //    and the second GELU
  if (a.le_mode == RL_LE_PARTIAL) {
    if (r == 0) {
      const float w0f = __ldg(a.lew), w1f = __ldg(a.lew + 1), w2f = __ldg(a.lew + 2);
      for (int t = tid; t < TM; t += RL_NT) {
        const int tl = t % L;
        const float p = (tl > 0) ? sh[(t - 1) * LDT] : 0.f;
        const float n = (tl + 1 < L) ? sh[(t + 1) * LDT] : 0.f;
        sfir[t] = w0f * p + w1f * sh[t * LDT] + w2f * n;
      }
    }
    __syncthreads();
    for (int i = tid; i < TM * TH; i += RL_NT) {
      const int t = i / TH, n = i % TH;
      const float f = (r == 0 && n == 0) ? sfir[t] : sh[t * LDT + n];
      sh[t * LDT + n] = gelu_f(f);
    }
    __syncthreads();
  }

  // 4. partial output of this hidden slice: yp = g2 W2[:, r*TH : (r+1)*TH]^T   -> su (LN output is dead)
  {
    MmaTile<TM, C> acc;
    acc.init();
    WStream<C, TH, B_NK>::run(acc, sh, LDT, sw, a.w2 + (size_t)r * TH, C, nullptr, HC);
    acc.epilogue([&](int t, int n, float v) { su[t * LDC + n] = v; });
  }

  // 5. DSMEM reduction: CTA r finishes window r
  cluster.sync();
  {
    const float* part[CL];
#pragma unroll
    for (int q = 0; q < CL; ++q) part[q] = cluster.map_shared_rank(su, q);
    const bool mine = (w0 + r) < a.B;
    const float* b2 = a.b2;
    const float* xr = xw + (size_t)r * L * C;
    const float* ex = a.extra ? a.extra + (tok0 + (size_t)r * L) * C : nullptr;
    float* yw = a.y + (tok0 + (size_t)r * L) * C;
    const bool resid = a.flags & RL_F_RESIDUAL;
    if (mine) {
      for (int i = tid; i < L * (C / 4); i += RL_NT) {
        const int t = i / (C / 4), c = (i % (C / 4)) * 4;
        const int off = (r * L + t) * LDC + c;
        float4 s = *reinterpret_cast<const float4*>(part[0] + off);
#pragma unroll
        for (int q = 1; q < CL; ++q) {
          const float4 p = *reinterpret_cast<const float4*>(part[q] + off);
          s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
        if (b2) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(b2 + c));
          s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
        }
        if (resid) {
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(xr + t * C + c));
          s.x += x4.x; s.y += x4.y; s.z += x4.z; s.w += x4.w;
        }
        if (ex) {
          const float4 e4 = __ldg(reinterpret_cast<const float4*>(ex + t * C + c));
          s.x += e4.x; s.y += e4.y; s.z += e4.z; s.w += e4.w;
        }
        *reinterpret_cast<float4*>(yw + t * C + c) = s;
      }
    }
  }
  cluster.sync();     // nobody may exit while its partial tile is still being read
}

// ---------------------------------------------------------------------------------------------
template <int C>
__host__ __device__ constexpr int bwd_swf() {
  return cmax(WStream<Geo<C>::TH, C, B_KN>::FLOATS, WStream<C, Geo<C>::TH, B_KN>::FLOATS);
}
template <int C>
constexpr size_t bwd_smem() {
  return sizeof(float) * ((size_t)Geo<C>::TM * Geo<C>::LDC + (size_t)Geo<C>::TM * Geo<C>::LDT + bwd_swf<C>() +
                          3 * Geo<C>::TM + 2 * C + 64);
}

template <int C>
__global__ void __launch_bounds__(RL_NT, RL_MINB) ffn_bwd_cluster_kernel(const rl_ffn_bwd_args a) {
  pdl_wait();
  pdl_trigger();
  using G = Geo<C>;
  constexpr int L = G::L, TM = G::TM, HC = G::HC, TH = G::TH, LDC = G::LDC, LDT = G::LDT;
  extern __shared__ __align__(16) float smem[];
  float* sg = smem;                       // dL/dy [TM][LDC]
  float* sp = sg;                         // partial du of this hidden slice (g is dead after the dg2 GEMM)
  float* sd = sg + TM * LDC;              // dh slice [TM][LDT]
  float* sw = sd + TM * LDT;
  float* sg10 = sw + bwd_swf<C>();        // g1[:,0]   (rank 0)
  float* sf0 = sg10 + TM;                 // fir(g1[:,0])
  float* sdf0 = sf0 + TM;                 // df[:,0]
  float* s_gb = sdf0 + TM;
  float* s_red = s_gb + 2 * C;
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int w0 = (blockIdx.x / CL) * CL;
  const int nvalid = min(CL, a.B - w0) * L;
  const int tid = threadIdx.x;
  const size_t tok0 = (size_t)w0 * L;
  const float* gw = a.g + tok0 * C;
  const float* xw = a.x + tok0 * C;
  const float* hw = a.h + tok0 * HC + r * TH;          // this CTA's hidden slice, row stride HC
  float* g2w = a.g2 + tok0 * HC + r * TH;
  float* dhw = a.dh + tok0 * HC + r * TH;
  const int mode = a.le_mode;
  const bool part0 = (mode == RL_LE_PARTIAL) && (r == 0);

  // 1. g -> sg; the convolved channel's g1 and FIR output (rank 0)
  for (int i = tid; i < TM * (C / 4); i += RL_NT) {
    const int t = i / (C / 4), c = (i % (C / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < nvalid) v = __ldg(reinterpret_cast<const float4*>(gw + t * C + c));
    *reinterpret_cast<float4*>(sg + t * LDC + c) = v;
  }
  for (int i = tid; i < 2 * C; i += RL_NT) s_gb[i] = 0.f;
  float lw0 = 0.f, lw1 = 0.f, lw2 = 0.f;
  if (part0) {
    lw0 = __ldg(a.lew); lw1 = __ldg(a.lew + 1); lw2 = __ldg(a.lew + 2);
    for (int t = tid; t < TM; t += RL_NT) sg10[t] = t < nvalid ? gelu_f(__ldg(hw + (size_t)t * HC)) : 0.f;
    __syncthreads();
    for (int t = tid; t < TM; t += RL_NT) {
      const int tl = t % L;
      const float p = (tl > 0) ? sg10[t - 1] : 0.f;
      const float n = (tl + 1 < L) ? sg10[t + 1] : 0.f;
      sf0[t] = lw0 * p + lw1 * sg10[t] + lw2 * n;
    }
  }
  __syncthreads();

  // 2. dg2 slice = g W2[:, slice]  (K = C), then GELU' / FIR^T / GELU' -> dh slice
  {
    MmaTile<TM, TH> acc;
    acc.init();
    WStream<TH, C, B_KN>::run(acc, sg, LDC, sw, a.w2 + (size_t)r * TH, 1 << 30, nullptr, HC);
    acc.epilogue([&](int t, int n, float v) {
      const bool ok = t < nvalid;
      float g1, d1;
      gelu_both(ok ? __ldg(hw + (size_t)t * HC + n) : 0.f, g1, d1);
      if (mode == RL_LE_NONE) {
        const float dh = v * d1;
        sd[t * LDT + n] = dh;
        if (ok) { g2w[(size_t)t * HC + n] = g1; dhw[(size_t)t * HC + n] = dh; }
      } else if (part0 && n == 0) {
        float g2, d2;
        gelu_both(sf0[t], g2, d2);
        if (ok) g2w[(size_t)t * HC] = g2;
        sdf0[t] = v * d2;
      } else {
        float g2, d2;
        gelu_both(g1, g2, d2);
        const float dh = v * d2 * d1;
        sd[t * LDT + n] = dh;
        if (ok) { g2w[(size_t)t * HC + n] = g2; dhw[(size_t)t * HC + n] = dh; }
      }
    });
    __syncthreads();
    if (part0) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      for (int t = tid; t < TM; t += RL_NT) {
        const int tl = t % L;
        const float d = sdf0[t];
        const float dn = (tl + 1 < L) ? sdf0[t + 1] : 0.f;
        const float dp = (tl > 0) ? sdf0[t - 1] : 0.f;
        const float dg1 = lw0 * dn + lw1 * d + lw2 * dp;            // adjoint of the 3-tap FIR
        const bool ok = t < nvalid;
        const float dh = ok ? dg1 * gelu_grad_f(__ldg(hw + (size_t)t * HC)) : 0.f;
        sd[t * LDT] = dh;
        if (ok) dhw[(size_t)t * HC] = dh;
        a0 += d * ((tl > 0) ? sg10[t - 1] : 0.f);
        a1 += d * sg10[t];
        a2 += d * ((tl + 1 < L) ? sg10[t + 1] : 0.f);
      }
      if (a.d_lew) {
        a0 = block_sum(a0, s_red);
        a1 = block_sum(a1, s_red);
        a2 = block_sum(a2, s_red);
        if (tid == 0) {
          atomicAdd(a.d_lew, a0);
          atomicAdd(a.d_lew + 1, a1);
          atomicAdd(a.d_lew + 2, a2);
        }
      }
    }
    __syncthreads();
  }

  // 3. partial du = dh_slice W1[slice rows]   (K = TH)
  {
    MmaTile<TM, C> acc;
    acc.init();
    WStream<C, TH, B_KN>::run(acc, sd, LDT, sw, a.w1 + (size_t)r * TH * C, 1 << 30, nullptr, C);
    acc.epilogue([&](int t, int n, float v) { sp[t * LDC + n] = v; });
  }

  // 4. DSMEM reduction + LayerNorm backward: CTA r finishes window r
  cluster.sync();
  {
    const float* part[CL];
#pragma unroll
    for (int q = 0; q < CL; ++q) part[q] = cluster.map_shared_rank(sp, q);
    const bool mine = (w0 + r) < a.B;
    const int row0 = r * L;
    float* dxw = a.dx + (tok0 + (size_t)row0) * C;
    float* uw = a.u + (tok0 + (size_t)row0) * C;
    const float* xr = xw + (size_t)row0 * C;
    const float* gr = gw + (size_t)row0 * C;
    const bool resid = a.flags & RL_F_RESIDUAL;
    auto du_at = [&](int t, int c) {
      const int off = (row0 + t) * LDC + c;
      float s = part[0][off];
#pragma unroll
      for (int q = 1; q < CL; ++q) s += part[q][off];
      return s;
    };
    if (mine) {
      if (a.flags & RL_F_PRENORM) {
        const float* lw = a.ln_w;
        const float* lb = a.ln_b;
        ln_backward_rows<C>(
            L, lw, s_gb, [&](int t, int c) { return __ldg(xr + t * C + c); }, du_at,
            [&](int t, int c, float dz, float zh) {
              dxw[t * C + c] = (resid ? __ldg(gr + t * C + c) : 0.f) + dz;
              uw[t * C + c] = fmaf(zh, __ldg(lw + c), __ldg(lb + c));
            });
      } else {
        for (int i = tid; i < L * C; i += RL_NT) {
          const int t = i / C, c = i % C;
          dxw[i] = du_at(t, c) + (resid ? __ldg(gr + i) : 0.f);
          uw[i] = __ldg(xr + i);
        }
      }
    }
    __syncthreads();
    if (mine && (a.flags & RL_F_PRENORM) && a.d_ln_w)
      for (int i = tid; i < C; i += RL_NT) {
        atomicAdd(a.d_ln_w + i, s_gb[i]);
        atomicAdd(a.d_ln_b + i, s_gb[C + i]);
      }
  }
  cluster.sync();
}

template <typename Args>
int launch_cluster(void (*kernel)(const Args), size_t smem, int B, const Args& a, cudaStream_t st, const char* name,
                   int C) {
  if (int rc = rl_set_smem(kernel, smem)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((B + CL - 1) / CL) * CL);
  cfg.blockDim = dim3(RL_NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  rl_prof_pre(st);
  cudaLaunchKernelEx(&cfg, kernel, a);
  return rl_check_launch(name, C);
}

}  // namespace

// returns 1 if the shape/mode is not handled here (caller falls through to the per-window kernels in ffn.cu)
int rl_ffn_fwd_cluster(const rl_ffn_fwd_args* a, cudaStream_t st) {
  if (a->L * a->C != 2048 || a->le_mode == RL_LE_DEPTHWISE) return 1;
  if (a->C == 128) return launch_cluster(ffn_fwd_cluster_kernel<128>, fwd_smem<128>(), a->B, *a, st, "ffn_fwd_cluster", 128);
  if (a->C == 64) return launch_cluster(ffn_fwd_cluster_kernel<64>, fwd_smem<64>(), a->B, *a, st, "ffn_fwd_cluster", 64);
  return 1;
}

int rl_ffn_bwd_cluster(const rl_ffn_bwd_args* a, cudaStream_t st) {
  if (a->L * a->C != 2048 || a->le_mode == RL_LE_DEPTHWISE) return 1;
  if (a->C == 128) return launch_cluster(ffn_bwd_cluster_kernel<128>, bwd_smem<128>(), a->B, *a, st, "ffn_bwd_cluster", 128);
  if (a->C == 64) return launch_cluster(ffn_bwd_cluster_kernel<64>, bwd_smem<64>(), a->B, *a, st, "ffn_bwd_cluster", 64);
  return 1;
}
