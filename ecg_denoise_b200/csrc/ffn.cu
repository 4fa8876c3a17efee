// ffn.cu -- local-enhancement feed-forward half of the RA-LENet TransformerBlock, one CTA per window.
//
//   y = x + fc2( GELU( leconv( GELU( fc1( LN2(x) ) ) ) ) )  (+ extra)
//
// Reference: model/transformer.py:392-395 (forward_part2), :149-161 (Mlp.forward), :54-59
// (PartialConv_1d.forward_split_cat), :146 (depthwise variant).  As shipped the "local enhancement" is a
// 3-tap FIR along the tokens on hidden channel 0 only, followed by a second GELU on all channels
// (SURVEY.md F5); RL_LE_DEPTHWISE applies per-channel FIRs to all 4C channels; RL_LE_NONE is the plain MLP.
// The L x 4C hidden activation (8192 floats per window) stays in shared memory between the two GEMMs.
#include "common.cuh"

namespace {

template <int C>
size_t ffn_fwd_smem(int L) {
  return sizeof(float) * ((size_t)L * lda_of<C>() + (size_t)L * (4 * C + 1) + SW_FLOATS + (size_t)L + 64);
}

template <int C, int WIN>
__global__ void __launch_bounds__(RL_NT) ffn_fwd_kernel(const rl_ffn_fwd_args a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LDA = lda_of<C>();
  constexpr int HC = 4 * C, LDH = HC + 1;
  const int L = a.L;
  float* su = smem;
  float* sh = su + L * LDA;
  float* sw = sh + L * LDH;
  float* sfir = sw + SW_FLOATS;
  const int tid = threadIdx.x;
  const size_t woff = (size_t)blockIdx.x * L * C;
  const float* xw = a.x + woff;

  // 1. LN2
  if (a.flags & RL_F_PRENORM) {
    const float* lw = a.ln_w;
    const float* lb = a.ln_b;
    ln_forward_rows<C>(
        L, [&](int t, int c) { return __ldg(xw + t * C + c); },
        [&](int t, int c, float zh) { su[t * LDA + c] = fmaf(zh, __ldg(lw + c), __ldg(lb + c)); });
  } else {
    for (int i = tid; i < L * C; i += RL_NT) su[(i / C) * LDA + (i % C)] = __ldg(xw + i);
  }
  __syncthreads();

  // 2. h = u W1^T + b1 (N = 4C, K = C);  g1 = GELU(h) -> sh
  {
    TileAcc<4 * WIN, 8> acc;
    acc.init(L, HC);
    const int KC = min(C, pow2_floor(SW_FLOATS / LDH));
    for (int k0 = 0; k0 < C; k0 += KC) {
      stage_wT(sw, LDH, a.w1, C, 0, HC, k0, KC);
      __syncthreads();
      acc.mac(su + k0, LDA, 1, sw, LDH, KC);
      __syncthreads();
    }
    const float* b1 = a.b1;
    float* hs = a.h ? a.h + (size_t)blockIdx.x * L * HC : nullptr;
    acc.epilogue([&](int t, int n, float v) {
      v += b1 ? __ldg(b1 + n) : 0.f;
      if (hs) hs[t * HC + n] = v;
      sh[t * LDH + n] = gelu_f(v);
    });
  }
  __syncthreads();

  // 3. local enhancement + second GELU
  if (a.le_mode == RL_LE_PARTIAL) {
    const float w0 = __ldg(a.lew), w1 = __ldg(a.lew + 1), w2 = __ldg(a.lew + 2);
    for (int t = tid; t < L; t += RL_NT) {
      const float p = (t > 0) ? sh[(t - 1) * LDH] : 0.f;
      const float n = (t + 1 < L) ? sh[(t + 1) * LDH] : 0.f;
      sfir[t] = w0 * p + w1 * sh[t * LDH] + w2 * n;
    }
    __syncthreads();
    for (int i = tid; i < L * HC; i += RL_NT) {
      const int t = i / HC, n = i % HC;
      const float f = (n == 0) ? sfir[t] : sh[t * LDH + n];
      sh[t * LDH + n] = gelu_f(f);
    }
    __syncthreads();
  } else if (a.le_mode == RL_LE_DEPTHWISE) {
    for (int c = tid; c < HC; c += RL_NT) {
      const float w0 = __ldg(a.lew + 3 * c), w1 = __ldg(a.lew + 3 * c + 1), w2 = __ldg(a.lew + 3 * c + 2);
      float prev = 0.f, cur = sh[c];
      for (int t = 0; t < L; ++t) {
        const float nxt = (t + 1 < L) ? sh[(t + 1) * LDH + c] : 0.f;
        sh[t * LDH + c] = gelu_f(w0 * prev + w1 * cur + w2 * nxt);
        prev = cur;
        cur = nxt;
      }
    }
    __syncthreads();
  }

  // 4. y = x + g2 W2^T + b2 (N = C, K = 4C)
  {
    TileAcc<4 * WIN, 2> acc;
    acc.init(L, C);
    const int ldd = C + 1;
    const int KC = min(HC, pow2_floor(SW_FLOATS / ldd));
    for (int k0 = 0; k0 < HC; k0 += KC) {
      stage_wT(sw, ldd, a.w2, HC, 0, C, k0, KC);
      __syncthreads();
      acc.mac(sh + k0, LDH, 1, sw, ldd, KC);
      __syncthreads();
    }
    const float* b2 = a.b2;
    const float* ex = a.extra ? a.extra + woff : nullptr;
    float* yw = a.y + woff;
    const bool resid = a.flags & RL_F_RESIDUAL;
    acc.epilogue([&](int t, int n, float v) {
      v += b2 ? __ldg(b2 + n) : 0.f;
      if (resid) v += __ldg(xw + t * C + n);
      if (ex) v += __ldg(ex + t * C + n);
      yw[t * C + n] = v;
    });
  }
}

// ---------------------------------------------------------------------------------------------
template <int C>
size_t ffn_bwd_smem(int L) {
  return sizeof(float) * (2 * (size_t)L * lda_of<C>() + 2 * (size_t)L * (4 * C + 1) + SW_FLOATS + 3 * (size_t)L +
                          2 * C + 64);
}

template <int C, int WIN>
__global__ void __launch_bounds__(RL_NT) ffn_bwd_kernel(const rl_ffn_bwd_args a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int LDA = lda_of<C>();
  constexpr int HC = 4 * C, LDH = HC + 1;
  const int L = a.L;
  float* sg = smem;                       // dL/dy, A operand
  float* su = sg + L * LDA;               // du (LN2 output gradient)
  float* sh = su + L * LDA;               // g1 = GELU(h)
  float* sd = sh + L * LDH;               // df / dh
  float* sw = sd + L * LDH;
  float* sg10 = sw + SW_FLOATS;           // g1[:,0]
  float* sf0 = sg10 + L;                  // fir(g1[:,0])
  float* sdf0 = sf0 + L;                  // df[:,0]
  float* s_gb = sdf0 + L;
  float* s_red = s_gb + 2 * C;
  const int tid = threadIdx.x;
  const size_t woff = (size_t)blockIdx.x * L * C;
  const size_t hoff = (size_t)blockIdx.x * L * HC;
  const float* gw = a.g + woff;
  const float* xw = a.x + woff;
  const float* hw = a.h + hoff;
  const int mode = a.le_mode;

  // 1. g -> sg, g1 = GELU(h) -> sh
  for (int i = tid; i < L * C; i += RL_NT) sg[(i / C) * LDA + (i % C)] = __ldg(gw + i);
  for (int i = tid; i < L * HC; i += RL_NT) sh[(i / HC) * LDH + (i % HC)] = gelu_f(__ldg(hw + i));
  for (int i = tid; i < 2 * C; i += RL_NT) s_gb[i] = 0.f;
  __syncthreads();

  // 2. second-GELU input f and g2 = GELU(f) (written to scratch for the fc2 weight gradient)
  float lw0 = 0.f, lw1 = 0.f, lw2 = 0.f;
  if (mode == RL_LE_PARTIAL) {
    lw0 = __ldg(a.lew); lw1 = __ldg(a.lew + 1); lw2 = __ldg(a.lew + 2);
    for (int t = tid; t < L; t += RL_NT) {
      const float p = (t > 0) ? sh[(t - 1) * LDH] : 0.f;
      const float n = (t + 1 < L) ? sh[(t + 1) * LDH] : 0.f;
      sg10[t] = sh[t * LDH];
      sf0[t] = lw0 * p + lw1 * sh[t * LDH] + lw2 * n;
    }
    __syncthreads();
  }
  {
    float* g2w = a.g2 + hoff;
    for (int i = tid; i < L * HC; i += RL_NT) {
      const int t = i / HC, n = i % HC;
      float v = sh[t * LDH + n];
      if (mode == RL_LE_PARTIAL) {
        v = gelu_f(n == 0 ? sf0[t] : v);
      } else if (mode == RL_LE_DEPTHWISE) {
        const float p = (t > 0) ? sh[(t - 1) * LDH + n] : 0.f;
        const float nx = (t + 1 < L) ? sh[(t + 1) * LDH + n] : 0.f;
        v = gelu_f(__ldg(a.lew + 3 * n) * p + __ldg(a.lew + 3 * n + 1) * v + __ldg(a.lew + 3 * n + 2) * nx);
      }
      g2w[i] = v;
    }
  }

  // 3. dg2 = g W2 (N = 4C, K = C; B(k,n) = W2[k][n] natural layout), then through GELU'/FIR^T/GELU'
  {
    TileAcc<4 * WIN, 8> acc;
    acc.init(L, HC);
    const int KC = min(C, pow2_floor(SW_FLOATS / LDH));
    for (int k0 = 0; k0 < C; k0 += KC) {
      stage_w(sw, LDH, a.w2, HC, k0, KC, 0, HC);
      __syncthreads();
      acc.mac(sg + k0, LDA, 1, sw, LDH, KC);
      __syncthreads();
    }
    float* dhw = a.dh + hoff;
    acc.epilogue([&](int t, int n, float v) {
      if (mode == RL_LE_NONE) {
        const float dh = v * gelu_grad_f(__ldg(hw + t * HC + n));
        sd[t * LDH + n] = dh;
        dhw[t * HC + n] = dh;
      } else if (mode == RL_LE_PARTIAL) {
        if (n == 0) {
          sdf0[t] = v * gelu_grad_f(sf0[t]);
        } else {
          const float dh = v * gelu_grad_f(sh[t * LDH + n]) * gelu_grad_f(__ldg(hw + t * HC + n));
          sd[t * LDH + n] = dh;
          dhw[t * HC + n] = dh;
        }
      } else {
        const float g1 = sh[t * LDH + n];
        const float p = (t > 0) ? sh[(t - 1) * LDH + n] : 0.f;
        const float nx = (t + 1 < L) ? sh[(t + 1) * LDH + n] : 0.f;
        const float f = __ldg(a.lew + 3 * n) * p + __ldg(a.lew + 3 * n + 1) * g1 + __ldg(a.lew + 3 * n + 2) * nx;
        sd[t * LDH + n] = v * gelu_grad_f(f);          // df, finished below
      }
    });
    __syncthreads();
    if (mode == RL_LE_PARTIAL) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
      for (int t = tid; t < L; t += RL_NT) {
        const float d = sdf0[t];
        const float dn = (t + 1 < L) ? sdf0[t + 1] : 0.f;
        const float dp = (t > 0) ? sdf0[t - 1] : 0.f;
        const float dg1 = lw0 * dn + lw1 * d + lw2 * dp;            // adjoint of the 3-tap FIR
        const float dh = dg1 * gelu_grad_f(__ldg(hw + t * HC));
        sd[t * LDH] = dh;
        dhw[t * HC] = dh;
        a0 += d * ((t > 0) ? sg10[t - 1] : 0.f);
        a1 += d * sg10[t];
        a2 += d * ((t + 1 < L) ? sg10[t + 1] : 0.f);
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
      __syncthreads();
    } else if (mode == RL_LE_DEPTHWISE) {
      for (int c = tid; c < HC; c += RL_NT) {
        const float w0 = __ldg(a.lew + 3 * c), w1 = __ldg(a.lew + 3 * c + 1), w2 = __ldg(a.lew + 3 * c + 2);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        float dprev = 0.f, dcur = sd[c];
        for (int t = 0; t < L; ++t) {
          const float dnxt = (t + 1 < L) ? sd[(t + 1) * LDH + c] : 0.f;
          const float dg1 = w0 * dnxt + w1 * dcur + w2 * dprev;
          a0 += dcur * ((t > 0) ? sh[(t - 1) * LDH + c] : 0.f);
          a1 += dcur * sh[t * LDH + c];
          a2 += dcur * ((t + 1 < L) ? sh[(t + 1) * LDH + c] : 0.f);
          const float dh = dg1 * gelu_grad_f(__ldg(hw + t * HC + c));
          sd[t * LDH + c] = dh;
          dhw[t * HC + c] = dh;
          dprev = dcur;
          dcur = dnxt;
        }
        if (a.d_lew) {
          atomicAdd(a.d_lew + 3 * c, a0);
          atomicAdd(a.d_lew + 3 * c + 1, a1);
          atomicAdd(a.d_lew + 3 * c + 2, a2);
        }
      }
      __syncthreads();
    }
  }

  // 4. du = dh W1 (N = C, K = 4C; B(k,n) = W1[k][n] natural layout)
  {
    TileAcc<4 * WIN, 2> acc;
    acc.init(L, C);
    const int ldd = C + 1;
    const int KC = min(HC, pow2_floor(SW_FLOATS / ldd));
    for (int k0 = 0; k0 < HC; k0 += KC) {
      stage_w(sw, ldd, a.w1, C, k0, KC, 0, C);
      __syncthreads();
      acc.mac(sd + k0, LDH, 1, sw, ldd, KC);
      __syncthreads();
    }
    acc.epilogue([&](int t, int n, float v) { su[t * LDA + n] = v; });
  }
  __syncthreads();

  // 5. LN2 backward + residual
  float* dxw = a.dx + woff;
  float* uw = a.u + woff;
  const bool resid = a.flags & RL_F_RESIDUAL;
  if (a.flags & RL_F_PRENORM) {
    const float* lw = a.ln_w;
    const float* lb = a.ln_b;
    ln_backward_rows<C>(
        L, lw, s_gb, [&](int t, int c) { return __ldg(xw + t * C + c); }, [&](int t, int c) { return su[t * LDA + c]; },
        [&](int t, int c, float dz, float zh) {
          dxw[t * C + c] = (resid ? sg[t * LDA + c] : 0.f) + dz;
          uw[t * C + c] = fmaf(zh, __ldg(lw + c), __ldg(lb + c));
        });
    __syncthreads();
    if (a.d_ln_w)
      for (int i = tid; i < C; i += RL_NT) {
        atomicAdd(a.d_ln_w + i, s_gb[i]);
        atomicAdd(a.d_ln_b + i, s_gb[C + i]);
      }
  } else {
    for (int i = tid; i < L * C; i += RL_NT) {
      const int t = i / C, c = i % C;
      dxw[i] = su[t * LDA + c] + (resid ? sg[t * LDA + c] : 0.f);
      uw[i] = __ldg(xw + i);
    }
  }
}

template <int C>
int launch_fwd(const rl_ffn_fwd_args* a, cudaStream_t st) {
  const size_t smem = ffn_fwd_smem<C>(a->L);
  if (a->L * C == 2048) {
    if (int rc = rl_set_smem(ffn_fwd_kernel<C, 1>, smem)) return rc;
    ffn_fwd_kernel<C, 1><<<a->B, RL_NT, smem, st>>>(*a);
  } else {
    if (int rc = rl_set_smem(ffn_fwd_kernel<C, 2>, smem)) return rc;
    ffn_fwd_kernel<C, 2><<<a->B, RL_NT, smem, st>>>(*a);
  }
  return rl_check_launch("ffn_fwd_kernel", C);
}

template <int C>
int launch_bwd(const rl_ffn_bwd_args* a, cudaStream_t st) {
  const size_t smem = ffn_bwd_smem<C>(a->L);
  if (a->L * C == 2048) {
    if (int rc = rl_set_smem(ffn_bwd_kernel<C, 1>, smem)) return rc;
    ffn_bwd_kernel<C, 1><<<a->B, RL_NT, smem, st>>>(*a);
  } else {
    if (int rc = rl_set_smem(ffn_bwd_kernel<C, 2>, smem)) return rc;
    ffn_bwd_kernel<C, 2><<<a->B, RL_NT, smem, st>>>(*a);
  }
  return rl_check_launch("ffn_bwd_kernel", C);
}

int check_shape(int B, int L, int C, int le) {
  RL_REQUIRE(B > 0, RL_ERR_SHAPE, "ffn: B=%d", B);
  RL_REQUIRE(C == 8 || C == 16 || C == 32 || C == 64 || C == 128, RL_ERR_SHAPE, "ffn: unsupported C=%d", C);
  RL_REQUIRE(L * C == 2048 || L * C == 4096, RL_ERR_SHAPE, "ffn: L*C must be 2048 or 4096 (L=%d, C=%d)", L, C);
  RL_REQUIRE(le >= RL_LE_NONE && le <= RL_LE_DEPTHWISE, RL_ERR_SHAPE, "ffn: bad le_mode %d", le);
  return RL_OK;
}

}  // namespace

extern "C" int ralenet_ffn_fwd(const rl_ffn_fwd_args* a, void* stream) {
  RL_REQUIRE(a, RL_ERR_NULL, "ffn_fwd: args is NULL");
  if (int rc = check_shape(a->B, a->L, a->C, a->le_mode)) return rc;
  RL_REQUIRE(a->x && a->y && a->w1 && a->w2, RL_ERR_NULL, "ffn_fwd: NULL tensor");
  RL_REQUIRE(!(a->flags & RL_F_PRENORM) || (a->ln_w && a->ln_b), RL_ERR_NULL, "ffn_fwd: prenorm needs ln");
  RL_REQUIRE(a->le_mode == RL_LE_NONE || a->lew, RL_ERR_NULL, "ffn_fwd: le_mode needs lew");
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->C) {
    case 8: return launch_fwd<8>(a, st);
    case 16: return launch_fwd<16>(a, st);
    case 32: return launch_fwd<32>(a, st);
    case 64: return launch_fwd<64>(a, st);
    case 128: return launch_fwd<128>(a, st);
  }
  return RL_ERR_SHAPE;
}

extern "C" int ralenet_ffn_bwd(const rl_ffn_bwd_args* a, void* stream) {
  RL_REQUIRE(a, RL_ERR_NULL, "ffn_bwd: args is NULL");
  if (int rc = check_shape(a->B, a->L, a->C, a->le_mode)) return rc;
  RL_REQUIRE(a->g && a->x && a->w1 && a->w2 && a->h && a->dx && a->dh && a->g2 && a->u, RL_ERR_NULL,
             "ffn_bwd: NULL tensor");
  RL_REQUIRE(!(a->flags & RL_F_PRENORM) || (a->ln_w && a->ln_b), RL_ERR_NULL, "ffn_bwd: prenorm needs ln");
  RL_REQUIRE(a->le_mode == RL_LE_NONE || a->lew, RL_ERR_NULL, "ffn_bwd: le_mode needs lew");
  RL_REQUIRE(!a->d_ln_w == !a->d_ln_b, RL_ERR_NULL, "ffn_bwd: d_ln_w/d_ln_b must be both set or both NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = RL_ERR_SHAPE;
  switch (a->C) {
    case 8: rc = launch_bwd<8>(a, st); break;
    case 16: rc = launch_bwd<16>(a, st); break;
    case 32: rc = launch_bwd<32>(a, st); break;
    case 64: rc = launch_bwd<64>(a, st); break;
    case 128: rc = launch_bwd<128>(a, st); break;
  }
  if (rc) return rc;
  const int M = a->B * a->L, C = a->C;
  if ((rc = rl_launch_wgrad(a->g, C, a->g2, 4 * C, M, C, 4 * C, a->d_w2, a->d_b2, st))) return rc;
  if ((rc = rl_launch_wgrad(a->dh, 4 * C, a->u, C, M, 4 * C, C, a->d_w1, a->d_b1, st))) return rc;
  return RL_OK;
}
