// attn_core.cuh -- tensor-core attention core for head_dim = 4 (model/transformer.py:289-323: softmax(0.5 q k^T +
// R-wave bias) v, and its backward), operating on one window held in shared memory.
//
// head_dim 4 leaves no room for a conventional K = 16/32 tile, so the split-precision operands are PACKED INTO THE
// UNUSED PART of the instruction shape.  Forward (legacy m16n8k8 TF32 MMA; the backward, further down, packs fp16
// hi/lo pairs into m16n8k16):
//   * contraction over the 4 head dims (S = q k^T):  A = [x_hi | x_lo] (k = 0..3 hi, 4..7 lo),
//     B = [y_hi ; y_hi] then [y_lo ; y_lo]  ->  two MMAs give (x_hi + x_lo)(y_hi + y_lo): fp32-grade logits.
//   * contraction over 8 keys (O = P v), N = 4 head dims:
//     B = [y_hi | y_lo] (n = 0..3 hi, 4..7 lo), A = p_hi then p_lo; the two column halves are added at the end.
//     The accumulator fragment of S (rows g, g+8; cols 2t, 2t+1) is reused directly as the A fragment by
//     contracting over the keys in the order (2t, 2t+1) -> (k = t, t+4), so P never touches shared memory.
// One warp owns one (head, 16-row tile) item and streams the other dimension in tiles of 8.
#pragma once
#include "common.cuh"

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// B fragment register for the [hi | lo]-packed N = 8 operand: lanes g < 4 feed the hi part, g >= 4 the lo part
__device__ __forceinline__ uint32_t pack_hl(float x, bool sel_hi) {
  uint32_t hi, lo;
  split_tf32(x, hi, lo);
  return sel_hi ? hi : lo;
}

// ---------------------------------------------------------------------------------------------
// forward, one (head, 16-query tile) item of one window: the calling warp streams the L keys of the window.
// sq / sk / sv: [L][LD] arrays of the window, the head's 4 channels at column hc; o overwrites q in place.
// stab_h: R-wave table column of this head (element (i - j + W - 1) at stab_h[(i - j + W - 1) * tab_stride]);
// lse_h: [L] log-sum-exp row of this head (global, log2 domain) or NULL.
template <int L, int LD>
__device__ __forceinline__ void attn_core_fwd_item(float* sq, const float* sk, const float* sv, int hc, int i0,
                                                   const float* stab_h, int tab_stride, int W, int c0,
                                                   float* __restrict__ lse_h) {
  constexpr int KT = L / 8, CH = (KT < 4) ? KT : 4;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float qs = 0.5f * RL_LOG2E;                  // head_dim^-0.5 (transformer.py:278) * log2(e)
  const bool sel_hi = g < 4;
  uint32_t qa[4];
  {
    const float* qp = sq + (i0 + g) * LD + hc + t;
    split_tf32x2(qp[0] * qs, qp[8 * LD] * qs, qa[0], qa[1], qa[2], qa[3]);
  }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  const bool qcen = (W > 0) && (i0 + 16 > c0) && (i0 < c0 + W);
  const float* kp = sk + g * LD + hc + t;
  const float* vp = sv + (2 * t) * LD + hc + (g & 3);
#pragma unroll 1
  for (int j0 = 0; j0 < L; j0 += 8 * CH) {
    float s[CH][4];
#pragma unroll
    for (int tt = 0; tt < CH; ++tt) {
      uint32_t kh, kl;
      split_tf32(kp[(j0 + 8 * tt) * LD], kh, kl);
      s[tt][0] = s[tt][1] = s[tt][2] = s[tt][3] = 0.f;
      const uint32_t bh[2] = {kh, kh}, bl[2] = {kl, kl};
      mma_tf32(s[tt], qa, bh);
      mma_tf32(s[tt], qa, bl);
    }
    if (qcen && (j0 + 8 * CH > c0) && (j0 < c0 + W)) {        // R-wave bias on the central W x W block
#pragma unroll
      for (int tt = 0; tt < CH; ++tt) {
        if ((j0 + 8 * tt + 8 <= c0) || (j0 + 8 * tt >= c0 + W)) continue;      // (warp-uniform) key tile outside
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = i0 + g + 8 * (e >> 1), j = j0 + 8 * tt + 2 * t + (e & 1);
          if ((unsigned)(i - c0) < (unsigned)W && (unsigned)(j - c0) < (unsigned)W)
            s[tt][e] += stab_h[(i - j + W - 1) * tab_stride];
        }
      }
    }
    float x0 = fmaxf(s[0][0], s[0][1]), x1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
    for (int tt = 1; tt < CH; ++tt) {
      x0 = fmaxf(x0, fmaxf(s[tt][0], s[tt][1]));
      x1 = fmaxf(x1, fmaxf(s[tt][2], s[tt][3]));
    }
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1));
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    const float n0 = fmaxf(m0, x0), n1 = fmaxf(m1, x1);
    const float r0 = fast_ex2(m0 - n0), r1 = fast_ex2(m1 - n1);
    // (packed fp32 arithmetic: same operations and rounding as the scalar form, half the issue slots)
    mul_f32x2(l0, l1, r0, r1, l0, l1);
    mul_f32x2(o[0], o[1], r0, r0, o[0], o[1]);
    mul_f32x2(o[2], o[3], r1, r1, o[2], o[3]);
    m0 = n0; m1 = n1;
#pragma unroll
    for (int tt = 0; tt < CH; ++tt) {
      float d0, d1, d2, d3;
      sub_f32x2(s[tt][0], s[tt][1], n0, n0, d0, d1);
      sub_f32x2(s[tt][2], s[tt][3], n1, n1, d2, d3);
      const float p0 = fast_ex2(d0), p1 = fast_ex2(d1), p2 = fast_ex2(d2), p3 = fast_ex2(d3);
      float q0, q1;
      add_f32x2(p0, p2, p1, p3, q0, q1);          // (p0 + p1, p2 + p3)
      add_f32x2(l0, l1, q0, q1, l0, l1);
      uint32_t ah[4], al[4];
      split_tf32x2(p0, p2, ah[0], ah[1], al[0], al[1]);
      split_tf32x2(p1, p3, ah[2], ah[3], al[2], al[3]);
      const float* vq = vp + (j0 + 8 * tt) * LD;
      const uint32_t b[2] = {pack_hl(vq[0], sel_hi), pack_hl(vq[LD], sel_hi)};
      mma_tf32(o, ah, b);
      mma_tf32(o, al, b);
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
#pragma unroll
  for (int e = 0; e < 4; ++e) o[e] += __shfl_xor_sync(0xffffffffu, o[e], 2);     // hi columns + lo columns
  __syncwarp();     // o overwrites the q rows other lanes of this warp read at the top (racecheck: write-after-read)
  if (t < 2) {
    const float v0 = 1.0f / l0, v1 = 1.0f / l1;
    *reinterpret_cast<float2*>(sq + (i0 + g) * LD + hc + 2 * t) = make_float2(o[0] * v0, o[1] * v0);
    *reinterpret_cast<float2*>(sq + (i0 + g + 8) * LD + hc + 2 * t) = make_float2(o[2] * v1, o[3] * v1);
  }
  if (lse_h != nullptr && t == 0) {
    lse_h[i0 + g] = m0 + log2f(l0);
    lse_h[i0 + g + 8] = m1 + log2f(l1);
  }
}

// forward over one window held as [L][ld_mk(C)] arrays: one warp per (head, 16-query tile) item.
// lse_g (global, [H][L] of this window, log2 domain) may be NULL
template <int C, int L>
__device__ __forceinline__ void attn_core_fwd(float* sq, const float* sk, const float* sv, const float* stab, int W,
                                              int c0, float* __restrict__ lse_g) {
  constexpr int H = C / RL_HD, LDC = ld_mk(C), NW = RL_NT / 32;
  constexpr int QT = L / 16, NITEM = H * QT;
  const int warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int item = warp; item < NITEM; item += NW) {
    const int h = item / QT, i0 = (item % QT) * 16;
    attn_core_fwd_item<L, LDC>(sq, sk, sv, 4 * h, i0, stab + h, H, W, c0, lse_g ? lse_g + h * L : nullptr);
  }
}

// =============================================================================================
// Single-pass backward core on fp16 hi/lo PAIRS (m16n8k16).
//
// Every operand x is carried as x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits, more than the
// truncated-tf32 split above (gradients are pre-scaled per window by a power of two so that fp16 cannot underflow).
// Four values are stored as four packed words  [hi(d0,d1) | hi(d2,d3) | lo(d0,d1) | lo(d2,d3)]  in the 16 bytes the
// fp32 head slice occupied ("row form": q, k, v, dO converted IN PLACE), and the 16 k-slots of one m16n8k16 MMA hold
// the four cross products of a head_dim-4 contraction:
//      slots 0-3: a_hi b_hi   4-7: a_hi b_lo   8-11: a_lo b_hi   12-15: a_lo b_lo        (one MMA = full product)
// One warp owns a (head, 16-key tile) and walks the queries once, 16 at a time:
//      S^T  = k q^T, dP^T = v dO^T                          (2 MMAs per 8 queries; rows = keys, cols = queries)
//      p    = exp2(S^T - lse), dS^T = p (dP^T - D)
//      dv  += P^T dO, dk += dS^T q                           (A = the accumulator fragments re-packed as hi / lo pairs)
//      dq^T = k^T dS^T                                       (B = dS^T pairs transposed in registers with movmatrix)
// The dq partial sums of the key tiles of one head are added into shared memory WITHOUT atomics (fp32 shared atomics
// are CAS loops): the warps that work on the same head walk the query blocks in a rotated order -- warp with key tile
// jt visits query block (jt + step) mod NQB -- and meet at a named barrier after every step, so at any time each
// query block of a head has exactly one writer.
// S and dP are computed once (a query-major pass for dq plus a key-major pass for dk, dv computed them twice) and no
// operand is split inside the loop.
#include <cuda_fp16.h>
#include <type_traits>

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
// (x0, x1) -> packed fp16 pairs hi = (fp16(x0), fp16(x1)), lo = the fp16 remainders; x0 in the low half
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  float r0, r1;
  sub_f32x2(x0, x1, __low2float(h), __high2float(h), r0, r1);
  const __half2 l = __floats2half2_rn(r0, r1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// convert the [rows][C] fp32 array (stride ld) to row form in place, scaling by `scale`
template <int C>
__device__ __forceinline__ void to_row_form(float* s, int ld, int rows, float scale) {
  constexpr int H = C / RL_HD;
  for (int i = threadIdx.x; i < rows * H; i += RL_NT) {
    float* p = s + (i / H) * ld + 4 * (i % H);
    const float4 v = *reinterpret_cast<const float4*>(p);
    uint4 o;
    split_h2(v.x * scale, v.y * scale, o.x, o.z);
    split_h2(v.z * scale, v.w * scale, o.y, o.w);
    *reinterpret_cast<uint4*>(p) = o;
  }
}
// one register of a "pairs along rows" operand built from two row-form words: the fp16 of dim-slot `slot`
// (0-3: hi of dims 0-3, 4-7: lo) of rows r and r+1 (row r in the low half)
// (the byte selector is a per-lane constant: one PRMT with a register selector instead of two predicated ones)
__device__ __forceinline__ uint32_t col_pair_sel(int slot) { return (slot & 1) ? 0x7632u : 0x5410u; }
__device__ __forceinline__ uint32_t col_pair(const uint32_t* base, int ld, int r, int slot, uint32_t sel) {
  const int word = (slot >> 2) * 2 + ((slot & 3) >> 1);
  return __byte_perm(base[r * ld + word], base[(r + 1) * ld + word], sel);
}

// QP (pre-scaled by 0.5 log2e), KP, VP, DP (pre-scaled by do_scale): row-form arrays; sD = -do_scale * D and
// sLse = -lse: both NEGATED by the caller (8 negations per step less in the walk);
// sdq: fp32, ZEROED, receives 0.5 dS k (plain read-modify-write, see above); sdk, sdv: fp32 outputs.
// sds_c (may be NULL): [H][W][W] fp32, receives dS of the central block for the gradient of the R-wave table -- plain
// stores; the diagonals are summed after the core (attn.cu).  (Round 2 added every element to its table entry with a
// shared-memory atomic: fp32 shared atomics are compare-and-swap loops, up to 8 lanes of a warp hit the same diagonal,
// and the 7 other key-tile warps of the head waited at the step barrier for the warp that owned a central tile.)
template <int C, int L>
__device__ __forceinline__ void attn_core_bwd_single(const float* QPf, const float* KPf, const float* VPf,
                                                     const float* DPf, const float* sD, const float* sLse, float* sdq,
                                                     float* sdk, float* sdv, const float* stab, float* sds_c,
                                                     int W, int c0, float inv_do_scale) {
  constexpr int H = C / RL_HD, LDC = ld_mk(C), NW = RL_NT / 32;
  constexpr int JT = L / 16, NITEM = H * JT;
  static_assert(NITEM % NW == 0 && (JT >= NW ? JT % NW == 0 : NW % JT == 0), "bwd core: items must tile the warps");
  constexpr int GW = (JT < NW) ? JT : NW;                       // warps that share a head within one round
  const uint32_t* QP = reinterpret_cast<const uint32_t*>(QPf);
  const uint32_t* KP = reinterpret_cast<const uint32_t*>(KPf);
  const uint32_t* VP = reinterpret_cast<const uint32_t*>(VPf);
  const uint32_t* DP = reinterpret_cast<const uint32_t*>(DPf);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const uint32_t psel = col_pair_sel(g);
  const float f_dq = 0.5f * inv_do_scale;                       // dq = 0.5 dS k
  const float f_dk = inv_do_scale / RL_LOG2E;                   // dk = 0.5 dS q, q carried as q * 0.5 log2e
#pragma unroll 1
  for (int item = warp; item < NITEM; item += NW) {
    const int h = item / JT, j0 = (item % JT) * 16;
    uint32_t ka[4], va[4], kt[4];
    {
      const int o0 = (j0 + g) * LDC + 4 * h + (t & 1);
      ka[0] = KP[o0]; ka[1] = KP[o0 + 8 * LDC]; ka[2] = KP[o0 + 2]; ka[3] = KP[o0 + 8 * LDC + 2];
      va[0] = VP[o0]; va[1] = VP[o0 + 8 * LDC]; va[2] = VP[o0 + 2]; va[3] = VP[o0 + 8 * LDC + 2];
      // k^T for dq^T = k^T dS^T: rows = dim slots (g < 8), k = keys (2t, 2t+1) and (2t+8, 2t+9); rows 8-15 unused
      kt[0] = col_pair(KP + 4 * h, LDC, j0 + 2 * t, g, psel);
      kt[2] = col_pair(KP + 4 * h, LDC, j0 + 2 * t + 8, g, psel);
      kt[1] = kt[3] = 0u;
    }
    float ak[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    const bool kcen = (W > 0) && (j0 + 16 > c0) && (j0 < c0 + W);
    const float* lsep = sLse + h * L + 2 * t;
    const float* Dp = sD + h * L + 2 * t;
    // the walk over the query blocks, compiled twice: key tiles outside the central block of the R-wave bias (most of
    // them) run a loop without any of the table logic (it was 7 % of the executed instructions of attn_bwd<32>)
    auto walk = [&](auto kcen_c) {
    constexpr bool KCEN = decltype(kcen_c)::value;
#pragma unroll 1
    for (int step = 0; step < JT; ++step) {
      const int i0 = (((item % JT) + step) % JT) * 16;          // rotated walk: one writer per query block and step
      uint32_t pah[4], pal[4], dah[4], dal[4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int ib = i0 + 8 * hf;
        // the accumulators start at -lse / -D (stored negated), so the MMAs deliver S^T - lse and dP^T - D directly
        const float2 ls = *reinterpret_cast<const float2*>(lsep + ib);
        const float2 Dd = *reinterpret_cast<const float2*>(Dp + ib);
        float s[4] = {ls.x, ls.y, ls.x, ls.y}, dp[4] = {Dd.x, Dd.y, Dd.x, Dd.y};
        {
          const uint32_t bq = QP[(ib + g) * LDC + 4 * h + t], bd = DP[(ib + g) * LDC + 4 * h + t];
          const uint32_t b0[2] = {bq, bq}, b1[2] = {bd, bd};
          mma_f16(s, ka, b0);
          mma_f16(dp, va, b1);
        }
        const bool cen = KCEN && (ib + 8 > c0) && (ib < c0 + W);
        if (cen) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = ib + 2 * t + (e & 1), j = j0 + g + 8 * (e >> 1);
            if ((unsigned)(i - c0) < (unsigned)W && (unsigned)(j - c0) < (unsigned)W) s[e] += stab[(i - j + W - 1) * H + h];
          }
        }
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) p[e] = fast_ex2(s[e]);
        mul_f32x2(p[0], p[1], dp[0], dp[1], ds[0], ds[1]);
        mul_f32x2(p[2], p[3], dp[2], dp[3], ds[2], ds[3]);
        if (cen && sds_c != nullptr) {     // dS of the central W x W block: every (h, i, j) has exactly one writer
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = ib + 2 * t + (e & 1), j = j0 + g + 8 * (e >> 1);
            if ((unsigned)(i - c0) < (unsigned)W && (unsigned)(j - c0) < (unsigned)W)
              sds_c[(h * W + (i - c0)) * W + (j - c0)] = ds[e] * inv_do_scale;
          }
        }
        // accumulator fragments -> A fragments of the query contraction (rows = keys, k = the queries of this half)
        split_h2(p[0], p[1], pah[2 * hf], pal[2 * hf]);
        split_h2(p[2], p[3], pah[2 * hf + 1], pal[2 * hf + 1]);
        split_h2(ds[0], ds[1], dah[2 * hf], dal[2 * hf]);
        split_h2(ds[2], ds[3], dah[2 * hf + 1], dal[2 * hf + 1]);
        // dq^T[dim slot][query] = sum_keys k^T dS^T: transpose the 8x8 (key, query) pair blocks into B fragments
        {
          const uint32_t bh[2] = {movmatrix_trans(dah[2 * hf]), movmatrix_trans(dah[2 * hf + 1])};
          const uint32_t bl[2] = {movmatrix_trans(dal[2 * hf]), movmatrix_trans(dal[2 * hf + 1])};
          float dq[4] = {0.f, 0.f, 0.f, 0.f};
          mma_f16(dq, kt, bh);
          mma_f16(dq, kt, bl);
          const float v0 = dq[0] + __shfl_xor_sync(0xffffffffu, dq[0], 16);     // hi-slot rows + lo-slot rows
          const float v1 = dq[1] + __shfl_xor_sync(0xffffffffu, dq[1], 16);
          // lanes g < 4 own (query 2t, dim g), lanes g >= 4 own (query 2t+1, dim g-4); exclusive by construction
          float* dst = sdq + (ib + 2 * t + (g >> 2)) * LDC + 4 * h + (g & 3);
          *dst += ((g < 4) ? v0 : v1) * f_dq;
        }
      }
      // dv += P^T dO, dk += dS^T q over the 16 queries: B = (query pairs) x (dim slots)
      {
        const uint32_t bd[2] = {col_pair(DP + 4 * h, LDC, i0 + 2 * t, g, psel),
                                col_pair(DP + 4 * h, LDC, i0 + 2 * t + 8, g, psel)};
        const uint32_t bq[2] = {col_pair(QP + 4 * h, LDC, i0 + 2 * t, g, psel),
                                col_pair(QP + 4 * h, LDC, i0 + 2 * t + 8, g, psel)};
        mma_f16(av, pah, bd);
        mma_f16(ak, dah, bq);
        mma_f16(av, pal, bd);
        mma_f16(ak, dal, bq);
      }
      if (GW > 1) {                                             // the GW warps of this head move to their next block
        if (GW == NW) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / GW), "n"(GW * 32) : "memory");
      }
    }
    };
    if (kcen) walk(std::true_type{}); else walk(std::false_type{});
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ak[e] += __shfl_xor_sync(0xffffffffu, ak[e], 2);          // hi dim slots (t = 0, 1) + lo dim slots (t = 2, 3)
      av[e] += __shfl_xor_sync(0xffffffffu, av[e], 2);
    }
    if (t < 2) {
      const int off = (j0 + g) * LDC + 4 * h + 2 * t;
      *reinterpret_cast<float2*>(sdk + off) = make_float2(f_dk * ak[0], f_dk * ak[1]);
      *reinterpret_cast<float2*>(sdk + off + 8 * LDC) = make_float2(f_dk * ak[2], f_dk * ak[3]);
      *reinterpret_cast<float2*>(sdv + off) = make_float2(inv_do_scale * av[0], inv_do_scale * av[1]);
      *reinterpret_cast<float2*>(sdv + off + 8 * LDC) = make_float2(inv_do_scale * av[2], inv_do_scale * av[3]);
    }
  }
}
