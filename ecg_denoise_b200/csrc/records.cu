// records.cu -- record <-> window pipeline around the network for long recordings (inference).
//
// The reference cuts 2-lead records into 256-sample windows on the host with numpy
// (local_utils/local_utils.py:47-65, 116-130) after per-lead z-normalisation (np_norm, :261-266) and never
// stitches them back.  Here: per-(record, lead) statistics, a fused z-norm + window gather, and a
// deterministic overlap-add stitch (gather form, no atomics) that also undoes the normalisation.
// All three are HBM-bound streaming kernels.
#include "common.cuh"

namespace {

// one CTA per (record, lead): mean and 1/std (population std, like numpy's default in np_norm)
__global__ void __launch_bounds__(RL_NT) record_stats_kernel(const float* __restrict__ x, int64_t T,
                                                             float* __restrict__ stats) {
  __shared__ double sd[2 * (RL_NT / 32)];
  const float* xr = x + (size_t)blockIdx.x * T;
  double s = 0.0, q = 0.0;
  for (int64_t i = threadIdx.x; i < T; i += RL_NT) {
    const double v = (double)__ldg(xr + i);
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sd[2 * w] = s; sd[2 * w + 1] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int k = 0; k < RL_NT / 32; ++k) { ts += sd[2 * k]; tq += sd[2 * k + 1]; }
    const double mean = ts / (double)T;
    const double var = fmax(tq / (double)T - mean * mean, 0.0);
    stats[2 * blockIdx.x] = (float)mean;
    stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(var + 1e-12));
  }
}

// win[(r*nper + w)][c][i] = (x[r][c][w*stride + i] - mean[r][c]) * rstd[r][c]
__global__ void __launch_bounds__(RL_NT) window_gather_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                              float* __restrict__ win, int C, int64_t T, int W,
                                                              int stride, int nper, int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * RL_NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * RL_NT) {
    const int i = (int)(idx % W);
    const int64_t rest = idx / W;
    const int c = (int)(rest % C);
    const int64_t wg = rest / C;                 // global window index
    const int w = (int)(wg % nper);
    const int64_t r = wg / nper;
    const size_t rc = (size_t)r * C + c;
    float v = __ldg(x + rc * T + (int64_t)w * stride + i);
    if (stats) v = (v - __ldg(stats + 2 * rc)) * __ldg(stats + 2 * rc + 1);
    win[idx] = v;
  }
}

// y[r][c][t] = mean over the windows covering t of win[...][c][t - w*stride], de-normalised; samples not covered
// by any full window pass the input through unchanged.
__global__ void __launch_bounds__(RL_NT) window_scatter_kernel(const float* __restrict__ win,
                                                               const float* __restrict__ x,
                                                               const float* __restrict__ stats, float* __restrict__ y,
                                                               int C, int64_t T, int W, int stride, int nper,
                                                               int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * RL_NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * RL_NT) {
    const int64_t t = idx % T;
    const size_t rc = (size_t)(idx / T);
    const int c = (int)(rc % C);
    const int64_t r = (int64_t)(rc / C);
    int64_t w_hi = t / stride;
    if (w_hi > nper - 1) w_hi = nper - 1;
    int64_t w_lo = (t - W + stride) / stride;    // ceil((t - W + 1) / stride) for t - W + 1 >= 0
    if (t - W + 1 <= 0) w_lo = 0;
    float s = 0.f;
    int n = 0;
    for (int64_t w = w_lo; w <= w_hi; ++w) {
      const int64_t off = t - w * stride;
      if (off >= 0 && off < W) {
        s += __ldg(win + (((size_t)(r * nper + w)) * C + c) * W + off);
        ++n;
      }
    }
    float out;
    if (n > 0) {
      out = s / (float)n;
      if (stats) out = out / __ldg(stats + 2 * rc + 1) + __ldg(stats + 2 * rc);
    } else {
      out = __ldg(x + idx);
    }
    y[idx] = out;
  }
}

}  // namespace

extern "C" int ralenet_record_stats(const float* x, int32_t RC, int64_t T, float* stats, void* stream) {
  RL_REQUIRE(x && stats, RL_ERR_NULL, "record_stats: NULL tensor");
  RL_REQUIRE(RC > 0 && T > 0, RL_ERR_SHAPE, "record_stats: RC=%d T=%lld", RC, (long long)T);
  record_stats_kernel<<<RC, RL_NT, 0, (cudaStream_t)stream>>>(x, T, stats);
  return rl_check_launch("record_stats_kernel");
}

static int rec_check(int32_t R, int32_t C, int64_t T, int32_t W, int32_t stride, int32_t* nper) {
  RL_REQUIRE(R > 0 && C > 0 && T >= W && W > 0 && stride > 0 && stride <= W, RL_ERR_SHAPE,
             "records: unsupported R=%d C=%d T=%lld W=%d stride=%d", R, C, (long long)T, W, stride);
  *nper = (int32_t)((T - W) / stride + 1);
  return RL_OK;
}

extern "C" int32_t ralenet_windows_per_record(int64_t T, int32_t W, int32_t stride) {
  if (T < W || W <= 0 || stride <= 0) return 0;
  return (int32_t)((T - W) / stride + 1);
}

extern "C" int ralenet_window_gather(const float* x, const float* stats, float* win, int32_t R, int32_t C, int64_t T,
                                     int32_t W, int32_t stride, void* stream) {
  RL_REQUIRE(x && win, RL_ERR_NULL, "window_gather: NULL tensor");
  int32_t nper = 0;
  if (int rc = rec_check(R, C, T, W, stride, &nper)) return rc;
  const int64_t total = (int64_t)R * nper * C * W;
  const int blocks = (int)((total + RL_NT * 4 - 1) / (RL_NT * 4) < 148 * 32 ? (total + RL_NT * 4 - 1) / (RL_NT * 4) : 148 * 32);
  window_gather_kernel<<<blocks < 1 ? 1 : blocks, RL_NT, 0, (cudaStream_t)stream>>>(x, stats, win, C, T, W, stride, nper,
                                                                                    total);
  return rl_check_launch("window_gather_kernel");
}

extern "C" int ralenet_window_scatter(const float* win, const float* x, const float* stats, float* y, int32_t R,
                                      int32_t C, int64_t T, int32_t W, int32_t stride, void* stream) {
  RL_REQUIRE(win && x && y, RL_ERR_NULL, "window_scatter: NULL tensor");
  int32_t nper = 0;
  if (int rc = rec_check(R, C, T, W, stride, &nper)) return rc;
  const int64_t total = (int64_t)R * C * T;
  const int blocks = (int)((total + RL_NT * 4 - 1) / (RL_NT * 4) < 148 * 32 ? (total + RL_NT * 4 - 1) / (RL_NT * 4) : 148 * 32);
  window_scatter_kernel<<<blocks < 1 ? 1 : blocks, RL_NT, 0, (cudaStream_t)stream>>>(win, x, stats, y, C, T, W, stride,
                                                                                     nper, total);
  return rl_check_launch("window_scatter_kernel");
}
