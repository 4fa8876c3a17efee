// comm.cu -- data-parallel exchange steps of the RA-LENet training step as hand-written kernels over NVLink 5 /
// NVSwitch peer memory (no NCCL on the step's critical path).
//
// The reference has no distributed code at all (main.py:1-3 picks one GPU), so the contract is
// "N ranks on a batch sharded in equal slices == one process on the global batch" (SURVEY.md section 8e).  The path
// has exactly three exchange steps per training step:
//   (1) the 17 BatchNorm forward statistics of the stem (model/transformer.py:570-574; sum, sum of squares, count),
//   (2) the 16 BatchNorm backward sums (+ the loss scalar, for logging),
//   (3) ONE sum of the flat gradient buffer (1,087,282 floats = 4.35 MB) in front of Adam (denoise_train.py:57).
// Every rank owns one SYMMETRIC buffer (allocated and exchanged by the host: torch symmetric memory = cuMemCreate +
// peer mappings, optionally a multicast (NVLS) mapping); the gradient kernels of the backward pass accumulate
// straight into its first `n` floats, so nothing is copied before the exchange:
//
//   [0, n)                      flat gradient buffer
//   [slot_off, +2*8*32)         BN exchange slots: set (fwd / bwd) x writer rank x 32 floats
//   [flag_off, ...)             uint32 barrier flags: (2 + RL_COMM_MAXG) barriers x 8 writer ranks
//
// ralenet_comm_allreduce_adam is the fused compute + collective kernel: CTA b of rank r
//   barrier A   (CTA b of every rank has arrived: all gradients are final)
//   reduce      chunk (b, r) of the buffer over all ranks -- `multimem.ld_reduce` through the switch when a multicast
//               mapping exists, else peer loads in rank order -- and broadcast of the sum into the buffers of all ranks
//               (`multimem.st`, else peer stores): every rank ends with bit-identical sums
//   barrier B   (the chunks (b, 0..W-1) have landed here; nobody reads this rank's buffer any more)
//   Adam        on chunks (b, 0..W-1) = one contiguous slice, from the local copy (torch.optim.Adam defaults).
// Both barriers are per-CTA flag exchanges (release stores into the peers' flag words, acquire spins on the local
// ones) with monotonically increasing epochs kept in device memory, so the kernel replays from a CUDA graph with no
// host involvement.  The grid is at most RL_COMM_MAXG <= 148 CTAs: all co-resident, so the spins cannot deadlock;
// a spin that does not complete within ~4 s traps instead of hanging the GPU.
#define RL_NT 512
#include "common.cuh"

namespace {

constexpr int MAXW = RL_COMM_MAXW;
constexpr int SLOT = 32;                       // floats per exchange slot

struct CommDev {
  float* peer[MAXW];
  float* mc;
  int world, rank;
  unsigned long long slot_off, flag_off;       // in floats
};

__host__ __device__ inline unsigned long long slot_off_of(unsigned long long n) { return (n + 63) & ~63ull; }
__host__ __device__ inline unsigned long long flag_off_of(unsigned long long n) { return slot_off_of(n) + 2 * MAXW * SLOT; }
inline unsigned long long total_floats(unsigned long long n) { return flag_off_of(n) + (2 + RL_COMM_MAXG) * MAXW + 64; }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float ld_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// NVLS: one load returns the sum over the copies of every rank, reduced inside the switch
__device__ __forceinline__ float4 mc_ld_reduce_v4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_v4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Barrier number `slot` among the CTAs that call it with the same slot on every rank.  Writes of this CTA issued
// before the call are visible to the peers' CTAs after it.
__device__ void xbarrier(const CommDev& c, float* const* peers, int slot, uint32_t epoch) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    uint32_t* remote = reinterpret_cast<uint32_t*>(peers[threadIdx.x] + c.flag_off) + slot * MAXW + c.rank;
    st_release_sys(remote, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(peers[c.rank] + c.flag_off) + slot * MAXW + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > (1ll << 33)) __trap();      // ~4 s at 2 GHz: a rank is missing; fail instead of hanging
    }
  }
  __syncthreads();
}

// ---- (1), (2): sum of n <= 32 floats over the ranks, in rank order (identical bits everywhere) ------------------
__global__ void __launch_bounds__(256) comm_exchange_kernel(CommDev c, int set, float* __restrict__ vals, int n,
                                                            uint32_t* __restrict__ epoch_ctr) {
  __shared__ float* peers[MAXW];
  if (threadIdx.x < MAXW) peers[threadIdx.x] = c.peer[threadIdx.x];
  const uint32_t e = epoch_ctr[set] + 1;
  __syncthreads();
  for (int idx = threadIdx.x; idx < c.world * n; idx += blockDim.x) {
    const int p = idx / n, i = idx - p * n;
    st_sys(peers[p] + c.slot_off + (size_t)(set * MAXW + c.rank) * SLOT + i, vals[i]);
  }
  xbarrier(c, peers, set, e);
  if ((int)threadIdx.x < n) {
    const float* base = peers[c.rank] + c.slot_off + (size_t)set * MAXW * SLOT + threadIdx.x;
    float s = 0.f;
    for (int q = 0; q < c.world; ++q) s += ld_sys(base + q * SLOT);
    vals[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) epoch_ctr[set] = e;
}

// ---- (3): gradient all-reduce fused in front of Adam -----------------------------------------------------------
__global__ void __launch_bounds__(RL_NT) comm_allreduce_adam_kernel(CommDev c, long long n4, float* __restrict__ p,
                                                                    float* __restrict__ m, float* __restrict__ v,
                                                                    float lr, float b1, float b2, float eps,
                                                                    const int32_t* __restrict__ step_dev, float gscale,
                                                                    uint32_t* __restrict__ epoch_ctr) {
  __shared__ float* peers[MAXW];
  if (threadIdx.x < MAXW) peers[threadIdx.x] = c.peer[threadIdx.x];
  const int b = blockIdx.x, G = gridDim.x, W = c.world;
  const uint32_t e0 = epoch_ctr[2 + b];
  const float s = (float)(*step_dev);
  const float bc1 = 1.f - powf(b1, s), bc2_sqrt = sqrtf(1.f - powf(b2, s));
  __syncthreads();
  xbarrier(c, peers, 2 + b, e0 + 1);

  // float4 index range of chunk j of G*W equal chunks
  const long long per = (n4 + (long long)G * W - 1) / ((long long)G * W);
  auto lo_of = [&](long long j) { const long long x = j * per; return x < n4 ? x : n4; };
  {
    const long long lo = lo_of((long long)b * W + c.rank), hi = lo_of((long long)b * W + c.rank + 1);
    if (c.mc != nullptr) {
      for (long long i = lo + threadIdx.x; i < hi; i += RL_NT) mc_st_v4(c.mc + 4 * i, mc_ld_reduce_v4(c.mc + 4 * i));
    } else {
      for (long long i = lo + threadIdx.x; i < hi; i += RL_NT) {
        float4 acc = ld_sys_v4(peers[0] + 4 * i);
        for (int q = 1; q < W; ++q) {
          const float4 t = ld_sys_v4(peers[q] + 4 * i);
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        for (int q = 0; q < W; ++q) st_sys_v4(peers[q] + 4 * i, acc);
      }
    }
  }
  xbarrier(c, peers, 2 + b, e0 + 2);

  const float* g = peers[c.rank];
  const long long lo = lo_of((long long)b * W), hi = lo_of((long long)(b + 1) * W);
  for (long long i = lo + threadIdx.x; i < hi; i += RL_NT) {
    const float4 g4 = ld_sys_v4(g + 4 * i);
    float4 p4 = *reinterpret_cast<float4*>(p + 4 * i);
    float4 m4 = *reinterpret_cast<float4*>(m + 4 * i);
    float4 v4 = *reinterpret_cast<float4*>(v + 4 * i);
    float* pp = &p4.x;
    const float* gg = &g4.x;
    float* mm = &m4.x;
    float* vv = &v4.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {          // same arithmetic as adam_kernel (stem_head.cu)
      const float gr = gg[k] * gscale;
      mm[k] = b1 * mm[k] + (1.f - b1) * gr;
      vv[k] = b2 * vv[k] + (1.f - b2) * gr * gr;
      pp[k] -= (lr / bc1) * mm[k] / (sqrtf(vv[k]) / bc2_sqrt + eps);
    }
    *reinterpret_cast<float4*>(p + 4 * i) = p4;
    *reinterpret_cast<float4*>(m + 4 * i) = m4;
    *reinterpret_cast<float4*>(v + 4 * i) = v4;
  }
  if (threadIdx.x == 0) epoch_ctr[2 + b] = e0 + 2;
}

__global__ void comm_step_inc_kernel(int32_t* s) { *s += 1; }

// CUDA loads kernels lazily on first launch, and that load may wait for running kernels.  A rank whose exchange kernel
// is spinning on its peers must never have its NEXT comm kernel blocked in the loader (with several ranks emulated in
// one process that is a deadlock), so every kernel of this file is loaded before the first one is launched.
int preload() {
  static thread_local int done_dev = -1;
  int dev = -1;
  cudaGetDevice(&dev);
  if (done_dev == dev) return RL_OK;
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, comm_exchange_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, comm_allreduce_adam_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, comm_step_inc_kernel);
  RL_REQUIRE(e == cudaSuccess, RL_ERR_CUDA, "comm: kernel preload failed: %s", cudaGetErrorString(e));
  done_dev = dev;
  return RL_OK;
}

int fill(CommDev* d, const rl_comm* c, const char* who) {
  if (int rc = preload()) return rc;
  RL_REQUIRE(c && c->epoch, RL_ERR_NULL, "%s: NULL comm / epoch counters", who);
  RL_REQUIRE(c->world >= 1 && c->world <= MAXW && c->rank >= 0 && c->rank < c->world, RL_ERR_SHAPE,
             "%s: world=%d rank=%d (1 <= world <= %d)", who, c->world, c->rank, MAXW);
  RL_REQUIRE(c->n > 0 && c->n % 4 == 0, RL_ERR_SHAPE, "%s: gradient length %llu must be a positive multiple of 4", who,
             (unsigned long long)c->n);
  for (int i = 0; i < MAXW; ++i) {
    d->peer[i] = i < c->world ? (float*)c->peer[i] : nullptr;
    RL_REQUIRE(i >= c->world || (d->peer[i] && (uintptr_t)d->peer[i] % 16 == 0), RL_ERR_NULL,
               "%s: peer mapping %d missing or misaligned", who, i);
  }
  d->mc = (float*)c->mc;
  d->world = c->world;
  d->rank = c->rank;
  d->slot_off = slot_off_of(c->n);
  d->flag_off = flag_off_of(c->n);
  return RL_OK;
}

}  // namespace

extern "C" uint64_t ralenet_comm_bytes(uint64_t n) { return (uint64_t)total_floats(n) * sizeof(float); }

extern "C" int ralenet_comm_exchange(const rl_comm* c, int32_t set, float* vals, int32_t n, void* stream) {
  CommDev d;
  if (int rc = fill(&d, c, "comm_exchange")) return rc;
  RL_REQUIRE(vals && n > 0 && n <= SLOT && (set == 0 || set == 1), RL_ERR_SHAPE, "comm_exchange: set=%d n=%d", set, n);
  rl_prof_pre((cudaStream_t)stream);
  comm_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(d, set, vals, n, c->epoch);
  return rl_check_launch("comm_exchange_kernel");
}

extern "C" int ralenet_comm_allreduce_adam(const rl_comm* c, float* p, float* m, float* v, float lr, float beta1,
                                           float beta2, float eps, int32_t* step_dev, float gscale, int32_t grid,
                                           void* stream) {
  CommDev d;
  if (int rc = fill(&d, c, "comm_allreduce_adam")) return rc;
  RL_REQUIRE(p && m && v && step_dev, RL_ERR_NULL, "comm_allreduce_adam: NULL tensor");
  RL_REQUIRE(((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) % 16 == 0, RL_ERR_SHAPE,
             "comm_allreduce_adam: buffers must be 16-byte aligned");
  if (grid <= 0) grid = RL_COMM_MAXG;
  RL_REQUIRE(grid <= RL_COMM_MAXG, RL_ERR_SHAPE, "comm_allreduce_adam: grid %d > %d (all CTAs must be co-resident)",
             grid, RL_COMM_MAXG);
  cudaStream_t st = (cudaStream_t)stream;
  rl_prof_pre(st);
  comm_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
  if (int rc = rl_check_launch("step_inc_kernel")) return rc;
  rl_prof_pre(st);
  comm_allreduce_adam_kernel<<<grid, RL_NT, 0, st>>>(d, (long long)(c->n / 4), p, m, v, lr, beta1, beta2, eps, step_dev,
                                                     gscale, c->epoch);
  return rl_check_launch("comm_allreduce_adam_kernel");
}
