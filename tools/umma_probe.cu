// umma_probe.cu -- stand-alone probe of the tcgen05 (UMMA) path used by the RA-LENet kernels:
//   * shared-memory matrix descriptors for the un-swizzled canonical layouts (K-major and MN-major),
//   * kind::tf32 operand handling (is the fp32 container truncated?) and the 3-pass split,
//   * where the rows of an M = 64 / M = 128 accumulator land in tensor memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
// Every wait is bounded, so a wrong descriptor can not hang the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Cfg {
  int M, N, K;       // UMMA M (64/128), N, total K (multiple of 8)
  int a_mn, b_mn;    // 0 = K-major, 1 = MN-major canonical layout
  int swap;          // swap the LBO / SBO fields of the descriptors
  int split;         // 1 = 3-pass split (hi*hi + lo*hi + hi*lo)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}

// byte offset of element (r, k) of an R x KT operand tile in the canonical un-swizzled layouts
__host__ __device__ inline int off_kmajor(int r, int k, int KT) {     // core matrix = 8 rows x 16 B, K-adjacent cores contiguous
  return (r % 8) * 16 + (k % 4) * 4 + (k / 4) * 128 + (r / 8) * (KT / 4) * 128;
}
__host__ __device__ inline int off_mnmajor(int r, int k, int R) {     // core matrix = 8 k x 16 B (4 rows), MN-adjacent cores contiguous
  return (r % 4) * 4 + (k % 8) * 16 + (r / 4) * 128 + (k / 8) * (R / 4) * 128;
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B,
                                            float* __restrict__ D, Cfg c, int* err) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int abytes = 128 * c.K * 4, bbytes = c.N * c.K * 4;
  unsigned char* sa = smem;
  unsigned char* sal = sa + abytes;
  unsigned char* sb = sal + abytes;
  unsigned char* sbl = sb + bbytes;
  for (int i = tid; i < (2 * abytes + 2 * bbytes) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  for (int i = tid; i < c.M * c.K; i += 128) {
    const int r = i / c.K, k = i % c.K;
    const float x = A[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const int o = c.a_mn ? off_mnmajor(r, k, c.M) : off_kmajor(r, k, c.K);
    *reinterpret_cast<float*>(sa + o) = x;
    *reinterpret_cast<float*>(sal + o) = x - hi;
  }
  for (int i = tid; i < c.N * c.K; i += 128) {
    const int r = i / c.K, k = i % c.K;
    const float x = B[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    const int o = c.b_mn ? off_mnmajor(r, k, c.N) : off_kmajor(r, k, c.K);
    *reinterpret_cast<float*>(sb + o) = x;
    *reinterpret_cast<float*>(sbl + o) = x - hi;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");      // generic-proxy smem writes -> visible to the async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tmem_base;

  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                     ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    // K-major: LBO = stride between the two 16-byte K chunks of one MMA, SBO = stride between 8-row groups
    // MN-major: SBO = stride between 4-row (16 B) groups, LBO = stride between 8-k groups
    const uint32_t a_lbo = c.a_mn ? (uint32_t)(c.M / 4) * 128 : 128, a_sbo = c.a_mn ? 128 : (uint32_t)(c.K / 4) * 128;
    const uint32_t b_lbo = c.b_mn ? (uint32_t)(c.N / 4) * 128 : 128, b_sbo = c.b_mn ? 128 : (uint32_t)(c.K / 4) * 128;
    const uint32_t a_kstep = c.a_mn ? a_lbo : 256, b_kstep = c.b_mn ? b_lbo : 256;   // bytes per 8 k
    uint32_t accum = 0;
    for (int ks = 0; ks < c.K / 8; ++ks) {
      for (int pass = 0; pass < (c.split ? 3 : 1); ++pass) {
        const unsigned char* pa = (pass == 1) ? sal : sa;
        const unsigned char* pb = (pass == 2) ? sbl : sb;
        const uint32_t aa = smem_u32(pa) + ks * a_kstep, bb = smem_u32(pb) + ks * b_kstep;
        const uint64_t da = c.swap ? make_desc(aa, a_sbo, a_lbo) : make_desc(aa, a_lbo, a_sbo);
        const uint64_t db = c.swap ? make_desc(bb, b_sbo, b_lbo) : make_desc(bb, b_lbo, b_sbo);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tb),
            "l"(da), "l"(db), "r"(idesc), "r"(accum));
        accum = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
  }
  // bounded wait for the MMAs
  {
    uint32_t done = 0;
    for (int it = 0; it < 2000000 && !done; ++it) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0));
    }
    if (!done && tid == 0) *err = 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  // dump all 128 lanes x N columns: warp w reads lanes 32w .. 32w+31
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) D[tid * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256));
}

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}
static float rna_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  const Cfg cfgs[] = {
      {128, 64, 32, 0, 0, 0, 0},  {128, 64, 32, 0, 0, 1, 0},  {128, 64, 32, 0, 0, 0, 1},
      {128, 64, 32, 1, 1, 0, 0},  {128, 64, 32, 1, 1, 1, 0},  {128, 64, 32, 1, 1, 0, 1},
      {128, 64, 32, 1, 0, 0, 0},  {128, 64, 32, 0, 1, 0, 0},  {64, 64, 32, 0, 0, 0, 0},
      {64, 64, 32, 1, 1, 0, 0},   {128, 256, 64, 0, 0, 0, 1}, {128, 8, 8, 0, 0, 0, 1},
      {128, 24, 8, 0, 0, 0, 1},   {128, 128, 64, 1, 1, 0, 1}};
  int* derr;
  cudaMalloc(&derr, 4);
  for (const Cfg& c : cfgs) {
    std::vector<float> A(128 * c.K, 0.f), B(c.N * c.K), D(128 * c.N, -777.f);
    srand(1234);
    for (int i = 0; i < c.M * c.K; ++i) A[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(derr, 0, 4);
    const size_t smem = 2 * (128 * c.K * 4) + 2 * (c.N * c.K * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, c, derr);
    cudaError_t e = cudaDeviceSynchronize();
    int herr = 0;
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    printf("cfg M=%d N=%d K=%d a_mn=%d b_mn=%d swap=%d split=%d : cuda=%s timeout=%d\n", c.M, c.N, c.K, c.a_mn, c.b_mn,
           c.swap, c.split, cudaGetErrorString(e), herr);
    if (e != cudaSuccess) return 1;
    // references: exact (double), truncated-tf32 operands, rna-tf32 operands
    std::vector<double> Rex(c.M * c.N), Rtr(c.M * c.N), Rrn(c.M * c.N);
    for (int i = 0; i < c.M; ++i)
      for (int j = 0; j < c.N; ++j) {
        double s0 = 0, s1 = 0, s2 = 0;
        for (int k = 0; k < c.K; ++k) {
          s0 += (double)A[i * c.K + k] * B[j * c.K + k];
          s1 += (double)trunc_tf32(A[i * c.K + k]) * trunc_tf32(B[j * c.K + k]);
          s2 += (double)rna_tf32(A[i * c.K + k]) * rna_tf32(B[j * c.K + k]);
        }
        Rex[i * c.N + j] = s0; Rtr[i * c.N + j] = s1; Rrn[i * c.N + j] = s2;
      }
    // lane mapping: for each matrix row find the lane whose dump matches best (identity expected for M = 128)
    double worst_ex = 0, worst_tr = 0, worst_rn = 0;
    int nonid = 0;
    std::vector<int> lane_of(c.M, -1);
    for (int i = 0; i < c.M; ++i) {
      double best = 1e30; int bl = -1;
      for (int l = 0; l < 128; ++l) {
        double d = 0;
        for (int j = 0; j < c.N; ++j) d = fmax(d, fabs(D[l * c.N + j] - Rex[i * c.N + j]));
        if (d < best) { best = d; bl = l; }
      }
      lane_of[i] = bl;
      if (bl != i) ++nonid;
      for (int j = 0; j < c.N; ++j) {
        worst_ex = fmax(worst_ex, fabs(D[bl * c.N + j] - Rex[i * c.N + j]));
        worst_tr = fmax(worst_tr, fabs(D[bl * c.N + j] - Rtr[i * c.N + j]));
        worst_rn = fmax(worst_rn, fabs(D[bl * c.N + j] - Rrn[i * c.N + j]));
      }
    }
    printf("   max|D-exact|=%.3e  max|D-trunc_ref|=%.3e  max|D-rna_ref|=%.3e  rows not on lane==row: %d\n", worst_ex,
           worst_tr, worst_rn, nonid);
    if (nonid) {
      printf("   row->lane:");
      for (int i = 0; i < c.M; i += 1) if (i % 8 == 0) printf(" %d:%d", i, lane_of[i]);
      printf("\n");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
