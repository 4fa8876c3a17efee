// micro-benchmark: legacy mma.sync throughput on sm_100a (TF32 m16n8k8, BF16 m16n8k16) vs FFMA
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void k_tf32(float* out, int iters) {
  float c[8][4] = {};
  uint32_t a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16(float* out, int iters) {
  float c[8][4] = {};
  uint32_t a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters) {
  float c[16]; for (int j = 0; j < 16; ++j) c[j] = j;
  float a = threadIdx.x * 1e-3f, b = 0.999f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = fmaf(c[j], b, a);
  }
  float s = 0; for (int j = 0; j < 16; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148, iters = 20000; float ms;
  for (int warps : {4, 8, 16, 32}) {
    k_tf32<<<sms, warps * 32>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_tf32<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double mac = (double)sms * warps * iters * 8 * 16 * 8 * 8;
    printf("tf32 m16n8k8  warps/SM=%2d: %.1f TFLOP/s (%.0f MAC/clk/SM @1.9GHz)\n", warps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / 1.9e9);
    cudaEventRecord(e0); k_bf16<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    mac = (double)sms * warps * iters * 8 * 16 * 8 * 16;
    printf("bf16 m16n8k16 warps/SM=%2d: %.1f TFLOP/s (%.0f MAC/clk/SM)\n", warps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / 1.9e9);
    cudaEventRecord(e0); k_ffma<<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    mac = (double)sms * warps * 32 * iters * 16;
    printf("ffma          warps/SM=%2d: %.1f TFLOP/s (%.0f FMA/clk/SM)\n", warps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / 1.9e9);
  }
  return 0;
}
