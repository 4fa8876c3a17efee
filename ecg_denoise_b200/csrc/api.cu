// api.cu -- error reporting, device check, launch accounting for libralenet_b200.so
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: backward runs on autograd threads

void rl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void rl_count_launch() { ++g_launches; }

int rl_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    rl_set_error("%s: %s", what, cudaGetErrorString(e));
    return RL_ERR_CUDA;
  }
  ++g_launches;
  return RL_OK;
}

extern "C" int ralenet_abi_version(void) { return RL_ABI_VERSION; }
extern "C" const char* ralenet_last_error(void) { return g_err; }

extern "C" int ralenet_check_device(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    rl_set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return RL_ERR_CUDA;
  }
  if (p.major != 10) {
    rl_set_error("device %d is sm_%d%d; libralenet_b200 is built for sm_100a only (no fallback path)", dev, p.major,
                 p.minor);
    return RL_ERR_ARCH;
  }
  return RL_OK;
}

extern "C" int64_t ralenet_launch_count(int32_t reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}
