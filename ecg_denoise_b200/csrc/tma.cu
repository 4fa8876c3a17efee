// tma.cu -- host side of tma.cuh: CUtensorMap construction through the driver entry point (no link-time dependency
// on libcuda), cached per weight matrix.  Weight pointers are stable for the life of a model (one flat parameter
// buffer), so a training step encodes nothing after its first iteration.
#include <mutex>
#include <unordered_map>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "tma.cuh"

namespace {

struct Key {
  const void* p;
  int rows, cols, ld, box;
  bool operator==(const Key& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box;
  }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    size_t h = std::hash<const void*>()(k.p);
    h ^= std::hash<long long>()(((long long)k.rows << 40) ^ ((long long)k.cols << 20) ^ ((long long)k.ld << 8) ^ k.box) +
         0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    return h;
  }
};
std::mutex g_mu;
std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

}  // namespace

int rl_tmap_weight(const float* base, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  RL_REQUIRE(base && ((uintptr_t)base % 16 == 0) && ld % 4 == 0 && cols % 32 == 0 && box_rows > 0 && box_rows <= 256,
             RL_ERR_SHAPE, "tensor map: unsupported weight block (rows %d cols %d ld %d box %d)", rows, cols, ld,
             box_rows);
  std::lock_guard<std::mutex> lock(g_mu);
  const Key key{base, rows, cols, ld, box_rows};
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return RL_OK;
  }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    RL_REQUIRE(e == cudaSuccess && q == cudaDriverEntryPointSuccess && fn, RL_ERR_CUDA,
               "cuTensorMapEncodeTiled is not available from this driver");
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1u, 1u};
  CUtensorMap tm;
  const CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box,
                              estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RL_REQUIRE(r == CUDA_SUCCESS, RL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d weight block", (int)r,
             rows, cols);
  if (g_cache.size() > 4096) g_cache.clear();
  g_cache.emplace(key, tm);
  *out = tm;
  return RL_OK;
}
