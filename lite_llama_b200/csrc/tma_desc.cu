#include "tma_desc.cuh"
#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_map>
#include <string.h>

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode = nullptr;
std::mutex g_mu;

struct Key {
  const void* base; uint64_t rows, cols, stride; uint32_t brows, bcols; int elem; int swz;
  bool operator==(const Key& o) const { return memcmp(this, &o, sizeof(Key)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(Key) / 8; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
  }
};
std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;

int resolve_encode() {
  if (g_encode) return B2L_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B2L_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  B2L_REQUIRE(qres == cudaDriverEntryPointSuccess && fn != nullptr,
              "cuTensorMapEncodeTiled is not available from this driver");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return B2L_OK;
}

}  // namespace

int b2l_make_tmap_2d(CUtensorMap* out, const void* base, TmaElem elem, uint64_t rows, uint64_t cols,
                     uint64_t row_stride_bytes, uint32_t box_rows, uint32_t box_cols, bool swizzle128) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = resolve_encode();
  if (rc != B2L_OK) return rc;
  Key key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rows = rows; key.cols = cols; key.stride = row_stride_bytes;
  key.brows = box_rows; key.bcols = box_cols; key.elem = (int)elem; key.swz = swizzle128 ? 1 : 0;
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return B2L_OK;
  }
  CUtensorMapDataType dt;
  uint32_t esz;
  switch (elem) {
    case TMA_U8: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; esz = 1; break;
    case TMA_F16: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; esz = 2; break;
    case TMA_BF16: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; esz = 2; break;
    case TMA_I32: dt = CU_TENSOR_MAP_DATA_TYPE_INT32; esz = 4; break;
    default: B2L_UNSUPPORTED("tma: element type %d", (int)elem);
  }
  B2L_REQUIRE((uintptr_t)base % 16 == 0, "tma: base pointer %p is not 16-byte aligned", base);
  B2L_REQUIRE(row_stride_bytes % 16 == 0, "tma: row stride %llu bytes is not a multiple of 16",
              (unsigned long long)row_stride_bytes);
  B2L_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols >= 1 && box_cols <= 256, "tma: box %ux%u out of range",
              box_rows, box_cols);
  B2L_REQUIRE(!swizzle128 || box_cols * esz <= 128, "tma: 128B swizzle needs an inner box of <= 128 bytes");
  B2L_REQUIRE((box_cols * esz) % 16 == 0, "tma: inner box must be a multiple of 16 bytes");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap tm;
  CUresult r = g_encode(&tm, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2L_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu stride=%llu box=%ux%u)",
              (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_bytes, box_rows,
              box_cols);
  if (g_cache.size() > 8192) g_cache.clear();
  g_cache.emplace(key, tm);
  *out = tm;
  return B2L_OK;
}
