#include "host_common.h"

#include <mutex>
#include <string.h>

#include "../../include/some_b200.h"

namespace some {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// Encoded tensor maps are pure functions of (address, shape, pitch, box, element size); encoding costs a driver call, and a
// forward pass re-uses the same ~40 (buffer, shape) combinations launch after launch, so they are memoised per thread in a
// small direct-mapped table (no locks; a collision just re-encodes).
namespace {
struct TmapKey {
  const void* base;
  uint64_t rows, cols, ld;
  uint32_t box_rows, box_cols, elem_bytes;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           box_cols == o.box_cols && elem_bytes == o.elem_bytes;
  }
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int kTmapSlots = 512;
thread_local TmapSlot g_tmaps[kTmapSlots];
inline uint32_t tmap_hash(const TmapKey& k) {
  uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
  h ^= (k.rows * 0xC2B2AE3D27D4EB4Full) ^ (k.cols << 17) ^ (k.ld << 29) ^ (uint64_t(k.box_rows) << 41) ^
       (uint64_t(k.box_cols) << 49) ^ (uint64_t(k.elem_bytes) << 57);
  h ^= h >> 29;
  return static_cast<uint32_t>(h % kTmapSlots);
}
}  // namespace

int make_tmap_2d(CUtensorMap* out, uint32_t elem_bytes, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols) {
  const TmapKey key{base, rows, cols, ld_elems, box_rows, box_cols, elem_bytes};
  TmapSlot& slot = g_tmaps[tmap_hash(key)];
  if (slot.valid && slot.key == key) {
    *out = slot.map;
    return 0;
  }
  PFN_encodeTiled enc = get_encode();
  SOME_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  SOME_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "TMA element size must be 2 (bf16) or 4 (f32)");
  SOME_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  SOME_REQUIRE((ld_elems * elem_bytes) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%llu x %u B)",
               (unsigned long long)ld_elems, elem_bytes);
  SOME_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols * elem_bytes == 128, "TMA box: rows<=256, inner extent 128 B");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SOME_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
               (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
  slot.key = key;
  slot.map = *out;
  slot.valid = true;
  return 0;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, 2, base, rows, cols, ld_elems, box_rows, box_cols);
}

int device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = device_index();
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v;
  }
  return n[dev];
}

static thread_local bool g_pdl = false;   // per launching thread (an engine's launches are serialised on one thread at a time)
bool pdl_enabled() { return g_pdl; }

}  // namespace some

extern "C" {
int some_version(void) { return SOME_B200_VERSION; }
int some_set_pdl(int on) {
  const int was = some::g_pdl ? 1 : 0;
  some::g_pdl = on != 0;
  return was;
}
const char* some_last_error(void) { return some::g_err; }
}
