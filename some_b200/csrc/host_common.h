// Host-side helpers shared by the C-ABI entry points: thread-local error string, launch checks and
// TMA tensor-map encoding (cuTensorMapEncodeTiled obtained through cudaGetDriverEntryPoint so the
// library needs no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace some {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// rows x cols bf16 row-major matrix (cols contiguous, `ld` elements between rows); the box is
// box_rows x 64 elements (128 B inner extent) with the 128-byte swizzle the UMMA descriptors expect.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols = 64);

// same for any element size (2 = bf16, 4 = f32): the inner box extent is always 128 bytes (64 bf16 / 32 f32 columns)
int make_tmap_2d(CUtensorMap* out, uint32_t elem_bytes, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols);

// Function attributes (dynamic shared-memory opt-in, carveout) and the SM count are per DEVICE: everything that caches them
// indexes by the current device so that engines on several GPUs of one process work (one slot per device ordinal).
constexpr int kMaxDevices = 64;
int device_index();
int num_sms();

#define SOME_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      some::set_error(__VA_ARGS__);  \
      return -1;                     \
    }                                \
  } while (0)

}  // namespace some
