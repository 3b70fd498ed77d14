// K-attn: per-clip (var-len) multi-head self-attention softmax(Q K^T / 8) V, no mask, 8 heads x 64
// (base_attention.py:34-45: both rearranges + F.scaled_dot_product_attention, attn_mask=None because
// conform_blocke never forwards a mask: Gconform.py:83-84,133).
//
// Input is the fused to_q|to_kv GEMM output qkv bf16 [M, 1536] = [q | k | v], heads contiguous 64-wide
// ('b t (h c)'), so no head-major copies are made; output bf16 [M, 512] is already 'b t (h c)'.
//
// Flash-attention forward: CTA = 128 query rows of one (clip, head), 8 warps x 16 rows; K/V tiles of 64
// keys double-buffered in shared memory with cp.async (zero-fill beyond the clip), XOR-swizzled 128-B rows
// for conflict-free ldmatrix; S and O accumulate in fp32 registers, online softmax in base 2.
// Tensor path: mma.sync.m16n8k16 bf16 (legacy HMMA path).
// Superseded on the product path by attention_tc.cu (tcgen05); exported as some_attention_varlen_mma and used by the
// tests as an independent cross-check of the tensor-core kernel.
#include "host_common.h"
#include "sm100_ptx.cuh"

#include "../../include/some_b200.h"

namespace some {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 64;
constexpr int ATT_LD = 3 * SOME_DIM;  // qkv row pitch (elements)
constexpr int ATT_SMEM = (ATT_BM * 64 + 4 * ATT_BN * 64) * 2;

struct AttnParams {
  const __nv_bfloat16* qkv[2];
  __nv_bfloat16* out[2];
  const int32_t* cu_frames;
  int tiles_per_clip;
};

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) attention_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + ATT_BM * 128;         // 2 buffers x 64 rows x 128 B
  const uint32_t sV = sK + 2 * ATT_BN * 128;

  const int clip = blockIdx.x / p.tiles_per_clip;
  const int qt = blockIdx.x - clip * p.tiles_per_clip;
  const int row_begin = p.cu_frames[clip];
  const int T = p.cu_frames[clip + 1] - row_begin;
  const int q0 = qt * ATT_BM;
  if (q0 >= T) return;
  const int head = blockIdx.y;
  const __nv_bfloat16* __restrict__ qkv = p.qkv[blockIdx.z] + (size_t)row_begin * ATT_LD + head * 64;
  __nv_bfloat16* __restrict__ out = p.out[blockIdx.z] + (size_t)row_begin * SOME_DIM + head * 64;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int num_kt = (T + ATT_BN - 1) / ATT_BN;

  // ---- async loads
  for (int i = tid; i < ATT_BM * 8; i += 256) {
    const int r = i >> 3, ch = i & 7;
    const bool ok = q0 + r < T;
    cp_async16(sQ + swz(r, ch), qkv + (size_t)(ok ? q0 + r : 0) * ATT_LD + ch * 8, ok);
  }
  auto load_kv = [&](int kt, int buf) {
    for (int i = tid; i < ATT_BN * 8; i += 256) {
      const int r = i >> 3, ch = i & 7;
      const int key = kt * ATT_BN + r;
      const bool ok = key < T;
      const __nv_bfloat16* src = qkv + (size_t)(ok ? key : 0) * ATT_LD + ch * 8;
      cp_async16(sK + buf * ATT_BN * 128 + swz(r, ch), src + SOME_DIM, ok);
      cp_async16(sV + buf * ATT_BN * 128 + swz(r, ch), src + 2 * SOME_DIM, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sc = 0.125f * 1.4426950408889634f;  // dim_head^-0.5 * log2(e)

  for (int kt = 0; kt < num_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < num_kt) load_kv(kt + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm_x4(sQ + swz(r, 2 * kk + (lane >> 4)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t kb = sK + buf * ATT_BN * 128;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int mat = lane >> 3;
        const int key = np * 16 + (lane & 7) + (mat >> 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + swz(key, 2 * kk + (mat & 1)), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qf[kk], b0, b1);
        mma_bf16(s[2 * np + 1], qf[kk], b2, b3);
      }
    }
    // ---- mask the ragged last key tile
    if ((kt + 1) * ATT_BN > T) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int key = kt * ATT_BN + nb * 8 + 2 * (lane & 3);
        if (key >= T) s[nb][0] = s[nb][2] = -INFINITY;
        if (key + 1 >= T) s[nb][1] = s[nb][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows g = lane / 4 and g + 8)
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], ms[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = exp2f((m_run[r] - mx[r]) * sc);  // first tile: exp2(-inf) = 0
      m_run[r] = mx[r];
      ms[r] = mx[r] * sc;
    }
    uint32_t pf[4][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = exp2f(fmaf(s[nb][0], sc, -ms[0]));
      const float p1 = exp2f(fmaf(s[nb][1], sc, -ms[0]));
      const float p2 = exp2f(fmaf(s[nb][2], sc, -ms[1]));
      const float p3 = exp2f(fmaf(s[nb][3], sc, -ms[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o[nb][0] *= corr[0];
      o[nb][1] *= corr[0];
      o[nb][2] *= corr[1];
      o[nb][3] *= corr[1];
    }
    // ---- O += P V
    const uint32_t vb = sV + buf * ATT_BN * 128;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        const int mat = lane >> 3;
        const int key = kk * 16 + (lane & 7) + (mat & 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vb + swz(key, 2 * dp + (mat >> 1)), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pf[kk], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }
  cp_async_wait<0>();

  // ---- finalise: row sums live in the 4 lanes of a quad
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int c = nb * 8 + 2 * (lane & 3);
    if (r0 < T) *reinterpret_cast<uint32_t*>(out + (size_t)r0 * SOME_DIM + c) = pack_bf16x2(o[nb][0] * inv0, o[nb][1] * inv0);
    if (r1 < T) *reinterpret_cast<uint32_t*>(out + (size_t)r1 * SOME_DIM + c) = pack_bf16x2(o[nb][2] * inv1, o[nb][3] * inv1);
  }
}

}  // namespace some

using namespace some;

extern "C" int some_attention_varlen_mma(const some_attn_args* a, cudaStream_t stream) {
  SOME_REQUIRE(a != nullptr && (a->groups == 1 || a->groups == 2), "some_attention_varlen: bad args");
  if (a->B <= 0 || a->max_frames <= 0) return 0;
  SOME_REQUIRE(a->cu_frames != nullptr, "some_attention_varlen: null cu_frames");
  AttnParams p;
  for (int g = 0; g < 2; ++g) {
    const int s = g < a->groups ? g : 0;
    SOME_REQUIRE(a->qkv[s] && a->out[s], "some_attention_varlen: null pointer in group %d", s);
    p.qkv[g] = reinterpret_cast<const __nv_bfloat16*>(a->qkv[s]);
    p.out[g] = reinterpret_cast<__nv_bfloat16*>(a->out[s]);
  }
  p.cu_frames = a->cu_frames;
  p.tiles_per_clip = (a->max_frames + ATT_BM - 1) / ATT_BM;
  static bool configured[kMaxDevices] = {};   // function attributes are per device
  const int dev_ = device_index();
  if (!configured[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    SOME_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
    configured[dev_] = true;
  }
  const long long gx = 1ll * p.tiles_per_clip * a->B;
  SOME_REQUIRE(gx < (1ll << 31), "some_attention_varlen: grid too large");
  dim3 grid(static_cast<unsigned>(gx), SOME_HEADS, a->groups);
  attention_kernel<<<grid, 256, ATT_SMEM, stream>>>(p);
  return check_launch("some_attention_varlen");
}
