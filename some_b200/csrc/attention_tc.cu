// K-attn (tcgen05): per-clip (var-len) multi-head self-attention softmax(Q K^T / 8) V, no mask, 8 heads x 64
// (base_attention.py:34-45; conform_blocke never forwards a mask: Gconform.py:83-84,133).
//
// Input: the fused to_q|to_kv GEMM output qkv bf16 [M, 1536] = [q | k | v] (heads 64-wide, contiguous);
// output bf16 [M, 512] = 'b h t c -> b t (h c)'.  No head-major copies are made: Q/K/V tiles are TMA boxes of
// 128 rows x 64 columns cut straight out of qkv.
//
// CTA = 128 query rows of one (clip, head); 2 CTAs per SM (112 KB smem, 256 TMEM columns each) so that one CTA's
// softmax overlaps the other's MMAs.  Roles (256 threads):
//   warp 0   TMA producer: Q once, then (K_j, V_j) 128-key tiles into a 2-stage ring (128-B swizzle)
//   warp 1   MMA issuer (one thread): S = Q K_j^T   (tcgen05.mma M128 N128 K16 x4, both operands K-major)
//                                      O += P_j V_j  (M128 N64 K16 x8, A = P from smem, B = V as MN-major operand)
//   warp 2   TMEM allocator: S fp32 [128 x 128] at column 0, O fp32 [128 x 64] at column 128
//   warps 4-7 softmax, thread = query row: tcgen05.ld S, online softmax in base 2 (ex2.approx), P -> bf16 -> smem in
//            the UMMA K-major swizzled layout; O stays in TMEM and is rescaled (tcgen05.ld / st) only when some row
//            max of the warp grew by more than 2^8 ("lazy rescale": the stale maximum is kept otherwise, P <= 256,
//            exact after the final division by the row sum).
// Rows of K/V beyond the clip end are masked (p = 0); rows beyond M are zero-filled by TMA.
#include "host_common.h"
#include "sm100_ptx.cuh"

#include "../../include/some_b200.h"

namespace some {

constexpr int TC_BM = 128;  // queries per CTA
constexpr int TC_BN = 128;  // keys per tile
constexpr int TC_TILE = 128 * 64 * 2;  // one 128 x 64 bf16 tile
constexpr int TC_SMEM = TC_TILE /*Q*/ + 2 * 2 * TC_TILE /*K,V x 2 stages*/ + 2 * TC_TILE /*P*/ + 128 /*barriers*/;
constexpr uint32_t TC_TMEM_COLS = 256;
constexpr uint32_t TC_O_COL = 128;

struct AttnTcParams {
  __nv_bfloat16* out[2];
  const int32_t* cu_frames;
  int tiles_per_clip;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                    const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + TC_TILE;                 // stage s: K at sKV + s * 2 * TILE, V right after it
  uint8_t* sP = smem + 5 * TC_TILE;              // 2 K-major atoms (keys 0..63, 64..127), 16 KB each
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TC_TILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* pv_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = blockIdx.x / p.tiles_per_clip;
  const int qt = blockIdx.x - clip * p.tiles_per_clip;
  const int row_begin = p.cu_frames[clip];
  const int T = p.cu_frames[clip + 1] - row_begin;
  const int q0 = qt * TC_BM;
  if (q0 >= T) return;  // whole CTA, before any barrier / TMEM use
  const int head = blockIdx.y;
  const int grp = blockIdx.z;
  const CUtensorMap* tm = grp == 0 ? &tm0 : &tm1;
  const int n_tiles = (T + TC_BN - 1) / TC_BN;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("some_b200: attention smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(tm);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<TC_TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, TC_TILE);
      tma_load_2d(sQ, tm, q_full, head * 64, row_begin + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TC_TILE);
        uint8_t* dst = sKV + s * 2 * TC_TILE;
        tma_load_2d(dst, tm, &kv_full[s], SOME_DIM + head * 64, row_begin + j * TC_BN);
        tma_load_2d(dst + TC_TILE, tm, &kv_full[s], 2 * SOME_DIM + head * 64, row_begin + j * TC_BN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16_f32(TC_BM, TC_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(TC_BM, 64, 0, 1);  // B = V is MN-major
      const uint64_t qdesc = umma_desc_kmajor_sw128(smem_u32(sQ));
      const uint64_t pdesc = umma_desc_kmajor_sw128(smem_u32(sP));
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t kaddr = smem_u32(sKV + s * 2 * TC_TILE);
        const uint64_t kdesc = umma_desc_kmajor_sw128(kaddr);
        // S = Q K^T: S_{j-1} has been fully read (p_full_{j-1} was waited before PV_{j-1} was issued)
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);  // P_j in smem, O rescaled
        tc_fence_after_sync();
        const uint64_t vdesc = umma_desc_mnmajor_sw128(kaddr + TC_TILE, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: 16 keys = 32 B inside the 64-key swizzle atom (+2), next atom +16 KB (+1024);  B: 16 key rows = 2 KB (+128)
          umma_bf16_ss(tmem_base + TC_O_COL, pdesc + (k >> 2) * 1024 + 2 * (k & 3), vdesc + 128 * k, idesc_pv,
                       (j | k) != 0);
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[s]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_o = t_s + TC_O_COL;
    const float c = 0.125f * 1.4426950408889634f;  // dim_head^-0.5 * log2(e)
    float m_used = -INFINITY, l = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    for (int j = 0; j < n_tiles; ++j) {
      const int valid = min(TC_BN, T - j * TC_BN);  // keys of this tile inside the clip
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      uint32_t v[32];
      // ---- pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        if (ch * 32 >= valid) break;  // warp-uniform
        tmem_ld_32x32(t_s + ch * 32, v);
        tmem_ld_wait();
        if ((ch + 1) * 32 <= valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ch * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      // ---- lazy rescale decision (warp-uniform)
      const float m_new = fmaxf(m_used, mx);
      const bool grow = (j == 0) || ((m_new - m_used) * c > 8.0f);
      const bool do_rescale = __any_sync(0xffffffffu, grow);
      float alpha = 1.0f;
      if (do_rescale) {
        alpha = (j == 0) ? 0.f : ex2_approx((m_used - m_new) * c);
        m_used = m_new;
        l *= alpha;
      }
      const float mc = m_used * c;
      // ---- pass 2: p = 2^(s c - m c), row sum, bf16 pack into registers
      uint32_t pk[64];
      float rs = 0.f;
      if (valid == TC_BN) {  // full tile (all but possibly the last): no per-key predicates
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          tmem_ld_32x32(t_s + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), c, -mc));
            const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mc));
            rs += p0 + p1;
            pk[ch * 16 + (i >> 1)] = pack_bf16x2(p0, p1);
          }
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (ch * 32 < valid) {
            tmem_ld_32x32(t_s + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), c, -mc));
              float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mc));
              if (ch * 32 + i >= valid) p0 = 0.f;
              if (ch * 32 + i + 1 >= valid) p1 = 0.f;
              rs += p0 + p1;
              pk[ch * 16 + (i >> 1)] = pack_bf16x2(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[ch * 16 + i] = 0u;
          }
        }
      }
      l += rs;
      // ---- O *= alpha (only when a maximum of this warp moved; PV_{j-1} is complete: s_full_j was committed after it)
      if (do_rescale && j > 0) {
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          tmem_ld_32x32(t_o + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(t_o + ch * 32, v);
        }
        tmem_st_wait();
      }
      // ---- P -> smem (UMMA K-major, 128-B swizzle: 16-byte chunk index XOR (row & 7))
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) {
        const int atom = cc >> 3, chunk = cc & 7;
        *reinterpret_cast<uint4*>(prow + atom * TC_TILE + ((chunk ^ sw) << 4)) =
            make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> out[row, head * 64 ..]
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after_sync();
    const float inv = 1.0f / l;
    const int qrow = q0 + r;
    __nv_bfloat16* dst = p.out[grp] + (size_t)(row_begin + qrow) * SOME_DIM + head * 64;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32(t_o + ch * 32, v);
      tmem_ld_wait();
      if (qrow < T) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<uint4*>(dst + ch * 32)[i] =
              make_uint4(pack_bf16x2(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv),
                         pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv),
                         pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv),
                         pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv));
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc<TC_TMEM_COLS>(tmem_base);
  }
}

}  // namespace some

using namespace some;

extern "C" int some_attention_varlen(const some_attn_args* a, cudaStream_t stream) {
  SOME_REQUIRE(a != nullptr && (a->groups == 1 || a->groups == 2), "some_attention_varlen: bad args");
  if (a->B <= 0 || a->max_frames <= 0 || a->M <= 0) return 0;
  SOME_REQUIRE(a->cu_frames != nullptr, "some_attention_varlen: null cu_frames");
  AttnTcParams p;
  CUtensorMap maps[2];
  for (int g = 0; g < 2; ++g) {
    const int s = g < a->groups ? g : 0;
    SOME_REQUIRE(a->qkv[s] && a->out[s], "some_attention_varlen: null pointer in group %d", s);
    if (make_tmap_bf16_2d(&maps[g], a->qkv[s], a->M, 3 * SOME_DIM, 3 * SOME_DIM, TC_BM)) return -1;
    p.out[g] = reinterpret_cast<__nv_bfloat16*>(a->out[s]);
  }
  p.cu_frames = a->cu_frames;
  p.tiles_per_clip = (a->max_frames + TC_BM - 1) / TC_BM;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    SOME_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long gx = 1ll * p.tiles_per_clip * a->B;
  SOME_REQUIRE(gx < (1ll << 31), "some_attention_varlen: grid too large");
  dim3 grid(static_cast<unsigned>(gx), SOME_HEADS, a->groups);
  attention_tc_kernel<<<grid, 256, TC_SMEM, stream>>>(maps[0], maps[1], p);
  return check_launch("some_attention_varlen");
}
