// EXPERIMENTAL (not part of libsome_b200.so; built only by tools/ab_bench.py --src): attention v7.
//
// v6 (some_b200/csrc/attention_tc.cu) leaves the MUFU pipe ~72 % busy because each softmax group owns ONE S buffer: after
// softmax(j) the group idles for the whole hand-off chain  p_full -> PV_j -> QK_{j+2} -> s_full  (~1100 of ~2850 clk per
// tile, profiles/r01_attention_tc_v6_notes.txt).  v7 keeps v6's TMEM budget (256 columns, 2 CTAs / SM, 4 softmax warps per
// SM sub-partition) but cuts the key tile to 32 keys, so each group's 64 S columns hold TWO tiles (double buffering): the
// round trip of tile j overlaps the softmax of tile j+2 of the same group.  Each group gets its own MMA-issuing warp
// (warps 1 and 10) so the two hand-off chains do not serialise on one thread.  P (32 keys = 16 columns) still overwrites the
// head of its own S buffer; QK_{j+4} refills that buffer right behind PV_j.  With double buffering s_full no longer implies
// that the group's previous PV retired, so the (rare) lazy O rescale waits on an alternating pv_done barrier, as v5 did.
// STATUS: compiles (ptxas: see tools/experimental/README.md); NOT yet run on hardware.
#include "../../some_b200/csrc/host_common.h"
#include "../../some_b200/csrc/sm100_ptx.cuh"

#include "../../include/some_b200.h"

namespace some {

constexpr int TC_BM = 128;                 // queries per CTA
constexpr int TC_BN = 32;                  // keys per tile
constexpr int TC_QTILE = 128 * 64 * 2;     // 16 KB
constexpr int TC_KTILE = TC_BN * 64 * 2;   // 4 KB (K or V tile)
constexpr int TC_STAGES = 10;              // 8 KB each: five 32-key tiles per group in flight
constexpr int TC_THREADS = 352;            // TMA | MMA group 0 | 4 + 4 softmax warps | MMA group 1
constexpr int TC_BAR_BYTES = 512;
constexpr int TC_SMEM = TC_QTILE + TC_STAGES * 2 * TC_KTILE + TC_BAR_BYTES + 2 * TC_BM * 8 /*group stats*/;
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
// 2^x for a packed pair on the FMA / ALU pipes instead of the MUFU pipe (the kernel's bottleneck): Cody-Waite split with
// the 1.5 * 2^23 rounding constant, degree-3 polynomial on [-0.5, 0.5] (max relative error 7.7e-5, well inside the bf16
// rounding of P), exponent spliced in with an integer add.  TC_POLY_OF_8 of every 8 scores take this path.
__device__ __forceinline__ void exp2_poly2(uint64_t y2, float& p0, float& p1) {
  float a, b;
  f2_unpack(y2, a, b);
  a = fmaxf(a, -126.f);
  b = fmaxf(b, -126.f);
  const uint64_t y = f2_pack(a, b);
  const uint64_t xf = f2_add(y, f2_pack(12582912.f, 12582912.f));
  const uint64_t n = f2_add(xf, f2_pack(-12582912.f, -12582912.f));
  const uint64_t r = f2_fma(n, f2_pack(-1.f, -1.f), y);
  uint64_t q = f2_fma(f2_pack(0.05508868396282196f, 0.05508868396282196f), r, f2_pack(0.24260404706001282f, 0.24260404706001282f));
  q = f2_fma(q, r, f2_pack(0.6932762265205383f, 0.6932762265205383f));
  q = f2_fma(q, r, f2_pack(0.9999289512634277f, 0.9999289512634277f));
  float qa, qb, xa, xb;
  f2_unpack(q, qa, qb);
  f2_unpack(xf, xa, xb);
  p0 = __int_as_float(__float_as_int(qa) + (__float_as_int(xa) << 23));
  p1 = __int_as_float(__float_as_int(qb) + (__float_as_int(xb) << 23));
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {  // 32 lanes x 16 columns
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// TC_POLY_MASK: bit q set -> the second pair of the q-th group of four scores (q = 0..3 inside every 16) takes the polynomial
// path: 0b1010 = 2 of 8 scores, 0b1110 = 3 of 8, 0b1111 = 4 of 8, 0 = none.
#ifndef TC_POLY_MASK
#define TC_POLY_MASK 0xA
#endif

#ifdef SOME_ATTN_TRACE
// debug build only (tools/attn_trace.py): SM-clock timestamps of one CTA's softmax groups and MMA thread
__device__ long long* g_attn_trace = nullptr;
#define ATTN_TRACE(role, tile, ev)                                                                  \
  do {                                                                                              \
    if (trace_on && (tile) < 64) g_attn_trace[(((role) * 64) + (tile)) * 4 + (ev)] = clock64();   \
  } while (0)
#else
#define ATTN_TRACE(role, tile, ev) do { } while (0)
#endif

__global__ void __launch_bounds__(TC_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmq0, const __grid_constant__ CUtensorMap tmkv0,
                    const __grid_constant__ CUtensorMap tmq1, const __grid_constant__ CUtensorMap tmkv1,
                    const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + TC_QTILE;                                  // stage s: K at +s * 2 * KTILE, V right after it
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + TC_STAGES * 2 * TC_KTILE);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                      // [TC_STAGES]
  uint64_t* kv_empty = kv_full + TC_STAGES;          // [TC_STAGES]
  uint64_t* s_full = kv_empty + TC_STAGES;           // [group][buffer]
  uint64_t* p_full = s_full + 4;                     // [group][buffer]
  uint64_t* pv_done = p_full + 4;                    // [group][buffer], alternating so a parity wait is never a phase behind
  uint64_t* grp_done = pv_done + 4;                  // [group]: every PV of the group retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(grp_done + 2);
  static_assert(8 * (1 + 2 * TC_STAGES + 14 + 1) <= TC_BAR_BYTES, "barrier block too small");
  float2* stats = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + TC_BAR_BYTES);  // [2][128] (max, row sum) per group

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int clip = blockIdx.x / p.tiles_per_clip;
  const int qt = blockIdx.x - clip * p.tiles_per_clip;
  const int row_begin = p.cu_frames[clip];
  const int T = p.cu_frames[clip + 1] - row_begin;
  const int q0 = qt * TC_BM;
  if (q0 >= T) return;  // whole CTA, before any barrier / TMEM use
  const int head = blockIdx.y;
  const int grp = blockIdx.z;
  const CUtensorMap* tmq = grp == 0 ? &tmq0 : &tmq1;
  const CUtensorMap* tmkv = grp == 0 ? &tmkv0 : &tmkv1;
  const int n_tiles = (T + TC_BN - 1) / TC_BN;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("some_b200: attention smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(tmq);
    tma_prefetch_desc(tmkv);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(&grp_done[0], 1);
    mbar_init(&grp_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<TC_TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, TC_QTILE);
      tma_load_2d(sQ, tmq, q_full, head * 64, row_begin + q0);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TC_KTILE);
        uint8_t* dst = sKV + s * 2 * TC_KTILE;
        tma_load_2d(dst, tmkv, &kv_full[s], SOME_DIM + head * 64, row_begin + j * TC_BN);
        tma_load_2d(dst + TC_KTILE, tmkv, &kv_full[s], 2 * SOME_DIM + head * 64, row_begin + j * TC_BN);
        if (++s == TC_STAGES) s = 0, ph ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 10) {
    // ---- one MMA issuer per softmax group: group g owns the 32-key tiles j = g, g + 2, g + 4, ... (local index it = j >> 1),
    //      S buffer it & 1 of its 64 S columns and the accumulator O[g]
    const int g = warp == 1 ? 0 : 1;
    if (elect_one_sync()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16_f32(TC_BM, TC_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(TC_BM, 64, 0, 1);  // B = V is MN-major
      const uint64_t qdesc = umma_desc_kmajor_sw128(smem_u32(sQ));
      const uint32_t s_tmem = tmem_base + g * 64;
      const uint32_t o_tmem = tmem_base + TC_O_COL + g * 64;
      auto issue_qk = [&](int it) {  // S[g][it & 1] = Q K_j^T, j = g + 2 it
        const int j = g + 2 * it;
        const int s = j % TC_STAGES;
        mbar_wait(&kv_full[s], (j / TC_STAGES) & 1);
        tc_fence_after_sync();
        const uint64_t kdesc = umma_desc_kmajor_sw128(smem_u32(sKV + s * 2 * TC_KTILE));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(s_tmem + (it & 1) * TC_BN, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[2 * g + (it & 1)]);
      };
      const int my_tiles = n_tiles > g ? (n_tiles - g + 1) / 2 : 0;
      mbar_wait(q_full, 0);
      if (my_tiles > 0) issue_qk(0);
      if (my_tiles > 1) issue_qk(1);
      for (int it = 0; it < my_tiles; ++it) {
        const int j = g + 2 * it;
        const int s = j % TC_STAGES;
        mbar_wait(&p_full[2 * g + (it & 1)], (it >> 1) & 1);  // P_j in TMEM (and O rescaled if it had to be)
        tc_fence_after_sync();
        const uint64_t vdesc = umma_desc_mnmajor_sw128(smem_u32(sKV + s * 2 * TC_KTILE + TC_KTILE), 1024);
        const uint32_t p_tmem = s_tmem + (it & 1) * TC_BN;  // P_j (bf16, two keys per column) over the head of its S buffer
#pragma unroll
        for (int k = 0; k < 2; ++k)  // 16 keys per MMA: A +8 TMEM columns, B +16 key rows = 2 KB (+128)
          umma_bf16_ts(o_tmem, p_tmem + 8 * k, vdesc + 128 * k, idesc_pv, it > 0 || k != 0);
        umma_commit(&pv_done[2 * g + (it & 1)]);
        umma_commit(&kv_empty[s]);
        // the buffer just consumed is refilled two local tiles ahead, right behind PV_j on the in-order tensor pipe
        if (it + 2 < my_tiles) issue_qk(it + 2);
      }
      umma_commit(&grp_done[g]);
    }
    __syncwarp();
  } else {
    const int g = (warp - 2) >> 2;  // softmax group: 0 = even key tiles, 1 = odd key tiles
    const int quad = warp & 3;      // the TMEM lane quadrant this warp may touch
    const int r = quad * 32 + lane; // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_sg = t_lane + g * 64;
    const uint32_t t_o = t_lane + TC_O_COL + g * 64;
    const float c = 0.125f * 1.4426950408889634f;  // dim_head^-0.5 * log2(e)
    const uint64_t c2 = f2_pack(c, c);
    float m_used = -INFINITY, l = 0.f;
    int it = 0;
    for (int j = g; j < n_tiles; j += 2, ++it) {
      const int valid = min(TC_BN, T - j * TC_BN);  // keys of this tile inside the clip
      const uint32_t t_s = t_sg + (it & 1) * TC_BN;
      mbar_wait(&s_full[2 * g + (it & 1)], (it >> 1) & 1);
      tc_fence_after_sync();
      uint32_t v[32];
      tmem_ld_32x32(t_s, v);  // the whole 32-key tile stays in registers for both passes
      tmem_ld_wait();
      // ---- pass 1: row maximum
      float mx = -INFINITY;
      if (valid == TC_BN) {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
          m2 = fmaxf(m2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      // ---- lazy rescale decision (warp-uniform)
      const float m_new = fmaxf(m_used, mx);
      const bool grow = (it == 0) || ((m_new - m_used) * c > 8.0f);
      const bool do_rescale = __any_sync(0xffffffffu, grow);
      if (do_rescale) {
        const float alpha = (it == 0) ? 0.f : ex2_approx((m_used - m_new) * c);
        m_used = m_new;
        l *= alpha;
        if (it > 0) {  // O[g] *= alpha (rare).  PV of the previous local tile may still be in flight: wait for it.
          mbar_wait(&pv_done[2 * g + ((it - 1) & 1)], ((it - 1) >> 1) & 1);
          tc_fence_after_sync();
          uint32_t o[32];
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            tmem_ld_32x32(t_o + 32 * h, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_o + 32 * h, o);
          }
          tmem_st_wait();
        }
      }
      const float mc = m_used * c;
      const uint64_t nmc2 = f2_pack(-mc, -mc);
      // ---- pass 2: p = 2^(s c - m c), row sum, bf16 pack (16 packed columns)
      uint32_t pk[16];
      uint64_t rs_a = f2_pack(0.f, 0.f), rs_b = rs_a;
      if (valid == TC_BN) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const uint64_t ya = f2_fma(f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), c2, nmc2);
          const uint64_t yb = f2_fma(f2_pack(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), c2, nmc2);
          float p0, p1, p2, p3;
          {
            float y0, y1;
            f2_unpack(ya, y0, y1);
            p0 = ex2_approx(y0);
            p1 = ex2_approx(y1);
          }
          if ((TC_POLY_MASK >> ((i >> 2) & 3)) & 1) {
            exp2_poly2(yb, p2, p3);
          } else {
            float y2, y3;
            f2_unpack(yb, y2, y3);
            p2 = ex2_approx(y2);
            p3 = ex2_approx(y3);
          }
          rs_a = f2_add(rs_a, f2_pack(p0, p1));
          rs_b = f2_add(rs_b, f2_pack(p2, p3));
          pk[i >> 1] = pack_bf16x2(p0, p1);
          pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), c, -mc));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mc));
          if (i >= valid) p0 = 0.f;
          if (i + 1 >= valid) p1 = 0.f;
          rs_a = f2_add(rs_a, f2_pack(p0, p1));
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
      }
      {
        float s0, s1, s2, s3;
        f2_unpack(rs_a, s0, s1);
        f2_unpack(rs_b, s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      // ---- P -> TMEM over the first 16 columns of this tile's S buffer (all 32 scores are in registers)
      tmem_st_32x16(t_s, pk);
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[2 * g + (it & 1)]);
    }
    // ---- merge the two groups and write O / l -> bf16 -> out[row, head * 64 ..]; group g writes channels [32 g, 32 g + 32)
    stats[g * TC_BM + r] = make_float2(m_used, l);
    mbar_wait(&grp_done[0], 0);
    mbar_wait(&grp_done[1], 0);
    tc_fence_after_sync();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float2 sa = stats[r], sb = stats[TC_BM + r];
    const bool two = n_tiles > 1;
    const float m = fmaxf(sa.x, sb.x);
    float wa = ex2_approx((sa.x - m) * c);
    float wb = two ? ex2_approx((sb.x - m) * c) : 0.f;
    const float inv = 1.0f / (sa.y * wa + sb.y * wb);
    wa *= inv;
    wb *= inv;
    const int qrow = q0 + r;
    __nv_bfloat16* dst = p.out[grp] + (size_t)(row_begin + qrow) * SOME_DIM + head * 64 + 32 * g;
    uint32_t oa[32], ob[32];
    tmem_ld_32x32(t_lane + TC_O_COL + 32 * g, oa);
    if (two) tmem_ld_32x32(t_lane + TC_O_COL + 64 + 32 * g, ob);
    tmem_ld_wait();
    if (qrow < T) {
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(oa[i]) * wa;
      if (two) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(__uint_as_float(ob[i]), wb, o[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16x2(o[8 * i], o[8 * i + 1]), pack_bf16x2(o[8 * i + 2], o[8 * i + 3]),
                                                      pack_bf16x2(o[8 * i + 4], o[8 * i + 5]), pack_bf16x2(o[8 * i + 6], o[8 * i + 7]));
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

#ifdef SOME_ATTN_TRACE
extern "C" int some_attention_set_trace(long long* buf) {
  return cudaMemcpyToSymbol(some::g_attn_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int some_attention_varlen(const some_attn_args* a, cudaStream_t stream) {
  SOME_REQUIRE(a != nullptr && (a->groups == 1 || a->groups == 2), "some_attention_varlen: bad args");
  if (a->B <= 0 || a->max_frames <= 0 || a->M <= 0) return 0;
  SOME_REQUIRE(a->cu_frames != nullptr, "some_attention_varlen: null cu_frames");
  AttnTcParams p;
  CUtensorMap maps[4];
  for (int g = 0; g < 2; ++g) {
    const int s = g < a->groups ? g : 0;
    SOME_REQUIRE(a->qkv[s] && a->out[s], "some_attention_varlen: null pointer in group %d", s);
    if (make_tmap_bf16_2d(&maps[2 * g], a->qkv[s], a->M, 3 * SOME_DIM, 3 * SOME_DIM, TC_BM)) return -1;
    if (make_tmap_bf16_2d(&maps[2 * g + 1], a->qkv[s], a->M, 3 * SOME_DIM, 3 * SOME_DIM, TC_BN)) return -1;
    p.out[g] = reinterpret_cast<__nv_bfloat16*>(a->out[s]);
  }
  p.cu_frames = a->cu_frames;
  p.tiles_per_clip = (a->max_frames + TC_BM - 1) / TC_BM;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    SOME_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e));
    // two CTAs per SM need the full shared-memory carveout (2 x 113 KB)
    e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    SOME_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(attention_tc carveout): %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long gx = 1ll * p.tiles_per_clip * a->B;
  SOME_REQUIRE(gx < (1ll << 31), "some_attention_varlen: grid too large");
  dim3 grid(static_cast<unsigned>(gx), SOME_HEADS, a->groups);
  attention_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  return check_launch("some_attention_varlen");
}
