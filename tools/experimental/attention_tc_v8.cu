// EXPERIMENTAL (built only by tools/ab_bench.py --src): attention v8.
//
// v6b (some_b200/csrc/attention_tc.cu) leaves every softmax group idle ~35 % of its period: a group owns ONE S buffer, so the
// chain  arrive -> PV_j -> QK_{j+2} -> s_full  (~950 clk) is exposed once per tile (profiles/r02_attention_notes.txt).  A second
// S buffer per group does not fit beside two O accumulators in 256 TMEM columns.  v8 makes both groups accumulate into ONE
// O (they share the per-row reference maximum R, so all P are on the same scale) and spends the freed 64 columns on a THIRD S
// buffer: tiles rotate through S0 | S1 | S2, tile j = buffer j % 3, group j & 1; QK_{j+3} refills a buffer right behind PV_j,
// so a group's next tile was issued half a period before it is needed.
//
// The shared reference maximum only ever changes at a CTA-wide rendezvous (rare after the first tiles): a warp that sees a
// score exceed R by more than 2^8 raises sync_flag and waits at a named barrier; every other softmax warp joins from its next
// safe point (the s_full polling loop or the drain loop at the end); the MMA thread is asked to quiesce (tcgen05.commit on a
// dedicated barrier) and parks; the rows' new maxima are exchanged through shared memory, O and the row sums are rescaled,
// and everything resumes.  Between rendezvous R is fixed, p = 2^((s - R) c) <= 2^8.
#include "../../some_b200/csrc/host_common.h"
#include "../../some_b200/csrc/sm100_ptx.cuh"

#include "../../include/some_b200.h"

namespace some {

constexpr int TC_BM = 128;                 // queries per CTA
constexpr int TC_BN = 64;                  // keys per tile
constexpr int TC_QTILE = 128 * 64 * 2;     // 16 KB
constexpr int TC_KTILE = TC_BN * 64 * 2;   // 8 KB (K or V tile)
constexpr int TC_STAGES = 5;                // slots of the K ring AND of the V ring (8 KB each): K_t is needed three iterations before V_t
constexpr int TC_THREADS = 320;
constexpr int TC_BAR_BYTES = 320;
constexpr int TC_SMEM = TC_QTILE + TC_STAGES * 2 * TC_KTILE + TC_BAR_BYTES + 2 * TC_BM * 4 /*mxs*/ + 2 * TC_BM * 4 /*row sums*/ + 64 /*flags, arrived[8]*/;
constexpr uint32_t TC_TMEM_COLS = 256;
constexpr uint32_t TC_O_COL = 192;         // S0 | S1 | S2 | O

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
// 2^x for a packed pair on the FMA / ALU pipes (see attention_tc.cu)
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
#ifndef TC_POLY_OF_8
#define TC_POLY_OF_8 2
#endif
#ifndef TC_POLL_NS
#define TC_POLL_NS 1000  // suspend-time hint of the MMA thread's blocking poll
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
  uint8_t* sK = smem + TC_QTILE;                                   // K ring: slot t % TC_STAGES
  uint8_t* sV = sK + TC_STAGES * TC_KTILE;                         // V ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + TC_STAGES * TC_KTILE);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                       // [TC_STAGES]
  uint64_t* k_empty = k_full + TC_STAGES;            // [TC_STAGES]
  uint64_t* v_full = k_empty + TC_STAGES;            // [TC_STAGES]
  uint64_t* v_empty = v_full + TC_STAGES;            // [TC_STAGES]
  uint64_t* s_full = v_empty + TC_STAGES;            // [3]  S buffer b holds Q K_j^T (tcgen05.commit)
  uint64_t* p_full = s_full + 3;                     // [3]  P_j written over the head of buffer b (4 warp arrivals)
  uint64_t* all_done = p_full + 3;                   // every PV retired
  uint64_t* sync_req = all_done + 1;                 // rendezvous: softmax -> MMA thread "quiesce please"
  uint64_t* quiesce = sync_req + 1;                  // MMA thread -> softmax: every MMA issued so far has retired
  uint64_t* resume = quiesce + 1;                    // softmax -> MMA thread: go on
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(resume + 1);
  static_assert(8 * (1 + 4 * TC_STAGES + 10 + 1) <= TC_BAR_BYTES, "barrier block too small");
  float* mxs = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + TC_BAR_BYTES);   // [2][128] row maxima offered at a rendezvous
  float* lsum = mxs + 2 * TC_BM;                                                             // [2][128] final row sums
  volatile int* sync_flag = reinterpret_cast<volatile int*>(lsum + 2 * TC_BM);               // a rendezvous has been requested
  volatile int* done_cnt = sync_flag + 1;                                                    // softmax warps that finished their tiles
  volatile int* qk_tile = sync_flag + 10;                                                    // [3] tile whose QK was last issued into S buffer b
  volatile int* arrived = sync_flag + 2;                                                     // [8] last tile each softmax warp arrived p_full for

  // Role index: 0 = TMA producer, 1 = MMA issuer, 2..9 = softmax.  The sub-partition arbiter favours the HIGHEST warp id, and
  // the producer / issuer threads sit on the kernel's critical hand-off chain, so with TC_MMA_HIGH_WARP they are hardware warps
  // 8 and 9 (softmax = hardware warps 0..7) instead of 0 and 1.
#ifdef TC_MMA_HIGH_WARP
  const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = hw_warp >= 8 ? hw_warp - 8 : hw_warp + 2;
#else
  const int hw_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = hw_warp;
#endif
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
#ifdef SOME_ATTN_TRACE
  const bool trace_on = g_attn_trace != nullptr && blockIdx.x == 7 && blockIdx.y == 3 && blockIdx.z == 0 &&
                        (lane == 0 || warp <= 1) && (warp <= 2 || warp == 6);
#endif

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("some_b200: attention smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(tmq);
    tma_prefetch_desc(tmkv);
    *sync_flag = 0;
    *done_cnt = 0;
    qk_tile[0] = qk_tile[1] = qk_tile[2] = -1;
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
    }
    mbar_init(all_done, 1);
    mbar_init(sync_req, 1);
    mbar_init(quiesce, 1);
    mbar_init(resume, 1);
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
      // load order = consumption order of the MMA thread: K_0 K_1 K_2, then (V_j, K_{j+3}) per iteration
      auto load_k = [&](int t) {
        const int s = t % TC_STAGES;
        mbar_wait(&k_empty[s], ((t / TC_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], TC_KTILE);
        tma_load_2d(sK + s * TC_KTILE, tmkv, &k_full[s], SOME_DIM + head * 64, row_begin + t * TC_BN);
      };
      auto load_v = [&](int t) {
        const int s = t % TC_STAGES;
        ATTN_TRACE(3, t, 3);                       // producer starts waiting for the slot of V_t
        mbar_wait(&v_empty[s], ((t / TC_STAGES) & 1) ^ 1);
        ATTN_TRACE(3, t, 2);                       // V_t load issued
        mbar_arrive_expect_tx(&v_full[s], TC_KTILE);
        tma_load_2d(sV + s * TC_KTILE, tmkv, &v_full[s], 2 * SOME_DIM + head * 64, row_begin + t * TC_BN);
      };
      for (int t = 0; t < 3 && t < n_tiles; ++t) load_k(t);
      for (int j = 0; j < n_tiles; ++j) {
        load_v(j);
        if (j + 3 < n_tiles) load_k(j + 3);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16_f32(TC_BM, TC_BN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_f32(TC_BM, 64, 0, 1);  // B = V is MN-major
      const uint64_t qdesc = umma_desc_kmajor_sw128(smem_u32(sQ));
      // Everything between "P_j is complete" and "PV_j, QK_{j+3} issued" is on the kernel's critical hand-off chain, so the
      // waits for the K / V tiles those MMAs read are taken EARLY: right after a tile of group g has been served, the thread
      // blocks (off the chain; the ring is five tiles deep, the data is normally long there) until V_{j+2} and K_{j+5} -- the
      // operands of the group's NEXT service -- have landed.
      auto wait_k = [&](int t) {
        if (t < n_tiles) mbar_wait(&k_full[t % TC_STAGES], (t / TC_STAGES) & 1);
      };
      auto wait_v = [&](int t) {
        if (t < n_tiles) mbar_wait(&v_full[t % TC_STAGES], (t / TC_STAGES) & 1);
      };
      auto qk_mmas = [&](int t) {  // S[t % 3] = Q K_t^T (K_t known to be in shared memory)
        const uint64_t kdesc = umma_desc_kmajor_sw128(smem_u32(sK + (t % TC_STAGES) * TC_KTILE));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + (t % 3) * TC_BN, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0);
      };
      mbar_wait(q_full, 0);
      for (int t = 0; t < 3 && t < n_tiles; ++t) {
        wait_k(t);
        tc_fence_after_sync();
        qk_mmas(t);
        umma_commit(&s_full[t % 3]);
        umma_commit(&k_empty[t % TC_STAGES]);
        qk_tile[t % 3] = t;
      }
      wait_v(0), wait_v(1), wait_k(3), wait_k(4);
      // The two groups' tiles are served in whatever order their P becomes ready (no head-of-line blocking): nxt[g] = the
      // group's next tile.  A rendezvous request is served only when neither group has a complete P waiting, so every P
      // that is complete at the rendezvous has gone through its PV before O is rescaled.
      int nxt[2] = {0, 1};
      int remaining = n_tiles;
      bool first = true;
      uint32_t served = 0, spins = 0;
      // mode 0: non-blocking poll (mbarrier.test_wait); mode 1: poll that may suspend for ~TC_POLL_NS (try_wait + time hint).
      // A plain try_wait may suspend the thread for a long, system-dependent time on ONE group's barrier while the other
      // group's P has long been complete.
      auto try_tile = [&](int g, int mode) -> bool {
        const int j = nxt[g];
        // p_full[j % 3] is shared with tile j - 3 (the other group's): its phase for tile j may only be polled once the phase
        // of tile j - 3 has been consumed, otherwise the parity test aliases with the phase before that one
        if (j >= n_tiles || (j >= 3 && nxt[g ^ 1] <= j - 3)) return false;
        if (mode == 0 ? !mbar_test_wait(&p_full[j % 3], (j / 3) & 1) : !mbar_try_wait_hint(&p_full[j % 3], (j / 3) & 1, TC_POLL_NS))
          return false;
        ATTN_TRACE(2, j, 1);
        tc_fence_after_sync();
        const int s = j % TC_STAGES, b = j % 3;
        const uint64_t vdesc = umma_desc_mnmajor_sw128(smem_u32(sV + s * TC_KTILE), 1024);
        const uint32_t p_tmem = tmem_base + b * TC_BN;  // P_j (bf16, two keys per column) over the first 32 columns of its S buffer
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 keys per MMA: A +8 TMEM columns, B +16 key rows = 2 KB (+128)
          umma_bf16_ts(tmem_base + TC_O_COL, p_tmem + 8 * k, vdesc + 128 * k, idesc_pv, !first || k != 0);
        first = false;
        ATTN_TRACE(2, j, 2);
        if (j + 3 < n_tiles) qk_mmas(j + 3);       // the buffer just consumed is refilled three tiles ahead
        umma_commit(&v_empty[s]);
        if (j + 3 < n_tiles) {
          umma_commit(&s_full[b]);
          umma_commit(&k_empty[(j + 3) % TC_STAGES]);
          qk_tile[b] = j + 3;
        }
        ATTN_TRACE(2, j, 3);
        nxt[g] = j + 2;
        --remaining;
        wait_v(j + 2);                              // operands of this group's next service (see above)
        ATTN_TRACE(3, j + 2, 1);                    // V_{j+2} seen in shared memory (an upper bound of its arrival time)
        wait_k(j + 5);
        return true;
      };
      while (remaining > 0) {
        // The tile with the lower index is the one expected first (the groups alternate): a SUSPENDED wait on it (wakes on
        // completion or after ~TC_POLL_NS), then a look at the other group's tile and at the rendezvous request.  No hot
        // spinning: this thread shares its sub-partition's issue slots with two softmax warps, and the slowest softmax warp
        // sets the pace of the whole CTA.
        const bool zero_first = nxt[0] < nxt[1];
        if (zero_first ? try_tile(0, 1) : try_tile(1, 1)) {
          spins = 0;
          continue;
        }
        if (zero_first ? try_tile(1, 0) : try_tile(0, 0)) {
          spins = 0;
          continue;
        }
        if (mbar_test_wait(sync_req, served & 1)) {
          // every softmax warp is parked: no more arrivals.  Drain what completed since the last look, then quiesce.
          while (try_tile(0, 0) || try_tile(1, 0)) {
          }
          umma_commit(quiesce);
          mbar_wait(resume, served & 1);
          tc_fence_after_sync();          // O (and pending P rows) were rewritten with tcgen05.st by the softmax warps
          ++served;
          continue;
        }
        if (++spins == (1u << 24)) {
          printf("some_b200: attention v8 MMA thread timeout block %d (next tiles %d %d of %d)\n", (int)blockIdx.x, nxt[0], nxt[1], n_tiles);
          __trap();
        }
      }
      umma_commit(all_done);
    }
    __syncwarp();
  } else {
    const int g = (warp - 2) >> 2;  // softmax group: 0 = even key tiles, 1 = odd key tiles
    const int quad = hw_warp & 3;   // the TMEM lane quadrant this warp may touch
    const int r = quad * 32 + lane; // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_o = t_lane + TC_O_COL;
    const float c = 0.125f * 1.4426950408889634f;  // dim_head^-0.5 * log2(e)
    const uint64_t c2 = f2_pack(c, c);
    float R = -INFINITY;   // shared reference maximum of this row (identical in both groups: it only changes at a rendezvous)
    float l = 0.f;         // this group's part of the row sum, relative to R
    uint32_t epoch = 0;    // rendezvous completed
    int my_arrived = g - 2;  // last tile this warp arrived p_full for (g - 2: none yet; keeps the tile parity of the group)

    // CTA-wide rendezvous of the eight softmax warps (+ the parked MMA thread).  offer = the row maximum this thread wants
    // the reference raised to (R itself when it has nothing to ask for).
    auto rendezvous = [&](float offer) {
      mxs[g * TC_BM + r] = offer;
      if (lane == 0) arrived[warp - 2] = my_arrived;
      __syncwarp();
      asm volatile("bar.sync 2, 256;" ::: "memory");            // every softmax warp is here: no new P arrivals from now on
      if (warp == 2 && lane == 0) {
        *sync_flag = 0;                                          // requests raised after this point start a new rendezvous
        mbar_arrive(sync_req);
      }
      mbar_wait(quiesce, epoch & 1);                             // every MMA issued so far has retired: O is quiescent
      tc_fence_after_sync();
      const float Rn = fmaxf(fmaxf(mxs[r], mxs[TC_BM + r]), R);
      const float alpha = (R == -INFINITY) ? 0.f : ex2_approx((R - Rn) * c);
      const bool touch_o = (R != -INFINITY) && (alpha != 1.0f);
      R = Rn;
      l *= alpha;
      // A P row this warp has written whose PV has not been issued (a sibling warp of the group has not arrived for that
      // tile yet: the MMA thread drained every COMPLETE tile before it quiesced) is still on the old scale: bring it along.
      {
        int gmin = my_arrived;
#pragma unroll
        for (int w = 0; w < 4; ++w) gmin = min(gmin, arrived[4 * g + w]);
        // with three S buffers a warp can be up to three of its group's tiles ahead of its slowest sibling: EVERY tile in
        // (gmin, my_arrived] is incomplete, its PV not issued, and this warp's rows of it are on the old scale
        if (__any_sync(0xffffffffu, touch_o)) {
#pragma unroll 1
          for (int t = gmin + 2; t <= my_arrived; t += 2) {
            const uint32_t t_p = t_lane + (t % 3) * TC_BN;
            uint32_t pp[32];
            tmem_ld_32x32(t_p, pp);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float lo = __uint_as_float(pp[i] << 16) * alpha, hi = __uint_as_float(pp[i] & 0xffff0000u) * alpha;
              pp[i] = pack_bf16x2(lo, hi);
            }
            tmem_st_32x32(t_p, pp);
            tmem_st_wait();
            tc_fence_before_sync();
          }
        }
      }
      if (g == 0 && __any_sync(0xffffffffu, touch_o)) {          // O rows of this quadrant *= alpha (group 0's warps own the job)
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
        tc_fence_before_sync();
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");            // O rescaled, maxima consumed
      if (warp == 2 && lane == 0) mbar_arrive(resume);
      ++epoch;
    };
    // wait on an mbarrier from a point where joining a rendezvous is safe
    auto wait_joining = [&](uint64_t* bar, uint32_t parity) {
      // warp-uniform decisions only: the lanes poll independently, and a lane that has seen the phase complete must not run
      // ahead (tcgen05.ld is .sync.aligned) while its siblings enter the rendezvous (bar.sync counts threads)
      uint32_t spins = 0;
      while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        if (__any_sync(0xffffffffu, *sync_flag != 0)) rendezvous(R);
        if (++spins == (1u << 24)) {
          printf("some_b200: attention v8 softmax wait timeout block %d warp %d\n", (int)blockIdx.x, warp);
          __trap();
        }
      }
    };

    for (int j = g; j < n_tiles; j += 2) {
      const int b = j % 3;
      const uint32_t t_s = t_lane + b * TC_BN;
      const int valid = min(TC_BN, T - j * TC_BN);  // keys of this tile inside the clip
      ATTN_TRACE(3, j, 0);   // role 3: when the group STARTED waiting for S_j
      // s_full[b] is shared with tile j - 3 (the OTHER group's).  A warp may run up to three tiles ahead of a slow sibling, i.e.
      // get here before Q K_{j-3}^T has been issued; the parity poll for tile j would then alias with the phase of tile j - 6
      // (same parity, long complete).  The MMA thread publishes the tile whose QK it last issued into each buffer: QK_j issued
      // implies PV_{j-3} issued, implies S_{j-3} was consumed, implies its phase is complete -- only then is the poll safe.
      {
        uint32_t spins = 0;
        while (__any_sync(0xffffffffu, qk_tile[b] < j)) {
          if (__any_sync(0xffffffffu, *sync_flag != 0)) rendezvous(R);
          if (++spins == (1u << 24)) {
            printf("some_b200: attention v8 qk_tile wait timeout block %d warp %d tile %d\n", (int)blockIdx.x, warp, j);
            __trap();
          }
        }
      }
      wait_joining(&s_full[b], (j / 3) & 1);
      ATTN_TRACE(g, j, 0);
      tc_fence_after_sync();
      uint32_t v[32];
      // ---- pass 1: row maximum (the scores are re-read from tensor memory in pass 2)
      float mx = -INFINITY;
      {
        uint32_t u[32];
        tmem_ld_32x32(t_s, v);
        tmem_ld_32x32(t_s + 32, u);
        tmem_ld_wait();
        if (valid == TC_BN) {
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          float m4 = -INFINITY, m5 = -INFINITY, m6 = -INFINITY, m7 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
            m4 = fmaxf(m4, fmaxf(__uint_as_float(u[i]), __uint_as_float(u[i + 1])));
            m5 = fmaxf(m5, fmaxf(__uint_as_float(u[i + 2]), __uint_as_float(u[i + 3])));
            m6 = fmaxf(m6, fmaxf(__uint_as_float(u[i + 4]), __uint_as_float(u[i + 5])));
            m7 = fmaxf(m7, fmaxf(__uint_as_float(u[i + 6]), __uint_as_float(u[i + 7])));
          }
          mx = fmaxf(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)), fmaxf(fmaxf(m4, m5), fmaxf(m6, m7)));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
            if (32 + i < valid) mx = fmaxf(mx, __uint_as_float(u[i]));
          }
        }
      }
      ATTN_TRACE(g, j, 1);
      // ---- the reference maximum may only move at a rendezvous (warp-uniform decision; R = -inf on the very first tile)
      while (__any_sync(0xffffffffu, (mx - R) * c > 8.0f || R == -INFINITY)) {
        if (lane == 0) *sync_flag = 1;
        rendezvous(fmaxf(R, mx));
      }
      const float mc = R * c;
      const uint64_t nmc2 = f2_pack(-mc, -mc);
      // ---- pass 2: p = 2^(s c - R c), row sum, bf16 pack
      uint32_t pk[32];
      uint64_t rs_a = f2_pack(0.f, 0.f), rs_b = rs_a;
      auto quarter = [&](const uint32_t(&x)[16], int q) {
        if (valid == TC_BN) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const uint64_t ya = f2_fma(f2_pack(__uint_as_float(x[i]), __uint_as_float(x[i + 1])), c2, nmc2);
            const uint64_t yb = f2_fma(f2_pack(__uint_as_float(x[i + 2]), __uint_as_float(x[i + 3])), c2, nmc2);
            float p0, p1, p2, p3;
            {
              float y0, y1;
              f2_unpack(ya, y0, y1);
              p0 = ex2_approx(y0);
              p1 = ex2_approx(y1);
            }
            if (TC_POLY_OF_8 >= 4 || ((i & 4) && TC_POLY_OF_8 >= 2)) {
              exp2_poly2(yb, p2, p3);
            } else {
              float y2, y3;
              f2_unpack(yb, y2, y3);
              p2 = ex2_approx(y2);
              p3 = ex2_approx(y3);
            }
            rs_a = f2_add(rs_a, f2_pack(p0, p1));
            rs_b = f2_add(rs_b, f2_pack(p2, p3));
            pk[8 * q + (i >> 1)] = pack_bf16x2(p0, p1);
            pk[8 * q + (i >> 1) + 1] = pack_bf16x2(p2, p3);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(x[i]), c, -mc));
            float p1 = ex2_approx(fmaf(__uint_as_float(x[i + 1]), c, -mc));
            if (16 * q + i >= valid) p0 = 0.f;
            if (16 * q + i + 1 >= valid) p1 = 0.f;
            rs_a = f2_add(rs_a, f2_pack(p0, p1));
            pk[8 * q + (i >> 1)] = pack_bf16x2(p0, p1);
          }
        }
      };
      {
        uint32_t xa[16], xb[16];
        tmem_ld_32x16(t_s, xa);
        tmem_ld_wait();
        tmem_ld_32x16(t_s + 16, xb);
        quarter(xa, 0);
        tmem_ld_wait();
        tmem_ld_32x16(t_s + 32, xa);
        quarter(xb, 1);
        tmem_ld_wait();
        tmem_ld_32x16(t_s + 48, xb);
        quarter(xa, 2);
        tmem_ld_wait();
        quarter(xb, 3);
      }
      {
        float s0, s1, s2, s3;
        f2_unpack(rs_a, s0, s1);
        f2_unpack(rs_b, s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      ATTN_TRACE(g, j, 2);
      // ---- P -> TMEM over the first 32 columns of this tile's S buffer; QK_{j+3} overwrites them only after PV_j
      tmem_st_32x32(t_s, pk);
      tmem_st_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[b]);
      ATTN_TRACE(g, j, 3);
      my_arrived = j;
    }
    // ---- drain: keep serving rendezvous until every softmax warp has finished its tiles
    __syncwarp();
    if (lane == 0) atomicAdd(const_cast<int*>(done_cnt), 1);
    {
      uint32_t spins = 0;
      while (__any_sync(0xffffffffu, *done_cnt < 8)) {
        if (__any_sync(0xffffffffu, *sync_flag != 0)) rendezvous(R);
        if (++spins == (1u << 26)) {
          printf("some_b200: attention v8 drain timeout block %d warp %d\n", (int)blockIdx.x, warp);
          __trap();
        }
      }
    }
    // ---- O / (l_A + l_B) -> bf16 -> out[row, head * 64 ..]; group g writes channels [32 g, 32 g + 32)
    lsum[g * TC_BM + r] = l;
    mbar_wait(all_done, 0);
    tc_fence_after_sync();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.0f / (lsum[r] + lsum[TC_BM + r]);
    const int qrow = q0 + r;
    __nv_bfloat16* dst = p.out[grp] + (size_t)(row_begin + qrow) * SOME_DIM + head * 64 + 32 * g;
    uint32_t oa[32];
    tmem_ld_32x32(t_o + 32 * g, oa);
    tmem_ld_wait();
    if (qrow < T) {
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(oa[i]) * inv;
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
  static bool configured[kMaxDevices] = {};   // function attributes are per device
  const int dev_ = device_index();
  if (!configured[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    SOME_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(attention_tc): %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    SOME_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(attention_tc carveout): %s", cudaGetErrorString(e));
    configured[dev_] = true;
  }
  const long long gx = 1ll * p.tiles_per_clip * a->B;
  SOME_REQUIRE(gx < (1ll << 31), "some_attention_varlen: grid too large");
  dim3 grid(static_cast<unsigned>(gx), SOME_HEADS, a->groups);
  attention_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  return check_launch("some_attention_varlen");
}
