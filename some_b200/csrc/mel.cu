// K-mel: fused zero-pad + framing + periodic-Hann window + 2048-point real FFT + magnitude + sparse
// triangular mel filterbank + log(clamp) -> [M, 80], written frame-major so the reference's
// `.transpose(1, 2)` (inference/me_infer.py:31) disappears.  Restates modules/rmvpe/spec.py:38-72
// (keyshift = 0, speed = 1, center = True).
//
// CTA = 16 consecutive frames of one clip (256 threads, two frames in flight, 128 threads each).
//   1. the (16 + 3) x 512 samples the frames overlap are staged once in shared memory with coalesced
//      float4 loads; samples outside [0, L) read as zero (the reference's F.pad 1024 / 1024);
//   2. per frame: z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1], in-place radix-4 DIF FFT of 1024 complex points in
//      shared memory (5 stages, host-computed double-precision per-stage twiddle tables), result in base-4 digit-reversed order;
//   3. real-FFT unpack for bins 0..371 only (mel weights above 8 kHz are zero), |X|;
//   4. mel[m] = sum over the filter's contiguous bin range (<= 24 bins), log(max(., clamp)).
// Algorithmic traffic: 512 x 4 B in + 80 x 4 B out per frame (2368 B); the FFT itself (~56 kFLOP/frame fp32,
// 80 KB of shared-memory traffic) makes this kernel shared-memory/FMA bound rather than HBM bound.
#include "host_common.h"
#include "sm100_ptx.cuh"

#include "../../include/some_b200.h"

namespace some {

constexpr int MEL_FPB = 16;                    // frames per CTA
constexpr int MEL_NS = (MEL_FPB + 3) * 512;    // staged samples
constexpr int MEL_MAXW = SOME_MEL_MAXW;
constexpr int MEL_TW = SOME_MEL_TW;             // per-stage twiddles (3 x 256 | 3 x 64 | 3 x 16 | 3 x 4) + 372 unpack
constexpr int MEL_SMEM = MEL_NS * 4 + MEL_TW * 8 /*twiddle*/ + 2048 * 4 /*window*/ + 2 * 1024 * 8 /*work*/ +
                         2 * SOME_MEL_BINS * 4 /*mag*/;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// position of X[k] after the in-place radix-4 DIF: reverse the five base-4 digits of k
__device__ __forceinline__ int rev4_10(int k) {
  const int b = __brev(static_cast<unsigned>(k)) >> 22;  // 10-bit reversal
  return ((b & 0x2AA) >> 1) | ((b & 0x155) << 1);        // swap the bits inside each digit back
}

// Shared-memory index swizzle of the 1024-point work buffer (float2 elements, 16 per 128-byte bank row): the low four
// index bits are XORed with bits 4-5 (twice) and bits 6-9, which keeps every access pattern of the kernel conflict-free:
// the strided butterflies of the last two radix-4 stages (threads differ in bits 2-5) and the digit-reversed reads of
// the real-FFT unpack (consecutive bins differ only in bits 6-9).
__device__ __forceinline__ int zsw(int i) { return i ^ (((i >> 4) & 3) * 5) ^ ((i >> 6) & 15); }

__global__ void __launch_bounds__(256)
mel_kernel(const float* __restrict__ wave, const int64_t* __restrict__ clip_start, const int64_t* __restrict__ clip_len,
           const int32_t* __restrict__ cu_frames,
           int tiles_per_clip, const int32_t* __restrict__ mel_start, const int32_t* __restrict__ mel_count,
           const float* __restrict__ mel_weights, const float* __restrict__ twiddle, const float* __restrict__ window,
           float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, float clamp) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* s_x = reinterpret_cast<float*>(smem_raw);
  float2* s_tw = reinterpret_cast<float2*>(s_x + MEL_NS);
  float* s_win = reinterpret_cast<float*>(s_tw + MEL_TW);
  float2* s_work = reinterpret_cast<float2*>(s_win + 2048);
  float* s_mag = reinterpret_cast<float*>(s_work + 2 * 1024);

  const int clip = blockIdx.x / tiles_per_clip;
  const int tile = blockIdx.x - clip * tiles_per_clip;
  const int row_begin = cu_frames[clip];
  const int T = cu_frames[clip + 1] - row_begin;
  const int frame0 = tile * MEL_FPB;
  if (frame0 >= T) return;
  const int64_t off = clip_start[clip];
  const int64_t L = clip_len[clip];
  const float* __restrict__ x = wave + off;
  const int nframes = min(MEL_FPB, T - frame0);

  // ---- stage samples [frame0 * 512 - 1024, +MEL_NS) of the clip; zero outside [0, L)
  const int64_t s0 = static_cast<int64_t>(frame0) * 512 - 1024;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  for (int i = threadIdx.x * 4; i < MEL_NS; i += 256 * 4) {
    const int64_t s = s0 + i;
    float4 v;
    if (vec_ok && s >= 0 && s + 3 < L) {
      v = *reinterpret_cast<const float4*>(x + s);
    } else {
      v.x = (s >= 0 && s < L) ? x[s] : 0.f;
      v.y = (s + 1 >= 0 && s + 1 < L) ? x[s + 1] : 0.f;
      v.z = (s + 2 >= 0 && s + 2 < L) ? x[s + 2] : 0.f;
      v.w = (s + 3 >= 0 && s + 3 < L) ? x[s + 3] : 0.f;
    }
    *reinterpret_cast<float4*>(s_x + i) = v;
  }
  for (int i = threadIdx.x; i < MEL_TW; i += 256) s_tw[i] = reinterpret_cast<const float2*>(twiddle)[i];
  for (int i = threadIdx.x; i < 2048; i += 256) s_win[i] = window[i];
  __syncthreads();

  const int slot = threadIdx.x >> 7;   // which of the two in-flight frames
  const int tid = threadIdx.x & 127;
  float2* z = s_work + slot * 1024;
  float* mag = s_mag + slot * SOME_MEL_BINS;

  for (int fpair = 0; fpair < MEL_FPB; fpair += 2) {
    const int f = fpair + slot;          // frame within the CTA
    const bool active = f < nframes;     // uniform per 128-thread half
    // ---- window + pack: z[n] = (w[2n] x[2n], w[2n+1] x[2n+1])
    if (active) {
      const float* xf = s_x + f * 512;
      for (int n = tid; n < 1024; n += 128) {
        const float2 xv = *reinterpret_cast<const float2*>(xf + 2 * n);
        const float2 wv = *reinterpret_cast<const float2*>(s_win + 2 * n);
        z[zsw(n)] = make_float2(xv.x * wv.x, xv.y * wv.y);
      }
    }
    __syncthreads();
    // ---- 5 radix-4 DIF stages, in place
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const int Lg = 1024 >> (2 * s);  // current group length
      const int q = Lg >> 2;
      if (active) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int b = tid + 128 * i;
          const int grp = b / q, j = b - grp * q;
          const int base = grp * Lg + j;
          const int i0 = zsw(base), i1 = zsw(base + q), i2 = zsw(base + 2 * q), i3 = zsw(base + 3 * q);
          const float2 a = z[i0], bb = z[i1], c = z[i2], d = z[i3];
          const float2 t0 = make_float2(a.x + c.x, a.y + c.y);
          const float2 t1 = make_float2(a.x - c.x, a.y - c.y);
          const float2 t2 = make_float2(bb.x + d.x, bb.y + d.y);
          const float2 t3 = make_float2(bb.y - d.y, -(bb.x - d.x));  // -i (b - d)
          float2 y0 = make_float2(t0.x + t2.x, t0.y + t2.y);
          float2 y1 = make_float2(t1.x + t3.x, t1.y + t3.y);
          float2 y2 = make_float2(t0.x - t2.x, t0.y - t2.y);
          float2 y3 = make_float2(t1.x - t3.x, t1.y - t3.y);
          if (s < 4) {  // last stage: all twiddles are 1
            // per-stage tables W_Lg^(m j), m = 1..3, j < q, contiguous in j: consecutive threads -> consecutive words
            // (a single strided W_2048 table made these reads up to 16-way bank conflicted)
            constexpr int kOff[4] = {0, 768, 960, 1008};
            const float2* tws = s_tw + kOff[s];
            y1 = cmul(y1, tws[j]);
            y2 = cmul(y2, tws[q + j]);
            y3 = cmul(y3, tws[2 * q + j]);
          }
          z[i0] = y0, z[i1] = y1, z[i2] = y2, z[i3] = y3;
        }
      }
      __syncthreads();
    }
    // ---- real-FFT unpack, bins 0..371:  X[k] = E + W_2048^k O,  E = (Z[k] + conj Z[N-k]) / 2,
    //      O = -i (Z[k] - conj Z[N-k]) / 2
    if (active) {
      for (int k = tid; k < SOME_MEL_BINS; k += 128) {
        const float2 zk = z[zsw(rev4_10(k))];
        const float2 zn = z[zsw(rev4_10((1024 - k) & 1023))];
        const float2 E = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        const float2 dlt = make_float2(zk.x - zn.x, zk.y + zn.y);  // Z[k] - conj(Z[N-k])
        const float2 O = make_float2(0.5f * dlt.y, -0.5f * dlt.x);  // -i/2 * dlt
        const float2 wo = cmul(s_tw[1020 + k], O);  // W_2048^k
        const float re = E.x + wo.x, im = E.y + wo.y;
        mag[k] = sqrtf(re * re + im * im);
      }
    }
    __syncthreads();
    // ---- mel + log
    if (active && tid < SOME_N_MELS) {
      const int st = __ldg(mel_start + tid), cn = __ldg(mel_count + tid);
      float acc = 0.f;
      for (int i = 0; i < cn; ++i) acc = fmaf(__ldg(mel_weights + tid * MEL_MAXW + i), mag[st + i], acc);
      const float v = logf(fmaxf(acc, clamp));
      const size_t o = static_cast<size_t>(row_begin + frame0 + f) * SOME_N_MELS + tid;
      if (out_f32 != nullptr) out_f32[o] = v;
      if (out_bf16 != nullptr) out_bf16[o] = __float2bfloat16_rn(v);
    }
    // the next iteration's pack overwrites z only after the unpack above (barrier), and mag only after the
    // stage barriers of the next FFT, so no extra barrier is needed here.
  }
}

}  // namespace some

using namespace some;

extern "C" int some_mel_logmel(const float* wave, const int64_t* clip_start, const int64_t* clip_len,
                               const int32_t* cu_frames, int B,
                               int max_frames, const int32_t* mel_start, const int32_t* mel_count,
                               const float* mel_weights, const float* twiddle, const float* window, float* out_f32,
                               uint16_t* out_bf16, float clamp, cudaStream_t stream) {
  SOME_REQUIRE(wave && clip_start && clip_len && cu_frames && mel_start && mel_count && mel_weights && twiddle && window,
               "some_mel_logmel: null pointer");
  SOME_REQUIRE(out_f32 || out_bf16, "some_mel_logmel: no output buffer");
  if (B <= 0 || max_frames <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEL_SMEM);
    SOME_REQUIRE(e == cudaSuccess, "cudaFuncSetAttribute(mel): %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tiles_per_clip = (max_frames + MEL_FPB - 1) / MEL_FPB;
  const long long grid = 1ll * tiles_per_clip * B;
  SOME_REQUIRE(grid < (1ll << 31), "some_mel_logmel: grid too large");
  mel_kernel<<<static_cast<unsigned>(grid), 256, MEL_SMEM, stream>>>(
      wave, clip_start, clip_len, cu_frames, tiles_per_clip, mel_start, mel_count, mel_weights, twiddle, window, out_f32,
      reinterpret_cast<__nv_bfloat16*>(out_bf16), clamp);
  return check_launch("some_mel_logmel");
}
