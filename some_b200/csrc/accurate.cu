// Reference-accurate fp32 path of the trunk (validation mode, SURVEY.md §7 hard part 3 (ii); north_star: "within 1e-3 fp32").
//
// The product path computes the dense contractions with bf16 operands on the tensor cores (1e-2 tolerance).  To assert the
// fp32 line of the contract — and to show that bf16 rounding is the ONLY reason decoded notes can differ from the
// reference's — the same launch sequence exists with fp32 operands on the CUDA cores: no tensor cores, no bf16 anywhere,
// exact expf-based SiLU / sigmoid / softmax.  It is a checker: ~100x slower than the product path and never used by
// infer() unless asked for (Engine.infer(..., accurate=True)).
//
//   K-sgemm      C = epi(A[M,K] . W[N,K]^T + bias)      nn.Linear / 1x1 Conv1d (same call sites as gemm.cu)
//   K-glu        out = y[:, :C] * sigmoid(y[:, C:]) (+ resid)   Gconform.py:15-18, base_conv.py:12-15
//   K-attn-f32   softmax(q k^T / 8) v per clip and head  base_attention.py:34-45
//   K-dwconv-f32 depthwise k = 31 + folded BatchNorm + SiLU   base_conv.py:66-68
//   K-head-f32   sigmoid / softmax of the logits         Gmidi_conform.py:33-37
// LayerNorm and the bound head reuse some_layernorm (fp32 output) / some_bound_head, which are fp32 already.
#include "host_common.h"

#include <math.h>

#include "../../include/some_b200.h"

namespace some {

constexpr int D = SOME_DIM;
constexpr int FFN = 4 * SOME_DIM;

// ---------------------------------------------------------------------------------------------- K-sgemm
// 64 x 64 tile, K step 16, 256 threads x (4 x 4) outputs, operands staged in shared memory.
enum { F32_EPI_BIAS = 0, F32_EPI_SILU = 1, F32_EPI_RESID = 2 };

template <int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ bias,
             float* C, int ldc, const float* resid, float alpha, int M, int N, int K) {   // resid may alias C
  __shared__ float sA[16][64 + 4];
  __shared__ float sW[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      sA[c][r] = (m0 + r < M && k0 + c < K) ? A[(size_t)(m0 + r) * lda + k0 + c] : 0.f;
      sW[c][r] = (n0 + r < N && k0 + c < K) ? W[(size_t)(n0 + r) * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i], w[i] = sW[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias != nullptr ? bias[n] : 0.f);
      if constexpr (EPI == F32_EPI_SILU) v = v / (1.0f + expf(-v));
      if constexpr (EPI == F32_EPI_RESID) v = fmaf(alpha, v, resid[(size_t)m * ldc + n]);
      C[(size_t)m * ldc + n] = v;
    }
  }
}

static int sgemm(int epi, const float* A, int lda, const float* W, const float* bias, float* C, int ldc, const float* resid,
                 float alpha, int M, int N, int K, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  if (epi == F32_EPI_BIAS) sgemm_kernel<F32_EPI_BIAS><<<grid, 256, 0, st>>>(A, lda, W, bias, C, ldc, resid, alpha, M, N, K);
  else if (epi == F32_EPI_SILU) sgemm_kernel<F32_EPI_SILU><<<grid, 256, 0, st>>>(A, lda, W, bias, C, ldc, resid, alpha, M, N, K);
  else sgemm_kernel<F32_EPI_RESID><<<grid, 256, 0, st>>>(A, lda, W, bias, C, ldc, resid, alpha, M, N, K);
  return check_launch("some_forward_f32(sgemm)");
}

// ---------------------------------------------------------------------------------------------- K-glu
__global__ void glu_f32_kernel(const float* __restrict__ y, float* __restrict__ out, const float* __restrict__ resid, int M, int Cc) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (size_t)M * Cc) return;
  const size_t m = i / Cc, c = i - m * Cc;
  const float a = y[m * 2 * Cc + c], g = y[m * 2 * Cc + Cc + c];
  const float v = a * (1.0f / (1.0f + expf(-g)));
  out[i] = resid != nullptr ? resid[i] + v : v;
}

// ---------------------------------------------------------------------------------------------- K-attn-f32
// One warp per query row of one (clip, head): lane = key inside a 32-key step (dot product over the 64 channels), online
// softmax in fp32, each lane owns output channels {lane, lane + 32}.
__global__ void __launch_bounds__(256)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, const int32_t* __restrict__ cu, int max_frames) {
  const int clip = blockIdx.z, head = blockIdx.y;
  const int row0 = cu[clip], T = cu[clip + 1] - row0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 8 + warp;
  if (qi >= T) return;
  __shared__ float sq[8][64];
  const float* q = qkv + (size_t)(row0 + qi) * 3 * D + head * 64;
  sq[warp][lane] = q[lane];
  sq[warp][lane + 32] = q[lane + 32];
  __syncwarp();
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int j0 = 0; j0 < T; j0 += 32) {
    const int j = j0 + lane;
    float s = -INFINITY;
    if (j < T) {
      const float* k = qkv + (size_t)(row0 + j) * 3 * D + D + head * 64;
      float acc = 0.f;
#pragma unroll 16
      for (int c = 0; c < 64; ++c) acc = fmaf(sq[warp][c], k[c], acc);
      s = acc * 0.125f;
    }
    float mx = s;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float m_new = fmaxf(m, mx);
    const float corr = expf(m - m_new);          // 0 on the first step (m = -inf)
    const float p = (j < T) ? expf(s - m_new) : 0.f;
    float ps = p;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
    l = l * corr + ps;
    o0 *= corr, o1 *= corr;
    const int nk = min(32, T - j0);
    for (int t = 0; t < nk; ++t) {
      const float pt = __shfl_sync(0xffffffffu, p, t);
      const float* v = qkv + (size_t)(row0 + j0 + t) * 3 * D + 2 * D + head * 64;
      o0 = fmaf(pt, v[lane], o0);
      o1 = fmaf(pt, v[lane + 32], o1);
    }
    m = m_new;
  }
  float* dst = out + (size_t)(row0 + qi) * D + head * 64;
  dst[lane] = o0 / l;
  dst[lane + 32] = o1 / l;
}

// ---------------------------------------------------------------------------------------------- K-dwconv-f32
__global__ void dwconv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                  float* __restrict__ out, const int32_t* __restrict__ cu, int B) {
  const int row = blockIdx.x, c = threadIdx.x + blockIdx.y * blockDim.x;   // one frame x 256 channels per block
  // clip of this row (B is small in validation runs: linear search)
  int clip = 0;
  while (clip + 1 < B && cu[clip + 1] <= row) ++clip;
  const int lo = cu[clip], hi = cu[clip + 1];
  float acc = b[c];
#pragma unroll
  for (int k = 0; k < SOME_CONV_K; ++k) {
    const int r = row + k - SOME_CONV_K / 2;
    if (r >= lo && r < hi) acc = fmaf(w[k * D + c], x[(size_t)r * D + c], acc);
  }
  out[(size_t)row * D + c] = acc / (1.0f + expf(-acc));
}

// ---------------------------------------------------------------------------------------------- K-head-f32
__global__ void head_f32_kernel(float* __restrict__ y, int M, int N, int head) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  float* r = y + (size_t)row * N;
  if (head == SOME_EPI_SIGMOID_F32) {
    for (int j = lane; j < N; j += 32) r[j] = 1.0f / (1.0f + expf(-r[j]));
  } else if (head == SOME_EPI_SOFTMAX_F32) {
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, r[j]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) sum += expf(r[j] - mx);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    for (int j = lane; j < N; j += 32) r[j] = expf(r[j] - mx) / sum;
  }
}

}  // namespace some

using namespace some;

namespace {

struct SeqF32 {
  const some_model_f32* m;
  const some_workspace_f32* ws;
  int M, B, max_frames;
  const int32_t* cu;
  cudaStream_t st;
  int rc = 0;

  void gemm(int epi, int s, const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int N, int K,
            float alpha = 1.0f) {
    if (rc) return;
    rc = sgemm(epi, A, lda, W, bias, C, ldc, epi == F32_EPI_RESID ? C : nullptr, alpha, M, N, K, st);
  }
  void ln(int s, const float* g, const float* b, float* out) {
    if (rc) return;
    some_ln_args a{};
    a.x[0] = ws->x[s], a.gamma[0] = g, a.beta[0] = b, a.out_f32[0] = out, a.groups = 1, a.M = M;
    rc = some_layernorm(&a, st);
  }
  void glu(const float* y, float* out, const float* resid) {
    if (rc) return;
    const size_t n = (size_t)M * D;
    glu_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, out, resid, M, D);
    rc = check_launch("some_forward_f32(glu)");
  }
  void block(int s, const some_block_weights_f32& w, bool last) {   // conform_blocke, Gconform.py:56-63, stream s
    float *x = ws->x[s], *a = ws->a[s], *h = ws->h[s], *qkv = ws->qkv[s], *g = ws->g[s], *y = ws->y[s];
    for (int f = 0; f < 2; ++f) {
      if (f == 1) {
        ln(s, w.ln_g[1], w.ln_b[1], a);
        gemm(F32_EPI_BIAS, s, a, D, w.w_qkv, nullptr, qkv, 3 * D, 3 * D, D);
        if (!rc) {
          dim3 grid((max_frames + 7) / 8, SOME_HEADS, B);
          attention_f32_kernel<<<grid, 256, 0, st>>>(qkv, a, cu, max_frames);
          rc = check_launch("some_forward_f32(attention)");
        }
        gemm(F32_EPI_RESID, s, a, D, w.w_out, w.b_out, x, D, D, D);
        ln(s, w.ln_g[2], w.ln_b[2], a);
        gemm(F32_EPI_BIAS, s, a, D, w.w_pw1, w.b_pw1, y, 2 * D, 2 * D, D);
        glu(y, g, nullptr);
        if (!rc) {
          dwconv_f32_kernel<<<dim3(M, D / 256), 256, 0, st>>>(g, w.w_dw, w.b_dw, a, cu, B);
          rc = check_launch("some_forward_f32(dwconv)");
        }
        gemm(F32_EPI_RESID, s, a, D, w.w_pw2, w.b_pw2, x, D, D, D);
      }
      ln(s, w.ln_g[f == 0 ? 0 : 3], w.ln_b[f == 0 ? 0 : 3], a);
      gemm(F32_EPI_SILU, s, a, D, w.ffn_w1[f], w.ffn_b1[f], h, FFN, FFN, D);
      gemm(F32_EPI_RESID, s, h, FFN, w.ffn_w2[f], w.ffn_b2[f], x, D, D, FFN, 0.5f);
    }
    if (!last) ln(s, w.ln_g[4], w.ln_b[4], x);        // norm5 in place: the residual stream of the next Gcf
  }
};

}  // namespace

extern "C" int some_forward_f32(const some_model_f32* m, const some_workspace_f32* ws, int M, int B, const int32_t* cu_frames,
                                int max_frames, int head, cudaStream_t stream) {
  SOME_REQUIRE(m != nullptr && ws != nullptr && cu_frames != nullptr && m->blocks != nullptr, "some_forward_f32: null argument");
  SOME_REQUIRE(head == SOME_EPI_SIGMOID_F32 || head == SOME_EPI_SOFTMAX_F32 || head == SOME_EPI_BIAS_F32,
               "some_forward_f32: bad head %d", head);
  if (M <= 0 || B <= 0) return 0;
  SeqF32 s{m, ws, M, B, max_frames, cu_frames, stream};
  for (int g = 0; g < 2; ++g) s.gemm(F32_EPI_BIAS, g, ws->units, SOME_N_MELS, m->w_in[g], m->b_in[g], ws->x[g], D, D, SOME_N_MELS);
  for (int i = 0; i < m->lay; ++i) {
    for (int g = 0; g < 2; ++g) s.block(g, m->blocks[2 * i + g], false);
    // Gcf.forward :85-87 (both use the norm5 outputs of this layer): midi += GLU(glu2(bound)); bound += GLU(glu1(midi))
    s.gemm(F32_EPI_BIAS, 0, ws->x[1], D, m->glu_w[2 * i + 1], m->glu_b[2 * i + 1], ws->y[0], 2 * D, 2 * D, D);
    s.gemm(F32_EPI_BIAS, 1, ws->x[0], D, m->glu_w[2 * i], m->glu_b[2 * i], ws->y[1], 2 * D, 2 * D, D);
    s.glu(ws->y[0], ws->x[0], ws->x[0]);
    s.glu(ws->y[1], ws->x[1], ws->x[1]);
  }
  for (int g = 0; g < 2; ++g) s.block(g, m->blocks[2 * m->lay + g], true);
  // heads: outln on norm5(midi); cutheard + sigmoid on norm5(bound)
  const some_block_weights_f32& w0 = m->blocks[2 * m->lay];
  const some_block_weights_f32& w1 = m->blocks[2 * m->lay + 1];
  s.ln(0, w0.ln_g[4], w0.ln_b[4], ws->a[0]);
  s.gemm(F32_EPI_BIAS, 0, ws->a[0], D, m->w_head, m->b_head, ws->probs, m->outdim, m->outdim, D);
  if (!s.rc && head != SOME_EPI_BIAS_F32) {
    head_f32_kernel<<<(M + 7) / 8, 256, 0, stream>>>(ws->probs, M, m->outdim, head);
    s.rc = check_launch("some_forward_f32(head)");
  }
  if (!s.rc) s.rc = some_bound_head(ws->x[1], w1.ln_g[4], w1.ln_b[4], m->w_cut, m->b_cut, M, ws->bounds, stream);
  return s.rc;
}
