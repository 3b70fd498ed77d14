// some_forward: the whole two-head conformer trunk (Gmidi_conform.forward, Gconform.py:119-140, plus the head
// activation of midi_conforms.forward, Gmidi_conform.py:30-40) as ONE C call that enqueues the kernel sequence on the
// caller's stream.  Same launches as some_b200/engine.py::Engine.run_trunk (which remains as the per-kernel profiling /
// debugging path); doing the sequencing natively takes the per-launch host cost from ~50 us of Python + ctypes to a few
// microseconds, which matters for small batches (one 10 s clip is ~70 launches of ~5-20 us each) and lets infer() cut
// big batches into more pipeline chunks.
//
// Per conform_blocke (Gconform.py:56-63), both streams (0 = midi / att1, 1 = bound / att2) in every launch:
//   LN1 -> GEMM(ffn1.ln1)+SiLU -> GEMM(ffn1.ln2)*0.5+x -> LN2 -> GEMM(to_q|to_kv) -> attention -> GEMM(to_out)+x ->
//   LN3 -> GEMM(pointwise_conv1)+GLU -> dwconv+BN+SiLU -> GEMM(pointwise_conv2)+x -> LN4 -> GEMM(ffn2.ln1)+SiLU ->
//   GEMM(ffn2.ln2)*0.5+x -> LN5
#include "host_common.h"

#include "../../include/some_b200.h"

namespace {

constexpr int D = SOME_DIM;
constexpr int FFN = 4 * SOME_DIM;

struct Seq {
  const some_model* m;
  const some_workspace* ws;
  int M, B, max_frames;
  const int32_t* cu;
  cudaStream_t st;
  int rc = 0;

  void gemm(const uint16_t* a0, const uint16_t* a1, const uint16_t* w0, const uint16_t* w1, const float* b0, const float* b1,
            void* o0, void* o1, const float* r0, const float* r1, int n, int k, int ld_out, int epi, float alpha = 1.0f,
            int groups = 2) {
    if (rc) return;
    some_gemm_args g{};
    g.A[0] = a0, g.A[1] = a1, g.W[0] = w0, g.W[1] = w1, g.bias[0] = b0, g.bias[1] = b1;
    g.out[0] = o0, g.out[1] = o1, g.resid[0] = r0, g.resid[1] = r1;
    g.groups = groups, g.M = M, g.N = n, g.K = k, g.lda = k, g.ld_out = ld_out, g.epilogue = epi, g.alpha = alpha;
    rc = some_gemm(&g, st);
  }
  void ln(const some_block_weights& w0, const some_block_weights& w1, int i, bool f32_out) {
    if (rc) return;
    some_ln_args a{};
    a.x[0] = ws->x[0], a.x[1] = ws->x[1];
    a.gamma[0] = w0.ln_g[i], a.gamma[1] = w1.ln_g[i], a.beta[0] = w0.ln_b[i], a.beta[1] = w1.ln_b[i];
    a.out_bf16[0] = ws->a[0], a.out_bf16[1] = ws->a[1];
    if (f32_out) a.out_f32[0] = ws->x[0], a.out_f32[1] = ws->x[1];
    a.groups = 2, a.M = M;
    rc = some_layernorm(&a, st);
  }
  void ffn(const some_block_weights& w0, const some_block_weights& w1, int i) {
    gemm(ws->a[0], ws->a[1], w0.ffn_w1[i], w1.ffn_w1[i], w0.ffn_b1[i], w1.ffn_b1[i], ws->h[0], ws->h[1], nullptr, nullptr,
         FFN, D, FFN, SOME_EPI_SILU_BF16);
    gemm(ws->h[0], ws->h[1], w0.ffn_w2[i], w1.ffn_w2[i], w0.ffn_b2[i], w1.ffn_b2[i], ws->x[0], ws->x[1], ws->x[0],
         ws->x[1], D, FFN, D, SOME_EPI_RESID_F32, 0.5f);
  }
  void block(const some_block_weights& w0, const some_block_weights& w1, bool last) {
    ln(w0, w1, 0, false);
    ffn(w0, w1, 0);                                                                      // Gconform.py:57
    ln(w0, w1, 1, false);
    gemm(ws->a[0], ws->a[1], w0.w_qkv, w1.w_qkv, nullptr, nullptr, ws->qkv[0], ws->qkv[1], nullptr, nullptr, 3 * D, D, 3 * D,
         SOME_EPI_STORE_BF16);
    if (!rc) {
      some_attn_args at{};
      at.qkv[0] = ws->qkv[0], at.qkv[1] = ws->qkv[1], at.out[0] = ws->a[0], at.out[1] = ws->a[1];
      at.groups = 2, at.B = B, at.M = M, at.cu_frames = cu, at.max_frames = max_frames;
      rc = some_attention_varlen(&at, st);
    }
    gemm(ws->a[0], ws->a[1], w0.w_out, w1.w_out, w0.b_out, w1.b_out, ws->x[0], ws->x[1], ws->x[0], ws->x[1], D, D, D,
         SOME_EPI_RESID_F32);                                                            // :60
    ln(w0, w1, 2, false);
    gemm(ws->a[0], ws->a[1], w0.w_pw1, w1.w_pw1, w0.b_pw1, w1.b_pw1, ws->g[0], ws->g[1], nullptr, nullptr, 2 * D, D, D,
         SOME_EPI_GLU_BF16);                                                             // base_conv.py:65
    if (!rc) {
      some_dwconv_args dw{};
      dw.x[0] = ws->g[0], dw.x[1] = ws->g[1], dw.w[0] = w0.w_dw, dw.w[1] = w1.w_dw, dw.b[0] = w0.b_dw, dw.b[1] = w1.b_dw;
      dw.out[0] = ws->a[0], dw.out[1] = ws->a[1];
      dw.groups = 2, dw.B = B, dw.cu_frames = cu, dw.max_frames = max_frames;
      rc = some_dwconv_bn_silu(&dw, st);                                                 // base_conv.py:66-68
    }
    gemm(ws->a[0], ws->a[1], w0.w_pw2, w1.w_pw2, w0.b_pw2, w1.b_pw2, ws->x[0], ws->x[1], ws->x[0], ws->x[1], D, D, D,
         SOME_EPI_RESID_F32);                                                            // base_conv.py:69 + Gconform.py:61
    ln(w0, w1, 3, false);
    ffn(w0, w1, 1);                                                                      // :62
    if (!last) {
      ln(w0, w1, 4, true);                                                               // :63, residual of the next Gcf
    } else if (!rc) {
      // final pair: midi stream -> normalised bf16 for outln; bound stream -> fused norm5 + cutheard + sigmoid
      some_ln_args a{};
      a.x[0] = a.x[1] = ws->x[0], a.gamma[0] = a.gamma[1] = w0.ln_g[4], a.beta[0] = a.beta[1] = w0.ln_b[4];
      a.out_bf16[0] = a.out_bf16[1] = ws->a[0];
      a.groups = 1, a.M = M;
      rc = some_layernorm(&a, st);
      if (!rc) rc = some_bound_head(ws->x[1], w1.ln_g[4], w1.ln_b[4], m->w_cut, m->b_cut, M, ws->bounds, st);
    }
  }
};

}  // namespace

extern "C" int some_forward(const some_model* m, const some_workspace* ws, int M, int B, const int32_t* cu_frames,
                            int max_frames, int head, cudaStream_t stream) {
  SOME_REQUIRE(m != nullptr && ws != nullptr && cu_frames != nullptr, "some_forward: null argument");
  SOME_REQUIRE(m->blocks != nullptr && m->lay >= 0 && m->outdim >= 1 && m->outdim <= 256, "some_forward: bad model");
  SOME_REQUIRE(head == SOME_EPI_SIGMOID_F32 || head == SOME_EPI_SOFTMAX_F32 || head == SOME_EPI_BIAS_F32,
               "some_forward: head must be SOME_EPI_SIGMOID_F32 / SOME_EPI_SOFTMAX_F32 / SOME_EPI_BIAS_F32 (got %d)", head);
  if (M <= 0 || B <= 0) return 0;
  Seq s{m, ws, M, B, max_frames, cu_frames, stream};
  // inln / inln1 (Gconform.py:122-125): both streams read the same units
  s.gemm(ws->units, ws->units, m->w_in[0], m->w_in[1], m->b_in[0], m->b_in[1], ws->x[0], ws->x[1], nullptr, nullptr, D,
         SOME_N_MELS, D, SOME_EPI_BIAS_F32);
  for (int i = 0; i < m->lay; ++i) {
    s.block(m->blocks[2 * i], m->blocks[2 * i + 1], false);
    // Gcf.forward :85-87: midi += GLU(glu2(bound)); bound += GLU(glu1(midi))   (ws->a = bf16 copies of the norm5 outputs)
    s.gemm(ws->a[1], ws->a[0], m->glu_w[2 * i + 1], m->glu_w[2 * i], m->glu_b[2 * i + 1], m->glu_b[2 * i], ws->x[0],
           ws->x[1], ws->x[0], ws->x[1], 2 * D, D, D, SOME_EPI_GLU_RESID_F32);
  }
  s.block(m->blocks[2 * m->lay], m->blocks[2 * m->lay + 1], true);
  s.gemm(ws->a[0], ws->a[0], m->w_head, m->w_head, m->b_head, m->b_head, ws->probs, ws->probs, nullptr, nullptr, m->outdim,
         D, m->outdim, head, 1.0f, 1);                                                   // outln (+ sigmoid / softmax)
  return s.rc;
}
