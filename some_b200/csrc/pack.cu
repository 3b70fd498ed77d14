// Host-side checkpoint packing for non-Python hosts: the transformations of some_b200/weights.py restated as plain C functions
// on HOST memory.  No CUDA call in this file: the caller reads the checkpoint (torch.load layout, base_infer.py:27-33), calls
// these, uploads the results and fills `some_model` / the mel-table arguments of `some_mel_logmel`.
//
//   some_pack_bf16        fp32 -> bf16, round to nearest even (what `.to(torch.bfloat16)` does)
//   some_pack_glu_rows    [2C, ...] rows (out 0..C-1 | gate C..2C-1) -> groups of 32 rows: 16 out rows then their 16 gates
//                         (the row order the GLU epilogues of some_gemm expect: Gconform.py:15-18, base_conv.py:12-15)
//   some_pack_dwconv_bn   depthwise Conv1d(k = 31) + BatchNorm1d(eval) folded into taps [K][C] and a bias [C] (base_conv.py:66-67)
//   some_pack_ln_fold     LayerNorm folded into the next Linear (SOME_EPI_LN_*): W' = bf16(W * gamma), s_n = sum_k W'[n][k],
//                         b' = W beta + b
//   some_mel_tables       librosa.filters.mel(htk=True, Slaney norm) as sparse per-filter tables + the FFT twiddles + the periodic
//                         Hann window (spec.py:8-36,45; the tables some_mel_logmel takes)
#include "host_common.h"

#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/some_b200.h"

using namespace some;

namespace {

inline uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x0040u);  // NaN stays NaN (quiet)
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
inline float bf16_to_f32(uint16_t h) {
  const uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// destination row of source row r under the GLU interleave (C = rows / 2 outputs)
inline int glu_dst_row(int r, int c) { return r < c ? (r / 16) * 32 + (r % 16) : ((r - c) / 16) * 32 + 16 + ((r - c) % 16); }

}  // namespace

extern "C" int some_pack_bf16(const float* src, long long n, uint16_t* dst) {
  SOME_REQUIRE(src != nullptr && dst != nullptr && n >= 0, "some_pack_bf16: bad arguments");
  for (long long i = 0; i < n; ++i) dst[i] = bf16_rne(src[i]);
  return 0;
}

extern "C" int some_pack_glu_rows(const void* src, int elem_bytes, int rows, long long row_elems, void* dst) {
  SOME_REQUIRE(src != nullptr && dst != nullptr && src != dst, "some_pack_glu_rows: null or aliased buffers");
  SOME_REQUIRE(elem_bytes > 0 && row_elems > 0 && rows > 0 && rows % 32 == 0,
               "some_pack_glu_rows: rows must be a positive multiple of 32 (got %d)", rows);
  const size_t row_bytes = static_cast<size_t>(elem_bytes) * static_cast<size_t>(row_elems);
  const int c = rows / 2;
  for (int r = 0; r < rows; ++r)
    memcpy(static_cast<uint8_t*>(dst) + row_bytes * glu_dst_row(r, c), static_cast<const uint8_t*>(src) + row_bytes * r, row_bytes);
  return 0;
}

extern "C" int some_pack_dwconv_bn(const float* dw_weight, const float* dw_bias, const float* bn_weight, const float* bn_bias,
                                   const float* bn_mean, const float* bn_var, int channels, int taps, float* out_taps,
                                   float* out_bias) {
  SOME_REQUIRE(dw_weight && dw_bias && bn_weight && bn_bias && bn_mean && bn_var && out_taps && out_bias,
               "some_pack_dwconv_bn: null pointer");
  SOME_REQUIRE(channels > 0 && taps > 0, "some_pack_dwconv_bn: bad shape");
  for (int c = 0; c < channels; ++c) {
    const double scale = static_cast<double>(bn_weight[c]) / sqrt(static_cast<double>(bn_var[c]) + 1e-5);   // BatchNorm1d eps
    for (int k = 0; k < taps; ++k)
      out_taps[static_cast<size_t>(k) * channels + c] = static_cast<float>(static_cast<double>(dw_weight[static_cast<size_t>(c) * taps + k]) * scale);
    out_bias[c] = static_cast<float>((static_cast<double>(dw_bias[c]) - static_cast<double>(bn_mean[c])) * scale + static_cast<double>(bn_bias[c]));
  }
  return 0;
}

extern "C" int some_pack_ln_fold(const float* w, const float* bias, const float* gamma, const float* beta, int n, int k,
                                 int glu_rows, uint16_t* w_out, float* s_out, float* b_out) {
  SOME_REQUIRE(w && gamma && beta && w_out && s_out && b_out, "some_pack_ln_fold: null pointer");
  SOME_REQUIRE(n > 0 && k > 0 && (!glu_rows || n % 32 == 0), "some_pack_ln_fold: bad shape N=%d K=%d", n, k);
  for (int r = 0; r < n; ++r) {
    const int d = glu_rows ? glu_dst_row(r, n / 2) : r;
    const float* wr = w + static_cast<size_t>(r) * k;
    double dot = 0.0, s = 0.0;
    for (int j = 0; j < k; ++j) {
      dot += static_cast<double>(wr[j]) * static_cast<double>(beta[j]);
      const uint16_t h = bf16_rne(static_cast<float>(static_cast<double>(wr[j]) * static_cast<double>(gamma[j])));
      w_out[static_cast<size_t>(d) * k + j] = h;
      s += static_cast<double>(bf16_to_f32(h));          // sums of the ROUNDED operand the tensor core sees
    }
    s_out[d] = static_cast<float>(s);
    b_out[d] = static_cast<float>(dot + (bias ? static_cast<double>(bias[r]) : 0.0));
  }
  return 0;
}

extern "C" int some_mel_tables(int sample_rate, int n_fft, int n_mels, double fmin, double fmax, int32_t* mel_start,
                               int32_t* mel_count, float* mel_weights, float* twiddle, float* window) {
  SOME_REQUIRE(mel_start && mel_count && mel_weights && twiddle && window, "some_mel_tables: null pointer");
  SOME_REQUIRE(n_fft == SOME_N_FFT && n_mels == SOME_N_MELS && sample_rate > 0,
               "some_mel_tables: the fused kernel is specialised for n_fft %d / %d mel bands", SOME_N_FFT, SOME_N_MELS);
  if (!(fmax > 0.0)) fmax = sample_rate / 2.0;
  const int n_bins = 1 + n_fft / 2;
  const double pi = 3.14159265358979323846;
  // librosa.filters.mel(htk=True), Slaney area normalisation (librosa 0.9: requirements.txt:10), all in float64
  auto to_mel = [](double f) { return 2595.0 * log10(1.0 + f / 700.0); };
  std::vector<double> edges(n_mels + 2);
  const double m0 = to_mel(fmin), m1 = to_mel(fmax), mstep = (m1 - m0) / (n_mels + 1);
  for (int i = 0; i < n_mels + 2; ++i) {
    const double mel = i == n_mels + 1 ? m1 : m0 + i * mstep;
    edges[i] = 700.0 * (pow(10.0, mel / 2595.0) - 1.0);
  }
  const double fstep = (sample_rate / 2.0) / (n_bins - 1);
  memset(mel_weights, 0, sizeof(float) * n_mels * SOME_MEL_MAXW);
  for (int m = 0; m < n_mels; ++m) {
    const double lo_w = edges[m + 1] - edges[m], hi_w = edges[m + 2] - edges[m + 1], norm = 2.0 / (edges[m + 2] - edges[m]);
    int first = -1, last = -1;
    std::vector<float> row(n_bins);
    for (int i = 0; i < n_bins; ++i) {
      const double f = i == n_bins - 1 ? sample_rate / 2.0 : i * fstep;
      const double rising = (f - edges[m]) / lo_w, falling = (edges[m + 2] - f) / hi_w;
      double tri_d = rising < falling ? rising : falling;
      if (tri_d < 0.0) tri_d = 0.0;
      const float tri = static_cast<float>(tri_d);
      row[i] = static_cast<float>(static_cast<double>(tri) * norm);
      if (row[i] != 0.f) {
        if (first < 0) first = i;
        last = i;
      }
    }
    mel_start[m] = 0, mel_count[m] = 0;
    if (first < 0) continue;
    SOME_REQUIRE(last < SOME_MEL_BINS && last - first + 1 <= SOME_MEL_MAXW,
                 "some_mel_tables: mel filter %d spans bins %d..%d: outside what the fused kernel keeps (%d bins, %d per filter)",
                 m, first, last, SOME_MEL_BINS, SOME_MEL_MAXW);
    mel_start[m] = first, mel_count[m] = last - first + 1;
    for (int i = first; i <= last; ++i) mel_weights[m * SOME_MEL_MAXW + (i - first)] = row[i];
  }
  // twiddles of the 32 x 32 register FFT (mel.cu): W_1024^(n2 k1) at [k1 * 32 + n2], then W_2048^k for the bins k < 372
  for (int k1 = 0; k1 < 32; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = -2.0 * pi * static_cast<double>(k1 * n2) / 1024.0;
      twiddle[2 * (k1 * 32 + n2)] = static_cast<float>(cos(a));
      twiddle[2 * (k1 * 32 + n2) + 1] = static_cast<float>(sin(a));
    }
  for (int kk = 0; kk < SOME_MEL_BINS; ++kk) {
    const double a = -2.0 * pi * static_cast<double>(kk) / 2048.0;
    twiddle[2 * (1024 + kk)] = static_cast<float>(cos(a));
    twiddle[2 * (1024 + kk) + 1] = static_cast<float>(sin(a));
  }
  // torch.hann_window(n_fft, periodic=True), float32 arithmetic (spec.py:45)
  const float wstep = static_cast<float>(2.0 * pi / n_fft);
  for (int i = 0; i < n_fft; ++i) window[i] = 0.5f - 0.5f * cosf(static_cast<float>(i) * wstep);
  return 0;
}
