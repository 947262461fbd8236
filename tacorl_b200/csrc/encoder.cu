// LMP vision encoder forward / backward (C-ABI entry points tacorl_lmp_encoder_{fwd,bwd}).
// Replaces the torch ops behind LMPVisionEncoder.forward
// (/root/reference/src/tacorl/networks/visual_encoders/encoder.py:369-419) and
// SpatialSoftArgmax.forward (networks/visual_encoders/utils.py:39-76):
//   conv(3->32,k8,s4)+ReLU, conv(32->64,k4,s2)+ReLU, conv(64->64,k3,s1)+ReLU,
//   spatial soft-argmax (learned temperature, pixel coordinates), FC 128->hidden+ReLU, FC hidden->latent.
// Saved activations are NHWC: fp32 on the fp32 parity path (im2col + SIMT GEMM); on the bf16 tensor-core
// path (`prec` = TACORL_PREC_BF16) y1 / y2 are bf16 (written by the tcgen05 GEMM epilogue) and y3 stays fp32.
#include "common.cuh"
#include "internal.h"
#include "../../include/tacorl_b200.h"
#include <cuda_bf16.h>
#include <cstdlib>

namespace tacorl {

struct EncGeom {
  int N, H, W, H1, W1, H2, W2, H3, W3;
  long long P1, P2, P3;
  bool ok;
  EncGeom(int n, int h, int w) : N(n), H(h), W(w) {
    H1 = (H - 8) / 4 + 1; W1 = (W - 8) / 4 + 1;
    H2 = (H1 - 4) / 2 + 1; W2 = (W1 - 4) / 2 + 1;
    H3 = H2 - 2; W3 = W2 - 2;
    P1 = (long long)H1 * W1; P2 = (long long)H2 * W2; P3 = (long long)H3 * W3;
    ok = H >= 8 && W >= 8 && H1 >= 4 && W1 >= 4 && H3 >= 1 && W3 >= 1;
  }
  long long col_floats_per_frame() const {
    long long a = P1 * 192, b = P2 * 512, c = P3 * 576;
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
  }
};

enum { P_W1 = 0, P_B1, P_W2, P_B2, P_W3, P_B3, P_TEMP, P_W4, P_B4, P_W5, P_B5, P_COUNT };
static const size_t kSplitKWs = 64ull << 20;

static size_t fwd_fixed_bytes(const EncGeom& g, int hidden, bool need_y12, bool need_rest) {
  size_t b = (64 * 512 + 64 * 576) * 4 + 1024;
  if (need_rest) b += (size_t)g.N * (g.P3 * 64 + 128 + 64 + 64 + hidden) * 4 + 2048;
  (void)need_y12;
  return b;
}

}  // namespace tacorl

using namespace tacorl;

extern "C" {

size_t tacorl_lmp_encoder_ws_bytes(int N, int H, int W, int hidden, int latent, int backward) {
  EncGeom g(N, H, W);
  if (!g.ok) return 0;
  int chunk = N < 128 ? N : 128;
  size_t per_frame = (size_t)g.col_floats_per_frame() * 4 + (g.P1 * 32 + g.P2 * 64) * 6 + 1024;
  size_t fixed = fwd_fixed_bytes(g, hidden, true, true) + kSplitKWs + (8 << 20);
  if (backward) fixed += (size_t)N * (hidden + 128 + g.P3 * 64 + 1) * 4 + 4 * (64 * 576) * 4 + (1 << 20);
  (void)latent;
  // implicit-GEMM tensor-core path: no col matrix, whole-batch bf16 gradient / s2d buffers instead
  size_t tc = (size_t)N * ((size_t)(g.H1 + 1) * (g.W1 + 1) * 128 + g.P1 * 64 + g.P2 * 128 + g.P3 * 64 * 6 +
                           (g.P1 * 64 + g.P2 * 128) + (size_t)(g.P2 - g.P3) * 128 + (size_t)(g.H1 + g.W1 + 1) * 64 + (128 + 64 + 64 + hidden + 128 + hidden) * 4 + 4096) +
              kSplitKWs + (32 << 20);
  size_t legacy = fixed + per_frame * chunk + (size_t)N * 3 * H * W * 4 + 4096;   // (+ fp32 copy of uint8 frames)
  return tc > legacy ? tc : legacy;
}

// ------------------------------------------------------------------------------------------ FC head, small batches
// Up to 256 frames (the 64 / 128-frame encoder calls of a TACO-RL / CQL step, rollouts): Linear(128, hidden) + ReLU +
// Linear(hidden, latent) and its whole backward run as one fused cluster kernel each way (mlp_chain.cu, fp32 FFMA)
// instead of 2 + 4 latency-bound GEMM launches with their casts, split-K reductions, column sums and ReLU gate.
// h4 keeps its meaning (the post-ReLU activation): TACORL_MLP_SAVE_ACTIVATED.
static bool fc_fused_ok(int N, int hidden, int latent) {
  return N >= 1 && N <= 256 && hidden % 4 == 0 && hidden <= 256 && latent >= 1 && latent <= 256;
}
static void fc_layers(tacorl_mlp_layer* ly, const float* const* params, int hidden, int latent, float* const* grads) {
  memset(ly, 0, 2 * sizeof(tacorl_mlp_layer));
  ly[0].W0 = params[P_W4]; ly[0].b0 = params[P_B4]; ly[0].n0 = hidden; ly[0].in = 128;
  ly[0].act = ACT_RELU | TACORL_MLP_SAVE_ACTIVATED;
  ly[1].W0 = params[P_W5]; ly[1].b0 = params[P_B5]; ly[1].n0 = latent; ly[1].in = hidden; ly[1].act = ACT_NONE;
  if (grads) {
    ly[0].dW0 = grads[P_W4]; ly[0].db0 = grads[P_B4];
    ly[1].dW0 = grads[P_W5]; ly[1].db0 = grads[P_B5];
  }
}
static int fc_head_fwd_fused(int N, const float* const* params, int hidden, int latent, const float* feat, float* h4,
                             float* emb, cudaStream_t st) {
  tacorl_mlp_layer ly[2];
  fc_layers(ly, params, hidden, latent, nullptr);
  return tacorl_mlp_chain_fwd(2, N, ly, feat, 128, 128, nullptr, 0, 0, h4, hidden, emb, latent, st);
}
static int fc_head_bwd_fused(int N, const float* const* params, int hidden, int latent, const float* feat,
                             const float* h4, const float* d_emb, float* const* grads, float* dfeat, void* ws,
                             size_t ws_bytes, cudaStream_t st) {
  tacorl_mlp_layer ly[2];
  fc_layers(ly, params, hidden, latent, grads);
  return tacorl_mlp_chain_bwd(2, N, ly, feat, 128, 128, nullptr, 0, 0, h4, hidden, d_emb, latent, dfeat, 128, nullptr, 0,
                              ws, ws_bytes, st);
}

// ------------------------------------------------------------------------------------------ tensor-core path
// conv stack as implicit GEMMs (conv_tc.cu): x -> space-to-depth bf16 -> conv1 -> conv2 -> conv3, NHWC bf16
// activations, packed bf16 weights resident in shared memory, no im2col buffers.
static int enc_fwd_tc(const void* xv, int x_u8, float x_scale, float x_shift, int N, int H, int W, const float* const* params, int hidden, int latent,
                      void* y1, void* y2, float* y3, float* feat, float* smax, float* ssum, float* h4, float* emb,
                      void* xs_save, void* ws, size_t ws_bytes, cudaStream_t st) {
  EncGeom g(N, H, W);
  Arena ar(ws, ws_bytes);
  __nv_bfloat16* wp1 = ar.take<__nv_bfloat16>(4 * 32 * 64);
  __nv_bfloat16* wp2 = ar.take<__nv_bfloat16>(8 * 64 * 64);
  __nv_bfloat16* wp3 = ar.take<__nv_bfloat16>(9 * 64 * 64);
  float* fcws = ar.take<float>((8 << 20) / 4);
  __nv_bfloat16* xs = xs_save ? (__nv_bfloat16*)xs_save : ar.take<__nv_bfloat16>((size_t)N * (g.H1 + 1) * (g.W1 + 1) * 64);
  __nv_bfloat16* y1b = y1 ? (__nv_bfloat16*)y1 : ar.take<__nv_bfloat16>((size_t)N * g.P1 * 32);
  __nv_bfloat16* y2b = y2 ? (__nv_bfloat16*)y2 : ar.take<__nv_bfloat16>((size_t)N * g.P2 * 64);
  if (!y3) y3 = ar.take<float>((size_t)N * g.P3 * 64);
  if (!feat) feat = ar.take<float>((size_t)N * 128);
  if (!smax) smax = ar.take<float>((size_t)N * 64);
  if (!ssum) ssum = ar.take<float>((size_t)N * 64);
  if (!h4) h4 = ar.take<float>((size_t)N * hidden);
  TACORL_REQUIRE(wp1 && wp2 && wp3 && fcws && xs && y1b && y2b && y3 && feat && smax && ssum && h4,
                 "lmp_encoder_fwd(bf16): workspace too small (%zu bytes)", ws_bytes);
  int rc;
  {
    const int modes[3] = {2, 1, 0};
    const float* Ws[3] = {params[P_W1], params[P_W2], params[P_W3]};
    void* Wps[3] = {wp1, wp2, wp3};
    if ((rc = conv_tc_pack_multi(3, modes, Ws, Wps, st))) return rc;
  }
  if ((rc = x_u8 ? conv_tc_s2d_u8((const unsigned char*)xv, N, H, W, g.H1 + 1, g.W1 + 1, x_scale, x_shift, xs, st)
                 : conv_tc_s2d((const float*)xv, N, H, W, g.H1 + 1, g.W1 + 1, xs, st))) return rc;
  if ((rc = conv_lin_conv1_fwd(xs, N, g.H1, g.W1, wp1, params[P_B1], y1b, st))) return rc;
  if ((rc = conv_tc_conv2_fwd(y1b, N, g.H1, g.W1, g.H2, g.W2, wp2, params[P_B2], y2b, st))) return rc;
  if ((rc = conv_lin_conv3_fwd(y2b, N, g.H2, g.W2, g.H3, g.W3, wp3, params[P_B3], y3, st))) return rc;
  if ((rc = softargmax_fwd_f32(y3, N, g.H3, g.W3, 64, params[P_TEMP], feat, smax, ssum, st))) return rc;
  if (fc_fused_ok(N, hidden, latent)) return fc_head_fwd_fused(N, params, hidden, latent, feat, h4, emb, st);
  GemmArgs f;
  f.transB = 1; f.M = N; f.N = hidden; f.K = 128; f.A = feat; f.lda = 128; f.B = params[P_W4]; f.ldb = 128;
  f.C = h4; f.ldc = hidden; f.bias = params[P_B4]; f.act = ACT_RELU; f.split_k = 1;
  if ((rc = gemm_tc_from_f32(f, fcws, 8 << 20, st))) return rc;
  f.N = latent; f.K = hidden; f.A = h4; f.lda = hidden; f.B = params[P_W5]; f.ldb = hidden; f.C = emb;
  f.ldc = latent; f.bias = params[P_B5]; f.act = ACT_NONE; f.split_k = 0;
  return gemm_tc_from_f32(f, fcws, 8 << 20, st);
}

static int enc_bwd_tc(const void* xv, int x_u8, float x_scale, float x_shift, int N, int H, int W, const float* const* params, int hidden, int latent,
                      const void* y1, const void* y2, const float* y3, const float* feat, const float* smax,
                      const float* ssum, const float* h4, const float* d_emb, float* const* grads, int accumulate,
                      const void* xs_saved, void* ws, size_t ws_bytes, cudaStream_t st) {
  EncGeom g(N, H, W);
  const float beta0 = accumulate ? 1.f : 0.f;
  Arena ar(ws, ws_bytes);
  __nv_bfloat16* wd3 = ar.take<__nv_bfloat16>(9 * 64 * 64);
  __nv_bfloat16* wd2 = ar.take<__nv_bfloat16>(16 * 32 * 64);
  float* dh4 = ar.take<float>((size_t)N * hidden);
  float* dfeat = ar.take<float>((size_t)N * 128);
  float* dtau = ar.take<float>(N);
  float* csws = ar.take<float>(592 * 64);
  float* skws = ar.take<float>(kSplitKWs / 4);
  // conv3's gradient lives at y2's pitch with a zero 2-pixel margin when the linear-shift weight gradient covers the shape
  const bool lin3 = conv_lin_conv3_wgrad_ok(g.W2);
  const int pad3 = lin3 ? 2 : 0;
  __nv_bfloat16* dy3b = ar.take<__nv_bfloat16>((size_t)N * (lin3 ? g.P2 : g.P3) * 64);
  __nv_bfloat16* dy2b = ar.take<__nv_bfloat16>((size_t)N * g.P2 * 64);
  // conv1's gradient at the s2d image's pitch (one zero column / row of margin) for the linear-shift weight gradient
  const bool lin1 = conv_lin_conv1_wgrad_ok(g.W1 + 1);
  __nv_bfloat16* dy1b = ar.take<__nv_bfloat16>((size_t)N * (lin1 ? (size_t)(g.H1 + 1) * (g.W1 + 1) : (size_t)g.P1) * 32);
  __nv_bfloat16* xs = xs_saved ? (__nv_bfloat16*)xs_saved : ar.take<__nv_bfloat16>((size_t)N * (g.H1 + 1) * (g.W1 + 1) * 64);
  TACORL_REQUIRE(wd3 && wd2 && dh4 && dfeat && dtau && csws && skws && dy3b && dy2b && dy1b && xs,
                 "lmp_encoder_bwd(bf16): workspace too small (%zu bytes)", ws_bytes);
  int rc;
  // ---- FC head (small batches: one fused kernel; else small GEMMs, fp32 operands staged to bf16)
  const bool fc_fused = !accumulate && fc_fused_ok(N, hidden, latent);
  if (fc_fused && (rc = fc_head_bwd_fused(N, params, hidden, latent, feat, h4, d_emb, grads, dfeat, skws, kSplitKWs, st)))
    return rc;
  GemmArgs a, b;
  if (!fc_fused) {
  a.transA = 1; a.transB = 0; a.M = latent; a.N = hidden; a.K = N; a.A = d_emb; a.lda = latent; a.B = h4;
  a.ldb = hidden; a.C = grads[P_W5]; a.ldc = hidden; a.beta = beta0; a.split_k = 0;
  if ((rc = gemm_tc_from_f32(a, skws, kSplitKWs, st))) return rc;
  if ((rc = colsum_f32(N, latent, d_emb, latent, grads[P_B5], accumulate, st))) return rc;
  b.M = N; b.N = hidden; b.K = latent; b.A = d_emb; b.lda = latent; b.B = params[P_W5]; b.ldb = hidden;
  b.C = dh4; b.ldc = hidden; b.split_k = 1;
  if ((rc = gemm_tc_from_f32(b, skws, kSplitKWs, st))) return rc;
  if ((rc = act_bwd_f32(ACT_RELU, (long long)N * hidden, dh4, h4, dh4, st))) return rc;
  a.M = hidden; a.N = 128; a.A = dh4; a.lda = hidden; a.B = feat; a.ldb = 128; a.C = grads[P_W4]; a.ldc = 128;
  if ((rc = gemm_tc_from_f32(a, skws, kSplitKWs, st))) return rc;
  if ((rc = colsum_f32(N, hidden, dh4, hidden, grads[P_B4], accumulate, st))) return rc;
  b.N = 128; b.K = hidden; b.A = dh4; b.lda = hidden; b.B = params[P_W4]; b.ldb = 128; b.C = dfeat; b.ldc = 128;
  if ((rc = gemm_tc_from_f32(b, skws, kSplitKWs, st))) return rc;
  }
  // ---- soft-argmax backward (applies conv3's ReLU mask, writes the bf16 operand directly) and temperature gradient
  if ((rc = softargmax_bwd_bf16out(y3, N, g.H3, g.W3, 64, params[P_TEMP], feat, smax, ssum, dfeat, dy3b, dtau, pad3, pad3, st)))
    return rc;
  if ((rc = colsum_f32(N, 1, dtau, 1, grads[P_TEMP], accumulate, st))) return rc;
  // ---- per layer: weight + bias gradient (one kernel), then the data gradient gated by the input's ReLU
  {
    const int modes[2] = {3, 5};
    const float* Ws[2] = {params[P_W3], params[P_W2]};
    void* Wps[2] = {wd3, wd2};
    if ((rc = conv_tc_pack_multi(2, modes, Ws, Wps, st))) return rc;
  }
  if (lin3) {
    if ((rc = conv_lin_conv3_wgrad(dy3b, y2, N, g.H2, g.W2, beta0, grads[P_W3], grads[P_B3], skws, kSplitKWs, st))) {
      if (rc == 1) set_last_error("lmp_encoder_bwd: linear-shift weight gradient refused a shape it was selected for");
      return rc;
    }
    // the margin of the padded gradient reads as zeros, exactly like the out-of-image taps TMA zero-fills
    if ((rc = conv_tc_conv3_dgrad(dy3b, N, g.H2, g.W2, g.H2, g.W2, wd3, y2, dy2b, st))) return rc;
  } else {
    if ((rc = conv_tc_wgrad(3, dy3b, y2, N, g.H2, g.W2, g.H3, g.W3, beta0, grads[P_W3], grads[P_B3], skws, kSplitKWs, st))) return rc;
    if ((rc = conv_tc_conv3_dgrad(dy3b, N, g.H2, g.W2, g.H3, g.W3, wd3, y2, dy2b, st))) return rc;
  }
  if ((rc = conv_tc_wgrad(2, dy2b, y1, N, g.H1, g.W1, g.H2, g.W2, beta0, grads[P_W2], grads[P_B2], skws, kSplitKWs, st))) return rc;
  if ((rc = conv_dgrad2_fused(dy2b, N, g.H1, g.W1, g.H2, g.W2, wd2, y1, dy1b, lin1 ? 1 : 0, st))) return rc;
  // ---- conv1 (weight gradient only; images receive no gradient)
  if (!xs_saved)
  if ((rc = x_u8 ? conv_tc_s2d_u8((const unsigned char*)xv, N, H, W, g.H1 + 1, g.W1 + 1, x_scale, x_shift, xs, st)
                 : conv_tc_s2d((const float*)xv, N, H, W, g.H1 + 1, g.W1 + 1, xs, st))) return rc;
  if (lin1) {
    rc = conv_lin_conv1_wgrad(dy1b, xs, N, g.H1 + 1, g.W1 + 1, beta0, grads[P_W1], grads[P_B1], skws, kSplitKWs, st);
    if (rc == 1) set_last_error("lmp_encoder_bwd: linear-shift weight gradient refused a shape it was selected for");
    return rc;
  }
  return conv_tc_wgrad(1, dy1b, xs, N, g.H1 + 1, g.W1 + 1, g.H1, g.W1, beta0, grads[P_W1], grads[P_B1], skws, kSplitKWs, st);
}

static int enc_fwd(const void* xv, int x_u8, float x_scale, float x_shift, int N, int H, int W,
                   const float* const* params, int hidden,
                   int latent, float* y1, float* y2, float* y3, float* feat, float* smax,
                   float* ssum, float* h4, float* emb, void* xs_save, void* ws, size_t ws_bytes, int prec,
                   cudaStream_t st) {
  EncGeom g(N, H, W);
  const float* x = (const float*)xv;
  TACORL_REQUIRE(g.ok, "lmp_encoder_fwd: image %dx%d too small", H, W);
  TACORL_REQUIRE(prec == PREC_F32 || prec == PREC_BF16, "lmp_encoder_fwd: unknown precision %d", prec);
  TACORL_REQUIRE(x && params && emb && ws, "lmp_encoder_fwd: null pointer");
  if (N == 0) return 0;
  if (prec == PREC_BF16) {
    TACORL_REQUIRE(W % 4 == 0, "lmp_encoder_fwd(bf16): image width must be a multiple of 4 (got %d)", W);
    return enc_fwd_tc(xv, x_u8, x_scale, x_shift, N, H, W, params, hidden, latent, (void*)y1, (void*)y2, y3, feat,
                      smax, ssum, h4, emb, xs_save, ws, ws_bytes, st);
  }
  Arena ar(ws, ws_bytes);
  if (x_u8) {   // fp32 parity path: materialise the normalised fp32 image once
    float* xf = ar.take<float>((size_t)N * 3 * H * W);
    TACORL_REQUIRE(xf, "lmp_encoder_fwd: workspace too small for the uint8 -> fp32 image");
    int rcu = u8_to_f32_normalized((long long)N * 3 * H * W, (const unsigned char*)xv, x_scale, x_shift, xf, st);
    if (rcu) return rcu;
    x = xf;
  }
  float* w2p = ar.take<float>(64 * 512);
  float* w3p = ar.take<float>(64 * 576);
  const bool save12 = (y1 != nullptr);   // y1/y2 null => chunk-local scratch (inference / no_grad)
  TACORL_REQUIRE((y1 == nullptr) == (y2 == nullptr), "lmp_encoder_fwd: y1/y2 must both be given or both null");
  if (!y3) y3 = ar.take<float>((size_t)N * g.P3 * 64);
  if (!feat) feat = ar.take<float>((size_t)N * 128);
  if (!smax) smax = ar.take<float>((size_t)N * 64);
  if (!ssum) ssum = ar.take<float>((size_t)N * 64);
  if (!h4) h4 = ar.take<float>((size_t)N * hidden);
  TACORL_REQUIRE(w2p && w3p && y3 && feat && smax && ssum && h4, "lmp_encoder_fwd: workspace too small");
  size_t per_frame = (size_t)g.col_floats_per_frame() * 4 + 512;
  if (!save12) per_frame += (g.P1 * 32 + g.P2 * 64) * 4 + 512;
  long long chunk = (long long)(ar.left() > 4096 ? (ar.left() - 4096) / per_frame : 0);
  if (chunk > N) chunk = N;
  TACORL_REQUIRE(chunk >= 1, "lmp_encoder_fwd: workspace too small for one frame (%zu bytes left)", ar.left());
  float* col = ar.take<float>((size_t)chunk * g.col_floats_per_frame());
  float* y1c = save12 ? nullptr : ar.take<float>((size_t)chunk * g.P1 * 32);
  float* y2c = save12 ? nullptr : ar.take<float>((size_t)chunk * g.P2 * 64);
  TACORL_REQUIRE(col && (save12 || (y1c && y2c)), "lmp_encoder_fwd: workspace carve failed");

  int rc;
  if ((rc = permute_conv_weight_f32(params[P_W2], w2p, 64, 32, 4, 4, 0, st))) return rc;
  if ((rc = permute_conv_weight_f32(params[P_W3], w3p, 64, 64, 3, 3, 0, st))) return rc;

  for (long long n0 = 0; n0 < N; n0 += chunk) {
    const int nf = (int)((N - n0) < chunk ? (N - n0) : chunk);
    float* y1p = save12 ? y1 + n0 * g.P1 * 32 : y1c;
    float* y2p = save12 ? y2 + n0 * g.P2 * 64 : y2c;
    float* y3p = y3 + n0 * g.P3 * 64;
    // conv1: NCHW input, K order (c,ky,kx) == torch weight layout
    if ((rc = im2col_f32(x + n0 * 3 * H * W, 3LL * H * W, (long long)H * W, W, 1, 3, 8, 8, 4, g.H1, g.W1, nf,
                         col, st, 0))) return rc;
    GemmArgs a;
    a.transB = 1; a.M = (int)(nf * g.P1); a.N = 32; a.K = 192;
    a.A = col; a.lda = 192; a.B = params[P_W1]; a.ldb = 192; a.C = y1p; a.ldc = 32;
    a.bias = params[P_B1]; a.act = ACT_RELU; a.split_k = 1;
    if ((rc = gemm_f32(a, nullptr, 0, st))) return rc;
    // conv2: NHWC y1, K order (ky,kx,c)
    if ((rc = im2col_f32(y1p, g.P1 * 32, 1, (long long)g.W1 * 32, 32, 32, 4, 4, 2, g.H2, g.W2, nf, col, st, 1)))
      return rc;
    a.M = (int)(nf * g.P2); a.N = 64; a.K = 512; a.lda = 512; a.B = w2p; a.ldb = 512; a.C = y2p; a.ldc = 64;
    a.bias = params[P_B2];
    if ((rc = gemm_f32(a, nullptr, 0, st))) return rc;
    // conv3
    if ((rc = im2col_f32(y2p, g.P2 * 64, 1, (long long)g.W2 * 64, 64, 64, 3, 3, 1, g.H3, g.W3, nf, col, st, 1)))
      return rc;
    a.M = (int)(nf * g.P3); a.N = 64; a.K = 576; a.lda = 576; a.B = w3p; a.ldb = 576; a.C = y3p; a.ldc = 64;
    a.bias = params[P_B3];
    if ((rc = gemm_f32(a, nullptr, 0, st))) return rc;
  }
  if ((rc = softargmax_fwd_f32(y3, N, g.H3, g.W3, 64, params[P_TEMP], feat, smax, ssum, st))) return rc;
  if (fc_fused_ok(N, hidden, latent)) return fc_head_fwd_fused(N, params, hidden, latent, feat, h4, emb, st);
  GemmArgs f;
  f.transB = 1; f.M = N; f.N = hidden; f.K = 128; f.A = feat; f.lda = 128; f.B = params[P_W4]; f.ldb = 128;
  f.C = h4; f.ldc = hidden; f.bias = params[P_B4]; f.act = ACT_RELU; f.split_k = 1;
  if ((rc = gemm_f32(f, nullptr, 0, st))) return rc;
  f.N = latent; f.K = hidden; f.A = h4; f.lda = hidden; f.B = params[P_W5]; f.ldb = hidden; f.C = emb;
  f.ldc = latent; f.bias = params[P_B5]; f.act = ACT_NONE;
  return gemm_f32(f, nullptr, 0, st);
}

int tacorl_lmp_encoder_fwd(const void* x, int x_dtype, float x_scale, float x_shift, int N, int H, int W,
                           const float* const* params, int hidden,
                           int latent, float* y1, float* y2, float* y3, float* feat, float* smax,
                           float* ssum, float* h4, float* emb, void* xs, void* ws, size_t ws_bytes, int prec,
                           void* stream) {
  TACORL_REQUIRE(x_dtype == 0 || x_dtype == 1, "lmp_encoder_fwd: x_dtype must be 0 (fp32) or 1 (uint8)");
  return enc_fwd(x, x_dtype, x_scale, x_shift, N, H, W, params, hidden, latent, y1, y2, y3, feat, smax, ssum, h4, emb,
                 xs, ws, ws_bytes, prec, (cudaStream_t)stream);
}

int tacorl_lmp_encoder_bwd(const void* xv, int x_dtype, float x_scale, float x_shift, int N, int H, int W,
                           const float* const* params, int hidden,
                           int latent, const float* y1, const float* y2, const float* y3,
                           const float* feat, const float* smax, const float* ssum, const float* h4,
                           const float* d_emb, float* const* grads, int accumulate, const void* xs, void* ws,
                           size_t ws_bytes, int prec, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  EncGeom g(N, H, W);
  const float* x = (const float*)xv;
  TACORL_REQUIRE(x_dtype == 0 || x_dtype == 1, "lmp_encoder_bwd: x_dtype must be 0 (fp32) or 1 (uint8)");
  TACORL_REQUIRE(g.ok, "lmp_encoder_bwd: image %dx%d too small", H, W);
  TACORL_REQUIRE(prec == PREC_F32 || prec == PREC_BF16, "lmp_encoder_bwd: unknown precision %d", prec);
  TACORL_REQUIRE(x && params && grads && y1 && y2 && y3 && feat && smax && ssum && h4 && d_emb && ws,
                 "lmp_encoder_bwd: null pointer");
  if (N == 0) return 0;
  if (prec == PREC_BF16)
    return enc_bwd_tc(xv, x_dtype, x_scale, x_shift, N, H, W, params, hidden, latent, (const void*)y1, (const void*)y2,
                      y3, feat, smax, ssum, h4, d_emb, grads, accumulate, xs, ws, ws_bytes, st);
  const float beta0 = accumulate ? 1.f : 0.f;
  Arena ar(ws, ws_bytes);
  if (x_dtype == 1) {   // fp32 parity path: normalised fp32 copy of the uint8 frames (conv1 weight gradient)
    float* xf = ar.take<float>((size_t)N * 3 * H * W);
    TACORL_REQUIRE(xf, "lmp_encoder_bwd: workspace too small for the uint8 -> fp32 image");
    int rcu = u8_to_f32_normalized((long long)N * 3 * H * W, (const unsigned char*)xv, x_scale, x_shift, xf, st);
    if (rcu) return rcu;
    x = xf;
  }
  float* w2p = ar.take<float>(64 * 512);
  float* w3p = ar.take<float>(64 * 576);
  float* dw2p = ar.take<float>(64 * 512);
  float* dw3p = ar.take<float>(64 * 576);
  float* dh4 = ar.take<float>((size_t)N * hidden);
  float* dfeat = ar.take<float>((size_t)N * 128);
  float* dy3 = ar.take<float>((size_t)N * g.P3 * 64);
  float* dtau = ar.take<float>(N);
  float* skws = ar.take<float>(kSplitKWs / 4);
  float* csws = ar.take<float>(592 * 64);
  TACORL_REQUIRE(w2p && w3p && dw2p && dw3p && dh4 && dfeat && dy3 && dtau && skws && csws,
                 "lmp_encoder_bwd: workspace too small");
  size_t per_frame = (size_t)g.col_floats_per_frame() * 4 + (g.P1 * 32 + g.P2 * 64) * 4 + 1024;
  long long chunk = (long long)(ar.left() > 4096 ? (ar.left() - 4096) / per_frame : 0);
  if (chunk > N) chunk = N;
  TACORL_REQUIRE(chunk >= 1, "lmp_encoder_bwd: workspace too small for one frame");
  float* col = ar.take<float>((size_t)chunk * g.col_floats_per_frame());
  float* dy1c = ar.take<float>((size_t)chunk * g.P1 * 32);
  float* dy2c = ar.take<float>((size_t)chunk * g.P2 * 64);
  TACORL_REQUIRE(col && dy1c && dy2c, "lmp_encoder_bwd: workspace carve failed");
  int rc;
  // ---- FC head
  const bool fc_fused = !accumulate && fc_fused_ok(N, hidden, latent);
  if (fc_fused && (rc = fc_head_bwd_fused(N, params, hidden, latent, feat, h4, d_emb, grads, dfeat, skws, kSplitKWs, st)))
    return rc;
  GemmArgs a, b;
  if (!fc_fused) {
  a.transA = 1; a.transB = 0; a.M = latent; a.N = hidden; a.K = N; a.A = d_emb; a.lda = latent; a.B = h4;
  a.ldb = hidden; a.C = grads[P_W5]; a.ldc = hidden; a.beta = beta0; a.split_k = 0;
  if ((rc = gemm_f32(a, skws, kSplitKWs, st))) return rc;                 // dW5 = d_emb^T h4
  if ((rc = colsum_f32(N, latent, d_emb, latent, grads[P_B5], accumulate, st))) return rc;
  b.M = N; b.N = hidden; b.K = latent; b.A = d_emb; b.lda = latent; b.B = params[P_W5]; b.ldb = hidden;
  b.C = dh4; b.ldc = hidden; b.split_k = 1;
  if ((rc = gemm_f32(b, skws, kSplitKWs, st))) return rc;                 // dh4 = d_emb W5
  if ((rc = act_bwd_f32(ACT_RELU, (long long)N * hidden, dh4, h4, dh4, st))) return rc;
  a.M = hidden; a.N = 128; a.A = dh4; a.lda = hidden; a.B = feat; a.ldb = 128; a.C = grads[P_W4]; a.ldc = 128;
  if ((rc = gemm_f32(a, skws, kSplitKWs, st))) return rc;                 // dW4 = dh4^T feat
  if ((rc = colsum_f32(N, hidden, dh4, hidden, grads[P_B4], accumulate, st))) return rc;
  b.N = 128; b.K = hidden; b.A = dh4; b.lda = hidden; b.B = params[P_W4]; b.ldb = 128; b.C = dfeat; b.ldc = 128;
  if ((rc = gemm_f32(b, skws, kSplitKWs, st))) return rc;                 // dfeat = dh4 W4
  }
  // ---- soft-argmax (also applies conv3's ReLU mask) and temperature grad
  if ((rc = softargmax_bwd_f32(y3, N, g.H3, g.W3, 64, params[P_TEMP], feat, smax, ssum, dfeat, dy3, dtau, st)))
    return rc;
  if ((rc = colsum_f32(N, 1, dtau, 1, grads[P_TEMP], accumulate, st))) return rc;
  if ((rc = colsum_tall_f32((long long)N * g.P3, 64, dy3, grads[P_B3], accumulate, csws, 592 * 64 * 4, st)))
    return rc;
  // ---- conv stack, chunked over frames
  if ((rc = permute_conv_weight_f32(params[P_W2], w2p, 64, 32, 4, 4, 0, st))) return rc;
  if ((rc = permute_conv_weight_f32(params[P_W3], w3p, 64, 64, 3, 3, 0, st))) return rc;
  if (accumulate) {
    if ((rc = permute_conv_weight_f32(grads[P_W2], dw2p, 64, 32, 4, 4, 0, st))) return rc;
    if ((rc = permute_conv_weight_f32(grads[P_W3], dw3p, 64, 64, 3, 3, 0, st))) return rc;
  }
  for (long long n0 = 0; n0 < N; n0 += chunk) {
    const int nf = (int)((N - n0) < chunk ? (N - n0) : chunk);
    const float betac = (n0 == 0) ? beta0 : 1.f;
    const int accc = (n0 == 0) ? accumulate : 1;
    const float* y1p = y1 + n0 * g.P1 * 32;
    const float* y2p = y2 + n0 * g.P2 * 64;
    const float* dy3p = dy3 + n0 * g.P3 * 64;
    // conv3 wgrad: dW3p[oc][k] (+)= sum_m dy3[m][oc] col3[m][k]
    if ((rc = im2col_f32(y2p, g.P2 * 64, 1, (long long)g.W2 * 64, 64, 64, 3, 3, 1, g.H3, g.W3, nf, col, st, 1)))
      return rc;
    GemmArgs w;
    w.transA = 1; w.M = 64; w.N = 576; w.K = (int)(nf * g.P3); w.A = dy3p; w.lda = 64; w.B = col; w.ldb = 576;
    w.C = dw3p; w.ldc = 576; w.beta = betac; w.split_k = 0;
    if ((rc = gemm_f32(w, skws, kSplitKWs, st))) return rc;
    // conv3 dgrad: dcol3 = dy3 W3p ; col2im with y2's ReLU mask
    GemmArgs d;
    d.M = (int)(nf * g.P3); d.N = 576; d.K = 64; d.A = dy3p; d.lda = 64; d.B = w3p; d.ldb = 576; d.C = col;
    d.ldc = 576; d.split_k = 1;
    if ((rc = gemm_f32(d, nullptr, 0, st))) return rc;
    if ((rc = col2im_f32(col, 64, g.H2, g.W2, 3, 3, 1, g.H3, g.W3, nf, y2p, dy2c, st))) return rc;
    if ((rc = colsum_tall_f32((long long)nf * g.P2, 64, dy2c, grads[P_B2], accc, csws, 592 * 64 * 4, st))) return rc;
    // conv2
    if ((rc = im2col_f32(y1p, g.P1 * 32, 1, (long long)g.W1 * 32, 32, 32, 4, 4, 2, g.H2, g.W2, nf, col, st, 1)))
      return rc;
    w.N = 512; w.K = (int)(nf * g.P2); w.A = dy2c; w.B = col; w.ldb = 512; w.C = dw2p; w.ldc = 512;
    if ((rc = gemm_f32(w, skws, kSplitKWs, st))) return rc;
    d.M = (int)(nf * g.P2); d.N = 512; d.A = dy2c; d.B = w2p; d.ldb = 512; d.ldc = 512;
    if ((rc = gemm_f32(d, nullptr, 0, st))) return rc;
    if ((rc = col2im_f32(col, 32, g.H1, g.W1, 4, 4, 2, g.H2, g.W2, nf, y1p, dy1c, st))) return rc;
    if ((rc = colsum_tall_f32((long long)nf * g.P1, 32, dy1c, grads[P_B1], accc, csws, 592 * 64 * 4, st))) return rc;
    // conv1 wgrad directly in torch layout (K order c,ky,kx)
    if ((rc = im2col_f32(x + n0 * 3 * H * W, 3LL * H * W, (long long)H * W, W, 1, 3, 8, 8, 4, g.H1, g.W1, nf,
                         col, st, 0))) return rc;
    w.M = 32; w.N = 192; w.K = (int)(nf * g.P1); w.A = dy1c; w.lda = 32; w.B = col; w.ldb = 192;
    w.C = grads[P_W1]; w.ldc = 192;
    if ((rc = gemm_f32(w, skws, kSplitKWs, st))) return rc;
  }
  if ((rc = permute_conv_weight_f32(dw2p, grads[P_W2], 64, 32, 4, 4, 1, st))) return rc;
  return permute_conv_weight_f32(dw3p, grads[P_W3], 64, 64, 3, 3, 1, st);
}

}  // extern "C"
