// ReLU Elman RNN layer (one direction), forward and BPTT backward, time-major buffers.
// C-ABI: tacorl_rnn_layer_{fwd,bwd}.  Replaces nn.RNN(nonlinearity="relu") as used by the
// plan-recognition BiRNN (/root/reference/src/tacorl/networks/plan_encoders/
// plan_recognition_tanh_net.py:23-31, plan_recognition_net.py:27-35) and the action decoder
// (networks/action_decoders/rnn_models.py:5-16):
//     h_t = relu(W_ih x_t + b_ih + W_hh h_{t-1} + b_hh),  h_init = h0 or 0.
// `n_steps` < T runs only the first n_steps of the recurrence (reverse: starting at t = T-1);
// the BiRNN's last-layer reverse direction only needs one step because only out[:, -1] is
// consumed (plan_recognition_tanh_net.py:43; SURVEY.md §0 finding 7-ii).
// fp32 parity path: one input GEMM for all steps + one recurrent GEMM per step.
#include "common.cuh"
#include "internal.h"
#include "../../include/tacorl_b200.h"

namespace tacorl {

__global__ void vec_add_kernel(int n, const float* a, const float* b, float* o) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

// rows x H block with row stride ld: p = relu(p)
__global__ void relu_rows_kernel(long long rows, int H, float* p, long long ld) {
  const long long total = rows * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float* q = p + (i / H) * ld + (i % H);
    *q = fmaxf(*q, 0.f);
  }
}

// d = (d + add?) * [out > 0]
__global__ void mask_rows_kernel(long long rows, int H, float* d, long long ldd, const float* out,
                                 long long ldo, const float* add, long long lda) {
  const long long total = rows * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H; const int c = (int)(i % H);
    float v = d[r * ldd + c];
    if (add) v += add[r * lda + c];
    d[r * ldd + c] = out[r * ldo + c] > 0.f ? v : 0.f;
  }
}

static inline int ew_blocks(long long n) { return (int)min((long long)1184, (n + 255) / 256); }

}  // namespace tacorl

using namespace tacorl;

extern "C" {

size_t tacorl_rnn_layer_ws_bytes(int T, int B, int I, int H) {
  (void)T; (void)I;
  // bias sum + split-K partials for the recurrent GEMM (<= 16 splits of BxH) and the weight grads
  size_t sk = (size_t)16 * B * H * 4;
  size_t wg = (size_t)4 * H * (size_t)(H > I ? H : I) * 4;
  return (size_t)H * 4 + 4096 + (sk > wg ? sk : wg) + (1 << 16);
}

int tacorl_rnn_layer_fwd(int T, int B, int I, int H, const float* x, long long ldx, const float* w_ih,
                         const float* w_hh, const float* b_ih, const float* b_hh, const float* h0,
                         int reverse, int n_steps, float* out, long long ldo, void* ws, size_t ws_bytes,
                         int prec, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TACORL_REQUIRE(prec == PREC_F32, "rnn_layer_fwd: precision %d not built into this entry point", prec);
  TACORL_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && out && ws, "rnn_layer_fwd: null pointer");
  TACORL_REQUIRE(n_steps >= 1 && n_steps <= T, "rnn_layer_fwd: n_steps %d out of range (T=%d)", n_steps, T);
  if (B == 0) return 0;
  Arena ar(ws, ws_bytes);
  float* bsum = ar.take<float>(H);
  TACORL_REQUIRE(bsum, "rnn_layer_fwd: workspace too small");
  float* sk = (float*)(ar.base + ar.off);
  size_t sk_bytes = ar.left();
  vec_add_kernel<<<cdiv(H, 256), 256, 0, st>>>(H, b_ih, b_hh, bsum);
  TACORL_LAUNCH_CHECK();
  const int t_lo = reverse ? T - n_steps : 0;
  int rc;
  GemmArgs in;   // pre-activations of all steps: out[t] = x[t] W_ih^T + (b_ih + b_hh)
  in.transB = 1; in.M = n_steps * B; in.N = H; in.K = I;
  in.A = x + (long long)t_lo * B * ldx; in.lda = ldx; in.B = w_ih; in.ldb = I;
  in.C = out + (long long)t_lo * B * ldo; in.ldc = ldo; in.bias = bsum; in.split_k = 0;
  if ((rc = gemm_f32(in, sk, sk_bytes, st))) return rc;
  for (int s = 0; s < n_steps; ++s) {
    const int t = reverse ? T - 1 - s : s;
    float* ot = out + (long long)t * B * ldo;
    const float* hp; long long ldh;
    if (s == 0) { hp = h0; ldh = H; }
    else { hp = out + (long long)(reverse ? t + 1 : t - 1) * B * ldo; ldh = ldo; }
    if (!hp) {
      relu_rows_kernel<<<ew_blocks((long long)B * H), 256, 0, st>>>(B, H, ot, ldo);
      TACORL_LAUNCH_CHECK();
      continue;
    }
    GemmArgs r;
    r.transB = 1; r.M = B; r.N = H; r.K = H; r.A = hp; r.lda = ldh; r.B = w_hh; r.ldb = H;
    r.C = ot; r.ldc = ldo; r.beta = 1.f; r.act = ACT_RELU; r.split_k = 0;
    if ((rc = gemm_f32(r, sk, sk_bytes, st))) return rc;
  }
  return 0;
}

int tacorl_rnn_layer_bwd(int T, int B, int I, int H, const float* x, long long ldx, const float* w_ih,
                         const float* w_hh, const float* h0, int reverse, int n_steps, const float* out,
                         long long ldo, float* dout, long long lddo, const float* dhn, float* dx,
                         long long lddx, int dx_accumulate, float* dw_ih, float* dw_hh, float* db_ih,
                         float* db_hh, int accumulate, float* dh0, void* ws, size_t ws_bytes, int prec,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  TACORL_REQUIRE(prec == PREC_F32, "rnn_layer_bwd: precision %d not built into this entry point", prec);
  TACORL_REQUIRE(x && w_ih && w_hh && out && dout && ws, "rnn_layer_bwd: null pointer");
  TACORL_REQUIRE(n_steps >= 1 && n_steps <= T, "rnn_layer_bwd: n_steps out of range");
  if (B == 0) return 0;
  float* sk = (float*)ws;
  size_t sk_bytes = ws_bytes;
  const float beta0 = accumulate ? 1.f : 0.f;
  const int t_lo = reverse ? T - n_steps : 0;
  int rc;
  // BPTT: walk the recurrence backwards; dout[t] becomes dpre[t] in place.
  for (int s = n_steps - 1; s >= 0; --s) {
    const int t = reverse ? T - 1 - s : s;
    float* dt = dout + (long long)t * B * lddo;
    const float* ot = out + (long long)t * B * ldo;
    mask_rows_kernel<<<ew_blocks((long long)B * H), 256, 0, st>>>(
        B, H, dt, lddo, ot, ldo, (s == n_steps - 1) ? dhn : nullptr, H);
    TACORL_LAUNCH_CHECK();
    if (s > 0) {  // dout[t_prev] += dpre[t] W_hh
      const int tp = reverse ? t + 1 : t - 1;
      GemmArgs c;
      c.M = B; c.N = H; c.K = H; c.A = dt; c.lda = lddo; c.B = w_hh; c.ldb = H;
      c.C = dout + (long long)tp * B * lddo; c.ldc = lddo; c.beta = 1.f; c.split_k = 0;
      if ((rc = gemm_f32(c, sk, sk_bytes, st))) return rc;
    } else if (dh0) {
      GemmArgs c;
      c.M = B; c.N = H; c.K = H; c.A = dt; c.lda = lddo; c.B = w_hh; c.ldb = H; c.C = dh0; c.ldc = H;
      c.split_k = 0;
      if ((rc = gemm_f32(c, sk, sk_bytes, st))) return rc;
    }
  }
  float* dpre = dout + (long long)t_lo * B * lddo;
  const long long rows = (long long)n_steps * B;
  // dW_hh (+)= sum_{s>=1} dpre[t(s)]^T h[t(s-1)]  (+ dpre[t(0)]^T h0)
  if (dw_hh) {
    bool wrote = false;
    if (n_steps > 1) {
      GemmArgs w;
      w.transA = 1; w.M = H; w.N = H; w.K = (n_steps - 1) * B;
      if (!reverse) { w.A = dpre + (long long)B * lddo; w.B = out; }
      else { w.A = dpre; w.B = out + (long long)(t_lo + 1) * B * ldo; }
      w.lda = lddo; w.ldb = ldo; w.C = dw_hh; w.ldc = H; w.beta = beta0; w.split_k = 0;
      if ((rc = gemm_f32(w, sk, sk_bytes, st))) return rc;
      wrote = true;
    }
    if (h0) {
      const int t0 = reverse ? T - 1 : 0;
      GemmArgs w;
      w.transA = 1; w.M = H; w.N = H; w.K = B; w.A = dout + (long long)t0 * B * lddo; w.lda = lddo;
      w.B = h0; w.ldb = H; w.C = dw_hh; w.ldc = H; w.beta = wrote ? 1.f : beta0; w.split_k = 0;
      if ((rc = gemm_f32(w, sk, sk_bytes, st))) return rc;
      wrote = true;
    }
    if (!wrote && !accumulate) TACORL_CHECK_CUDA(cudaMemsetAsync(dw_hh, 0, (size_t)H * H * 4, st));
  }
  if (dw_ih) {
    GemmArgs w;
    w.transA = 1; w.M = H; w.N = I; w.K = (int)rows; w.A = dpre; w.lda = lddo;
    w.B = x + (long long)t_lo * B * ldx; w.ldb = ldx; w.C = dw_ih; w.ldc = I; w.beta = beta0; w.split_k = 0;
    if ((rc = gemm_f32(w, sk, sk_bytes, st))) return rc;
  }
  if (db_ih) if ((rc = colsum_f32((int)rows, H, dpre, lddo, db_ih, accumulate, st))) return rc;
  if (db_hh) if ((rc = colsum_f32((int)rows, H, dpre, lddo, db_hh, accumulate, st))) return rc;
  if (dx) {
    if (!dx_accumulate && n_steps < T) {
      // rows outside the active range receive no gradient
      const long long lo_rows = (long long)t_lo * B, hi_rows = (long long)(T - t_lo - n_steps) * B;
      if (lo_rows) TACORL_CHECK_CUDA(cudaMemset2DAsync(dx, lddx * 4, 0, (size_t)I * 4, lo_rows, st));
      if (hi_rows) TACORL_CHECK_CUDA(cudaMemset2DAsync(dx + (long long)(t_lo + n_steps) * B * lddx, lddx * 4, 0,
                                                       (size_t)I * 4, hi_rows, st));
    }
    GemmArgs d;
    d.M = (int)rows; d.N = I; d.K = H; d.A = dpre; d.lda = lddo; d.B = w_ih; d.ldb = I;
    d.C = dx + (long long)t_lo * B * lddx; d.ldc = lddx; d.beta = dx_accumulate ? 1.f : 0.f; d.split_k = 0;
    if ((rc = gemm_f32(d, sk, sk_bytes, st))) return rc;
  }
  return 0;
}

}  // extern "C"
