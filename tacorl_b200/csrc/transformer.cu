// Small fused kernels of the transformer plan recogniser
// (/root/reference/src/tacorl/networks/plan_encoders/plan_recognition_transformer.py:70-105 with torch's
// nn.TransformerEncoderLayer defaults: post-norm, ReLU, eps 1e-5, sequence-first (T,B,D) layout).
// The linear layers run through tacorl_gemm; these kernels cover everything between them:
//   position-embedding add (+pad, +transpose, +dropout), multi-head attention core (head_dim 4..16, T <= 32),
//   residual + dropout + LayerNorm, mean over time.  Dropout masks are inputs (pre-scaled keep masks, or NULL).
#include "common.cuh"
#include "internal.h"
#include "../../include/tacorl_b200.h"

namespace tacorl {

constexpr int kMaxT = 32;
constexpr int kMaxHD = 16;

// x[t][b][d] = (emb[b][t][d] (0 for d >= D0) + pos[t][d]) * mask[t][b][d]
__global__ void posemb_fwd_kernel(int B, int T, int D0, int D, const float* __restrict__ emb,
                                  const float* __restrict__ pos, const float* __restrict__ mask,
                                  float* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * B * D) return;
  const int d = i % D, b = (i / D) % B, t = i / (D * B);
  float v = (d < D0 ? emb[((long long)b * T + t) * D0 + d] : 0.f) + pos[t * D + d];
  x[i] = mask ? v * mask[i] : v;
}
// demb[b][t][d] = dx[t][b][d]*mask ; dpos[t][d] = sum_b dx[t][b][d]*mask   (one thread per (t,d))
__global__ void posemb_bwd_kernel(int B, int T, int D0, int D, const float* __restrict__ dx,
                                  const float* __restrict__ mask, float* __restrict__ demb, float* __restrict__ dpos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * D) return;
  const int d = i % D, t = i / D;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const long long j = ((long long)t * B + b) * D + d;
    const float g = mask ? dx[j] * mask[j] : dx[j];
    s += g;
    if (demb && d < D0) demb[((long long)b * T + t) * D0 + d] = g;
  }
  if (dpos) dpos[i] = s;
}

// Attention core.  qkv: (T,B,3D) = [q | k | v]; one warp per (b, head), lane i = query row i.
// P (B*heads,T,T): softmax probabilities (pre-dropout), saved for backward.  amask: (B*heads,T,T) or NULL.
__global__ void attn_fwd_kernel(int B, int T, int D, int heads, const float* __restrict__ qkv,
                                const float* __restrict__ amask, float* __restrict__ P, float* __restrict__ out) {
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads, hd = D / heads, i = threadIdx.x;
  __shared__ float ks[kMaxT][kMaxHD], vs[kMaxT][kMaxHD];
  if (i < T)
    for (int d = 0; d < hd; ++d) {
      ks[i][d] = qkv[((long long)i * B + b) * 3 * D + D + h * hd + d];
      vs[i][d] = qkv[((long long)i * B + b) * 3 * D + 2 * D + h * hd + d];
    }
  __syncthreads();
  if (i >= T) return;
  const float scale = rsqrtf((float)hd);
  float q[kMaxHD], s[kMaxT], o[kMaxHD];
  for (int d = 0; d < hd; ++d) { q[d] = qkv[((long long)i * B + b) * 3 * D + h * hd + d] * scale; o[d] = 0.f; }
  float mx = -INFINITY;
  for (int j = 0; j < T; ++j) {
    float a = 0.f;
    for (int d = 0; d < hd; ++d) a = fmaf(q[d], ks[j][d], a);
    s[j] = a; mx = fmaxf(mx, a);
  }
  float den = 0.f;
  for (int j = 0; j < T; ++j) { s[j] = expf(s[j] - mx); den += s[j]; }
  const float inv = 1.f / den;
  for (int j = 0; j < T; ++j) {
    const float p = s[j] * inv;
    P[((long long)bh * T + i) * T + j] = p;
    const float pd = amask ? p * amask[((long long)bh * T + i) * T + j] : p;
    for (int d = 0; d < hd; ++d) o[d] = fmaf(pd, vs[j][d], o[d]);
  }
  for (int d = 0; d < hd; ++d) out[((long long)i * B + b) * D + h * hd + d] = o[d];
}

// dqkv from dout: dS = P * (dP - sum_j dP P), dP = (dO V^T) * amask
__global__ void attn_bwd_kernel(int B, int T, int D, int heads, const float* __restrict__ qkv,
                                const float* __restrict__ amask, const float* __restrict__ P,
                                const float* __restrict__ dout, float* __restrict__ dqkv) {
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads, hd = D / heads, i = threadIdx.x;
  __shared__ float qs[kMaxT][kMaxHD], ks[kMaxT][kMaxHD], vs[kMaxT][kMaxHD], dos[kMaxT][kMaxHD];
  __shared__ float dS[kMaxT][kMaxT + 1], Pd[kMaxT][kMaxT + 1];
  const float scale = rsqrtf((float)hd);
  if (i < T)
    for (int d = 0; d < hd; ++d) {
      const long long base = ((long long)i * B + b) * 3 * D + h * hd + d;
      qs[i][d] = qkv[base]; ks[i][d] = qkv[base + D]; vs[i][d] = qkv[base + 2 * D];
      dos[i][d] = dout[((long long)i * B + b) * D + h * hd + d];
    }
  __syncthreads();
  if (i < T) {
    float dot = 0.f, dp[kMaxT];
    for (int j = 0; j < T; ++j) {
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a = fmaf(dos[i][d], vs[j][d], a);
      const long long pj = ((long long)bh * T + i) * T + j;
      const float m = amask ? amask[pj] : 1.f;
      const float p = P[pj];
      Pd[i][j] = p * m;
      dp[j] = a * m;
      dot = fmaf(dp[j], p, dot);
    }
    float dq[kMaxHD];
    for (int d = 0; d < hd; ++d) dq[d] = 0.f;
    for (int j = 0; j < T; ++j) {
      const float ds = P[((long long)bh * T + i) * T + j] * (dp[j] - dot);
      dS[i][j] = ds;
      for (int d = 0; d < hd; ++d) dq[d] = fmaf(ds, ks[j][d], dq[d]);
    }
    for (int d = 0; d < hd; ++d) dqkv[((long long)i * B + b) * 3 * D + h * hd + d] = dq[d] * scale;
  }
  __syncthreads();
  if (i < T) {   // now lane i = key/value row
    float dk[kMaxHD], dv[kMaxHD];
    for (int d = 0; d < hd; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int r = 0; r < T; ++r) {
      const float ds = dS[r][i], pd = Pd[r][i];
      for (int d = 0; d < hd; ++d) { dk[d] = fmaf(ds, qs[r][d], dk[d]); dv[d] = fmaf(pd, dos[r][d], dv[d]); }
    }
    for (int d = 0; d < hd; ++d) {
      const long long base = ((long long)i * B + b) * 3 * D + h * hd + d;
      dqkv[base + D] = dk[d] * scale;
      dqkv[base + 2 * D] = dv[d];
    }
  }
}

// y = LayerNorm(x + r*mask) * w + b over the last dim D (<= 128); one warp per row.  Saves xhat and rstd.
__global__ void add_ln_fwd_kernel(int rows, int D, const float* __restrict__ x, const float* __restrict__ r,
                                  const float* __restrict__ mask, const float* __restrict__ w,
                                  const float* __restrict__ bvec, float eps, float* __restrict__ y,
                                  float* __restrict__ xhat, float* __restrict__ rstd) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[4];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = lane + 32 * k;
    v[k] = 0.f;
    if (d < D) {
      const long long j = (long long)row * D + d;
      v[k] = x[j] + (r ? (mask ? r[j] * mask[j] : r[j]) : 0.f);
      sum += v[k];
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) if (lane + 32 * k < D) { const float c = v[k] - mean; var = fmaf(c, c, var); }
  const float rs = rsqrtf(warp_sum(var) / (float)D + eps);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = lane + 32 * k;
    if (d < D) {
      const long long j = (long long)row * D + d;
      const float xh = (v[k] - mean) * rs;
      xhat[j] = xh;
      y[j] = xh * w[d] + bvec[d];
    }
  }
  if (lane == 0) rstd[row] = rs;
}

// dx = rstd * (g - mean(g) - xhat*mean(g*xhat)), g = dy*w;  dr = dx*mask;  dw/db accumulated with atomics
__global__ void add_ln_bwd_kernel(int rows, int D, const float* __restrict__ dy, const float* __restrict__ xhat,
                                  const float* __restrict__ rstd, const float* __restrict__ w,
                                  const float* __restrict__ mask, float* __restrict__ dx, float* __restrict__ dr,
                                  float* __restrict__ dw, float* __restrict__ db) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float g[4], xh[4];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = lane + 32 * k;
    g[k] = 0.f; xh[k] = 0.f;
    if (d < D) {
      const long long j = (long long)row * D + d;
      xh[k] = xhat[j];
      const float dyv = dy[j];
      g[k] = dyv * w[d];
      s1 += g[k]; s2 = fmaf(g[k], xh[k], s2);
      atomicAdd(dw + d, dyv * xh[k]);
      atomicAdd(db + d, dyv);
    }
  }
  const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
  const float rs = rstd[row];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = lane + 32 * k;
    if (d < D) {
      const long long j = (long long)row * D + d;
      const float v = rs * (g[k] - m1 - xh[k] * m2);
      dx[j] = v;
      if (dr) dr[j] = mask ? v * mask[j] : v;
    }
  }
}

// y[b][c] = mean_t x[b][t][c]   /   dx[b][t][c] = dy[b][c] / T
__global__ void mean_t_fwd_kernel(int B, int T, int C, const float* __restrict__ x, float* __restrict__ y) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * C) return;
  const int c = (int)(i % C); const long long b = i / C;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += x[(b * T + t) * C + c];
  y[i] = s / (float)T;
}
__global__ void mean_t_bwd_kernel(int B, int T, int C, const float* __restrict__ dy, float* __restrict__ dx) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * C) return;
  const int c = (int)(i % C); const long long b = i / ((long long)T * C);
  dx[i] = dy[b * C + c] / (float)T;
}

__global__ void mul_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = a[i] * b[i];
}

}  // namespace tacorl

using namespace tacorl;

extern "C" {

int tacorl_posemb_fwd(int B, int T, int D0, int D, const float* emb, const float* pos, const float* mask, float* x,
                      void* stream) {
  TACORL_REQUIRE(emb && pos && x && D >= D0, "posemb_fwd: bad arguments");
  posemb_fwd_kernel<<<cdiv((long long)T * B * D, 256), 256, 0, (cudaStream_t)stream>>>(B, T, D0, D, emb, pos, mask, x);
  TACORL_LAUNCH_CHECK();
  return 0;
}
int tacorl_posemb_bwd(int B, int T, int D0, int D, const float* dx, const float* mask, float* demb, float* dpos,
                      void* stream) {
  TACORL_REQUIRE(dx, "posemb_bwd: bad arguments");
  posemb_bwd_kernel<<<cdiv(T * D, 128), 128, 0, (cudaStream_t)stream>>>(B, T, D0, D, dx, mask, demb, dpos);
  TACORL_LAUNCH_CHECK();
  return 0;
}
int tacorl_attn_fwd(int B, int T, int D, int heads, const float* qkv, const float* amask, float* P, float* out,
                    void* stream) {
  TACORL_REQUIRE(T <= kMaxT && D % heads == 0 && D / heads <= kMaxHD, "attn_fwd: needs T <= 32 and head_dim <= 16");
  attn_fwd_kernel<<<B * heads, 32, 0, (cudaStream_t)stream>>>(B, T, D, heads, qkv, amask, P, out);
  TACORL_LAUNCH_CHECK();
  return 0;
}
int tacorl_attn_bwd(int B, int T, int D, int heads, const float* qkv, const float* amask, const float* P,
                    const float* dout, float* dqkv, void* stream) {
  TACORL_REQUIRE(T <= kMaxT && D % heads == 0 && D / heads <= kMaxHD, "attn_bwd: needs T <= 32 and head_dim <= 16");
  attn_bwd_kernel<<<B * heads, 32, 0, (cudaStream_t)stream>>>(B, T, D, heads, qkv, amask, P, dout, dqkv);
  TACORL_LAUNCH_CHECK();
  return 0;
}
int tacorl_add_ln_fwd(int rows, int D, const float* x, const float* r, const float* mask, const float* w,
                      const float* b, float eps, float* y, float* xhat, float* rstd, void* stream) {
  TACORL_REQUIRE(D <= 128, "add_ln_fwd: D must be <= 128");
  add_ln_fwd_kernel<<<cdiv((long long)rows * 32, 128), 128, 0, (cudaStream_t)stream>>>(rows, D, x, r, mask, w, b, eps, y,
                                                                                     xhat, rstd);
  TACORL_LAUNCH_CHECK();
  return 0;
}
/* dw, db (D floats each) are ACCUMULATED into: zero them first */
int tacorl_add_ln_bwd(int rows, int D, const float* dy, const float* xhat, const float* rstd, const float* w,
                      const float* mask, float* dx, float* dr, float* dw, float* db, void* stream) {
  TACORL_REQUIRE(D <= 128, "add_ln_bwd: D must be <= 128");
  add_ln_bwd_kernel<<<cdiv((long long)rows * 32, 128), 128, 0, (cudaStream_t)stream>>>(rows, D, dy, xhat, rstd, w, mask,
                                                                                     dx, dr, dw, db);
  TACORL_LAUNCH_CHECK();
  return 0;
}
int tacorl_mean_t_fwd(int B, int T, int C, const float* x, float* y, void* stream) {
  mean_t_fwd_kernel<<<cdiv((long long)B * C, 256), 256, 0, (cudaStream_t)stream>>>(B, T, C, x, y);
  TACORL_LAUNCH_CHECK();
  return 0;
}
int tacorl_mean_t_bwd(int B, int T, int C, const float* dy, float* dx, void* stream) {
  mean_t_bwd_kernel<<<cdiv((long long)B * T * C, 256), 256, 0, (cudaStream_t)stream>>>(B, T, C, dy, dx);
  TACORL_LAUNCH_CHECK();
  return 0;
}
int tacorl_mul(long long n, const float* a, const float* b, float* out, void* stream) {
  if (n == 0) return 0;
  mul_kernel<<<(int)min((long long)1184, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, a, b, out);
  TACORL_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
