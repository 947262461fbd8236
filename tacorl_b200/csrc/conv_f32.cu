// fp32 building blocks of the LMP vision encoder (parity path): im2col / col2im around the
// fp32 GEMM, weight-layout permutes, and the fused ReLU-aware spatial soft-argmax fwd/bwd.
// Activations are kept NHWC internally: (frame, oy, ox, channel), channel fastest.
// Reference: networks/visual_encoders/encoder.py:369-419, utils.py:39-76.
#include "common.cuh"
#include "internal.h"
#include <cuda_bf16.h>

namespace tacorl {

// col[m][k], m = (n, oy, ox) over `nframes` frames; k = (ky, kx, c) (c fastest) when korder == 1,
// k = (c, ky, kx) (the torch weight layout) when korder == 0.
// Input element (n, c, iy, ix) lives at x[n*sn + c*sc + iy*sh + ix*sw] (NCHW or NHWC by strides).
__device__ __forceinline__ void store_as(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_as(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ float load_as(const float* p) { return *p; }
__device__ __forceinline__ float load_as(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename OutT>
__global__ void im2col_kernel(const float* __restrict__ x, long long sn, long long sc, long long sh,
                              long long sw, int C, int KH, int KW, int stride, int OH, int OW,
                              int nframes, OutT* __restrict__ col, int korder) {
  const int K = KH * KW * C;
  const long long total = (long long)nframes * OH * OW * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const long long m = idx / K;
    int c, kx, ky;
    if (korder) { c = k % C; kx = (k / C) % KW; ky = k / (C * KW); }
    else { kx = k % KW; ky = (k / KW) % KH; c = k / (KW * KH); }
    const int ox = (int)(m % OW), oy = (int)((m / OW) % OH);
    const long long n = m / ((long long)OW * OH);
    store_as(col + idx, __ldg(x + n * sn + c * sc + (long long)(oy * stride + ky) * sh + (long long)(ox * stride + kx) * sw));
  }
}

int im2col_f32(const float* x, long long sn, long long sc, long long sh, long long sw, int C, int KH,
               int KW, int stride, int OH, int OW, int nframes, float* col, cudaStream_t st, int korder) {
  long long total = (long long)nframes * OH * OW * KH * KW * C;
  if (total == 0) return 0;
  int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
  im2col_kernel<float><<<blocks, 256, 0, st>>>(x, sn, sc, sh, sw, C, KH, KW, stride, OH, OW, nframes, col, korder);
  TACORL_LAUNCH_CHECK();
  return 0;
}

// dX (NHWC) = mask(Y>0) * sum over the kernel taps that touch (iy, ix) of dcol.
// dcol[m][k] with the same (ky,kx,c) ordering as im2col.  Gather form => deterministic.
template <typename InT>
__global__ void col2im_kernel(const InT* __restrict__ dcol, int C, int H, int W, int KH, int KW,
                              int stride, int OH, int OW, int nframes, const float* __restrict__ ymask,
                              float* __restrict__ dx, __nv_bfloat16* __restrict__ dxb) {
  const int K = KH * KW * C;
  const long long total = (long long)nframes * H * W * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int ix = (int)((idx / C) % W), iy = (int)((idx / ((long long)C * W)) % H);
    const long long n = idx / ((long long)C * W * H);
    float s = 0.f;
    if (!ymask || ymask[idx] > 0.f) {
      for (int ky = 0; ky < KH; ++ky) {
        const int ty = iy - ky;
        if (ty < 0 || ty % stride) continue;
        const int oy = ty / stride;
        if (oy >= OH) continue;
        for (int kx = 0; kx < KW; ++kx) {
          const int tx = ix - kx;
          if (tx < 0 || tx % stride) continue;
          const int ox = tx / stride;
          if (ox >= OW) continue;
          s += load_as(dcol + ((n * OH + oy) * OW + ox) * K + (ky * KW + kx) * C + c);
        }
      }
    }
    dx[idx] = s;
    if (dxb) dxb[idx] = __float2bfloat16(s);
  }
}

int col2im_f32(const float* dcol, int C, int H, int W, int KH, int KW, int stride, int OH, int OW,
               int nframes, const float* ymask, float* dx, cudaStream_t st) {
  long long total = (long long)nframes * H * W * C;
  if (total == 0) return 0;
  int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
  col2im_kernel<float><<<blocks, 256, 0, st>>>(dcol, C, H, W, KH, KW, stride, OH, OW, nframes, ymask, dx, nullptr);
  TACORL_LAUNCH_CHECK();
  return 0;
}

// (oc, c, ky, kx) <-> (oc, ky, kx, c).  dir 0: torch -> khwc; dir 1: khwc -> torch.
__global__ void permute_w_kernel(const float* __restrict__ src, float* __restrict__ dst, int OC, int C,
                                 int KH, int KW, int dir) {
  const int total = OC * C * KH * KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // i indexes the torch layout
    const int kx = i % KW, ky = (i / KW) % KH, c = (i / (KW * KH)) % C, oc = i / (KW * KH * C);
    const int j = ((oc * KH + ky) * KW + kx) * C + c;
    if (dir == 0) dst[j] = src[i]; else dst[i] = src[j];
  }
}

int permute_conv_weight_f32(const float* src, float* dst, int OC, int C, int KH, int KW, int dir,
                            cudaStream_t st) {
  int total = OC * C * KH * KW;
  permute_w_kernel<<<cdiv(total, 256), 256, 0, st>>>(src, dst, OC, C, KH, KW, dir);
  TACORL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Spatial soft-argmax over an NHWC feature map y (already ReLU'd): per (frame, channel)
//   p = softmax_i(y_i / tau),  f[2c] = sum p_i * col_i,  f[2c+1] = sum p_i * row_i.
// One CTA per frame, blockDim = (C, G): thread (c, g) scans positions g, g+G, ... with an online
// softmax, then the G partials are merged through shared memory.  Saves (max, sumexp) per (n,c).
template <int G>
__global__ void softargmax_fwd_kernel(const float* __restrict__ y, int P, int OW, int C,
                                      const float* __restrict__ temperature, float* __restrict__ feat,
                                      float* __restrict__ smax, float* __restrict__ ssum) {
  extern __shared__ float sm[];  // 4 * G * C partials, then the (col, row) coordinate of every position
  const int c = threadIdx.x, g = threadIdx.y;
  const long long n = blockIdx.x;
  const float inv_t = 1.f / __ldg(temperature);
  const float* yp = y + n * (long long)P * C;
  float2* xy = reinterpret_cast<float2*>(sm + 4 * G * C);   // replaces an integer divide per element
  for (int p = g * C + c; p < P; p += G * C) xy[p] = make_float2((float)(p % OW), (float)(p / OW));
  __syncthreads();
  float m = -INFINITY, s = 0.f, sx = 0.f, sy = 0.f;
  // batches of 8 positions: all loads in flight first, one max, then the exponentials (merged into the running
  // online-softmax state), instead of a load -> compare -> exp chain per position
  constexpr int UB = 8;
  for (int p0 = g; p0 < P; p0 += G * UB) {
    float v[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      v[j] = p < P ? yp[(long long)p * C + c] * inv_t : -INFINITY;
    }
    float bm = v[0];
#pragma unroll
    for (int j = 1; j < UB; ++j) bm = fmaxf(bm, v[j]);
    if (bm > m) {
      const float sc = __expf(m - bm);     // exp(-inf) = 0 on the first batch
      s *= sc; sx *= sc; sy *= sc; m = bm;
    }
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      const float e = __expf(v[j] - m);    // 0 for the padded tail
      const float2 q = xy[p < P ? p : 0];
      s += e; sx += e * q.x; sy += e * q.y;
    }
  }
  float* q = sm + (g * C + c) * 4;
  q[0] = m; q[1] = s; q[2] = sx; q[3] = sy;
  __syncthreads();
  if (g == 0) {
    float M = -INFINITY;
    for (int i = 0; i < G; ++i) M = fmaxf(M, sm[(i * C + c) * 4]);
    float S = 0.f, SX = 0.f, SY = 0.f;
    for (int i = 0; i < G; ++i) {
      const float* r = sm + (i * C + c) * 4;
      if (r[1] > 0.f) {
        const float sc = expf(r[0] - M);
        S += r[1] * sc; SX += r[2] * sc; SY += r[3] * sc;
      }
    }
    feat[n * 2 * C + 2 * c] = SX / S;
    feat[n * 2 * C + 2 * c + 1] = SY / S;
    smax[n * C + c] = M;
    ssum[n * C + c] = S;
  }
}

// ---- vectorised variants (C % 4 == 0): thread (cq, g) owns 4 consecutive channels (one 16-byte load per position)
// and scans positions g, g+G, ...; 256-thread CTAs (C = 64, G = 16) so that several CTAs share an SM and every thread
// keeps 8 x 16 bytes in flight: the scalar kernels above run 1024-thread CTAs at one CTA per SM (register limit) and
// reach 1.7 TB/s (ncu, round 1); these reach 3.2 TB/s.  (Tried and measured slower, round 2: 512-thread CTAs that put the
// whole 113 KB frame in flight in one batch per thread, 54 / 58 us against 36 / 44 us: with one CTA per SM the load phase
// of one frame no longer overlaps the exponentials of another.)
template <int G>
__global__ void __launch_bounds__(256)
softargmax_fwd_v4_kernel(const float* __restrict__ y, int P, int OW, int C, const float* __restrict__ temperature,
                         float* __restrict__ feat, float* __restrict__ smax, float* __restrict__ ssum) {
  extern __shared__ float sm[];  // 16 * G * C/4 partials, then the (col, row) coordinate of every position
  const int C4 = C >> 2;
  const int cq = threadIdx.x, g = threadIdx.y;
  const long long n = blockIdx.x;
  const float inv_t = 1.f / __ldg(temperature);
  const float4* yp = reinterpret_cast<const float4*>(y + n * (long long)P * C);
  float2* xy = reinterpret_cast<float2*>(sm + 16 * G * C4);
  for (int p = g * C4 + cq; p < P; p += G * C4) xy[p] = make_float2((float)(p % OW), (float)(p / OW));
  __syncthreads();
  float m[4], s[4], sx[4], sy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { m[k] = -INFINITY; s[k] = 0.f; sx[k] = 0.f; sy[k] = 0.f; }
  constexpr int UB = 8;
  for (int p0 = g; p0 < P; p0 += G * UB) {
    float4 v[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      v[j] = p < P ? yp[(long long)p * C4 + cq] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      v[j].x *= inv_t; v[j].y *= inv_t; v[j].z *= inv_t; v[j].w *= inv_t;
      bm[0] = fmaxf(bm[0], v[j].x); bm[1] = fmaxf(bm[1], v[j].y); bm[2] = fmaxf(bm[2], v[j].z); bm[3] = fmaxf(bm[3], v[j].w);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (bm[k] > m[k]) {
        const float sc = __expf(m[k] - bm[k]);     // exp(-inf) = 0 on the first batch
        s[k] *= sc; sx[k] *= sc; sy[k] *= sc; m[k] = bm[k];
      }
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      const float2 q = xy[p < P ? p : 0];
      const float e0 = __expf(v[j].x - m[0]), e1 = __expf(v[j].y - m[1]), e2 = __expf(v[j].z - m[2]), e3 = __expf(v[j].w - m[3]);
      s[0] += e0; sx[0] += e0 * q.x; sy[0] += e0 * q.y;
      s[1] += e1; sx[1] += e1 * q.x; sy[1] += e1 * q.y;
      s[2] += e2; sx[2] += e2 * q.x; sy[2] += e2 * q.y;
      s[3] += e3; sx[3] += e3 * q.x; sy[3] += e3 * q.y;
    }
  }
  // partials [g][c][4]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float* q = sm + ((g * C + 4 * cq + k) << 2);
    q[0] = m[k]; q[1] = s[k]; q[2] = sx[k]; q[3] = sy[k];
  }
  __syncthreads();
  const int tid = g * C4 + cq;
  if (tid < C) {
    const int c = tid;
    float M = -INFINITY;
    for (int i = 0; i < G; ++i) M = fmaxf(M, sm[(i * C + c) * 4]);
    float S = 0.f, SX = 0.f, SY = 0.f;
    for (int i = 0; i < G; ++i) {
      const float* r = sm + (i * C + c) * 4;
      if (r[1] > 0.f) {
        const float sc = expf(r[0] - M);
        S += r[1] * sc; SX += r[2] * sc; SY += r[3] * sc;
      }
    }
    feat[n * 2 * C + 2 * c] = SX / S;
    feat[n * 2 * C + 2 * c + 1] = SY / S;
    smax[n * C + c] = M;
    ssum[n * C + c] = S;
  }
}

__device__ __forceinline__ void sa_store4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void sa_store4(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

template <int G, typename OutT>
__global__ void __launch_bounds__(256)
softargmax_bwd_v4_kernel(const float* __restrict__ y, int P, int OW, int C, const float* __restrict__ temperature,
                         const float* __restrict__ feat, const float* __restrict__ smax, const float* __restrict__ ssum,
                         const float* __restrict__ dfeat, OutT* __restrict__ dy, float* __restrict__ dtau_part,
                         int out_w, int out_p) {
  // out_w / out_p: row pitch (pixels) and pixels per frame of dy; larger than OW / P = dy is written into a frame with a
  // zero right / bottom margin (the pitch of conv3's input: linear-shift weight gradient, conv_tc.cu)
  __shared__ float red[32];
  extern __shared__ float sm[];   // (col, row) coordinate of every position
  const int C4 = C >> 2;
  const int cq = threadIdx.x, g = threadIdx.y;
  const long long n = blockIdx.x;
  const float tau = __ldg(temperature), inv_t = 1.f / tau;
  const float4* yp = reinterpret_cast<const float4*>(y + n * (long long)P * C);
  OutT* dyp = dy + n * (long long)out_p * C;
  float2* xy = reinterpret_cast<float2*>(sm);
  for (int p = g * C4 + cq; p < P; p += G * C4) xy[p] = make_float2((float)(p % OW), (float)(p / OW));
  __syncthreads();
  if (out_p != P) {                                      // zero margin: right of every row, then the rows below
    const int OH = P / OW, padw = out_w - OW, right = OH * padw, all = right + (out_p - OH * out_w);
    for (int m = g; m < all; m += G) {
      const int pp = m < right ? (m / padw) * out_w + OW + m % padw : OH * out_w + (m - right);
      sa_store4(dyp + (long long)pp * C + 4 * cq, 0.f, 0.f, 0.f, 0.f);
    }
  }
  float gx[4], gy[4], M[4], invS[4], dotg[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = 4 * cq + k;
    gx[k] = dfeat[n * 2 * C + 2 * c]; gy[k] = dfeat[n * 2 * C + 2 * c + 1];
    const float fx = feat[n * 2 * C + 2 * c], fy = feat[n * 2 * C + 2 * c + 1];
    M[k] = smax[n * C + c]; invS[k] = 1.f / ssum[n * C + c];
    dotg[k] = gx[k] * fx + gy[k] * fy;
  }
  float dt = 0.f;
  constexpr int UB = 8;
  for (int p0 = g; p0 < P; p0 += G * UB) {
    float4 yv[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      yv[j] = p < P ? yp[(long long)p * C4 + cq] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      if (p >= P) break;
      const float2 q = xy[p];
      const float in[4] = {yv[j].x, yv[j].y, yv[j].z, yv[j].w};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float pr = __expf(in[k] * inv_t - M[k]) * invS[k];
        const float dz = pr * (gx[k] * q.x + gy[k] * q.y - dotg[k]);
        dt += dz * in[k];
        o[k] = in[k] > 0.f ? dz * inv_t : 0.f;
      }
      const int po = p + (int)q.y * (out_w - OW);
      sa_store4(dyp + (long long)po * C + 4 * cq, o[0], o[1], o[2], o[3]);
    }
  }
  dt = warp_sum(dt);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nwarps = (blockDim.x * blockDim.y + 31) / 32;
  if ((tid & 31) == 0) red[tid >> 5] = dt;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < nwarps; ++i) t += red[i];
    dtau_part[n] = -t * inv_t * inv_t;
  }
}

static bool sa_v4_ok(int C, const void* y, const void* dy, int dy_elt) {
  return C % 4 == 0 && (C / 4) * 16 <= 256 && ((C / 4) * 16) % 32 == 0 && ((uintptr_t)y & 15) == 0 &&
         (dy == nullptr || ((uintptr_t)dy & (dy_elt * 4 - 1)) == 0);
}

int softargmax_fwd_f32(const float* y, int N, int OH, int OW, int C, const float* temperature,
                       float* feat, float* smax, float* ssum, cudaStream_t st) {
  if (N == 0) return 0;
  constexpr int G = 16;
  if (sa_v4_ok(C, y, nullptr, 0)) {
    const size_t smem4 = (4 * (size_t)G * C + 2 * (size_t)OH * OW) * sizeof(float);
    if (smem4 <= 48 * 1024) {
      softargmax_fwd_v4_kernel<G><<<N, dim3(C / 4, G), smem4, st>>>(y, OH * OW, OW, C, temperature, feat, smax, ssum);
      TACORL_LAUNCH_CHECK();
      return 0;
    }
  }
  TACORL_REQUIRE(C * G <= 1024, "softargmax: too many channels");
  const size_t smem = (4 * (size_t)G * C + 2 * (size_t)OH * OW) * sizeof(float);
  TACORL_REQUIRE(smem <= 48 * 1024, "softargmax: feature map of %d x %d positions is too large", OH, OW);
  softargmax_fwd_kernel<G><<<N, dim3(C, G), smem, st>>>(y, OH * OW, OW, C, temperature, feat, smax, ssum);
  TACORL_LAUNCH_CHECK();
  return 0;
}

// Backward: dz_i = p_i * (gx*col_i + gy*row_i - (gx*fx + gy*fy));  dy_i = dz_i / tau * [y_i > 0]
// dtau = sum_i dz_i * (-y_i / tau^2), reduced per frame into dtau_part[n] (summed by a colsum).
__device__ __forceinline__ void sa_store(float* p, float v) { *p = v; }
__device__ __forceinline__ void sa_store(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <int G, typename OutT>
__global__ void softargmax_bwd_kernel(const float* __restrict__ y, int P, int OW, int C,
                                      const float* __restrict__ temperature,
                                      const float* __restrict__ feat, const float* __restrict__ smax,
                                      const float* __restrict__ ssum, const float* __restrict__ dfeat,
                                      OutT* __restrict__ dy, float* __restrict__ dtau_part) {
  __shared__ float red[32];
  extern __shared__ float sm[];   // (col, row) coordinate of every position
  const int c = threadIdx.x, g = threadIdx.y;
  const long long n = blockIdx.x;
  const float tau = __ldg(temperature), inv_t = 1.f / tau;
  const float* yp = y + n * (long long)P * C;
  OutT* dyp = dy + n * (long long)P * C;
  float2* xy = reinterpret_cast<float2*>(sm);
  for (int p = g * C + c; p < P; p += G * C) xy[p] = make_float2((float)(p % OW), (float)(p / OW));
  __syncthreads();
  const float gx = dfeat[n * 2 * C + 2 * c], gy = dfeat[n * 2 * C + 2 * c + 1];
  const float fx = feat[n * 2 * C + 2 * c], fy = feat[n * 2 * C + 2 * c + 1];
  const float M = smax[n * C + c], invS = 1.f / ssum[n * C + c];
  const float dotg = gx * fx + gy * fy;
  float dt = 0.f;
  constexpr int UB = 8;
  for (int p0 = g; p0 < P; p0 += G * UB) {
    float yv[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      yv[j] = p < P ? yp[(long long)p * C + c] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < UB; ++j) {
      const int p = p0 + j * G;
      if (p >= P) break;
      const float pr = __expf(yv[j] * inv_t - M) * invS;
      const float2 q = xy[p];
      const float dz = pr * (gx * q.x + gy * q.y - dotg);
      dt += dz * yv[j];
      sa_store(dyp + (long long)p * C + c, yv[j] > 0.f ? dz * inv_t : 0.f);
    }
  }
  dt = warp_sum(dt);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int nwarps = (blockDim.x * blockDim.y + 31) / 32;
  if ((tid & 31) == 0) red[tid >> 5] = dt;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < nwarps; ++i) t += red[i];
    dtau_part[n] = -t * inv_t * inv_t;
  }
}

int softargmax_bwd_f32(const float* y, int N, int OH, int OW, int C, const float* temperature,
                       const float* feat, const float* smax, const float* ssum, const float* dfeat,
                       float* dy, float* dtau_part, cudaStream_t st) {
  if (N == 0) return 0;
  constexpr int G = 4;
  TACORL_REQUIRE((size_t)OH * OW * 8 <= 40 * 1024, "softargmax bwd: feature map of %d x %d positions is too large", OH, OW);
  if (sa_v4_ok(C, y, dy, 4)) {
    softargmax_bwd_v4_kernel<16, float><<<N, dim3(C / 4, 16), (size_t)OH * OW * 8, st>>>(y, OH * OW, OW, C, temperature, feat,
                                                                                       smax, ssum, dfeat, dy, dtau_part,
                                                                                       OW, OH * OW);
    TACORL_LAUNCH_CHECK();
    return 0;
  }
  TACORL_REQUIRE(C * G <= 1024 && (C * G) % 32 == 0, "softargmax bwd: unsupported channel count");
  softargmax_bwd_kernel<G, float><<<N, dim3(C, G), (size_t)OH * OW * 8, st>>>(y, OH * OW, OW, C, temperature, feat, smax,
                                                                            ssum, dfeat, dy, dtau_part);
  TACORL_LAUNCH_CHECK();
  return 0;
}

// same, gradient written as bf16 (operand of the tensor-core conv3 weight / data gradient kernels)
int softargmax_bwd_bf16out(const float* y, int N, int OH, int OW, int C, const float* temperature,
                           const float* feat, const float* smax, const float* ssum, const float* dfeat,
                           void* dy_bf16, float* dtau_part, int pad_h, int pad_w, cudaStream_t st) {
  if (N == 0) return 0;
  constexpr int G = 16;
  TACORL_REQUIRE((size_t)OH * OW * 8 <= 40 * 1024, "softargmax bwd: feature map of %d x %d positions is too large", OH, OW);
  if (sa_v4_ok(C, y, dy_bf16, 2)) {
    softargmax_bwd_v4_kernel<G, __nv_bfloat16><<<N, dim3(C / 4, G), (size_t)OH * OW * 8, st>>>(
        y, OH * OW, OW, C, temperature, feat, smax, ssum, dfeat, (__nv_bfloat16*)dy_bf16, dtau_part, OW + pad_w,
        (OH + pad_h) * (OW + pad_w));
    TACORL_LAUNCH_CHECK();
    return 0;
  }
  TACORL_REQUIRE(pad_h == 0 && pad_w == 0, "softargmax bwd: padded gradient layout needs the 4-channel kernel");
  TACORL_REQUIRE(C * G <= 1024 && (C * G) % 32 == 0, "softargmax bwd: unsupported channel count");
  softargmax_bwd_kernel<G, __nv_bfloat16><<<N, dim3(C, G), (size_t)OH * OW * 8, st>>>(
      y, OH * OW, OW, C, temperature, feat, smax, ssum, dfeat, (__nv_bfloat16*)dy_bf16, dtau_part);
  TACORL_LAUNCH_CHECK();
  return 0;
}

}  // namespace tacorl

// ==========================================================================================
namespace tacorl {

// column sums of a tall bf16 matrix [M][N] (bias gradients), two-stage like colsum_tall_kernel
__global__ void colsum_tall_bf16_kernel(long long M, int N, const __nv_bfloat16* __restrict__ X, float* __restrict__ part) {
  extern __shared__ float sm[];
  const int n = threadIdx.x, r = threadIdx.y, R = blockDim.y;
  const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
  const long long beg = blockIdx.x * rows_per, end = min(M, beg + rows_per);
  float s = 0.f;
  for (long long m = beg + r; m < end; m += R) s += __bfloat162float(X[m * N + n]);
  sm[r * N + n] = s;
  __syncthreads();
  if (r == 0) {
    float t = 0.f;
    for (int i = 0; i < R; ++i) t += sm[i * N + n];
    part[(long long)blockIdx.x * N + n] = t;
  }
}

int colsum_tall_bf16(long long M, int N, const void* X, float* out, int accumulate, float* ws, size_t ws_bytes,
                     cudaStream_t st) {
  TACORL_REQUIRE(N <= 256 && 256 % N == 0, "colsum_tall_bf16: N must divide 256");
  int blocks = (int)min((long long)592, (M + 255) / 256);
  if (blocks < 1) blocks = 1;
  TACORL_REQUIRE(ws && ws_bytes >= (size_t)blocks * N * sizeof(float), "colsum_tall_bf16: workspace too small");
  colsum_tall_bf16_kernel<<<blocks, dim3(N, 256 / N), 256 * sizeof(float), st>>>(M, N, (const __nv_bfloat16*)X, ws);
  TACORL_LAUNCH_CHECK();
  return colsum_f32(blocks, N, ws, N, out, accumulate, st);
}

}  // namespace tacorl
