// Optimiser-side kernels over flat fp32 parameter buffers: global-norm clip + Adam in one pass,
// Polyak target update, squared-norm reduction.  HBM-bound streaming kernels (grid = multiple of
// the 148 SMs, float4 accesses).  Reference call sites: torch.optim.Adam at
// /root/reference/src/tacorl/modules/play_lmp/play_lmp_for_rl.py:362-368 and
// modules/cql/cql_offline_lightning.py:553-574; clip_grad_norm_ at :522-537;
// soft_update_from_to at :229-232.
#include "common.cuh"
#include <cuda_bf16.h>
#include "internal.h"
#include "../../include/tacorl_b200.h"

namespace tacorl {

constexpr int kOptBlocks = 148 * 4;

// Adam (no weight decay / amsgrad), torch semantics:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// g is first multiplied by grad_scale * clip_coef, clip_coef = min(1, max_norm / (sqrt(*sqnorm)*grad_scale + 1e-6))
// when sqnorm != nullptr (clip_grad_norm_ semantics on the already-scaled gradient).
template <int UN>
__global__ void adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, float lr, float b1, float b2,
                            float eps, float bc1, float bc2_sqrt, float grad_scale,
                            const float* __restrict__ sqnorm, float max_norm, const int* __restrict__ step_dev,
                            __nv_bfloat16* __restrict__ shadow) {
  if (step_dev) {   // CUDA-graph friendly: bias corrections from a device-resident step counter
    const float t = (float)(*step_dev);
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  float gs = grad_scale;
  if (sqnorm) {
    const float norm = sqrtf(*sqnorm) * grad_scale;
    gs *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float step = lr / bc1;
  const long long n4 = n >> 2;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  float4* p4 = (float4*)p; const float4* g4 = (const float4*)g; float4* m4 = (float4*)m; float4* v4 = (float4*)v;
  for (long long i0 = tid; i0 < n4; i0 += nth * UN) {
    float4 pp[UN], gg[UN], mm[UN], vv[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {          // every load of the iteration in flight before the first use
      const long long i = i0 + u * nth;
      if (i < n4) { pp[u] = p4[i]; gg[u] = g4[i]; mm[u] = m4[i]; vv[u] = v4[i]; }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = i0 + u * nth;
      if (i >= n4) break;
#define ADAMC(c)                                                               \
      { const float gx = gg[u].c * gs;                                         \
        mm[u].c = b1 * mm[u].c + (1.f - b1) * gx;                              \
        vv[u].c = b2 * vv[u].c + (1.f - b2) * gx * gx;                         \
        pp[u].c -= step * mm[u].c / (sqrtf(vv[u].c) / bc2_sqrt + eps); }
      ADAMC(x) ADAMC(y) ADAMC(z) ADAMC(w)
#undef ADAMC
      p4[i] = pp[u]; m4[i] = mm[u]; v4[i] = vv[u];
      if (shadow) {     // bf16 copy of the updated parameters: the tensor-core operands of the next step
        const __nv_bfloat162 lo = __floats2bfloat162_rn(pp[u].x, pp[u].y), hi = __floats2bfloat162_rn(pp[u].z, pp[u].w);
        reinterpret_cast<uint2*>(shadow)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo),
                                                          *reinterpret_cast<const uint32_t*>(&hi));
      }
    }
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) {
    const float gx = g[i] * gs;
    const float mm = b1 * m[i] + (1.f - b1) * gx;
    const float vv = b2 * v[i] + (1.f - b2) * gx * gx;
    m[i] = mm; v[i] = vv;
    const float pn = p[i] - step * mm / (sqrtf(vv) / bc2_sqrt + eps);
    p[i] = pn;
    if (shadow) shadow[i] = __float2bfloat16_rn(pn);
  }
}

__global__ void increment_kernel(int* p) { *p += 1; }

__global__ void polyak_kernel(long long n, float* __restrict__ tgt, const float* __restrict__ src, float tau) {
  const long long n4 = n >> 2;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  float4* t4 = (float4*)tgt; const float4* s4 = (const float4*)src;
  const float k = 1.f - tau;
  for (long long i = tid; i < n4; i += nth) {
    float4 t = t4[i], s = s4[i];
    t.x = t.x * k + s.x * tau; t.y = t.y * k + s.y * tau; t.z = t.z * k + s.z * tau; t.w = t.w * k + s.w * tau;
    t4[i] = t;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) tgt[i] = tgt[i] * k + src[i] * tau;
}

// deterministic two-stage sum of squares: part[blockIdx] then a single-CTA finish
__global__ void sqnorm_part_kernel(long long n, const float* __restrict__ x, float* __restrict__ part) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) { const float v = x[i]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) part[blockIdx.x] = t;
  }
}
__global__ void sum_finish_kernel(int n, const float* __restrict__ part, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) out[0] = t;
  }
}

}  // namespace tacorl

using namespace tacorl;

extern "C" {

int tacorl_adam_step_range(long long n, float* p, const float* g, float* m, float* v, float lr, float beta1,
                           float beta2, float eps, int step, int* step_dev, int increment_step, int background,
                           float grad_scale, const float* sqnorm, float max_norm, void* shadow_bf16, void* stream);

int tacorl_adam_step(long long n, float* p, const float* g, float* m, float* v, float lr, float beta1,
                     float beta2, float eps, int step, int* step_dev, float grad_scale, const float* sqnorm,
                     float max_norm, void* shadow_bf16, void* stream) {
  return tacorl_adam_step_range(n, p, g, m, v, lr, beta1, beta2, eps, step, step_dev, 1, 0, grad_scale, sqnorm, max_norm,
                                shadow_bf16, stream);
}

int tacorl_adam_step_range(long long n, float* p, const float* g, float* m, float* v, float lr, float beta1,
                           float beta2, float eps, int step, int* step_dev, int increment_step, int background,
                           float grad_scale, const float* sqnorm, float max_norm, void* shadow_bf16, void* stream) {
  if (n == 0) return 0;
  TACORL_REQUIRE(p && g && m && v && (step >= 1 || step_dev), "adam_step: bad arguments");
  if (step_dev && increment_step) {
    increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    TACORL_LAUNCH_CHECK();
  }
  if (step_dev && step < 1) step = 1;
  TACORL_REQUIRE(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                     ((uintptr_t)v % 16 == 0) && ((uintptr_t)shadow_bf16 % 8 == 0),
                 "adam_step: buffers must be 16-byte aligned");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  if (background) {
    // meant to run UNDER other kernels (the encoder backward): one 256-thread CTA per SM (the foreground grid of
    // 4 CTAs per SM takes 57 K of an SM's 64 K registers and starves everything else), two 16-byte loads per array
    // in flight per thread instead
    int blocks = (int)min((long long)148, (n / 8 + 255) / 256 + 1);
    adam_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, p, g, m, v, lr, beta1, beta2, eps, bc1, sqrtf(bc2),
                                                            grad_scale, sqnorm, max_norm, step_dev,
                                                            (__nv_bfloat16*)shadow_bf16);
  } else {
    int blocks = (int)min((long long)kOptBlocks, (n / 4 + 255) / 256 + 1);
    adam_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(n, p, g, m, v, lr, beta1, beta2, eps, bc1, sqrtf(bc2),
                                                            grad_scale, sqnorm, max_norm, step_dev,
                                                            (__nv_bfloat16*)shadow_bf16);
  }
  TACORL_LAUNCH_CHECK();
  return 0;
}

int tacorl_polyak_update(long long n, float* target, const float* source, float tau, void* stream) {
  if (n == 0) return 0;
  TACORL_REQUIRE(target && source, "polyak_update: null pointer");
  TACORL_REQUIRE(((uintptr_t)target % 16 == 0) && ((uintptr_t)source % 16 == 0), "polyak_update: misaligned");
  int blocks = (int)min((long long)kOptBlocks, (n / 4 + 255) / 256 + 1);
  polyak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, target, source, tau);
  TACORL_LAUNCH_CHECK();
  return 0;
}

// out[0] = sum x^2 ; ws needs >= 592 floats
int tacorl_sqnorm(long long n, const float* x, float* out, float* ws, void* stream) {
  TACORL_REQUIRE(x && out && ws, "sqnorm: null pointer");
  int blocks = (int)min((long long)kOptBlocks, (n + 255) / 256);
  if (blocks < 1) blocks = 1;
  sqnorm_part_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, x, ws);
  TACORL_LAUNCH_CHECK();
  sum_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(blocks, ws, out);
  TACORL_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
