// Shared helpers for the tacorl_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>

namespace tacorl {

void set_last_error(const char* fmt, ...);
void note_launch();   // counts kernel launches issued by this library (bench.py's gpu_launches)

#define TACORL_CHECK_CUDA(expr)                                                        \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::tacorl::set_last_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,          \
                               cudaGetErrorName(_e), cudaGetErrorString(_e));          \
      return -2;                                                                       \
    }                                                                                  \
  } while (0)

#define TACORL_REQUIRE(cond, ...)                                                      \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      ::tacorl::set_last_error(__VA_ARGS__);                                           \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

#define TACORL_LAUNCH_CHECK()            \
  do {                                   \
    ::tacorl::note_launch();             \
    TACORL_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };
enum Prec { PREC_F32 = 0, PREC_BF16 = 1 };

struct GemmArgs {
  int transA = 0, transB = 0;   // op(A): MxK, op(B): KxN, all row-major with leading dims
  int M = 0, N = 0, K = 0;
  float alpha = 1.f;
  const float* A = nullptr; long long lda = 0;
  const float* B = nullptr; long long ldb = 0;
  float beta = 0.f;             // C = act(alpha*op(A)op(B) + beta*C + bias)
  float* C = nullptr; long long ldc = 0;
  const float* bias = nullptr;  // per output column (N)
  int act = ACT_NONE;
  float* Cpre = nullptr; long long ldpre = 0;   // optional pre-activation output
  int split_k = 1;              // >1: partials in ws, deterministic reduce
  // optional dense bf16 copies of A / B as stored (pitch = stored column count); used by the tensor-core path
  const void* A_bf16 = nullptr; const void* B_bf16 = nullptr;
};

// fp32 SIMT GEMM (parity path).  ws is only needed when split_k > 1 (split_k*M*N floats).
int gemm_f32(const GemmArgs& g, float* ws, size_t ws_bytes, cudaStream_t st);
// out[n] (+)= sum_m X[m*ldx + n]
int colsum_f32(int M, int N, const float* X, long long ldx, float* out, int accumulate, cudaStream_t st);
int colsum2_f32(int M, int N, const float* X, long long ldx, float* out, int accumulate, float* ws, size_t ws_bytes,
                cudaStream_t st);
// dZ = dY * act'(.)  (relu: uses Y (post-act); silu: uses pre-activation)
int act_bwd_f32(int act, long long n, const float* dY, const float* YorPre, float* dZ, cudaStream_t st);

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) {
  // matches F.softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace tacorl
