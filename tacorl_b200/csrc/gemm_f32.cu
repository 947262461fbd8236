// fp32 SIMT GEMM family (the parity path: every contraction of the PlayLMP / TACO-RL step can
// run through here in full fp32, matching the reference's fp32 PyTorch math to ~1e-6).
// The bf16 tcgen05 kernels (gemm_bf16_tc.cu) replace these on the performance path.
#include "common.cuh"
#include <cstdarg>
#include <atomic>

namespace tacorl {

static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_last_error; }

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

__device__ __forceinline__ float apply_act(int act, float v) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SILU) return v / (1.f + __expf(-v));
  return v;
}

// C tile BMxBN per CTA, TMxTN per thread, K stepped by BK through shared memory.
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(GemmArgs g, int k_chunk, float* __restrict__ partial) {
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_chunk;
  const int kend = min(g.K, kbeg + k_chunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  constexpr int LA = (BM * BK + NT - 1) / NT, LB = (BN * BK + NT - 1) / NT;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // all global loads of the K tile are issued before the first shared-memory store (one memory latency per
    // tile instead of one per element)
    float ra[LA], rb[LB];
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int e = tid + i * NT;
      int m, k;
      if (g.transA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      const int gm = m0 + m, gk = k0 + k;
      const bool ok = e < BM * BK && gm < g.M && gk < kend;
      const long long off = g.transA ? (long long)gk * g.lda + gm : (long long)gm * g.lda + gk;
      ra[i] = ok ? __ldg(g.A + off) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int e = tid + i * NT;
      int n, k;
      if (g.transB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      const bool ok = e < BN * BK && gn < g.N && gk < kend;
      const long long off = g.transB ? (long long)gn * g.ldb + gk : (long long)gk * g.ldb + gn;
      rb[i] = ok ? __ldg(g.B + off) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int e = tid + i * NT;
      int m, k;
      if (g.transA) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
      if (e < BM * BK) As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int e = tid + i * NT;
      int n, k;
      if (g.transB) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      if (e < BN * BK) Bs[k][n] = rb[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  if (gridDim.z > 1) {
    float* P = partial + (long long)blockIdx.z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int gm = m0 + ty * TM + i;
      if (gm >= g.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int gn = n0 + tx * TN + j;
        if (gn < g.N) P[(long long)gm * g.N + gn] = acc[i][j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.beta != 0.f) v = fmaf(g.beta, g.C[(long long)gm * g.ldc + gn], v);
      if (g.bias) v += __ldg(g.bias + gn);
      if (g.Cpre) g.Cpre[(long long)gm * g.ldpre + gn] = v;
      g.C[(long long)gm * g.ldc + gn] = apply_act(g.act, v);
    }
  }
}

__global__ void splitk_reduce_kernel(GemmArgs g, int splits, const float* __restrict__ partial) {
  const long long total = (long long)g.M * g.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int gm = (int)(idx / g.N), gn = (int)(idx % g.N);
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(long long)z * total + idx];
    float v = g.alpha * s;
    if (g.beta != 0.f) v = fmaf(g.beta, g.C[(long long)gm * g.ldc + gn], v);
    if (g.bias) v += __ldg(g.bias + gn);
    if (g.Cpre) g.Cpre[(long long)gm * g.ldpre + gn] = v;
    g.C[(long long)gm * g.ldc + gn] = apply_act(g.act, v);
  }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_sgemm(const GemmArgs& g, int splits, float* ws, cudaStream_t st) {
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), splits);
  int k_chunk = cdiv(cdiv(g.K, splits), BK) * BK;
  sgemm_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g, k_chunk, ws);
  TACORL_LAUNCH_CHECK();
  if (splits > 1) {
    long long total = (long long)g.M * g.N;
    int blocks = (int)min((long long)1184, (total + 255) / 256);
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(g, splits, ws);
    TACORL_LAUNCH_CHECK();
  }
  return 0;
}

int gemm_f32(const GemmArgs& g, float* ws, size_t ws_bytes, cudaStream_t st) {
  TACORL_REQUIRE(g.M >= 0 && g.N >= 0 && g.K >= 0, "gemm_f32: negative dims");
  if (g.M == 0 || g.N == 0) return 0;
  TACORL_REQUIRE(g.A && g.B && g.C, "gemm_f32: null operand");
  constexpr int BK = 16;
  // tile choice by output width, then by how many CTAs the grid gets
  int BM, BN;
  if (g.N <= 32) { BM = 256; BN = 32; }
  else if (g.N <= 64) { BM = 128; BN = 64; }
  else { BM = 128; BN = 128; }
  long long ctas = (long long)cdiv(g.M, BM) * cdiv(g.N, BN);
  bool small = false;
  if (ctas < 148) {
    BM = 64; BN = 64;
    ctas = (long long)cdiv(g.M, BM) * cdiv(g.N, BN);
    small = true;   // latency-bound: deep K tiles (64 per iteration, every load of a tile in flight at once)
  }
  int splits = g.split_k;
  if (splits <= 0) {  // auto: fill ~2 waves of the 148 SMs when the output grid is small
    splits = 1;
    if (ws && small && g.K >= 128) {          // one or two 64-deep iterations per CTA
      splits = (int)min((long long)cdiv(296, ctas), (long long)cdiv(g.K, 64));
      if (splits < 1) splits = 1;
    }
  }
  if (splits > 1) {
    size_t need = (size_t)splits * g.M * g.N * sizeof(float);
    if (!ws || need > ws_bytes) {
      splits = (ws && ws_bytes >= 2ull * g.M * g.N * sizeof(float))
                   ? (int)(ws_bytes / ((size_t)g.M * g.N * sizeof(float)))
                   : 1;
      if (g.split_k > 1 && splits < 2) {
        set_last_error("gemm_f32: split-K workspace too small (%zu < %zu)", ws_bytes, need);
        return -1;
      }
    }
  }
  if (BM == 256) return launch_sgemm<256, 32, BK, 8, 4>(g, splits, ws, st);
  if (BM == 128 && BN == 64) return launch_sgemm<128, 64, BK, 8, 4>(g, splits, ws, st);
  if (BM == 128) return launch_sgemm<128, 128, BK, 8, 8>(g, splits, ws, st);
  if (small) return launch_sgemm<64, 64, 64, 4, 4>(g, splits, ws, st);
  return launch_sgemm<64, 64, BK, 4, 4>(g, splits, ws, st);
}

// ------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] (+)= sum_m X[m][n].  One CTA of 32 x R threads per 32 columns.
template <int R>
__global__ void colsum_kernel(int M, int N, const float* __restrict__ X, long long ldx,
                              float* __restrict__ out, int accumulate) {
  __shared__ float red[R][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < M; m += R) s += X[(long long)m * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < R; ++i) t += red[i][threadIdx.x];
    out[n] = accumulate ? out[n] + t : t;
  }
}

// N == 1: plain deterministic sum of a strided vector by one 1024-thread CTA (loss reductions)
__global__ void vecsum_kernel(int M, const float* __restrict__ X, long long ldx, float* __restrict__ out, int accumulate) {
  __shared__ float red[32];
  float s = 0.f;
  for (int m = threadIdx.x; m < M; m += 1024) s += X[(long long)m * ldx];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + t : t;
  }
}

// slab-parallel variant: grid (N/32, slabs); part[slab][N]
__global__ void colsum_slab_kernel(int M, int N, const float* __restrict__ X, long long ldx, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int beg = blockIdx.y * rows_per, end = min(M, beg + rows_per);
  float s = 0.f;
  if (n < N)
    for (int m = beg + threadIdx.y; m < end; m += 8) s += X[(long long)m * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    part[(long long)blockIdx.y * N + n] = t;
  }
}

int colsum_f32(int M, int N, const float* X, long long ldx, float* out, int accumulate, cudaStream_t st) {
  if (N == 0) return 0;
  if (N == 1) vecsum_kernel<<<1, 1024, 0, st>>>(M, X, ldx, out, accumulate);
  else if (M >= 256) colsum_kernel<32><<<cdiv(N, 32), dim3(32, 32), 0, st>>>(M, N, X, ldx, out, accumulate);
  else colsum_kernel<8><<<cdiv(N, 32), dim3(32, 8), 0, st>>>(M, N, X, ldx, out, accumulate);
  TACORL_LAUNCH_CHECK();
  return 0;
}

// deterministic two-stage column sum that fills the machine when M is large and N moderate
int colsum2_f32(int M, int N, const float* X, long long ldx, float* out, int accumulate, float* ws, size_t ws_bytes,
                cudaStream_t st) {
  if (N == 0) return 0;
  int slabs = min(16, max(1, M / 64));
  if (!ws || ws_bytes < (size_t)slabs * N * sizeof(float) || slabs == 1) return colsum_f32(M, N, X, ldx, out, accumulate, st);
  colsum_slab_kernel<<<dim3(cdiv(N, 32), slabs), dim3(32, 8), 0, st>>>(M, N, X, ldx, ws);
  TACORL_LAUNCH_CHECK();
  return colsum_f32(slabs, N, ws, N, out, accumulate, st);
}

// two-stage column sum for very tall matrices (conv bias grads: M ~ 2.4M, N = 32/64)
__global__ void colsum_tall_kernel(long long M, int N, const float* __restrict__ X, float* __restrict__ part) {
  // blockDim = (N, 256/N); each CTA reduces a row-slab to part[blockIdx.x][N]
  extern __shared__ float sm[];
  const int n = threadIdx.x, r = threadIdx.y, R = blockDim.y;
  const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
  const long long beg = blockIdx.x * rows_per, end = min(M, beg + rows_per);
  float s = 0.f;
  for (long long m = beg + r; m < end; m += R) s += X[m * N + n];
  sm[r * N + n] = s;
  __syncthreads();
  if (r == 0) {
    float t = 0.f;
    for (int i = 0; i < R; ++i) t += sm[i * N + n];
    part[(long long)blockIdx.x * N + n] = t;
  }
}

int colsum_tall_f32(long long M, int N, const float* X, float* out, int accumulate, float* ws,
                    size_t ws_bytes, cudaStream_t st) {
  TACORL_REQUIRE(N <= 256 && 256 % N == 0, "colsum_tall: N must divide 256");
  int blocks = (int)min((long long)592, (M + 255) / 256);
  if (blocks < 1) blocks = 1;
  TACORL_REQUIRE(ws && ws_bytes >= (size_t)blocks * N * sizeof(float), "colsum_tall: workspace too small");
  colsum_tall_kernel<<<blocks, dim3(N, 256 / N), 256 * sizeof(float), st>>>(M, N, X, ws);
  TACORL_LAUNCH_CHECK();
  return colsum_f32(blocks, N, ws, N, out, accumulate, st);
}

// ------------------------------------------------------------------------------------------
__global__ void act_bwd_kernel(int act, long long n, const float* __restrict__ dY,
                               const float* __restrict__ Y, float* __restrict__ dZ) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float g = dY[i], y = Y[i], r;
    if (act == ACT_RELU) r = y > 0.f ? g : 0.f;
    else if (act == ACT_SILU) { float s = 1.f / (1.f + __expf(-y)); r = g * s * (1.f + y * (1.f - s)); }
    else r = g;
    dZ[i] = r;
  }
}

int act_bwd_f32(int act, long long n, const float* dY, const float* YorPre, float* dZ, cudaStream_t st) {
  if (n == 0) return 0;
  int blocks = (int)min((long long)1184, (n + 255) / 256);
  act_bwd_kernel<<<blocks, 256, 0, st>>>(act, n, dY, YorPre, dZ);
  TACORL_LAUNCH_CHECK();
  return 0;
}

}  // namespace tacorl
