// Fused small-MLP chains (fp32 FFMA): a whole Linear -> act -> Linear -> ... stack, forward or backward, in ONE launch.
//
// Replaces the per-layer GEMM + split-K reduce + bias-gradient column sum + activation-gradient launches behind
//   VisualGoalEncoder.forward      /root/reference/src/tacorl/networks/visual_encoders/goal_encoder.py:29-33
//   MLPPolicy.forward              networks/actor_critic/actor.py:252-270   (3 x Linear+SiLU, fc_mean | fc_log_std)
//   MLPQNetwork.forward            networks/actor_critic/critic.py:92-97    (3 x Linear+SiLU, out)
// whose tensors are tiny (<= 256 wide, 64 .. 832 rows, < 1 MB of weights): those launches are latency-bound (a PlayLMP
// step spent 0.29 ms in ~60 of them, a TACO-RL step 0.97 ms in ~200).
//
// One cluster of 8 CTAs per block of 16 rows (64 rows = 4 clusters = 32 SMs).  Forward, per layer: every CTA holds the
// layer's full input activations (16 x in) in shared memory and computes out/8 of the output columns (its slice of W
// staged in shared memory, prefetched a layer ahead); every output value is pushed, already activated, into the
// next-input buffer of all 8 CTAs through distributed shared memory, and ONE cluster barrier per layer publishes the
// pushes: no global-memory round trip between layers.  Pre-activations also go to global memory (the saved tensors of
// the backward pass).  Backward, per layer: dZ = dY * act'(z) (every CTA, full), dW / db for the CTA's slice of output
// rows (complete sums when there is one row block, else per-row-block partials reduced by a second small launch in a
// fixed order), dX for the CTA's slice of input columns, pushed to the peers the same way.  fp32 throughout: this is
// also the parity path.  Up to 256 rows: for the 13 x 64 = 832 rows of the CQL twin-Q passes the per-layer tensor-core GEMMs
// are as fast (measured: TACO-RL step 3.86 ms with them, 3.91 ms with 52-cluster chains), so larger batches keep them.
#include "common.cuh"
#include "internal.h"
#include "../../include/tacorl_b200.h"

namespace tacorl {

constexpr int MC_ROWS = 16;        // rows per cluster (small: these chains are latency-bound, more clusters = more SMs)
constexpr int MC_CTAS = 8;         // CTAs per cluster
constexpr int MC_THREADS = 256;
constexpr int MC_MAXW = 256;       // widest layer
constexpr int MC_PAD = 4;
constexpr int MC_LD = MC_MAXW + MC_PAD;
constexpr int MC_MAXL = 4;

constexpr int MC_SEGS = 3;
struct McLayer {
  const float* W[MC_SEGS]; const float* b[MC_SEGS];   // up to three weight segments stacked along the output dim
                                                      // (fc_mean | fc_log_std | gripper_action)
  int n[MC_SEGS];                         // output rows of each segment (n[1], n[2] may be 0)
  int in, out, act;                       // act applied to this layer's output (ACT_NONE for the last layer)
  int zpost;                              // store act(z) instead of z in Z (ReLU only: the backward pass cannot tell)
  float* dW[MC_SEGS]; float* db[MC_SEGS]; // backward outputs (may be null: gradient w.r.t. the input only)
  long long part_off;                     // offset of this layer's [out][in + 1] block inside one partial slab
};
struct McChain {
  int L, rows;
  McLayer layer[MC_MAXL];
  const float* X[2]; int xin[2]; long long ldx[2];      // input = concat of up to two tensors along the feature dim
  float* Z; long long ldz; int zoff[MC_MAXL];           // saved pre-activations of layers 0 .. L-2, packed per row
  float* out; long long ldo;                            // last layer's output
  // backward only
  const float* dOut; long long lddo;
  float* dX[2]; long long lddx[2];                      // may be null
  float* dybuf[2];                                      // scratch [rows][MC_MAXW] ping-pong
  float* part; long long part_stride;                   // per-row-block partial slabs (row blocks > 1), else null
};

__device__ __forceinline__ float mc_act(int act, float v) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SILU) return v / (1.f + __expf(-v));
  return v;
}
__device__ __forceinline__ float mc_dact(int act, float z) {
  if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == ACT_SILU) { const float s = 1.f / (1.f + __expf(-z)); return s * (1.f + z * (1.f - s)); }
  return 1.f;
}
__device__ __forceinline__ void mc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// output row n of a layer -> (segment, row inside the segment)
__device__ __forceinline__ int mc_seg(const McLayer& l, int n, int& nn) {
  if (n < l.n[0]) { nn = n; return 0; }
  if (n < l.n[0] + l.n[1]) { nn = n - l.n[0]; return 1; }
  nn = n - l.n[0] - l.n[1];
  return 2;
}
__device__ __forceinline__ const float* mc_wrow(const McLayer& l, int n) {
  int nn;
  const int sg = mc_seg(l, n, nn);
  return l.W[sg] + (long long)nn * l.in;
}

// Tile loader: dst[r][4*q .. 4*q+3] = src(r, q) for r < nrows, q < ncols4, with MC_UN independent 16-byte loads in flight
// per thread (their latency, not their bandwidth, is the cost)
constexpr int MC_UN = 4;
template <typename F>
__device__ __forceinline__ void mc_tile_load4(float* dst, int nrows, int ncols4, F src) {
  const int total = nrows * ncols4;
  for (int base = threadIdx.x; base < total; base += MC_THREADS * MC_UN) {
    float4 v[MC_UN];
#pragma unroll
    for (int u = 0; u < MC_UN; ++u) {
      const int i = base + u * MC_THREADS;
      if (i < total) { const int r = i / ncols4; v[u] = src(r, i - r * ncols4); }
    }
#pragma unroll
    for (int u = 0; u < MC_UN; ++u) {
      const int i = base + u * MC_THREADS;
      if (i < total) { const int r = i / ncols4; *reinterpret_cast<float4*>(dst + r * MC_LD + 4 * (i - r * ncols4)) = v[u]; }
    }
  }
}
__device__ __forceinline__ float4 mc_act4(int act, float4 v) {
  return make_float4(mc_act(act, v.x), mc_act(act, v.y), mc_act(act, v.z), mc_act(act, v.w));
}
// store v at the same shared-memory offset in every CTA of the cluster (distributed shared memory)
__device__ __forceinline__ void mc_push_all(float* local, float v) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
#pragma unroll
  for (int p = 0; p < MC_CTAS; ++p) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(p));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
  }
}

// layer-0 input (concat of up to two tensors) for rows [r0, r0 + MC_ROWS) -> As; rows >= rows are zero-filled
__device__ void mc_load_x(const McChain& c, int r0, float* As) {
  const int in = c.layer[0].in;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = (c.xin[0] & 3) == 0 && (c.ldx[0] & 3) == 0 && (c.ldx[1] & 3) == 0 && ((uintptr_t)c.X[0] & 15) == 0 &&
                   (c.xin[1] == 0 || ((uintptr_t)c.X[1] & 15) == 0);
  if (vec) {
    const int q0 = c.xin[0] >> 2;
    mc_tile_load4(As, MC_ROWS, in >> 2, [&](int r, int q) {
      if (r0 + r >= c.rows) return zero;
      return q < q0 ? *reinterpret_cast<const float4*>(c.X[0] + (long long)(r0 + r) * c.ldx[0] + 4 * q)
                    : *reinterpret_cast<const float4*>(c.X[1] + (long long)(r0 + r) * c.ldx[1] + 4 * (q - q0));
    });
    return;
  }
  for (int i = threadIdx.x; i < MC_ROWS * in; i += MC_THREADS) {
    const int r = i / in, k = i - r * in;
    float v = 0.f;
    if (r0 + r < c.rows)
      v = k < c.xin[0] ? c.X[0][(long long)(r0 + r) * c.ldx[0] + k] : c.X[1][(long long)(r0 + r) * c.ldx[1] + (k - c.xin[0])];
    As[r * MC_LD + k] = v;
  }
}
// input activations of layer li: the concat input (li == 0) or act(saved pre-activation of layer li - 1)
__device__ void mc_load_input(const McChain& c, int li, int r0, float* As) {
  if (li == 0) { mc_load_x(c, r0, As); return; }
  const int act = c.layer[li - 1].act;
  const float* zb = c.Z + c.zoff[li - 1];
  mc_tile_load4(As, MC_ROWS, c.layer[li].in >> 2, [&](int r, int q) {
    return r0 + r < c.rows ? mc_act4(act, *reinterpret_cast<const float4*>(zb + (long long)(r0 + r) * c.ldz + 4 * q))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  });
}
// my slice of a layer's weight rows [n0, n0 + nc) -> Ws[nc][MC_LD]
__device__ void mc_load_wslice(const McLayer& l, int n0, int nc, float* Ws) {
  mc_tile_load4(Ws, nc, l.in >> 2, [&](int n, int q) { return *reinterpret_cast<const float4*>(mc_wrow(l, n0 + n) + 4 * q); });
}

// ------------------------------------------------------------------------------------------ forward
// thread = (column j of my <= 32 columns, rows 2g and 2g+1); every output value is pushed, already activated, into the
// next-input buffer of all 8 CTAs through distributed shared memory; the cluster barrier that publishes the pushes is the
// only synchronisation per layer.  Pre-activations also go to global memory: the backward pass's saved tensors.
__global__ void __launch_bounds__(MC_THREADS, 2) mlp_chain_fwd_kernel(const __grid_constant__ McChain c) {
  extern __shared__ __align__(16) float mc_smem[];
  float* Ab[2] = {mc_smem, mc_smem + MC_ROWS * MC_LD};                                          // 2 x [16][MC_LD]
  float* Wb[2] = {mc_smem + 2 * MC_ROWS * MC_LD, mc_smem + (2 * MC_ROWS + 32) * MC_LD};         // 2 x [32][MC_LD]
  const int cta = blockIdx.x, r0 = blockIdx.y * MC_ROWS, tid = threadIdx.x;
  {
    const McLayer& l = c.layer[0];
    const int per = (l.out + MC_CTAS - 1) / MC_CTAS, n0 = cta * per;
    mc_load_wslice(l, n0, max(0, min(per, l.out - n0)), Wb[0]);
    mc_load_x(c, r0, Ab[0]);
  }
  // nobody may push into a peer before that peer has started (its shared memory is live only then)
  mc_cluster_sync();
  for (int li = 0; li < c.L; ++li) {
    const McLayer& l = c.layer[li];
    const float* Ws = Wb[li & 1];
    const float* As = Ab[li & 1];
    float* An = Ab[(li + 1) & 1];
    const int per = (l.out + MC_CTAS - 1) / MC_CTAS, n0 = cta * per, nc = max(0, min(per, l.out - n0));
    const int j = tid & 31, g = tid >> 5;
    if (j < nc) {
      float acc0 = 0.f, acc1 = 0.f;
      const float* wp = Ws + j * MC_LD;
      const float* ap = As + (2 * g) * MC_LD;
      for (int k = 0; k < l.in; k += 4) {       // (every layer width is a multiple of 4; checked on the host)
        const float4 w = *reinterpret_cast<const float4*>(wp + k);
        const float4 a0 = *reinterpret_cast<const float4*>(ap + k);
        const float4 a1 = *reinterpret_cast<const float4*>(ap + MC_LD + k);
        acc0 = fmaf(a0.x, w.x, acc0); acc0 = fmaf(a0.y, w.y, acc0); acc0 = fmaf(a0.z, w.z, acc0); acc0 = fmaf(a0.w, w.w, acc0);
        acc1 = fmaf(a1.x, w.x, acc1); acc1 = fmaf(a1.y, w.y, acc1); acc1 = fmaf(a1.z, w.z, acc1); acc1 = fmaf(a1.w, w.w, acc1);
      }
      const int n = n0 + j;
      int bn;
      const int bsg = mc_seg(l, n, bn);
      const float bias = l.b[bsg] ? l.b[bsg][bn] : 0.f;
      const float z[2] = {acc0 + bias, acc1 + bias};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int lr = 2 * g + r, row = r0 + lr;
        if (li == c.L - 1) {
          if (row < c.rows) c.out[(long long)row * c.ldo + n] = z[r];
        } else {
          if (row < c.rows) c.Z[(long long)row * c.ldz + c.zoff[li] + n] = l.zpost ? mc_act(l.act, z[r]) : z[r];
          mc_push_all(An + lr * MC_LD + n, mc_act(l.act, z[r]));
        }
      }
    }
    if (li + 1 < c.L) {
      // the next layer's weight slice does not depend on this layer: fetch it before waiting for the peers
      const McLayer& ln = c.layer[li + 1];
      const int pern = (ln.out + MC_CTAS - 1) / MC_CTAS, nn0 = cta * pern;
      mc_load_wslice(ln, nn0, max(0, min(pern, ln.out - nn0)), Wb[(li + 1) & 1]);
      mc_cluster_sync();
    }
  }
  // (the last pushes of the kernel precede the last barrier above: no CTA can exit while a peer still writes into it)
}

// ------------------------------------------------------------------------------------------ backward
__global__ void __launch_bounds__(MC_THREADS, 2) mlp_chain_bwd_kernel(const __grid_constant__ McChain c) {
  extern __shared__ __align__(16) float mc_smem[];
  float* As = mc_smem;                                                         // input activations of the layer [16][MC_LD]
  float* Db[2] = {mc_smem + MC_ROWS * MC_LD, mc_smem + 2 * MC_ROWS * MC_LD};   // dY / dZ, double buffered (peers push dX)
  float* Ws = mc_smem + 3 * MC_ROWS * MC_LD;                                   // W^T slice [32][MC_LD]
  const int cta = blockIdx.x, rb = blockIdx.y, r0 = rb * MC_ROWS, tid = threadIdx.x;
  {   // dOut -> Db[(L-1) & 1]
    const McLayer& l = c.layer[c.L - 1];
    float* Ds = Db[(c.L - 1) & 1];
    for (int i = tid; i < MC_ROWS * l.out; i += MC_THREADS) {
      const int r = i / l.out, n = i - r * l.out;
      Ds[r * MC_LD + n] = r0 + r < c.rows ? c.dOut[(long long)(r0 + r) * c.lddo + n] : 0.f;
    }
  }
  mc_cluster_sync();
  for (int li = c.L - 1; li >= 0; --li) {
    const McLayer& l = c.layer[li];
    float* Ds = Db[li & 1];
    float* Dn = Db[(li + 1) & 1];        // receives the peers' dX = dY of layer li - 1
    // ---- dZ = dY * act'(z) in place (every CTA, full [16][out]); input activations of the layer
    if (li < c.L - 1) {
      const float* zb = c.Z + c.zoff[li];
      const int act = l.act, o4 = l.out >> 2;
      for (int i = tid; i < MC_ROWS * o4; i += MC_THREADS) {
        const int r = i / o4, q = i - r * o4;
        float4 d = *reinterpret_cast<float4*>(Ds + r * MC_LD + 4 * q);
        if (r0 + r < c.rows) {
          const float4 z = *reinterpret_cast<const float4*>(zb + (long long)(r0 + r) * c.ldz + 4 * q);
          d.x *= mc_dact(act, z.x); d.y *= mc_dact(act, z.y); d.z *= mc_dact(act, z.z); d.w *= mc_dact(act, z.w);
        } else {
          d = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        *reinterpret_cast<float4*>(Ds + r * MC_LD + 4 * q) = d;
      }
    }
    if (l.out & 3)       // the dX loop reads dZ four columns at a time: the tail of the last group must be zero, not stale
      for (int i = tid; i < MC_ROWS * 4; i += MC_THREADS) {
        const int r = i >> 2, n = (l.out & ~3) + (i & 3);
        if (n >= l.out) Ds[r * MC_LD + n] = 0.f;
      }
    mc_load_input(c, li, r0, As);
    // ---- W^T slice for my input columns: Ws[k][n] = W[n][k0 + k]
    const bool need_dx = li > 0 || c.dX[0] || c.dX[1];
    const int perk = (l.in + MC_CTAS - 1) / MC_CTAS, k0 = cta * perk, kc = max(0, min(perk, l.in - k0));
    if (need_dx) {
      if ((kc & 3) == 0 && (k0 & 3) == 0) {
        const int kc4 = kc >> 2, total = l.out * kc4;
        for (int base = tid; base < total; base += MC_THREADS * MC_UN) {
          float4 v[MC_UN];
#pragma unroll
          for (int u = 0; u < MC_UN; ++u) {
            const int i = base + u * MC_THREADS;
            if (i < total) { const int n = i / kc4; v[u] = *reinterpret_cast<const float4*>(mc_wrow(l, n) + k0 + 4 * (i - n * kc4)); }
          }
#pragma unroll
          for (int u = 0; u < MC_UN; ++u) {
            const int i = base + u * MC_THREADS;
            if (i < total) {
              const int n = i / kc4, k = 4 * (i - n * kc4);
              Ws[k * MC_LD + n] = v[u].x; Ws[(k + 1) * MC_LD + n] = v[u].y; Ws[(k + 2) * MC_LD + n] = v[u].z; Ws[(k + 3) * MC_LD + n] = v[u].w;
            }
          }
        }
      } else {
        for (int i = tid; i < l.out * kc; i += MC_THREADS) {
          const int n = i / kc, k = i - n * kc;
          Ws[k * MC_LD + n] = mc_wrow(l, n)[k0 + k];
        }
      }
      const int o4r = (l.out + 3) & ~3;
      for (int i = tid; i < kc * (o4r - l.out); i += MC_THREADS) {      // zero the tail of a group of 4 when out % 4 != 0
        const int k = i / (o4r - l.out), n = l.out + i - k * (o4r - l.out);
        Ws[k * MC_LD + n] = 0.f;
      }
    }
    __syncthreads();
    // ---- dW / db for my slice of output rows: thread = input column k, 8 output rows at a time
    const int per = (l.out + MC_CTAS - 1) / MC_CTAS, n0 = cta * per, nc = max(0, min(per, l.out - n0));
    if (l.dW[0]) {
      for (int k = tid; k < l.in; k += MC_THREADS) {
        for (int nb = 0; nb < nc; nb += 8) {
          float acc[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll 4
          for (int r = 0; r < MC_ROWS; ++r) {
            const float a = As[r * MC_LD + k];
            const float* d = Ds + r * MC_LD + n0 + nb;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(d[q], a, acc[q]);      // (columns >= my slice: computed, never stored)
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int n = n0 + nb + q;
            if (nb + q < nc) {
              if (c.part) c.part[(long long)rb * c.part_stride + l.part_off + (long long)n * (l.in + 1) + k] = acc[q];
              else { int nn; const int sg = mc_seg(l, n, nn); l.dW[sg][(long long)nn * l.in + k] = acc[q]; }
            }
          }
        }
      }
      if (tid < nc) {
        const int n = n0 + tid;
        float s = 0.f;
        for (int r = 0; r < MC_ROWS; ++r) s += Ds[r * MC_LD + n];
        if (c.part) c.part[(long long)rb * c.part_stride + l.part_off + (long long)n * (l.in + 1) + l.in] = s;
        else { int nn; const int sg = mc_seg(l, n, nn); if (l.db[sg]) l.db[sg][nn] = s; }
      }
    }
    // ---- dX for my slice of input columns: dX[r][k] = sum_n dZ[r][n] W[n][k]; thread = (column j, rows 2g, 2g+1)
    if (need_dx) {
      const int j = tid & 31, g = tid >> 5;
      if (j < kc) {
        float acc0 = 0.f, acc1 = 0.f;
        const float* wp = Ws + j * MC_LD;
        const float* dp = Ds + (2 * g) * MC_LD;
        for (int n = 0; n < l.out; n += 4) {
          const float4 w = *reinterpret_cast<const float4*>(wp + n);
          const float4 d0 = *reinterpret_cast<const float4*>(dp + n);
          const float4 d1 = *reinterpret_cast<const float4*>(dp + MC_LD + n);
          acc0 = fmaf(d0.x, w.x, acc0); acc0 = fmaf(d0.y, w.y, acc0); acc0 = fmaf(d0.z, w.z, acc0); acc0 = fmaf(d0.w, w.w, acc0);
          acc1 = fmaf(d1.x, w.x, acc1); acc1 = fmaf(d1.y, w.y, acc1); acc1 = fmaf(d1.z, w.z, acc1); acc1 = fmaf(d1.w, w.w, acc1);
        }
        const int k = k0 + j;
        const float v[2] = {acc0, acc1};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int lr = 2 * g + r, row = r0 + lr;
          if (li > 0) mc_push_all(Dn + lr * MC_LD + k, v[r]);
          else if (row < c.rows) {
            if (k < c.xin[0]) { if (c.dX[0]) c.dX[0][(long long)row * c.lddx[0] + k] = v[r]; }
            else if (c.dX[1]) c.dX[1][(long long)row * c.lddx[1] + (k - c.xin[0])] = v[r];
          }
        }
      }
    }
    if (li > 0) mc_cluster_sync();
  }
}

// partial slabs -> gradients, fixed summation order over the row blocks
__global__ void mlp_chain_reduce_kernel(const McChain c, int row_blocks) {
  for (int li = 0; li < c.L; ++li) {
    const McLayer& l = c.layer[li];
    if (!l.dW[0]) continue;
    const long long total = (long long)l.out * (l.in + 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      float s = 0.f;
      for (int rb = 0; rb < row_blocks; ++rb) s += c.part[(long long)rb * c.part_stride + l.part_off + i];
      const int n = (int)(i / (l.in + 1)), k = (int)(i - (long long)n * (l.in + 1));
      int nn;
      const int seg = mc_seg(l, n, nn);
      if (k < l.in) l.dW[seg][(long long)nn * l.in + k] = s;
      else if (l.db[seg]) l.db[seg][nn] = s;
    }
  }
}

static int mc_check(const McChain& c, bool bwd) {
  TACORL_REQUIRE(c.L >= 1 && c.L <= MC_MAXL, "mlp_chain: 1..%d layers supported (got %d)", MC_MAXL, c.L);
  TACORL_REQUIRE(c.rows >= 1, "mlp_chain: empty batch");
  int in = c.xin[0] + c.xin[1];
  for (int i = 0; i < c.L; ++i) {
    const McLayer& l = c.layer[i];
    TACORL_REQUIRE(l.in == in, "mlp_chain: layer %d expects %d inputs, gets %d", i, l.in, in);
    TACORL_REQUIRE(l.in % 4 == 0 && l.in <= MC_MAXW && l.out >= 1 && l.out <= MC_MAXW && l.n[0] + l.n[1] + l.n[2] == l.out,
                   "mlp_chain: layer %d has unsupported dims (in %d, out %d)", i, l.in, l.out);
    TACORL_REQUIRE(l.W[0] && (l.n[1] == 0 || l.W[1]) && (l.n[2] == 0 || l.W[2]), "mlp_chain: null weight pointer in layer %d", i);
    TACORL_REQUIRE(l.n[1] > 0 || l.n[2] == 0, "mlp_chain: layer %d uses segment 2 without segment 1", i);
    TACORL_REQUIRE(i == c.L - 1 || l.out % 4 == 0, "mlp_chain: hidden width %d must be a multiple of 4", l.out);
    in = l.out;
  }
  TACORL_REQUIRE(c.X[0] && (c.xin[1] == 0 || c.X[1]) && c.out && (c.L == 1 || c.Z), "mlp_chain: null tensor");
  if (bwd) TACORL_REQUIRE(c.dOut, "mlp_chain_bwd: null tensor");
  TACORL_REQUIRE(c.rows <= 256, "mlp_chain: up to 256 rows (got %d)", c.rows);
  return 0;
}

static int mc_launch(const void* kern, const McChain& c, size_t smem, cudaStream_t st) {
  const int row_blocks = cdiv(c.rows, MC_ROWS);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(MC_CTAS, row_blocks); cfg.blockDim = dim3(MC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = MC_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  void* args[] = {(void*)&c};
  TACORL_CHECK_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
  note_launch();
  return 0;
}

}  // namespace tacorl

using namespace tacorl;

extern "C" {

/* One descriptor per layer, plain C (mirrors McLayer): see include/tacorl_b200.h */
static int mc_build(McChain& c, int L, int rows, const tacorl_mlp_layer* layers, const float* x0, int xin0, long long ldx0,
                    const float* x1, int xin1, long long ldx1, float* z, long long ldz, float* out, long long ldo) {
  memset(&c, 0, sizeof(c));
  TACORL_REQUIRE(L >= 1 && L <= MC_MAXL && layers, "mlp_chain: 1..%d layers supported", MC_MAXL);
  c.L = L; c.rows = rows;
  int zoff = 0;
  long long poff = 0;
  for (int i = 0; i < L; ++i) {
    McLayer& l = c.layer[i];
    const tacorl_mlp_layer& s = layers[i];
    l.W[0] = s.W0; l.W[1] = s.W1; l.W[2] = s.W2; l.b[0] = s.b0; l.b[1] = s.b1; l.b[2] = s.b2;
    l.n[0] = s.n0; l.n[1] = s.n1; l.n[2] = s.n2;
    l.in = s.in; l.out = s.n0 + s.n1 + s.n2; l.act = i == L - 1 ? ACT_NONE : (s.act & 0xFF);
    l.zpost = (s.act & TACORL_MLP_SAVE_ACTIVATED) ? 1 : 0;
    TACORL_REQUIRE(!l.zpost || l.act == ACT_RELU || l.act == ACT_NONE,
                   "mlp_chain: TACORL_MLP_SAVE_ACTIVATED needs a ReLU layer (layer %d)", i);
    l.dW[0] = s.dW0; l.dW[1] = s.dW1; l.dW[2] = s.dW2; l.db[0] = s.db0; l.db[1] = s.db1; l.db[2] = s.db2;
    l.part_off = poff; poff += (long long)l.out * (l.in + 1);
    c.zoff[i] = zoff; if (i < L - 1) zoff += l.out;
  }
  c.part_stride = poff;
  c.X[0] = x0; c.xin[0] = xin0; c.ldx[0] = ldx0; c.X[1] = x1; c.xin[1] = xin1; c.ldx[1] = ldx1;
  c.Z = z; c.ldz = ldz; c.out = out; c.ldo = ldo;
  TACORL_REQUIRE(L == 1 || ldz >= zoff, "mlp_chain: pre-activation buffer pitch %lld < %d", ldz, zoff);
  return 0;
}

size_t tacorl_mlp_chain_ws_bytes(int L, int rows, const tacorl_mlp_layer* layers) {
  size_t slab = 0;
  for (int i = 0; i < L; ++i) slab += (size_t)(layers[i].n0 + layers[i].n1 + layers[i].n2) * (layers[i].in + 1);
  const size_t rb = (size_t)cdiv(rows, MC_ROWS);
  return 2 * (size_t)rows * MC_MAXW * 4 + (rb > 1 ? rb * slab * 4 : 0) + 1024;
}

int tacorl_mlp_chain_fwd(int L, int rows, const tacorl_mlp_layer* layers, const float* x0, int xin0, long long ldx0,
                         const float* x1, int xin1, long long ldx1, float* z, long long ldz, float* out, long long ldo,
                         void* stream) {
  McChain c;
  int rc;
  if ((rc = mc_build(c, L, rows, layers, x0, xin0, ldx0, x1, xin1, ldx1, z, ldz, out, ldo))) return rc;
  if ((rc = mc_check(c, false))) return rc;
  const size_t smem = (size_t)(2 * MC_ROWS + 64) * MC_LD * 4;
  static bool configured = false;
  if (!configured) {
    TACORL_CHECK_CUDA(cudaFuncSetAttribute(mlp_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  return mc_launch((const void*)mlp_chain_fwd_kernel, c, smem, (cudaStream_t)stream);
}

int tacorl_mlp_chain_bwd(int L, int rows, const tacorl_mlp_layer* layers, const float* x0, int xin0, long long ldx0,
                         const float* x1, int xin1, long long ldx1, const float* z, long long ldz, const float* d_out,
                         long long lddo, float* dx0, long long lddx0, float* dx1, long long lddx1, void* ws,
                         size_t ws_bytes, void* stream) {
  McChain c;
  int rc;
  float dummy_out;
  if ((rc = mc_build(c, L, rows, layers, x0, xin0, ldx0, x1, xin1, ldx1, const_cast<float*>(z), ldz, &dummy_out, 0))) return rc;
  c.dOut = d_out; c.lddo = lddo; c.dX[0] = dx0; c.lddx[0] = lddx0; c.dX[1] = dx1; c.lddx[1] = lddx1;
  TACORL_REQUIRE(ws && ws_bytes >= tacorl_mlp_chain_ws_bytes(L, rows, layers), "mlp_chain_bwd: workspace too small");
  Arena ar(ws, ws_bytes);
  c.dybuf[0] = ar.take<float>((size_t)rows * MC_MAXW);
  c.dybuf[1] = ar.take<float>((size_t)rows * MC_MAXW);
  const int row_blocks = cdiv(rows, MC_ROWS);
  bool any_dw = false;
  for (int i = 0; i < L; ++i) any_dw |= c.layer[i].dW[0] != nullptr;
  if (row_blocks > 1 && any_dw) {
    c.part = ar.take<float>((size_t)row_blocks * c.part_stride);
    TACORL_REQUIRE(c.part, "mlp_chain_bwd: workspace too small for the partial slabs");
  }
  if ((rc = mc_check(c, true))) return rc;
  const size_t smem = (size_t)(3 * MC_ROWS + 32) * MC_LD * 4;
  static bool configured = false;
  if (!configured) {
    TACORL_CHECK_CUDA(cudaFuncSetAttribute(mlp_chain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  if ((rc = mc_launch((const void*)mlp_chain_bwd_kernel, c, smem, (cudaStream_t)stream))) return rc;
  if (c.part) {
    mlp_chain_reduce_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(c, row_blocks);
    TACORL_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
