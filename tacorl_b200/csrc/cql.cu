// CQL scalar-loss kernels: every Bellman / conservative / Lagrange / actor / alpha term of one
// TACO-RL update, value and gradient, in two single-CTA launches (the tensors are (B,13) at most).
// Reference: /root/reference/src/tacorl/modules/cql/cql_offline_lightning.py
//   compute_critic_loss :284-314, compute_conservative_loss :316-406,
//   compute_actor_and_alpha_loss :439-468.
#include "common.cuh"
#include "internal.h"
#include "../../include/tacorl_b200.h"

namespace tacorl {

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// q_all_i layout: [data (B) | rand (n,B) | curr (n,B) | next (n,B)]  (sample-major like the reference's
// (n*bs,1)->(n,bs) views, :255-257).  scalars out: see tacorl_b200.h TACORL_CQL_*.
__global__ void cql_critic_loss_kernel(int B, int n, const float* __restrict__ q1, const float* __restrict__ q2,
                                       const float* __restrict__ lp_curr, const float* __restrict__ lp_next,
                                       const float* __restrict__ tq1, const float* __restrict__ tq2,
                                       const float* __restrict__ reward, const float* __restrict__ done,
                                       const float* __restrict__ log_alpha_prime, float rand_density,
                                       float discount, float reward_scale, float gap, float cw, float temp,
                                       int with_lagrange, float* __restrict__ scalars, float* __restrict__ dq1,
                                       float* __restrict__ dq2, float* __restrict__ d_lap) {
  __shared__ float red[32];
  const float invB = 1.f / (float)B, inv_t = 1.f / temp;
  float ap = 1.f, dap = 0.f;
  if (with_lagrange) {
    const float e = expf(*log_alpha_prime);
    ap = fminf(fmaxf(e, 0.f), 1e6f);
    dap = (e >= 0.f && e <= 1e6f) ? e : 0.f;
  }
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float y = reward_scale * reward[b] + (1.f - done[b]) * discount * fminf(tq1[b], tq2[b]);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float* q = c ? q2 : q1;
      float* dq = c ? dq2 : dq1;
      const float qd = q[b];
      const float e = qd - y;
      acc[c] += e * e;                       // bellman
      acc[4 + c] += qd;                      // q_data
      float mx = -INFINITY;
      for (int j = 0; j < n; ++j) {
        const float r = (q[B + j * B + b] - rand_density) * inv_t;
        const float cu = (q[B + (n + j) * B + b] - lp_curr[j * B + b]) * inv_t;
        const float ne = (q[B + (2 * n + j) * B + b] - lp_next[j * B + b]) * inv_t;
        mx = fmaxf(mx, fmaxf(r, fmaxf(cu, ne)));
        acc[6 + c] += q[B + j * B + b];          // q_random
        acc[8 + c] += q[B + (n + j) * B + b];    // q_policy (current obs)
      }
      float se = 0.f;
      for (int j = 0; j < n; ++j) {
        se += expf((q[B + j * B + b] - rand_density) * inv_t - mx);
        se += expf((q[B + (n + j) * B + b] - lp_curr[j * B + b]) * inv_t - mx);
        se += expf((q[B + (2 * n + j) * B + b] - lp_next[j * B + b]) * inv_t - mx);
      }
      const float lse = mx + logf(se);
      acc[2 + c] += lse;
      if (dq) {
        const float gw = ap * cw * invB;
        dq[b] = 2.f * e * invB - gw;
        for (int j = 0; j < n; ++j) {
          dq[B + j * B + b] = gw * expf((q[B + j * B + b] - rand_density) * inv_t - lse);
          dq[B + (n + j) * B + b] = gw * expf((q[B + (n + j) * B + b] - lp_curr[j * B + b]) * inv_t - lse);
          dq[B + (2 * n + j) * B + b] = gw * expf((q[B + (2 * n + j) * B + b] - lp_next[j * B + b]) * inv_t - lse);
        }
      }
    }
  }
  float tot[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) tot[i] = block_sum(acc[i], red);
  if (threadIdx.x == 0) {
    const float bell1 = tot[0] * invB, bell2 = tot[1] * invB;
    const float raw1 = tot[2] * invB * cw * temp - tot[4] * invB * cw;
    const float raw2 = tot[3] * invB * cw * temp - tot[5] * invB * cw;
    const float cons1 = with_lagrange ? ap * (raw1 - gap) : raw1;
    const float cons2 = with_lagrange ? ap * (raw2 - gap) : raw2;
    scalars[0] = bell1; scalars[1] = bell2; scalars[2] = cons1; scalars[3] = cons2;
    scalars[4] = ap; scalars[5] = (-cons1 - cons2) * 0.5f;
    scalars[6] = bell1 + cons1; scalars[7] = bell2 + cons2;
    scalars[8] = tot[4] * invB; scalars[9] = tot[6] * invB / (float)n; scalars[10] = tot[8] * invB / (float)n;
    scalars[11] = tot[5] * invB; scalars[12] = tot[7] * invB / (float)n; scalars[13] = tot[9] * invB / (float)n;
    if (d_lap) d_lap[0] = with_lagrange ? -0.5f * ((raw1 - gap) + (raw2 - gap)) * dap : 0.f;
  }
}

// mode 0: alpha_loss = -mean(log_alpha * (log_pi + target_entropy))            -> out[0], d_log_alpha
// mode 1: actor_loss (BC)  = mean(alpha*log_pi - plp)                           -> out[0], out[1]=alpha
// mode 2: actor_loss (Q)   = mean(alpha*log_pi - min(q1,q2))
__global__ void cql_actor_loss_kernel(int mode, int B, const float* __restrict__ log_pi,
                                      const float* __restrict__ a, const float* __restrict__ b2,
                                      const float* __restrict__ log_alpha, float target_entropy,
                                      float* __restrict__ out, float* __restrict__ d_log_alpha,
                                      float* __restrict__ d_log_pi, float* __restrict__ da, float* __restrict__ db) {
  __shared__ float red[32];
  const float invB = 1.f / (float)B;
  const float la = *log_alpha, alpha = expf(la);
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float lp = log_pi[i];
    if (mode == 0) {
      s += lp + target_entropy;
    } else if (mode == 1) {
      s += alpha * lp - a[i];
      if (d_log_pi) { d_log_pi[i] = alpha * invB; da[i] = -invB; }
    } else {
      const float x = a[i], y = b2[i];
      s += alpha * lp - fminf(x, y);
      if (d_log_pi) {
        d_log_pi[i] = alpha * invB;
        da[i] = x < y ? -invB : (x == y ? -0.5f * invB : 0.f);
        db[i] = y < x ? -invB : (x == y ? -0.5f * invB : 0.f);
      }
    }
  }
  const float t = block_sum(s, red);
  if (threadIdx.x == 0) {
    if (mode == 0) { out[0] = -la * t * invB; if (d_log_alpha) d_log_alpha[0] = -t * invB; }
    else { out[0] = t * invB; out[1] = alpha; }
  }
}

}  // namespace tacorl

using namespace tacorl;

extern "C" {

int tacorl_cql_critic_loss(int B, int n, const float* q1_all, const float* q2_all, const float* lp_curr,
                           const float* lp_next, const float* tq1, const float* tq2, const float* reward,
                           const float* done, const float* log_alpha_prime, float rand_density, float discount,
                           float reward_scale, float gap, float conservative_weight, float temp,
                           int with_lagrange, float* scalars, float* dq1_all, float* dq2_all,
                           float* d_log_alpha_prime, void* stream) {
  TACORL_REQUIRE(B > 0 && n > 0, "cql_critic_loss: empty batch");
  TACORL_REQUIRE(q1_all && q2_all && lp_curr && lp_next && tq1 && tq2 && reward && done && scalars,
                 "cql_critic_loss: null pointer");
  TACORL_REQUIRE(!with_lagrange || log_alpha_prime, "cql_critic_loss: lagrange needs log_alpha_prime");
  cql_critic_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(
      B, n, q1_all, q2_all, lp_curr, lp_next, tq1, tq2, reward, done, log_alpha_prime, rand_density, discount,
      reward_scale, gap, conservative_weight, temp, with_lagrange, scalars, dq1_all, dq2_all, d_log_alpha_prime);
  TACORL_LAUNCH_CHECK();
  return 0;
}

int tacorl_cql_actor_loss(int mode, int B, const float* log_pi, const float* a, const float* b,
                          const float* log_alpha, float target_entropy, float* out, float* d_log_alpha,
                          float* d_log_pi, float* da, float* db, void* stream) {
  TACORL_REQUIRE(B > 0 && log_pi && log_alpha && out, "cql_actor_loss: bad arguments");
  TACORL_REQUIRE(mode >= 0 && mode <= 2, "cql_actor_loss: bad mode");
  cql_actor_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(mode, B, log_pi, a, b, log_alpha, target_entropy, out,
                                                            d_log_alpha, d_log_pi, da, db);
  TACORL_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
