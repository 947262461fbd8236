// Generic C-ABI entry points: error reporting, GEMM / linear-layer pieces.
#include "common.cuh"
#include "internal.h"
#include "../../include/tacorl_b200.h"

using namespace tacorl;

namespace tacorl {
static int g_sm_reserve = -1;
int persistent_ctas() {
  if (g_sm_reserve < 0) { const char* e = getenv("TACORL_SM_RESERVE"); g_sm_reserve = e ? atoi(e) : 0; }
  const int n = 148 - g_sm_reserve;
  return n < 8 ? 8 : n;
}
void set_sm_reserve(int n) { g_sm_reserve = n < 0 ? 0 : (n > 140 ? 140 : n); }
}  // namespace tacorl

extern "C" {

const char* tacorl_last_error(void) { return last_error(); }

int tacorl_abi_version(void) { return TACORL_B200_ABI_VERSION; }

unsigned long long tacorl_launch_count(void) { return launch_count(); }

unsigned tacorl_rnn_seq_timeouts(void) { return rnn_seq_timeouts(); }

int tacorl_set_sm_reserve(int n) { set_sm_reserve(n); return persistent_ctas(); }

int tacorl_rnn_seq_enable(int on) { const int was = rnn_seq_mode(); rnn_seq_set_enabled(on); return was; }

int tacorl_gemm_ex(int transA, int transB, int M, int N, int K, float alpha, const float* A, long long lda,
                   const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, int act,
                   float* Cpre, long long ldpre, const void* A_bf16, const void* B_bf16, void* ws, size_t ws_bytes,
                   int prec, void* stream) {
  TACORL_REQUIRE(prec == PREC_F32 || prec == PREC_BF16, "gemm: unknown precision %d", prec);
  GemmArgs g;
  g.transA = transA; g.transB = transB; g.M = M; g.N = N; g.K = K; g.alpha = alpha;
  g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.beta = beta; g.C = C; g.ldc = ldc; g.bias = bias;
  g.act = act; g.Cpre = Cpre; g.ldpre = ldpre; g.split_k = 0; g.A_bf16 = A_bf16; g.B_bf16 = B_bf16;
  return gemm_any(prec, g, (float*)ws, ws_bytes, (cudaStream_t)stream);
}

int tacorl_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, long long lda,
                const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, int act,
                float* Cpre, long long ldpre, void* ws, size_t ws_bytes, int prec, void* stream) {
  return tacorl_gemm_ex(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, act, Cpre, ldpre, nullptr,
                        nullptr, ws, ws_bytes, prec, stream);
}

int tacorl_colsum(int M, int N, const float* X, long long ldx, float* out, int accumulate, void* stream) {
  TACORL_REQUIRE(X && out, "colsum: null pointer");
  return colsum_f32(M, N, X, ldx, out, accumulate, (cudaStream_t)stream);
}

int tacorl_act_bwd(int act, long long n, const float* dY, const float* y_or_pre, float* dZ, void* stream) {
  TACORL_REQUIRE(dY && y_or_pre && dZ, "act_bwd: null pointer");
  return act_bwd_f32(act, n, dY, y_or_pre, dZ, (cudaStream_t)stream);
}

}  // extern "C"
