// Internal (non-ABI) declarations shared between the .cu translation units.
#pragma once
#include "common.cuh"

namespace tacorl {

const char* last_error();
unsigned long long launch_count();

int colsum_tall_f32(long long M, int N, const float* X, float* out, int accumulate, float* ws,
                    size_t ws_bytes, cudaStream_t st);

int im2col_f32(const float* x, long long sn, long long sc, long long sh, long long sw, int C, int KH,
               int KW, int stride, int OH, int OW, int nframes, float* col, cudaStream_t st, int korder);
int col2im_f32(const float* dcol, int C, int H, int W, int KH, int KW, int stride, int OH, int OW,
               int nframes, const float* ymask, float* dx, cudaStream_t st);
int permute_conv_weight_f32(const float* src, float* dst, int OC, int C, int KH, int KW, int dir,
                            cudaStream_t st);
int softargmax_fwd_f32(const float* y, int N, int OH, int OW, int C, const float* temperature,
                       float* feat, float* smax, float* ssum, cudaStream_t st);
int softargmax_bwd_f32(const float* y, int N, int OH, int OW, int C, const float* temperature,
                       const float* feat, const float* smax, const float* ssum, const float* dfeat,
                       float* dy, float* dtau_part, cudaStream_t st);

// simple bump allocator over a caller-provided workspace (256-byte aligned slices)
struct Arena {
  char* base; size_t cap; size_t off = 0;
  Arena(void* p, size_t n) : base((char*)p), cap(n) {}
  template <typename T> T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    if (off + bytes > cap) return nullptr;
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
  size_t left() const { return cap - off; }
};

}  // namespace tacorl
