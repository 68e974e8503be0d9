// Shared host/device helpers for libsrk (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/srk.h"

namespace srk {

void set_error(const char* fmt, ...);

#define SRK_FAIL(...)            \
  do {                           \
    srk::set_error(__VA_ARGS__); \
    return 1;                    \
  } while (0)

#define SRK_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) SRK_FAIL(__VA_ARGS__); \
  } while (0)

#define SRK_CUDA_LAUNCH_CHECK(name)                                              \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) SRK_FAIL("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// Strided view of a 4-D tensor: element (n, c, y, x) lives at p[off + n*sn + y*sh + x*sw + c*sc].
// Covers both the NCHW fp32 image layout and the zero-bordered channels-last activation layout.
struct View {
  void* p;
  long long off, sn, sh, sw, sc;
  int N, C, H, W;
  int dtype;
  int is_act;
};

inline View make_view(const srk_tensor* t) {
  View v;
  v.p = t->data;
  v.N = t->n; v.C = t->c; v.H = t->h; v.W = t->w;
  v.dtype = t->dtype;
  v.is_act = (t->layout == SRK_LAYOUT_ACT);
  if (v.is_act) {
    long long Hp = t->h + 2, Wp = t->w + 2, C = t->c;
    v.sn = Hp * Wp * C; v.sh = Wp * C; v.sw = C; v.sc = 1; v.off = (Wp + 1) * C;
  } else {
    long long H = t->h, W = t->w, C = t->c;
    v.sn = C * H * W; v.sc = H * W; v.sh = W; v.sw = 1; v.off = 0;
  }
  return v;
}

inline bool same_geometry(const srk_tensor* a, const srk_tensor* b) {
  return a->n == b->n && a->c == b->c && a->h == b->h && a->w == b->w;
}
inline int64_t act_elems(const srk_tensor* t) {
  return (int64_t)t->n * (t->h + 2) * (t->w + 2) * t->c;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum (blockDim.x multiple of 32, <= 1024); result valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem32[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}
__device__ __forceinline__ double block_sum_d(double v, double* smem32) {
  v = warp_sum_d(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? smem32[lane] : 0.0;
    r = warp_sum_d(r);
  }
  return r;
}

constexpr int kNumSMs = 148;  // B200

}  // namespace srk
