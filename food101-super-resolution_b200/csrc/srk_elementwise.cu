// HBM-bound kernels on the zero-bordered channels-last activation layout: BatchNorm (training and
// eval), PReLU/ReLU backward (+ pixel-unshuffle), squeeze-excite gate, layout conversion, residual
// add, bicubic upsampling.  All are coalesced along the channel dimension with 16-byte vectors when
// C % 8 == 0; reductions go warp-shuffle -> shared memory -> one atomic per block and channel.
#include "srk_common.cuh"

#include <cstdlib>

namespace srk {

// ---- vector helpers: VEC contiguous channels -------------------------------------------------
template <typename T, int VEC> struct Vec;
template <> struct Vec<float, 1> {
  typedef float Raw;
  static __device__ __forceinline__ Raw ldraw(const float* p) { return p[0]; }
  static __device__ __forceinline__ void unpack(const Raw& r, float* o) { o[0] = r; }
  static __device__ __forceinline__ void ld(const float* p, float* o) { o[0] = p[0]; }
  static __device__ __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
};
template <> struct Vec<__nv_bfloat16, 1> {
  typedef __nv_bfloat16 Raw;
  static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) { return p[0]; }
  static __device__ __forceinline__ void unpack(const Raw& r, float* o) { o[0] = __bfloat162float(r); }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) { o[0] = __bfloat162float(p[0]); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) { p[0] = __float2bfloat16_rn(v[0]); }
};
template <> struct Vec<float, 8> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw ldraw(const float* p) {
    Raw r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float* o) {
    o[0] = r.a.x; o[1] = r.a.y; o[2] = r.a.z; o[3] = r.a.w; o[4] = r.b.x; o[5] = r.b.y; o[6] = r.b.z; o[7] = r.b.w;
  }
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec<__nv_bfloat16, 8> {
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& r, float* o) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* o) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* v) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

struct Geo {
  int N, Hp, Wp, C;
  long long pixels;  // N*Hp*Wp
};
inline Geo geo_of(const srk_tensor* t) {
  Geo g; g.N = t->n; g.Hp = t->h + 2; g.Wp = t->w + 2; g.C = t->c;
  g.pixels = (long long)g.N * g.Hp * g.Wp;
  return g;
}
__device__ __forceinline__ bool is_border(const Geo& g, long long q) {
  int xx = (int)(q % g.Wp);
  int yy = (int)((q / g.Wp) % g.Hp);
  return xx == 0 || xx == g.Wp - 1 || yy == 0 || yy == g.Hp - 1;
}

// Pixel-tiled iteration: a block owns `ppb` consecutive padded pixels; thread t handles channel
// vector cv = t % CV for pixel rows t / CV, t / CV + rows, ...  The walker tracks the padded (x, y) of its
// pixel incrementally: one 64-bit division per thread instead of two per pixel.
struct PixWalk {
  long long q, q_end;
  int xx, yy, rows, Wp, Hp;
  __device__ __forceinline__ PixWalk(const Geo& g, int ppb, int rows_, int prow) {
    const long long q_begin = (long long)blockIdx.x * ppb;
    q_end = q_begin + ppb;
    if (q_end > g.pixels) q_end = g.pixels;
    q = q_begin + prow;
    rows = rows_; Wp = g.Wp; Hp = g.Hp;
    xx = (int)(q % g.Wp);
    yy = (int)((q / g.Wp) % g.Hp);
  }
  __device__ __forceinline__ bool valid() const { return q < q_end; }
  __device__ __forceinline__ bool border() const { return xx == 0 || xx == Wp - 1 || yy == 0 || yy == Hp - 1; }
  __device__ __forceinline__ void next() {
    q += rows; xx += rows;
    while (xx >= Wp) { xx -= Wp; if (++yy == Hp) yy = 0; }
  }
};
#define SRK_PIXEL_LOOP(g, VEC)                                                      \
  const int CV = (g).C / (VEC);                                                     \
  const int rows = blockDim.x / CV;                                                 \
  const int cv = threadIdx.x % CV;                                                  \
  const int prow = threadIdx.x / CV;                                                \
  _Pragma("unroll 4")                                                               \
  for (PixWalk pw((g), ppb, rows, prow); prow < rows && pw.valid(); pw.next())

// Batched form of the same ownership for the bandwidth-critical passes: EW_U pixels per thread and pass, all of
// their 16-byte loads issued before the first use (memory-level parallelism per thread instead of per SM only:
// the one-pixel-at-a-time loops above reach ~3.7 TB/s on cold data, the batched ones are the HBM-bound ones).
constexpr int EW_U = 4;

// Same ownership without the (x, y) bookkeeping, for reductions whose border terms vanish.
#define SRK_FLAT_LOOP(g, VEC)                                                       \
  const int CV = (g).C / (VEC);                                                     \
  const int rows = blockDim.x / CV;                                                 \
  const int cv = threadIdx.x % CV;                                                  \
  const int prow = threadIdx.x / CV;                                                \
  const long long q_begin__ = (long long)blockIdx.x * ppb;                          \
  const long long q_end__ = q_begin__ + ppb < (g).pixels ? q_begin__ + ppb : (g).pixels; \
  _Pragma("unroll 4")                                                               \
  for (long long q = q_begin__ + prow; prow < rows && q < q_end__; q += rows)

// Reductions end in one atomic per block and channel: fewer, longer blocks win (measured on B200 at the C2 layer
// shape: bn_bwd_reduce 17.0 us at 2 blocks / SM, 23.0 at 8, 31.5 at 16); the map kernels want 4-8 blocks / SM.
inline void reduce_grid(const Geo& g, int& blocks, int& ppb) {
  static int per_sm = 0;
  if (!per_sm) { const char* e = getenv("SRK_EW_REDUCE_PER_SM"); per_sm = e ? atoi(e) : 2; if (per_sm < 1) per_sm = 2; }
  long long target = 148LL * per_sm;
  long long p = (g.pixels + target - 1) / target;
  if (p < 64) p = 64;
  ppb = (int)p;
  blocks = (int)((g.pixels + ppb - 1) / ppb);
}

// The BatchNorm apply passes (forward and backward) read two tensors and write one: 4 blocks / SM measured best
// (bn_apply + residual 20.6 vs 21.7 us, bn_bwd_apply 21.0 vs 23.7 us at 4 vs 8 blocks / SM, C2 layer shape).
// resident: blocks of the kernel that fit on an SM (0 = not queried): a grid of 4 blocks / SM whose blocks only fit three
// at a time runs as 1.33 waves (bn_bwd_apply at 79 registers: 33.5 us; sized to what is resident: 25.4 us).
inline void bn_map_grid(const Geo& g, int& blocks, int& ppb, int resident = 0) {
  static int per_sm_env = 0;
  if (!per_sm_env) { const char* e = getenv("SRK_EW_BN_PER_SM"); per_sm_env = e ? atoi(e) : 4; if (per_sm_env < 1) per_sm_env = 4; }
  const int per_sm = (resident > 0 && resident < per_sm_env) ? resident : per_sm_env;
  long long target = 148LL * per_sm;
  long long p = (g.pixels + target - 1) / target;
  if (p < 64) p = 64;
  ppb = (int)p;
  blocks = (int)((g.pixels + ppb - 1) / ppb);
}

inline void pixel_grid(const Geo& g, int& blocks, int& ppb) {
  // ~4 blocks per SM worth of work, at least 64 pixels per block
  static int per_sm = 0;
  if (!per_sm) { const char* e = getenv("SRK_EW_BLOCKS_PER_SM"); per_sm = e ? atoi(e) : 8; if (per_sm < 1) per_sm = 8; }
  long long target = 148LL * per_sm;
  long long p = (g.pixels + target - 1) / target;
  if (p < 64) p = 64;
  ppb = (int)p;
  blocks = (int)((g.pixels + ppb - 1) / ppb);
}

// Reduce per-thread VEC partials across the rows of a block into dst[C] (shared memory), in a fixed order.
// When the channel-vector count divides the warp size the lanes that share a channel vector are first folded
// with shuffles, so only one row per warp goes through shared memory.  Ends with a __syncthreads(): dst is
// complete and visible, `smem` may be reused.
template <int VEC>
__device__ __forceinline__ void block_channel_sum(float* part, float* smem, int CV, int rows,
                                                  int cv, int prow, float* dst) {
  const int C = CV * VEC;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  __syncthreads();
  if ((32 % CV) == 0 && (blockDim.x % CV) == 0) {
    for (int o = CV; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
    }
    if (lane < CV) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) smem[warp * C + cv * VEC + j] = part[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < nwarps; ++r) s += smem[r * C + c];
      dst[c] = s;
    }
    __syncthreads();
    return;
  }
  // smem: [rows][CV*VEC]
  if (prow < rows) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) smem[prow * C + cv * VEC + j] = part[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += smem[r * C + c];
    dst[c] = s;
  }
  __syncthreads();
}

// Dynamic shared memory of the reducing kernels: [row scratch: red_rows_floats][vals: NV (padded to 4)][fold
// scratch: blockDim float4].  Cross-block stage: ordered_fold (srk_common.cuh).
struct RedArgs {
  unsigned* tickets;   // ticket base of this launch (one per reduction slot)
  float* partials;     // partial rows of this launch
};
__device__ __forceinline__ void sync_block() { __syncthreads(); }

// ---- BatchNorm ---------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ y, Geo g, int ppb,
                                                       float* __restrict__ sums /* [2][C]: sum | sum of squares */,
                                                       RedArgs ra, int row_floats) {
  pdl_wait();      // the inputs come from the previous kernel of the stream (see launch_dep)
  pdl_trigger();
  extern __shared__ __align__(16) float smem[];
  float* vals = smem + row_floats;
  float4* scratch = reinterpret_cast<float4*>(vals + ((2 * g.C + 3) & ~3));
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  // the border is zero by invariant, so it contributes nothing: plain strided walk over all padded pixels
  SRK_FLAT_LOOP(g, VEC) {
    float v[VEC];
    Vec<T, VEC>::ld(y + q * g.C + cv * VEC, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
  }
  block_channel_sum<VEC>(s1, smem, CV, rows, cv, prow, vals);
  block_channel_sum<VEC>(s2, smem, CV, rows, cv, prow, vals + g.C);
  ordered_fold(vals, 2 * g.C, ra.tickets, gridDim.x, blockIdx.x, ra.partials, scratch, threadIdx.x, blockDim.x,
               sync_block, [&](int i, float v) { sums[i] = v; });
}

__global__ void bn_finalize_kernel(const float* sum, const float* sumsq, int C, double count, float eps,
                                   float momentum, float* running_mean, float* running_var,
                                   long long* nbt, float* mean, float* invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    double m = (double)sum[c] / count;
    double var = (double)sumsq[c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
      double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
  if (c == 0 && nbt != nullptr) nbt[0] += 1;
}

__global__ void bn_eval_params_kernel(const float* rm, const float* rv, int C, float eps, float* mean,
                                      float* invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = rm[c]; invstd[c] = rsqrtf(rv[c] + eps); }
}

// Training-mode statistics folded into the apply pass (one launch less on the critical path of every BN layer):
// when `fin.sum` is set, every block derives mean / invstd from the per-channel sums itself (double precision, as
// bn_finalize_kernel does) and block 0 also publishes them and updates the running statistics.
struct BnFinalize {
  const float* sum; const float* sumsq;
  double count; float eps, momentum;
  float* running_mean; float* running_var; long long* nbt;
  float* mean_out; float* invstd_out;
  unsigned long long* acc;   // the sums as an exact integer accumulator (srk_common.cuh) instead of sum / sumsq
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ y, Geo g, int ppb,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ alpha_p, const T* __restrict__ res,
    T* __restrict__ out, const BnFinalize fin) {
  pdl_wait();      // the inputs come from the previous kernel of the stream (see launch_dep)
  pdl_trigger();
  extern __shared__ float bn_smem[];   // fused statistics: [C] mean, [C] invstd
  const float alpha = alpha_p ? alpha_p[0] : 1.f;
  bool clear_acc = false;
  if (fin.sum || fin.acc) {
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) {
      const double m = (fin.acc ? acc_read(fin.acc, c) : (double)fin.sum[c]) / fin.count;
      double var = (fin.acc ? acc_read(fin.acc, g.C + c) : (double)fin.sumsq[c]) / fin.count - m * m;
      if (var < 0.0) var = 0.0;
      const float mf = (float)m, isf = (float)(1.0 / sqrt(var + (double)fin.eps));
      bn_smem[c] = mf; bn_smem[g.C + c] = isf;
      if (blockIdx.x == 0) {
        fin.mean_out[c] = mf; fin.invstd_out[c] = isf;
        if (fin.running_mean != nullptr) {
          const double unbiased = fin.count > 1.0 ? var * fin.count / (fin.count - 1.0) : var;
          fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * mf;
          fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unbiased;
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && fin.nbt != nullptr) fin.nbt[0] += 1;
    __syncthreads();
    mean = bn_smem; invstd = bn_smem + g.C;
    // every block has its statistics (the loads have returned: their values are in shared memory): take the ticket
    // now, look at it after the pixel loop - the block that drew the last one resets the accumulator
    if (fin.acc && threadIdx.x == 0) clear_acc = acc_ticket(fin.acc, gridDim.x);
  }
  // per-channel scale / shift of this thread's channel vector, hoisted out of the pixel loop
  float sc[VEC], sh[VEC];
  {
    const int c0 = (threadIdx.x % (g.C / VEC)) * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      sc[j] = gamma[c0 + j] * invstd[c0 + j];
      sh[j] = beta[c0 + j] - mean[c0 + j] * sc[j];
    }
  }
  SRK_PIXEL_LOOP(g, VEC) {
    const long long q = pw.q;
    float o[VEC];
    const long long e = q * g.C + cv * VEC;
    if (pw.border()) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    } else {
      float v[VEC];
      Vec<T, VEC>::ld(y + e, v);
      float r[VEC];
      if (res) Vec<T, VEC>::ld(res + e, r);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float b = fmaf(v[j], sc[j], sh[j]);
        o[j] = (alpha_p && b < 0.f) ? alpha * b : b;
      }
      if (res) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] += r[j];
      }
    }
    Vec<T, VEC>::st(out + e, o);
  }
  if (fin.acc) {
    if (__syncthreads_or(clear_acc)) acc_clear(fin.acc, threadIdx.x, blockDim.x);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(const T* __restrict__ dout,
    const T* __restrict__ y, Geo g, int ppb, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ alpha_p, float* __restrict__ dgamma, float* __restrict__ dbeta,
    float* __restrict__ dalpha, RedArgs ra, int row_floats) {
  pdl_wait();      // the inputs come from the previous kernel of the stream (see launch_dep)
  pdl_trigger();
  extern __shared__ __align__(16) float smem[];
  float* vals = smem + row_floats;
  float4* scratch = reinterpret_cast<float4*>(vals + ((2 * g.C + 1 + 3) & ~3));
  const float alpha = alpha_p ? alpha_p[0] : 1.f;
  float dg[VEC], db[VEC];
  float da = 0.f;
  // xhat = v * is - mi,  bn output b = v * sc + sh   (4 constants per channel)
  float is[VEC], mi[VEC], sc[VEC], sh[VEC];
  {
    const int c0 = (threadIdx.x % (g.C / VEC)) * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      dg[j] = 0.f; db[j] = 0.f;
      is[j] = invstd[c0 + j]; mi[j] = mean[c0 + j] * is[j];
      sc[j] = gamma[c0 + j] * is[j]; sh[j] = beta[c0 + j] - mean[c0 + j] * sc[j];
    }
  }
  // dout is zero on the border (layout invariant), so border pixels add exactly zero to every sum
  const int CV = g.C / VEC, rows = blockDim.x / CV, cv = threadIdx.x % CV, prow = threadIdx.x / CV;
  typedef typename Vec<T, VEC>::Raw Raw;
  const long long q_begin = (long long)blockIdx.x * ppb;
  const long long q_end = q_begin + ppb < g.pixels ? q_begin + ppb : g.pixels;
  for (long long q = q_begin + prow; prow < rows && q < q_end; q += (long long)EW_U * rows) {
    Raw rv[EW_U], rd[EW_U];
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      const long long qu = q + (long long)u * rows;
      if (qu < q_end) {
        const long long e = qu * g.C + cv * VEC;
        rv[u] = Vec<T, VEC>::ldraw(y + e);
        rd[u] = Vec<T, VEC>::ldraw(dout + e);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_U; ++u) {
      if (q + (long long)u * rows >= q_end) continue;
      float v[VEC], d[VEC];
      Vec<T, VEC>::unpack(rv[u], v);
      Vec<T, VEC>::unpack(rd[u], d);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float xh = fmaf(v[j], is[j], -mi[j]);
        float gd = d[j];
        if (alpha_p) {
          float b = fmaf(v[j], sc[j], sh[j]);
          if (b < 0.f) { da = fmaf(gd, b, da); gd *= alpha; }
        }
        dg[j] = fmaf(gd, xh, dg[j]);
        db[j] += gd;
      }
    }
  }
  block_channel_sum<VEC>(dg, smem, CV, rows, cv, prow, vals);
  block_channel_sum<VEC>(db, smem, CV, rows, cv, prow, vals + g.C);
  {
    __shared__ float red[32];
    const float t = block_sum(da, red);
    if (threadIdx.x == 0) vals[2 * g.C] = t;
    __syncthreads();
  }
  const int C = g.C;
  ordered_fold(vals, 2 * C + 1, ra.tickets, gridDim.x, blockIdx.x, ra.partials, scratch, threadIdx.x, blockDim.x,
               sync_block, [&](int i, float v) {
                 if (i < C) dgamma[i] = v;
                 else if (i < 2 * C) dbeta[i - C] = v;
                 else if (alpha_p && dalpha) dalpha[0] = v;
               });
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ dout,
    const T* __restrict__ y, Geo g, int ppb, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ alpha_p, const float* __restrict__ dgamma, const float* __restrict__ dbeta,
    float inv_count, int batch_stats, T* __restrict__ dy, float* __restrict__ dgamma_out,
    unsigned long long* acc, float* __restrict__ dbeta_out, float* __restrict__ dalpha_out) {
  pdl_wait();      // the inputs come from the previous kernel of the stream (see launch_dep)
  pdl_trigger();
  const float alpha = alpha_p ? alpha_p[0] : 1.f;
  // raw mode (dgamma_out != null): the reduction came out of a conv epilogue as (sum g, sum g*z) in (dbeta, dgamma);
  // sum g*xhat = invstd * (sum g*z - mean * sum g).  Block 0 publishes the finished dgamma.
  // With `acc` the raw sums arrive as an exact integer accumulator [sum g C | sum g*z C | dalpha] (srk_common.cuh):
  // converted once per block through shared memory; block 0 also publishes dbeta = sum g and dalpha.
  __shared__ float racc[kAccNV];
  bool clear_acc = false;
  if (acc != nullptr) {
    for (int i = threadIdx.x; i < 2 * g.C + 1; i += blockDim.x) racc[i] = (float)acc_read(acc, i);
    __syncthreads();
    if (threadIdx.x == 0) clear_acc = acc_ticket(acc, gridDim.x);   // looked at after the pixel loop
    dbeta = racc; dgamma = racc + g.C;
    if (blockIdx.x == 0) {
      for (int c = threadIdx.x; c < g.C; c += blockDim.x) dbeta_out[c] = racc[c];
      if (threadIdx.x == 0 && dalpha_out != nullptr) dalpha_out[0] = racc[2 * g.C];
    }
  }
  if (dgamma_out != nullptr && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < g.C; c += blockDim.x) dgamma_out[c] = invstd[c] * (dgamma[c] - mean[c] * dbeta[c]);
  }
  // dy = a1 * g' + a2 * v + a3 with g' = PReLU-masked dout; mask from b = v * a1 + sh   (4 constants / channel)
  float a1[VEC], a2[VEC], a3[VEC], sh[VEC];
  {
    const int c0 = (threadIdx.x % (g.C / VEC)) * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float is = invstd[c0 + j], mu = mean[c0 + j], ga = gamma[c0 + j];
      const float dg = dgamma_out != nullptr ? is * (dgamma[c0 + j] - mu * dbeta[c0 + j]) : dgamma[c0 + j];
      const float k1 = batch_stats ? dbeta[c0 + j] * inv_count : 0.f;
      const float k2 = batch_stats ? dg * inv_count : 0.f;
      a1[j] = ga * is;
      a2[j] = -a1[j] * is * k2;
      a3[j] = -a1[j] * k1 - a2[j] * mu;
      sh[j] = beta[c0 + j] - mu * a1[j];
    }
  }
  SRK_PIXEL_LOOP(g, VEC) {
    const long long q = pw.q;
    float o[VEC];
    const long long e = q * g.C + cv * VEC;
    float v[VEC], d[VEC];
    Vec<T, VEC>::ld(y + e, v);
    Vec<T, VEC>::ld(dout + e, d);
    const bool border = pw.border();
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float gd = d[j];
      if (alpha_p && fmaf(v[j], a1[j], sh[j]) < 0.f) gd *= alpha;
      o[j] = border ? 0.f : fmaf(a1[j], gd, fmaf(a2[j], v[j], a3[j]));
    }
    Vec<T, VEC>::st(dy + e, o);
  }
  if (acc != nullptr) {
    if (__syncthreads_or(clear_acc)) acc_clear(acc, threadIdx.x, blockDim.x);
  }
}

// generic consumer of an accumulator: out[i] = value i (i < nv), then the accumulator is reset
__global__ void acc_read_kernel(unsigned long long* acc, int nv, float* __restrict__ out) {
  for (int i = threadIdx.x; i < nv; i += blockDim.x) out[i] = (float)acc_read(acc, i);
  __syncthreads();
  acc_clear(acc, threadIdx.x, blockDim.x);
}

// ---- activation backward -------------------------------------------------------------------------
// PReLU backward needs the sign (and, for the slope gradient, the value) of the PRE-activation z.  For alpha > 0
// both follow from the saved output (sign(out) = sign(z), z = out / alpha on the negative side).  For alpha <= 0
// they do not (alpha < 0 maps negative z to positive outputs, alpha = 0 maps them to 0): the forward epilogues then
// store z itself into `zsave` (same geometry as out) and this kernel reads it instead - a uniform branch on the
// device-resident slope, so the common case costs nothing and no host decision depends on a parameter value.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ out,
    const T* __restrict__ zsave, Geo g, int ppb, int act, const float* __restrict__ alpha_p,
    float* __restrict__ dalpha, RedArgs ra, T* __restrict__ dz) {
  const float alpha = (act == SRK_ACT_PRELU) ? alpha_p[0] : 0.f;
  const bool use_z = act == SRK_ACT_PRELU && zsave != nullptr && !(alpha > 0.f);
  const float inv_alpha = (act == SRK_ACT_PRELU && alpha != 0.f) ? 1.f / alpha : 0.f;
  const T* __restrict__ src = use_z ? zsave : out;
  const float zscale = use_z ? 1.f : inv_alpha;     // z = src * zscale on the negative side
  float da = 0.f;
  SRK_PIXEL_LOOP(g, VEC) {
    const long long q = pw.q;
    float o[VEC];
    const long long e = q * g.C + cv * VEC;
    if (pw.border()) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    } else {
      float v[VEC], d[VEC];
      Vec<T, VEC>::ld(src + e, v);
      Vec<T, VEC>::ld(dout + e, d);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const bool pos = use_z ? !(v[j] < 0.f) : (v[j] > 0.f);
        if (pos) o[j] = d[j];
        else { o[j] = alpha * d[j]; da = fmaf(d[j], v[j] * zscale, da); }
      }
    }
    Vec<T, VEC>::st(dz + e, o);
  }
  if (act == SRK_ACT_PRELU && dalpha) {
    __shared__ float red[32];
    __shared__ __align__(16) float vals[4];
    __shared__ float4 scratch[256];
    const float t = block_sum(da, red);
    if (threadIdx.x == 0) vals[0] = t;
    __syncthreads();
    ordered_fold(vals, 1, ra.tickets, gridDim.x, blockIdx.x, ra.partials, scratch, threadIdx.x, blockDim.x,
                 sync_block, [&](int, float v) { dalpha[0] = v; });
  }
}

// dout/out: [N][2H+2][2W+2][C]; dz: [N][H+2][W+2][4C].  Iterates over dz.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) act_bwd_unshuffle_kernel(const T* __restrict__ dout,
    const T* __restrict__ out_, const T* __restrict__ zsave, Geo gz, int ppb, int C, int act,
    const float* __restrict__ alpha_p, float* __restrict__ dalpha, RedArgs ra, int perm_tc, T* __restrict__ dz) {
  const float alpha = (act == SRK_ACT_PRELU) ? alpha_p[0] : 0.f;
  const bool use_z = act == SRK_ACT_PRELU && zsave != nullptr && !(alpha > 0.f);   // see act_bwd_kernel
  const float inv_alpha = (act == SRK_ACT_PRELU && alpha != 0.f) ? 1.f / alpha : 0.f;
  const T* __restrict__ out = use_z ? zsave : out_;
  const float zscale = use_z ? 1.f : inv_alpha;
  const int Wp2 = 2 * (gz.Wp - 2) + 2, Hp2 = 2 * (gz.Hp - 2) + 2;
  float da = 0.f;
  SRK_PIXEL_LOOP(gz, VEC) {
    const long long q = pw.q;
    float o[VEC];
    const long long e = q * gz.C + cv * VEC;
    if (pw.border()) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    } else {
      int xx = (int)(q % gz.Wp) - 1;
      long long t = q / gz.Wp;
      int yy = (int)(t % gz.Hp) - 1;
      int n = (int)(t / gz.Hp);
      float v[VEC], d[VEC];
      if (perm_tc && VEC == 8) {
        int cop = cv * VEC, sub = cop / C, c = cop - sub * C;
        long long src = (((long long)n * Hp2 + (2 * yy + (sub >> 1) + 1)) * Wp2 + (2 * xx + (sub & 1) + 1)) * C + c;
        Vec<T, VEC>::ld(out + src, v);
        Vec<T, VEC>::ld(dout + src, d);
      } else if (!perm_tc && VEC == 8 && sizeof(T) == 2) {
        // reference channel order co = 4c + sub: this thread's 8 channels are c = 2cv, 2cv+1 for the four
        // sub-pixels -> one 32-bit load (two adjacent bf16 channels) per sub-pixel and tensor; across the
        // warp each sub-pixel's 64 channels are one contiguous 128-byte row
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          long long src = (((long long)n * Hp2 + (2 * yy + (sub >> 1) + 1)) * Wp2 + (2 * xx + (sub & 1) + 1)) * C + 2 * cv;
          const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(out + src);
          const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(dout + src);
          const float2 af = __bfloat1622float2(a), bf = __bfloat1622float2(b);
          v[sub] = af.x; v[4 + sub] = af.y;
          d[sub] = bf.x; d[4 + sub] = bf.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          int cop = cv * VEC + j, sub, c;
          if (perm_tc) { sub = cop / C; c = cop - sub * C; } else { c = cop >> 2; sub = cop & 3; }
          long long src = (((long long)n * Hp2 + (2 * yy + (sub >> 1) + 1)) * Wp2 + (2 * xx + (sub & 1) + 1)) * C + c;
          v[j] = to_f<T>(out[src]);
          d[j] = to_f<T>(dout[src]);
        }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const bool pos = act == SRK_ACT_NONE || (use_z ? !(v[j] < 0.f) : (v[j] > 0.f));
        if (pos) o[j] = d[j];
        else { o[j] = alpha * d[j]; da = fmaf(d[j], v[j] * zscale, da); }
      }
    }
    Vec<T, VEC>::st(dz + e, o);
  }
  if (act == SRK_ACT_PRELU && dalpha) {
    __shared__ float red[32];
    __shared__ __align__(16) float vals[4];
    __shared__ float4 scratch[256];
    const float t = block_sum(da, red);
    if (threadIdx.x == 0) vals[0] = t;
    __syncthreads();
    ordered_fold(vals, 1, ra.tickets, gridDim.x, blockIdx.x, ra.partials, scratch, threadIdx.x, blockDim.x,
                 sync_block, [&](int, float v) { dalpha[0] = v; });
  }
}

// ---- squeeze-excite ------------------------------------------------------------------------------
// per-image channel sums of a (optionally times b): grid = (blocks_per_image, N)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) image_channel_sum_kernel(const T* __restrict__ a,
    const T* __restrict__ b, Geo g, int ppb, float scale, float* __restrict__ out, RedArgs ra, int row_floats) {
  extern __shared__ __align__(16) float smem[];
  float* vals = smem + row_floats;
  float4* scratch = reinterpret_cast<float4*>(vals + ((g.C + 3) & ~3));
  const int n = blockIdx.y;
  const long long img_pixels = (long long)g.Hp * g.Wp;
  const int CV = g.C / VEC, rows = blockDim.x / CV, cv = threadIdx.x % CV, prow = threadIdx.x / CV;
  long long q_begin = (long long)blockIdx.x * ppb, q_end = q_begin + ppb;
  if (q_end > img_pixels) q_end = img_pixels;
  float s[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s[j] = 0.f;
  // the border is zero by invariant (in a, and in b's partner), so it adds exact zeros: plain strided walk, EW_U
  // pixels per thread and pass with all their 16-byte loads in flight before the first use
  typedef typename Vec<T, VEC>::Raw Raw;
  if (prow < rows)
    for (long long ql = q_begin + prow; ql < q_end; ql += (long long)EW_U * rows) {
      Raw ra_[EW_U], rb_[EW_U];
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        const long long qu = ql + (long long)u * rows;
        if (qu < q_end) {
          const long long e = (n * img_pixels + qu) * g.C + cv * VEC;
          ra_[u] = Vec<T, VEC>::ldraw(a + e);
          if (b) rb_[u] = Vec<T, VEC>::ldraw(b + e);
        }
      }
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        if (ql + (long long)u * rows >= q_end) continue;
        float v[VEC];
        Vec<T, VEC>::unpack(ra_[u], v);
        if (b) {
          float w[VEC];
          Vec<T, VEC>::unpack(rb_[u], w);
#pragma unroll
          for (int j = 0; j < VEC; ++j) s[j] = fmaf(v[j], w[j], s[j]);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) s[j] += v[j];
        }
      }
    }
  block_channel_sum<VEC>(s, smem, CV, rows, cv, prow, vals);
  // one reduction slot per image: ticket n, partial rows [n][gridDim.x][C]
  float* dst = out + (long long)n * g.C;
  ordered_fold(vals, g.C, ra.tickets + (size_t)n * red_tickets_needed(gridDim.x), gridDim.x, blockIdx.x,
               ra.partials + (size_t)n * red_partial_floats(gridDim.x, g.C), scratch, threadIdx.x, blockDim.x, sync_block,
               [&](int i, float v) { dst[i] = v * scale; });
}

__global__ void se_fc_kernel(const float* __restrict__ pool, const float* __restrict__ w1,
                             const float* __restrict__ w2, int C, int Cr, float* __restrict__ hidden,
                             float* __restrict__ gate) {
  extern __shared__ float sh[];  // pool[C] + hidden[Cr]
  float* sp = sh; float* hd = sh + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sp[c] = pool[(long long)n * C + c];
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < Cr; j += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w1[j * C + c], sp[c], s);
    s = warp_sum(s);
    if (lane == 0) { s = fmaxf(s, 0.f); hd[j] = s; hidden[(long long)n * Cr + j] = s; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(w2[c * Cr + j], hd[j], s);
    gate[(long long)n * C + c] = 1.f / (1.f + expf(-s));
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) se_apply_kernel(const T* __restrict__ x, const T* __restrict__ r,
    Geo g, int ppb, const float* __restrict__ gate, float scale, T* __restrict__ out) {
  const long long img_pixels = (long long)g.Hp * g.Wp;
  SRK_PIXEL_LOOP(g, VEC) {
    const long long q = pw.q;
    float o[VEC];
    const long long e = q * g.C + cv * VEC;
    if (pw.border()) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    } else {
      int n = (int)(q / img_pixels);
      float v[VEC];
      Vec<T, VEC>::ld(r + e, v);
      if (x) Vec<T, VEC>::ld(x + e, o);
      else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = fmaf(scale * gate[(long long)n * g.C + cv * VEC + j], v[j], o[j]);
    }
    Vec<T, VEC>::st(out + e, o);
  }
}

// single block, every operand staged in shared memory (N*C up to a few thousand floats): the five small products run
// out of shared memory instead of as chains of dependent global loads (22.8 -> ~5 us at N = 32, C = 96, Cr = 6).  Same
// loop order per output as se_fc_bwd_kernel below, so the results are bit-identical to it.
__global__ void __launch_bounds__(1024) se_fc_bwd_smem_kernel(const float* __restrict__ dgate_raw,
    const float* __restrict__ gate, const float* __restrict__ hidden, const float* __restrict__ pool,
    const float* __restrict__ w1, const float* __restrict__ w2, int N, int C, int Cr, float scale,
    float* __restrict__ dw1, float* __restrict__ dw2, float* __restrict__ dpool) {
  extern __shared__ float sfc[];
  float* dz2 = sfc;                 // [N][C]
  float* pl = dz2 + N * C;          // [N][C]
  float* hd = pl + N * C;           // [N][Cr]
  float* dh = hd + N * Cr;          // [N][Cr]
  float* sw1 = dh + N * Cr;         // [Cr][C]
  float* sw2 = sw1 + Cr * C;        // [C][Cr]
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    const float sg = gate[i];
    dz2[i] = scale * dgate_raw[i] * sg * (1.f - sg);
    pl[i] = pool[i];
  }
  for (int i = threadIdx.x; i < N * Cr; i += blockDim.x) hd[i] = hidden[i];
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) { sw1[i] = w1[i]; sw2[i] = w2[i]; }
  __syncthreads();
  for (int i = threadIdx.x; i < N * Cr; i += blockDim.x) {
    const int n = i / Cr, j = i - n * Cr;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(sw2[c * Cr + j], dz2[n * C + c], s);
    dh[i] = hd[i] > 0.f ? s : 0.f;
  }
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
    const int c = i / Cr, j = i - c * Cr;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dz2[n * C + c], hd[n * Cr + j], s);
    dw2[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) {
    const int j = i / C, c = i - j * C;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dh[n * Cr + j], pl[n * C + c], s);
    dw1[i] = s;
  }
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    const int n = i / C, c = i - n * C;
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(sw1[j * C + c], dh[n * Cr + j], s);
    dpool[i] = s;
  }
}

// single block; writes dw1 / dw2 (sums over the batch) and dpool
__global__ void se_fc_bwd_kernel(const float* __restrict__ dgate_raw, const float* __restrict__ gate,
    const float* __restrict__ hidden, const float* __restrict__ pool, const float* __restrict__ w1,
    const float* __restrict__ w2, int N, int C, int Cr, float scale, float* __restrict__ dw1,
    float* __restrict__ dw2, float* __restrict__ dpool, float* __restrict__ dz2_ws /* [N][C] */,
    float* __restrict__ dh_ws /* [N][Cr] */) {
  // dz2[n][c] = scale * dgate_raw * gate * (1-gate)
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    float s = gate[i];
    dz2_ws[i] = scale * dgate_raw[i] * s * (1.f - s);
  }
  __syncthreads();
  // dh[n][j] = [hidden>0] * sum_c w2[c][j] dz2[n][c]
  for (int i = threadIdx.x; i < N * Cr; i += blockDim.x) {
    int n = i / Cr, j = i - n * Cr;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(w2[c * Cr + j], dz2_ws[n * C + c], s);
    dh_ws[i] = hidden[i] > 0.f ? s : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
    int c = i / Cr, j = i - c * Cr;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dz2_ws[n * C + c], hidden[n * Cr + j], s);
    dw2[i] = s;
  }
  for (int i = threadIdx.x; i < Cr * C; i += blockDim.x) {
    int j = i / C, c = i - j * C;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dh_ws[n * Cr + j], pool[n * C + c], s);
    dw1[i] = s;
  }
  for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
    int n = i / C, c = i - n * C;
    float s = 0.f;
    for (int j = 0; j < Cr; ++j) s = fmaf(w1[j * C + c], dh_ws[n * Cr + j], s);
    dpool[i] = s;
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) se_bwd_apply_kernel(const T* __restrict__ dout, Geo g, int ppb,
    const float* __restrict__ gate, const float* __restrict__ dpool, float scale, float inv_hw,
    T* __restrict__ dr) {
  const long long img_pixels = (long long)g.Hp * g.Wp;
  SRK_PIXEL_LOOP(g, VEC) {
    const long long q = pw.q;
    float o[VEC];
    const long long e = q * g.C + cv * VEC;
    if (pw.border()) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    } else {
      int n = (int)(q / img_pixels);
      float d[VEC];
      Vec<T, VEC>::ld(dout + e, d);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        long long gi = (long long)n * g.C + cv * VEC + j;
        o[j] = scale * gate[gi] * d[j] + dpool[gi] * inv_hw;
      }
    }
    Vec<T, VEC>::st(dr + e, o);
  }
}

// ---- layout / misc ---------------------------------------------------------------------------------
template <typename T>
__global__ void image_to_act_kernel(const float* __restrict__ img, T* __restrict__ act, int N, int C,
                                    int H, int W) {
  const int Hp = H + 2, Wp = W + 2;
  long long total = (long long)N * Hp * Wp * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C); long long t = i / C;
    int xx = (int)(t % Wp); t /= Wp;
    int yy = (int)(t % Hp); int n = (int)(t / Hp);
    float v = 0.f;
    if (xx > 0 && xx < Wp - 1 && yy > 0 && yy < Hp - 1)
      v = img[(((long long)n * C + c) * H + (yy - 1)) * W + (xx - 1)];
    act[i] = from_f<T>(v);
  }
}
template <typename T>
__global__ void act_to_image_kernel(const T* __restrict__ act, float* __restrict__ img, int N, int C,
                                    int H, int W) {
  const int Hp = H + 2, Wp = W + 2;
  long long total = (long long)N * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % W); long long t = i / W;
    int y = (int)(t % H); t /= H;
    int c = (int)(t % C); int n = (int)(t / C);
    img[i] = to_f<T>(act[(((long long)n * Hp + y + 1) * Wp + x + 1) * C + c]);
  }
}
template <typename T>
__global__ void act_add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o,
                               long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    o[i] = from_f<T>(to_f<T>(a[i]) + to_f<T>(b[i]));
}

// F.interpolate bicubic, align_corners=False, A=-0.75, border taps clamped (ATen upsample_bicubic2d)
__device__ __forceinline__ void cubic_coeffs(float t, float* w) {
  const float A = -0.75f;
  float x;
  x = t + 1.f; w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;       w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t; w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t; w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}
__global__ void bicubic_kernel(const float* __restrict__ in, float* __restrict__ out, int NC, int H,
                               int W, int OH, int OW, float sh, float sw) {
  long long total = (long long)NC * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ox = (int)(i % OW); long long t = i / OW;
    int oy = (int)(t % OH); int nc = (int)(t / OH);
    float fy = sh * (oy + 0.5f) - 0.5f, fx = sw * (ox + 0.5f) - 0.5f;
    int iy = (int)floorf(fy), ix = (int)floorf(fx);
    float wy[4], wx[4];
    cubic_coeffs(fy - iy, wy);
    cubic_coeffs(fx - ix, wx);
    const float* p = in + (long long)nc * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int yy = min(max(iy - 1 + a, 0), H - 1);
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        int xx = min(max(ix - 1 + b, 0), W - 1);
        row = fmaf(wx[b], p[(long long)yy * W + xx], row);
      }
      acc = fmaf(wy[a], row, acc);
    }
    out[i] = acc;
  }
}

}  // namespace srk

using namespace srk;

#define ACT_CHECK(t, name)                                                                  \
  SRK_REQUIRE((t) && (t)->layout == SRK_LAYOUT_ACT && ((t)->dtype == SRK_F32 || (t)->dtype == SRK_BF16), \
              "%s: expected ACT tensor", name)

// dispatch on dtype and vector width
// ---- 2x2 max pooling, stride 2, floor mode (nn.MaxPool2d(2, 2) inside VGG19.features, loss.py:23-24) ----------
// One thread per (padded output pixel, channel vector).  Forward: border pixels of `out` are written as zeros.
// Backward: the gradient of an output pixel goes to the first maximum of its window in row-major scan order (the
// index ATen's max_pool2d_with_indices records); every input pixel - including rows / columns the floor mode
// leaves uncovered, and the border - is written exactly once (zero where no gradient arrives).
template <typename T, int VEC>
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C,
                                                           T* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2, CV = C / VEC;
  const long long total = (long long)N * (Ho + 2) * (Wo + 2) * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long t = i / CV;
    const int xo = (int)(t % (Wo + 2)); t /= (Wo + 2);
    const int yo = (int)(t % (Ho + 2));
    const int n = (int)(t / (Ho + 2));
    float o[VEC];
    if (xo == 0 || xo == Wo + 1 || yo == 0 || yo == Ho + 1) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = 0.f;
    } else {
      const long long base = (((long long)n * (H + 2) + (2 * (yo - 1) + 1)) * (W + 2) + (2 * (xo - 1) + 1)) * C + cv * VEC;
      float a[VEC], b[VEC], c[VEC], d[VEC];
      Vec<T, VEC>::ld(x + base, a);
      Vec<T, VEC>::ld(x + base + C, b);
      Vec<T, VEC>::ld(x + base + (long long)(W + 2) * C, c);
      Vec<T, VEC>::ld(x + base + (long long)(W + 2) * C + C, d);
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(c[j], d[j]));
    }
    Vec<T, VEC>::st(out + i * VEC, o);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dout,
                                                           int N, int H, int W, int C, T* __restrict__ dx) {
  const int Ho = H / 2, Wo = W / 2, CV = C / VEC;
  // one thread per 2x2 block of the PADDED input grid, aligned to the pooling windows: block by covers padded rows
  // 2by-1 and 2by, so rows 0 .. H+1 need by = 0 .. (H+2)/2; out-of-range quarters are skipped
  const int By = (H + 4) / 2, Bx = (W + 4) / 2;
  const long long total = (long long)N * By * Bx * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long t = i / CV;
    const int bx = (int)(t % Bx); t /= Bx;
    const int by = (int)(t % By);
    const int n = (int)(t / By);
    // window (yo, xo) of the pooling covers padded input pixels (2yo-1 .. 2yo, 2xo-1 .. 2xo), yo in [1, Ho]; blocks
    // are aligned to those windows: block (by, bx) = padded pixels (2by-1 .. 2by, 2bx-1 .. 2bx)
    const bool win = by >= 1 && by <= Ho && bx >= 1 && bx <= Wo;
    float g[VEC], v[4][VEC];
    if (win) {
      Vec<T, VEC>::ld(dout + ((((long long)n * (Ho + 2) + by) * (Wo + 2) + bx) * C + cv * VEC), g);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int py = 2 * by - 1 + (k >> 1), px = 2 * bx - 1 + (k & 1);
        Vec<T, VEC>::ld(x + ((((long long)n * (H + 2) + py) * (W + 2) + px) * C + cv * VEC), v[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int py = 2 * by - 1 + (k >> 1), px = 2 * bx - 1 + (k & 1);
      if (py < 0 || px < 0 || py > H + 1 || px > W + 1) continue;
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float r = 0.f;
        if (win) {
          // first maximum in scan order: element k wins if it is >= all later ones and > all earlier ones
          bool is_max = true;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if (m < k) is_max = is_max && (v[k][j] > v[m][j]);
            if (m > k) is_max = is_max && (v[k][j] >= v[m][j]);
          }
          r = is_max ? g[j] : 0.f;
        }
        o[j] = r;
      }
      Vec<T, VEC>::st(dx + ((((long long)n * (H + 2) + py) * (W + 2) + px) * C + cv * VEC), o);
    }
  }
}

// blocks of 256 threads of `kernel` that fit on an SM; queried once per (kernel, dynamic shared memory)
template <typename K>
static int resident_blocks(K kernel, int smem = 0) {
  struct Entry { const void* k; int smem, occ; };
  static Entry cache[32];
  static int n = 0;
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < n; ++i)
    if (cache[i].k == key && cache[i].smem == smem) return cache[i].occ;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ < 1) occ = 1;
  if (n < 32) cache[n++] = {key, smem, occ};   // (a race between two host threads at worst repeats the query)
  return occ;
}

#define DISPATCH_T_VEC(t, ...)                                            \
  do {                                                                    \
    const bool vec8__ = ((t)->c % 8 == 0) && ((t)->c / 8 <= 256);         \
    if ((t)->dtype == SRK_BF16) {                                         \
      using T = __nv_bfloat16;                                            \
      if (vec8__) { constexpr int VEC = 8; __VA_ARGS__; }                 \
      else { constexpr int VEC = 1; __VA_ARGS__; }                        \
    } else {                                                              \
      using T = float;                                                    \
      if (vec8__) { constexpr int VEC = 8; __VA_ARGS__; }                 \
      else { constexpr int VEC = 1; __VA_ARGS__; }                        \
    }                                                                     \
  } while (0)

// Experiment knob: ask for the maximum shared-memory carveout for the BN backward kernels so that their blocks can
// share an SM with the (shared-memory heavy) tcgen05 wgrad kernel running on the side stream.
static void ew_carveout_once() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* e = getenv("SRK_EW_CARVEOUT");
  if (!e || atoi(e) == 0) return;
  const int v = atoi(e) == 1 ? (int)cudaSharedmemCarveoutMaxShared : atoi(e);
  cudaFuncSetAttribute(bn_bwd_reduce_kernel<__nv_bfloat16, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
  cudaFuncSetAttribute(bn_bwd_apply_kernel<__nv_bfloat16, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, v);
}

static inline size_t red_smem(const srk_tensor* t) {
  // [rows][C] floats with rows <= 256
  int vec = (t->c % 8 == 0 && t->c / 8 <= 256) ? 8 : 1;
  int cv = t->c / vec;
  int rows = 256 / cv; if (rows < 1) rows = 1;
  return (size_t)rows * t->c * sizeof(float);
}
// dynamic shared memory of a reducing kernel with NV values per block (layout: see RedArgs)
static inline int red_row_floats(const srk_tensor* t) { return (int)((red_smem(t) / sizeof(float) + 3) & ~(size_t)3); }
static inline size_t red_total_smem(const srk_tensor* t, int nv) {
  return ((size_t)red_row_floats(t) + (size_t)((nv + 3) & ~3) + 4 * 256) * sizeof(float);
}
static inline bool c_ok(const srk_tensor* t) {
  int vec = (t->c % 8 == 0 && t->c / 8 <= 256) ? 8 : 1;
  return t->c / vec <= 256 && red_total_smem(t, 2 * t->c + 1) <= 48 * 1024;
}
// partial rows of `slots` reductions with `nblk` blocks and `nv` values each must fit the reduce workspace
static inline bool red_fits(long long slots, long long nblk, int nv) {
  return slots * red_tickets_needed((int)nblk) <= kRedTickets &&
         slots * red_partial_floats((int)nblk, nv) <= (long long)kRedPartialFloats;
}
#define RED_WS_CHECK(ws, slots, nblk, nv, name)                                                              \
  do {                                                                                                       \
    SRK_REQUIRE((ws) != nullptr, "%s: reduce_ws is required (srk_reduce_workspace_bytes() bytes, zero-filled once)", name); \
    SRK_REQUIRE(red_fits(slots, nblk, nv), "%s: reduction of %lld x %lld rows x %d values exceeds the reduce workspace", \
                name, (long long)(slots), (long long)(nblk), (int)(nv));                                      \
  } while (0)

extern "C" int64_t srk_reduce_workspace_bytes(void) { return (int64_t)kRedWsBytes; }

extern "C" int srk_bn_stats(const srk_tensor* y, float* sums, void* reduce_ws, void* stream) {
  ACT_CHECK(y, "srk_bn_stats");
  SRK_REQUIRE(c_ok(y), "srk_bn_stats: unsupported channel count %d", y->c);
  SRK_REQUIRE(sums != nullptr, "srk_bn_stats: null output");
  Geo g = geo_of(y); int blocks, ppb; reduce_grid(g, blocks, ppb);
  RED_WS_CHECK(reduce_ws, 1, blocks, 2 * y->c, "srk_bn_stats");
  const RedArgs ra = {red_tickets(reduce_ws), red_partials(reduce_ws)};
  DISPATCH_T_VEC(y, (launch_dep(PDL_BN, bn_stats_kernel<T, VEC>, dim3(blocks), dim3(256), red_total_smem(y, 2 * y->c),
                        (cudaStream_t)stream, (const T*)y->data, g, ppb, sums, ra, red_row_floats(y))));
  SRK_CUDA_LAUNCH_CHECK("bn_stats");
  return 0;
}

extern "C" int srk_bn_finalize(const float* sum, const float* sumsq, int c, int64_t count, float eps,
                               float momentum, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, float* mean, float* invstd, void* stream) {
  SRK_REQUIRE(count > 0, "srk_bn_finalize: empty batch");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      sum, sumsq, c, (double)count, eps, momentum, running_mean, running_var,
      (long long*)num_batches_tracked, mean, invstd);
  SRK_CUDA_LAUNCH_CHECK("bn_finalize");
  return 0;
}

extern "C" int srk_bn_eval_params(const float* running_mean, const float* running_var, int c, float eps,
                                  float* mean, float* invstd, void* stream) {
  bn_eval_params_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, c, eps, mean, invstd);
  SRK_CUDA_LAUNCH_CHECK("bn_eval_params");
  return 0;
}

extern "C" int srk_bn_apply(const srk_tensor* y, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, const float* alpha,
                            const srk_tensor* residual, const srk_tensor* out, void* stream) {
  ACT_CHECK(y, "srk_bn_apply"); ACT_CHECK(out, "srk_bn_apply");
  SRK_REQUIRE(same_geometry(y, out) && y->dtype == out->dtype, "srk_bn_apply: geometry mismatch");
  if (residual) SRK_REQUIRE(same_geometry(y, residual) && residual->dtype == y->dtype && residual->layout == SRK_LAYOUT_ACT, "srk_bn_apply: residual mismatch");
  SRK_REQUIRE(c_ok(y), "srk_bn_apply: unsupported channel count %d", y->c);
  Geo g = geo_of(y); int blocks = 0, ppb = 0;
  BnFinalize fin = {};
  DISPATCH_T_VEC(y, (bn_map_grid(g, blocks, ppb, resident_blocks(bn_apply_kernel<T, VEC>)),
                     launch_dep(PDL_BN, bn_apply_kernel<T, VEC>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                        (const T*)y->data, g, ppb, mean, invstd, gamma, beta, alpha,
                        residual ? (const T*)residual->data : nullptr, (T*)out->data, fin)));
  SRK_CUDA_LAUNCH_CHECK("bn_apply");
  return 0;
}

extern "C" int srk_bn_apply_train(const srk_tensor* y, const float* sum, const float* sumsq, int64_t count, float eps,
                                  float momentum, float* running_mean, float* running_var,
                                  int64_t* num_batches_tracked, float* mean, float* invstd, const float* gamma,
                                  const float* beta, const float* alpha, const srk_tensor* residual,
                                  const srk_tensor* out, void* acc, void* stream) {
  ACT_CHECK(y, "srk_bn_apply_train"); ACT_CHECK(out, "srk_bn_apply_train");
  SRK_REQUIRE(same_geometry(y, out) && y->dtype == out->dtype, "srk_bn_apply_train: geometry mismatch");
  if (residual) SRK_REQUIRE(same_geometry(y, residual) && residual->dtype == y->dtype && residual->layout == SRK_LAYOUT_ACT, "srk_bn_apply_train: residual mismatch");
  SRK_REQUIRE(c_ok(y), "srk_bn_apply_train: unsupported channel count %d", y->c);
  SRK_REQUIRE(((sum && sumsq) || acc) && mean && invstd && count > 0, "srk_bn_apply_train: statistics buffers required");
  SRK_REQUIRE(acc == nullptr || 2 * y->c + 1 <= kAccNV, "srk_bn_apply_train: an accumulator holds at most %d values", kAccNV);
  SRK_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "srk_bn_apply_train: running_mean / running_var go together");
  Geo g = geo_of(y); int blocks = 0, ppb = 0;
  BnFinalize fin = {sum, sumsq, (double)count, eps, momentum, running_mean, running_var,
                    (long long*)num_batches_tracked, mean, invstd, (unsigned long long*)acc};
  const size_t smem = 2 * (size_t)y->c * sizeof(float);
  DISPATCH_T_VEC(y, (bn_map_grid(g, blocks, ppb, resident_blocks(bn_apply_kernel<T, VEC>, (int)smem)),
                     launch_dep(PDL_BN, bn_apply_kernel<T, VEC>, dim3(blocks), dim3(256), smem, (cudaStream_t)stream,
                        (const T*)y->data, g, ppb, nullptr, nullptr, gamma, beta, alpha,
                        residual ? (const T*)residual->data : nullptr, (T*)out->data, fin)));
  SRK_CUDA_LAUNCH_CHECK("bn_apply_train");
  return 0;
}

extern "C" int srk_bn_bwd_reduce(const srk_tensor* dout, const srk_tensor* y, const float* mean,
                                 const float* invstd, const float* gamma, const float* beta,
                                 const float* alpha, float* dgamma, float* dbeta, float* dalpha,
                                 void* reduce_ws, void* stream) {
  ACT_CHECK(y, "srk_bn_bwd_reduce"); ACT_CHECK(dout, "srk_bn_bwd_reduce");
  SRK_REQUIRE(same_geometry(y, dout) && y->dtype == dout->dtype, "srk_bn_bwd_reduce: geometry mismatch");
  SRK_REQUIRE(c_ok(y), "srk_bn_bwd_reduce: unsupported channel count %d", y->c);
  ew_carveout_once();
  Geo g = geo_of(y); int blocks, ppb; reduce_grid(g, blocks, ppb);
  RED_WS_CHECK(reduce_ws, 1, blocks, 2 * y->c + 1, "srk_bn_bwd_reduce");
  const RedArgs ra = {red_tickets(reduce_ws), red_partials(reduce_ws)};
  DISPATCH_T_VEC(y, (launch_dep(PDL_BN, bn_bwd_reduce_kernel<T, VEC>, dim3(blocks), dim3(256),
                        red_total_smem(y, 2 * y->c + 1), (cudaStream_t)stream,
                        (const T*)dout->data, (const T*)y->data, g, ppb, mean, invstd, gamma, beta,
                        alpha, dgamma, dbeta, dalpha, ra, red_row_floats(y))));
  SRK_CUDA_LAUNCH_CHECK("bn_bwd_reduce");
  return 0;
}

extern "C" int srk_bn_bwd_apply(const srk_tensor* dout, const srk_tensor* y, const float* mean,
                                const float* invstd, const float* gamma, const float* beta,
                                const float* alpha, const float* dgamma_b, const float* dbeta_b,
                                int batch_stats, const srk_tensor* dy, void* stream) {
  ACT_CHECK(y, "srk_bn_bwd_apply"); ACT_CHECK(dout, "srk_bn_bwd_apply"); ACT_CHECK(dy, "srk_bn_bwd_apply");
  SRK_REQUIRE(same_geometry(y, dout) && same_geometry(y, dy) && y->dtype == dout->dtype && y->dtype == dy->dtype,
              "srk_bn_bwd_apply: geometry mismatch");
  SRK_REQUIRE(c_ok(y), "srk_bn_bwd_apply: unsupported channel count %d", y->c);
  Geo g = geo_of(y); int blocks = 0, ppb = 0;
  float inv_count = 1.f / ((float)y->n * y->h * y->w);
  DISPATCH_T_VEC(y, (bn_map_grid(g, blocks, ppb, resident_blocks(bn_bwd_apply_kernel<T, VEC>)),
                     launch_dep(PDL_BN, bn_bwd_apply_kernel<T, VEC>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                        (const T*)dout->data, (const T*)y->data, g, ppb, mean, invstd, gamma, beta,
                        alpha, dgamma_b, dbeta_b, inv_count, batch_stats, (T*)dy->data, nullptr, nullptr, nullptr, nullptr)));
  SRK_CUDA_LAUNCH_CHECK("bn_bwd_apply");
  return 0;
}

extern "C" int srk_bn_bwd_apply_raw(const srk_tensor* dout, const srk_tensor* y, const float* mean,
                                    const float* invstd, const float* gamma, const float* beta,
                                    const float* alpha, const float* sum_g, const float* sum_gz,
                                    int batch_stats, float* dgamma_out, const srk_tensor* dy, void* acc,
                                    float* dbeta_out, float* dalpha_out, void* stream) {
  ACT_CHECK(y, "srk_bn_bwd_apply_raw"); ACT_CHECK(dout, "srk_bn_bwd_apply_raw"); ACT_CHECK(dy, "srk_bn_bwd_apply_raw");
  SRK_REQUIRE(same_geometry(y, dout) && same_geometry(y, dy) && y->dtype == dout->dtype && y->dtype == dy->dtype,
              "srk_bn_bwd_apply_raw: geometry mismatch");
  SRK_REQUIRE(c_ok(y), "srk_bn_bwd_apply_raw: unsupported channel count %d", y->c);
  SRK_REQUIRE(((sum_g && sum_gz) || (acc && dbeta_out)) && dgamma_out,
              "srk_bn_bwd_apply_raw: sums (or an accumulator and dbeta_out) and dgamma_out are required");
  SRK_REQUIRE(acc == nullptr || 2 * y->c + 1 <= kAccNV, "srk_bn_bwd_apply_raw: an accumulator holds at most %d values", kAccNV);
  Geo g = geo_of(y); int blocks = 0, ppb = 0;
  float inv_count = 1.f / ((float)y->n * y->h * y->w);
  DISPATCH_T_VEC(y, (bn_map_grid(g, blocks, ppb, resident_blocks(bn_bwd_apply_kernel<T, VEC>)),
                     launch_dep(PDL_BN, bn_bwd_apply_kernel<T, VEC>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                        (const T*)dout->data, (const T*)y->data, g, ppb, mean, invstd, gamma, beta,
                        alpha, sum_gz, sum_g, inv_count, batch_stats, (T*)dy->data, dgamma_out,
                        (unsigned long long*)acc, dbeta_out, dalpha_out)));
  SRK_CUDA_LAUNCH_CHECK("bn_bwd_apply_raw");
  return 0;
}

extern "C" int64_t srk_acc_bytes(void) { return (int64_t)kAccBytes; }

extern "C" int srk_acc_read(void* acc, int nv, float* out, void* stream) {
  SRK_REQUIRE(acc && out && nv >= 1 && nv <= kAccNV, "srk_acc_read: bad arguments");
  acc_read_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((unsigned long long*)acc, nv, out);
  SRK_CUDA_LAUNCH_CHECK("acc_read");
  return 0;
}

extern "C" int srk_act_bwd(const srk_tensor* dout, const srk_tensor* out, const srk_tensor* zsave,
                           const srk_tensor* dz, int act, const float* alpha, float* dalpha, int pixel_unshuffle,
                           int perm_tc, void* reduce_ws, void* stream) {
  ACT_CHECK(dout, "srk_act_bwd"); ACT_CHECK(out, "srk_act_bwd"); ACT_CHECK(dz, "srk_act_bwd");
  SRK_REQUIRE(same_geometry(dout, out) && dout->dtype == out->dtype && dz->dtype == out->dtype, "srk_act_bwd: geometry mismatch");
  SRK_REQUIRE(act != SRK_ACT_PRELU || alpha != nullptr, "srk_act_bwd: PReLU needs alpha");
  SRK_REQUIRE(c_ok(dz), "srk_act_bwd: unsupported channel count %d", dz->c);
  if (zsave) {
    ACT_CHECK(zsave, "srk_act_bwd");
    SRK_REQUIRE(same_geometry(zsave, out) && zsave->dtype == out->dtype, "srk_act_bwd: zsave must match out");
  }
  Geo g = geo_of(dz); int blocks, ppb; pixel_grid(g, blocks, ppb);
  RedArgs ra = {nullptr, nullptr};
  if (act == SRK_ACT_PRELU && dalpha) {
    RED_WS_CHECK(reduce_ws, 1, blocks, 1, "srk_act_bwd");
    ra.tickets = red_tickets(reduce_ws); ra.partials = red_partials(reduce_ws);
  }
  if (pixel_unshuffle == 2) {
    SRK_REQUIRE(dz->c == 4 * out->c && out->h == 2 * dz->h && out->w == 2 * dz->w && dz->n == out->n,
                "srk_act_bwd: unshuffle geometry mismatch");
    SRK_REQUIRE(!perm_tc || out->c % 8 == 0, "srk_act_bwd: perm_tc needs C %% 8 == 0");
    DISPATCH_T_VEC(dz, (act_bwd_unshuffle_kernel<T, VEC><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                           (const T*)dout->data, (const T*)out->data, zsave ? (const T*)zsave->data : nullptr, g, ppb,
                           out->c, act, alpha, dalpha, ra, perm_tc, (T*)dz->data)));
  } else {
    SRK_REQUIRE(pixel_unshuffle == 0 && same_geometry(dz, out), "srk_act_bwd: geometry mismatch");
    DISPATCH_T_VEC(dz, (act_bwd_kernel<T, VEC><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                           (const T*)dout->data, (const T*)out->data, zsave ? (const T*)zsave->data : nullptr, g, ppb,
                           act, alpha, dalpha, ra, (T*)dz->data)));
  }
  SRK_CUDA_LAUNCH_CHECK("act_bwd");
  return 0;
}

static int image_sum_launch(const srk_tensor* a, const srk_tensor* b, float scale, float* out, void* reduce_ws,
                            cudaStream_t st, const char* name) {
  Geo g = geo_of(a);
  long long img_pixels = (long long)g.Hp * g.Wp;
  int bpi = (int)((148LL * 4 + g.N - 1) / g.N);
  long long maxb = (img_pixels + 63) / 64;
  if (bpi > maxb) bpi = (int)maxb;
  if (bpi < 1) bpi = 1;
  int ppb = (int)((img_pixels + bpi - 1) / bpi);
  bpi = (int)((img_pixels + ppb - 1) / ppb);
  dim3 grid(bpi, g.N);
  RED_WS_CHECK(reduce_ws, g.N, bpi, a->c, name);
  const RedArgs ra = {red_tickets(reduce_ws), red_partials(reduce_ws)};
  DISPATCH_T_VEC(a, (image_channel_sum_kernel<T, VEC><<<grid, 256, red_total_smem(a, a->c), st>>>(
                        (const T*)a->data, b ? (const T*)b->data : nullptr, g, ppb, scale, out, ra, red_row_floats(a))));
  SRK_CUDA_LAUNCH_CHECK(name);
  return 0;
}

extern "C" int srk_se_pool(const srk_tensor* r, float* pool, void* reduce_ws, void* stream) {
  ACT_CHECK(r, "srk_se_pool");
  SRK_REQUIRE(c_ok(r), "srk_se_pool: unsupported channel count %d", r->c);
  return image_sum_launch(r, nullptr, 1.f / ((float)r->h * r->w), pool, reduce_ws, (cudaStream_t)stream, "srk_se_pool");
}

extern "C" int srk_se_fc(const float* pool, const float* w1, const float* w2, int n, int c, int cr,
                         float* hidden, float* gate, void* stream) {
  SRK_REQUIRE(cr >= 1 && c >= 1, "srk_se_fc: bad sizes");
  se_fc_kernel<<<n, 128, sizeof(float) * (c + cr), (cudaStream_t)stream>>>(pool, w1, w2, c, cr, hidden, gate);
  SRK_CUDA_LAUNCH_CHECK("se_fc");
  return 0;
}

extern "C" int srk_se_apply(const srk_tensor* x, const srk_tensor* r, const float* gate, float scale,
                            const srk_tensor* out, void* stream) {
  ACT_CHECK(r, "srk_se_apply"); ACT_CHECK(out, "srk_se_apply");
  SRK_REQUIRE(same_geometry(r, out) && r->dtype == out->dtype, "srk_se_apply: geometry mismatch");
  if (x) SRK_REQUIRE(same_geometry(r, x) && x->dtype == r->dtype && x->layout == SRK_LAYOUT_ACT, "srk_se_apply: x mismatch");
  SRK_REQUIRE(c_ok(r), "srk_se_apply: unsupported channel count %d", r->c);
  Geo g = geo_of(r); int blocks, ppb; pixel_grid(g, blocks, ppb);
  DISPATCH_T_VEC(r, (se_apply_kernel<T, VEC><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                        x ? (const T*)x->data : nullptr, (const T*)r->data, g, ppb, gate, scale, (T*)out->data)));
  SRK_CUDA_LAUNCH_CHECK("se_apply");
  return 0;
}

extern "C" int srk_se_bwd_reduce(const srk_tensor* dout, const srk_tensor* r, float* dgate_raw, void* reduce_ws,
                                 void* stream) {
  ACT_CHECK(r, "srk_se_bwd_reduce"); ACT_CHECK(dout, "srk_se_bwd_reduce");
  SRK_REQUIRE(same_geometry(r, dout) && r->dtype == dout->dtype, "srk_se_bwd_reduce: geometry mismatch");
  SRK_REQUIRE(c_ok(r), "srk_se_bwd_reduce: unsupported channel count %d", r->c);
  return image_sum_launch(dout, r, 1.f, dgate_raw, reduce_ws, (cudaStream_t)stream, "srk_se_bwd_reduce");
}

extern "C" int srk_se_fc_bwd(const float* dgate_raw, const float* gate, const float* hidden,
                             const float* pool, const float* w1, const float* w2, int n, int c, int cr,
                             float scale, float* dw1, float* dw2, float* dpool, void* stream) {
  // scratch: dz2 [n][c] + dh [n][cr]; borrowed from dpool's tail is not possible -> use a static
  // per-call device allocation-free trick: dpool is [n][c]; we need n*c + n*cr more floats.  The
  // caller provides dpool with room for 2*n*c + n*cr floats (documented in the Python binding).
  const size_t smem = sizeof(float) * (2 * (size_t)n * c + 2 * (size_t)n * cr + 2 * (size_t)c * cr);
  if (smem <= 200 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(se_fc_bwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    se_fc_bwd_smem_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(dgate_raw, gate, hidden, pool, w1, w2, n, c, cr, scale,
                                                                  dw1, dw2, dpool);
    SRK_CUDA_LAUNCH_CHECK("se_fc_bwd");
    return 0;
  }
  float* dz2 = dpool + (size_t)n * c;
  float* dh = dz2 + (size_t)n * c;
  se_fc_bwd_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(dgate_raw, gate, hidden, pool, w1, w2, n, c, cr,
                                                         scale, dw1, dw2, dpool, dz2, dh);
  SRK_CUDA_LAUNCH_CHECK("se_fc_bwd");
  return 0;
}

extern "C" int srk_se_bwd_apply(const srk_tensor* dout, const float* gate, const float* dpool,
                                float scale, const srk_tensor* dr, void* stream) {
  ACT_CHECK(dr, "srk_se_bwd_apply"); ACT_CHECK(dout, "srk_se_bwd_apply");
  SRK_REQUIRE(same_geometry(dr, dout) && dr->dtype == dout->dtype, "srk_se_bwd_apply: geometry mismatch");
  SRK_REQUIRE(c_ok(dr), "srk_se_bwd_apply: unsupported channel count %d", dr->c);
  Geo g = geo_of(dr); int blocks, ppb; pixel_grid(g, blocks, ppb);
  float inv_hw = 1.f / ((float)dr->h * dr->w);
  DISPATCH_T_VEC(dr, (se_bwd_apply_kernel<T, VEC><<<blocks, 256, 0, (cudaStream_t)stream>>>(
                         (const T*)dout->data, g, ppb, gate, dpool, scale, inv_hw, (T*)dr->data)));
  SRK_CUDA_LAUNCH_CHECK("se_bwd_apply");
  return 0;
}

static inline int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int srk_image_to_act(const srk_tensor* img, const srk_tensor* act, void* stream) {
  ACT_CHECK(act, "srk_image_to_act");
  SRK_REQUIRE(img && img->layout == SRK_LAYOUT_IMAGE && same_geometry(img, act), "srk_image_to_act: geometry mismatch");
  long long total = act_elems(act);
  if (act->dtype == SRK_BF16)
    image_to_act_kernel<__nv_bfloat16><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)img->data, (__nv_bfloat16*)act->data, act->n, act->c, act->h, act->w);
  else
    image_to_act_kernel<float><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)img->data, (float*)act->data, act->n, act->c, act->h, act->w);
  SRK_CUDA_LAUNCH_CHECK("image_to_act");
  return 0;
}

extern "C" int srk_act_to_image(const srk_tensor* act, const srk_tensor* img, void* stream) {
  ACT_CHECK(act, "srk_act_to_image");
  SRK_REQUIRE(img && img->layout == SRK_LAYOUT_IMAGE && same_geometry(img, act), "srk_act_to_image: geometry mismatch");
  long long total = (long long)act->n * act->c * act->h * act->w;
  if (act->dtype == SRK_BF16)
    act_to_image_kernel<__nv_bfloat16><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)act->data, (float*)img->data, act->n, act->c, act->h, act->w);
  else
    act_to_image_kernel<float><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)act->data, (float*)img->data, act->n, act->c, act->h, act->w);
  SRK_CUDA_LAUNCH_CHECK("act_to_image");
  return 0;
}

extern "C" int srk_act_add(const srk_tensor* a, const srk_tensor* b, const srk_tensor* out, void* stream) {
  ACT_CHECK(a, "srk_act_add"); ACT_CHECK(b, "srk_act_add"); ACT_CHECK(out, "srk_act_add");
  SRK_REQUIRE(same_geometry(a, b) && same_geometry(a, out) && a->dtype == b->dtype && a->dtype == out->dtype, "srk_act_add: geometry mismatch");
  long long total = act_elems(a);
  if (a->dtype == SRK_BF16)
    act_add_kernel<__nv_bfloat16><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a->data, (const __nv_bfloat16*)b->data, (__nv_bfloat16*)out->data, total);
  else
    act_add_kernel<float><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)a->data, (const float*)b->data, (float*)out->data, total);
  SRK_CUDA_LAUNCH_CHECK("act_add");
  return 0;
}

extern "C" int srk_maxpool2_fwd(const srk_tensor* x, const srk_tensor* out, void* stream) {
  ACT_CHECK(x, "srk_maxpool2_fwd"); ACT_CHECK(out, "srk_maxpool2_fwd");
  SRK_REQUIRE(out->n == x->n && out->c == x->c && out->h == x->h / 2 && out->w == x->w / 2 && out->dtype == x->dtype &&
                  out->h >= 1 && out->w >= 1,
              "srk_maxpool2_fwd: output must be [N, C, H/2, W/2] of the input dtype");
  const long long total = (long long)out->n * (out->h + 2) * (out->w + 2) * (x->c % 8 == 0 ? x->c / 8 : x->c);
  if (x->c % 8 == 0) {
    if (x->dtype == SRK_BF16)
      maxpool2_fwd_kernel<__nv_bfloat16, 8><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->data, x->n, x->h, x->w, x->c, (__nv_bfloat16*)out->data);
    else
      maxpool2_fwd_kernel<float, 8><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)x->data, x->n, x->h, x->w, x->c, (float*)out->data);
  } else {
    if (x->dtype == SRK_BF16)
      maxpool2_fwd_kernel<__nv_bfloat16, 1><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->data, x->n, x->h, x->w, x->c, (__nv_bfloat16*)out->data);
    else
      maxpool2_fwd_kernel<float, 1><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)x->data, x->n, x->h, x->w, x->c, (float*)out->data);
  }
  SRK_CUDA_LAUNCH_CHECK("maxpool2_fwd");
  return 0;
}

extern "C" int srk_maxpool2_bwd(const srk_tensor* x, const srk_tensor* dout, const srk_tensor* dx, void* stream) {
  ACT_CHECK(x, "srk_maxpool2_bwd"); ACT_CHECK(dout, "srk_maxpool2_bwd"); ACT_CHECK(dx, "srk_maxpool2_bwd");
  SRK_REQUIRE(same_geometry(x, dx) && x->dtype == dx->dtype && dout->dtype == x->dtype && dout->n == x->n &&
                  dout->c == x->c && dout->h == x->h / 2 && dout->w == x->w / 2,
              "srk_maxpool2_bwd: geometry mismatch");
  const long long total = (long long)x->n * ((x->h + 4) / 2) * ((x->w + 4) / 2) * (x->c % 8 == 0 ? x->c / 8 : x->c);
  if (x->c % 8 == 0) {
    if (x->dtype == SRK_BF16)
      maxpool2_bwd_kernel<__nv_bfloat16, 8><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->data, (const __nv_bfloat16*)dout->data, x->n, x->h, x->w, x->c, (__nv_bfloat16*)dx->data);
    else
      maxpool2_bwd_kernel<float, 8><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)x->data, (const float*)dout->data, x->n, x->h, x->w, x->c, (float*)dx->data);
  } else {
    if (x->dtype == SRK_BF16)
      maxpool2_bwd_kernel<__nv_bfloat16, 1><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->data, (const __nv_bfloat16*)dout->data, x->n, x->h, x->w, x->c, (__nv_bfloat16*)dx->data);
    else
      maxpool2_bwd_kernel<float, 1><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)x->data, (const float*)dout->data, x->n, x->h, x->w, x->c, (float*)dx->data);
  }
  SRK_CUDA_LAUNCH_CHECK("maxpool2_bwd");
  return 0;
}

extern "C" int srk_bicubic_upsample(const srk_tensor* in, const srk_tensor* out, void* stream) {
  SRK_REQUIRE(in && out && in->layout == SRK_LAYOUT_IMAGE && out->layout == SRK_LAYOUT_IMAGE, "srk_bicubic_upsample: IMAGE tensors expected");
  SRK_REQUIRE(in->n == out->n && in->c == out->c && out->h >= 1 && out->w >= 1, "srk_bicubic_upsample: geometry mismatch");
  // F.interpolate(scale_factor=s) with recompute_scale_factor=None uses 1/s as the coordinate scale
  float sh = (float)((double)in->h / (double)out->h), sw = (float)((double)in->w / (double)out->w);
  long long total = (long long)out->n * out->c * out->h * out->w;
  bicubic_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const float*)in->data, (float*)out->data, in->n * in->c, in->h, in->w, out->h, out->w, sh, sw);
  SRK_CUDA_LAUNCH_CHECK("bicubic");
  return 0;
}
