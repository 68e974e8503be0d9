// Shared host/device helpers for libsrk (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "../../include/srk.h"

#ifndef SRK_PDL_DEFAULT
#define SRK_PDL_DEFAULT 2
#endif

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

// ---- deterministic cross-block reductions ------------------------------------------------------------------
// Float atomics make a sum depend on the order in which blocks happen to retire; two runs of the same training
// step then differ in the last bits of every BatchNorm statistic and, amplified by bf16 rounding, by percents
// on a gradient.  Every cross-block reduction of the library therefore goes through ordered_fold(): each block
// stores its partial row, takes a ticket, and the block that draws the LAST ticket sums all rows in block-index
// order (which block is last does not matter: the order of the additions is fixed by the code alone).
//
// Workspace (caller-owned, srk_reduce_workspace_bytes() bytes, zero-filled once when it is allocated, one per
// stream on which libsrk kernels may run concurrently):  [kRedTickets x u32 tickets][float partial rows].
// Tickets reset themselves, partial rows need no initialisation.
constexpr int kRedTickets = 4096;
constexpr size_t kRedWsBytes = (size_t)8 << 20;
constexpr size_t kRedPartialFloats = (kRedWsBytes - kRedTickets * sizeof(unsigned)) / sizeof(float);
inline unsigned* red_tickets(void* ws) { return reinterpret_cast<unsigned*>(ws); }
inline float* red_partials(void* ws) { return reinterpret_cast<float*>(reinterpret_cast<unsigned*>(ws) + kRedTickets); }

// Two levels keep the serial tail short: blocks are grouped by kRedGroup consecutive indices; the last block of a
// group to arrive folds the group's rows (a few KB) into one group row, and the last GROUP to finish folds the group
// rows.  One SM pulling all rows of a 296-block reduction (150 KB) at the end of the kernel measured +14 us; the two
// small folds cost about two L2 round trips.  Tickets of a reduction: [0] level 2, [1 + g] group g.
constexpr int kRedGroup = 8;
constexpr int kRedFlat = 32;   // grids up to this size fold in one level
__host__ __device__ inline int red_groups(int nblk) { return (nblk + kRedGroup - 1) / kRedGroup; }
__host__ __device__ inline int red_tickets_needed(int nblk) { return 1 + red_groups(nblk); }
// partial floats of one reduction: block rows followed by group rows
__host__ __device__ inline long long red_partial_floats(int nblk, int nv) {
  return (long long)(nblk + red_groups(nblk)) * ((nv + 3) & ~3);
}

// Called by the T threads `tid` = 0..T-1 of a block (a whole block with sync = __syncthreads, or a warp-aligned
// subset with a named barrier).  vals[NV]: this block's partial sums in shared memory, complete and visible to the
// T threads.  part: red_partial_floats(nblk, NV) floats of this reduction; bidx in [0, nblk): this block's row.
// scratch: T float4 of shared memory (may alias vals).  finish(i, sum) is called once per value, by one block.
template <typename Sync, typename Finish>
__device__ __forceinline__ void ordered_fold(const float* vals, int NV, unsigned* ticket, int nblk, int bidx,
                                             float* part, float4* scratch, int tid, int T, Sync sync, Finish finish) {
  const int NV4 = (NV + 3) >> 2;
  float4* rows = reinterpret_cast<float4*>(part);
  float4* prow = rows + (size_t)bidx * NV4;
  for (int i = tid; i < NV4; i += T) {
    float4 v;
    v.x = vals[4 * i];
    v.y = 4 * i + 1 < NV ? vals[4 * i + 1] : 0.f;
    v.z = 4 * i + 2 < NV ? vals[4 * i + 2] : 0.f;
    v.w = 4 * i + 3 < NV ? vals[4 * i + 3] : 0.f;
    __stcg(prow + i, v);
  }
  __threadfence();
  sync();
  int* flag = reinterpret_cast<int*>(scratch);
  // grids of up to kRedFlat blocks (the per-image reductions of the squeeze-excite gate, small layers) are ONE group:
  // the last block to arrive folds every row in block order and finishes - one ticket and one round of loads
  const int gs = nblk <= kRedFlat ? nblk : kRedGroup;
  const int grp = bidx / gs, ngrp = (nblk + gs - 1) / gs;
  const int g0 = grp * gs, gsize = nblk - g0 < gs ? nblk - g0 : gs;
  if (tid == 0) {
    const unsigned k = atomicAdd(ticket + 1 + grp, 1u);
    const int last = (k == (unsigned)gsize - 1u);
    if (last) ticket[1 + grp] = 0u;     // the whole group has arrived: ready for the next launch
    flag[0] = last;
  }
  sync();
  bool last = flag[0] != 0;
  sync();                       // flag[0] is about to be overwritten
  if (!last) return;
  __threadfence();
  // level 1: this group's rows, in block order, into the group row
  float4* grow = rows + (size_t)(nblk + grp) * NV4;
  for (int i = tid; i < NV4; i += T) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int b = 0; b < gsize; ++b) {
      const float4 v = __ldcg(rows + (size_t)(g0 + b) * NV4 + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (ngrp == 1) {      // single group: these are the totals
      const int k = 4 * i;
      finish(k, acc.x);
      if (k + 1 < NV) finish(k + 1, acc.y);
      if (k + 2 < NV) finish(k + 2, acc.z);
      if (k + 3 < NV) finish(k + 3, acc.w);
    } else {
      __stcg(grow + i, acc);
    }
  }
  if (ngrp == 1) return;
  __threadfence();
  sync();
  if (tid == 0) {
    const unsigned k = atomicAdd(ticket, 1u);
    const int l2 = (k == (unsigned)ngrp - 1u);
    if (l2) ticket[0] = 0u;
    flag[0] = l2;
  }
  sync();
  last = flag[0] != 0;
  sync();
  if (!last) return;
  __threadfence();
  // level 2: the group rows, in group order
  const float4* gro = rows + (size_t)nblk * NV4;
  for (int c0 = 0; c0 < NV4; c0 += T) {
    const int ncol = NV4 - c0 < T ? NV4 - c0 : T;
    const int G = T / ncol;     // row sets: thread (g, col) sums rows g, g + G, ... in that order
    const int col = tid % ncol, g = tid / ncol;
    if (g < G) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* p = gro + c0 + col;
#pragma unroll 4
      for (int b = g; b < ngrp; b += G) {
        const float4 v = __ldcg(p + (size_t)b * NV4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      scratch[g * ncol + col] = acc;
    }
    sync();
    // fixed binary tree over the row sets
    int top = 1;
    while (top < G) top <<= 1;
    for (int h = top >> 1; h >= 1; h >>= 1) {
      if (g < h && g + h < G) {
        float4 a = scratch[g * ncol + col];
        const float4 b = scratch[(g + h) * ncol + col];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        scratch[g * ncol + col] = a;
      }
      sync();
    }
    if (tid < ncol) {
      const float4 s = scratch[tid];
      const int i = 4 * (c0 + tid);
      finish(i, s.x);
      if (i + 1 < NV) finish(i + 1, s.y);
      if (i + 2 < NV) finish(i + 2, s.z);
      if (i + 3 < NV) finish(i + 3, s.w);
    }
    sync();
  }
}

// ---- exact cross-block sums: accumulators ("acc") ---------------------------------------------------------
// ordered_fold() ends a kernel with a chain of dependent global-memory steps (row store, fence, ticket, row loads,
// second level): measured at the end of the C2 trunk conv each step costs 1.2-2 us - 5 us on a 23 us kernel, on 65
// launches of a training step.  The convs therefore hand their per-CTA partial sums over as INTEGERS instead: a float
// is split exactly into radix-2^40 digits at 2^-94, 2^-54, 2^-14, 2^26, 2^66 (a 24-bit mantissa touches at most two
// of them; what lies below 2^-94 is dropped) and each non-zero digit is added to a 64-bit counter with a fire-and-
// forget red.global.add.u64.  Integer addition is associative, so the total does not depend on the order in which
// the CTAs arrive - bit-reproducible like ordered_fold, and exact - and nothing waits for a reply.  A sixth counter
// counts non-finite contributions and partial sums >= 2^120 (the sum then reads as NaN; below that bound 2^9
// contributors cannot overflow a counter: tests/test_acc_model.py checks the arithmetic on the host).  The CONSUMER kernel (bn_apply_train /
// bn_bwd_apply_raw / acc_read) converts the digits back in its prologue, and the consumer block that draws the last
// ticket - taken right after the prologue, its latency hidden behind the block's main loop - zero-fills the
// accumulator for its next use.  Contract: zero-filled once by the caller, then producer and consumer alternate.
// Layout (kAccBytes): u64 [6][kAccNV] digits (value-minor) | u32 consumer ticket.
constexpr int kAccNV = 132, kAccLimbs = 5;
constexpr size_t kAccBytes = 8192;
__device__ __forceinline__ double acc_unit(int k) {   // 2^(-94 + 40 k)
  return __longlong_as_double((long long)(1023 - 94 + 40 * k) << 52);
}
__device__ __forceinline__ void acc_add(unsigned long long* acc, int i, float p) {
  // NaN, Inf and partial sums of 2^120 and more (148 of them would overflow the top counter - and the float32 total
  // anyway) count as non-finite
  if (!(fabsf(p) < 1.329227995784916e36f)) { atomicAdd(acc + kAccLimbs * kAccNV + i, 1ull); return; }
  double r = (double)p;
#pragma unroll
  for (int k = kAccLimbs - 1; k >= 0; --k) {
    const long long d = __double2ll_rz(r * __longlong_as_double((long long)(1023 + 94 - 40 * k) << 52));   // r / unit
    r -= (double)d * acc_unit(k);   // exact: d has at most 24 significant bits
    if (d != 0) atomicAdd(acc + k * kAccNV + i, (unsigned long long)d);
  }
}
__device__ __forceinline__ double acc_read(const unsigned long long* acc, int i) {
  if (__ldcg(acc + kAccLimbs * kAccNV + i) != 0ull) return __longlong_as_double(0x7ff8000000000000ll);
  double s = 0.0;
#pragma unroll
  for (int k = kAccLimbs - 1; k >= 0; --k) s += (double)(long long)__ldcg(acc + k * kAccNV + i) * acc_unit(k);
  return s;
}
// Consumer side, called by thread 0 of EVERY block after the values of the block's acc_read()s have been USED (stored
// to shared memory, a __syncthreads() behind them): returns true in the block that must zero-fill the accumulator
// (acc_clear, at the end of that block).  No fence: the loads have returned their values before the ticket is issued,
// and the zero-fill is issued after the last ticket has been observed, so it cannot reach a load that is still in
// flight; the next producer is ordered behind the zero-fill by the kernel boundary.
__device__ __forceinline__ bool acc_ticket(unsigned long long* acc, int nblk) {
  unsigned* ticket = reinterpret_cast<unsigned*>(acc + (kAccLimbs + 1) * kAccNV);
  return atomicAdd(ticket, 1u) == (unsigned)nblk - 1u;
}
__device__ __forceinline__ void acc_clear(unsigned long long* acc, int tid, int T) {
  for (int i = tid; i < (kAccLimbs + 1) * kAccNV + 1; i += T) acc[i] = 0ull;   // digits, flags and the ticket word
}

// ---- programmatic dependent launch ---------------------------------------------------------------------
// A training step is ~270 launches of 7-60 us each, so the gap between two kernels of a stream (launch latency,
// the tail of the first, the prologue of the second: barrier init, TMEM allocation, descriptor prefetch) is a few
// percent of the step.  Kernels that take part call pdl_wait() before their first access to global memory that an
// earlier kernel may have written (or may still read) and pdl_trigger() right after it; launch_dep() adds the
// programmatic-serialization attribute, so such a kernel may become resident and run its prologue while its
// predecessor drains.  griddepcontrol.wait returns only when the predecessor grid has completed and its writes are
// visible, so ordering is unchanged; without the attribute both instructions are no-ops.
// Measured (B200, one box, graph replay): it pays where one operator is a chain of short dependent launches - the
// multi-pass weight gradients of 96- and 256-channel layers (AttentionSR step 16.23 -> 15.67 ms) - and is neutral to
// slightly negative between the 25-60 us persistent kernels of the ResNet trunk (7.18-7.24 -> 7.19-7.38 ms), whose
// CTAs cannot co-reside anyway (shared memory, register file).  Default: multi-pass wgrads only; SRK_PDL=<bits>
// selects classes (1 trunk convs, 2 multi-pass wgrads, 4 BatchNorm kernels, 8 single-pass wgrads), 0 = none.
// kernel classes, bits of SRK_PDL: trunk convs, multi-pass weight gradients, BatchNorm kernels, single-pass wgrads
enum { PDL_CONV = 1, PDL_WGRAD = 2, PDL_BN = 4, PDL_WGRAD_SINGLE = 8 };
inline bool pdl_enabled(int cls) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SRK_PDL"); v = e ? atoi(e) : SRK_PDL_DEFAULT; }
  return (v & cls) != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dep(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled(cls) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace srk
