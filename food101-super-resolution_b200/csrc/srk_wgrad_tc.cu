// tcgen05 weight-gradient of the 3x3 convolutions on the zero-bordered channels-last bf16 layout.
//
//   dW[tap][ci][co] = sum_p  X[p + d(tap)][ci] * dZ[p][co]        (reference: convolution_backward of
//   models.py:46,49,65,67,113,117,120; p over all padded pixels - dZ is zero on the border)
//
// GEMM view per 16-pixel K step: D[(tap pair, ci) = 128][co = 64] += A^T B with BOTH operands "MN-major"
// (pixels are the contraction dimension and the slow axis of the NHWC tiles TMA delivers):
//   A = halo slab of X  [pixels][64 ci], SWIZZLE_128B; the two 64-row halves of M are two taps of the same
//       slab, separated by the descriptor's leading-byte-offset (d(t2) - d(t1)) * 128 B
//   B = dZ tile         [128 pixels][64 co], SWIZZLE_128B
// Five accumulators (tap pairs (0,1)(2,3)(4,5)(6,7)(8,-)) live in TMEM for the CTA's whole pixel range
// (split-K over persistent CTAs); each CTA stores its partial [9][64 ci][64 co] block once and a small
// kernel adds the partials into the OIHW fp32 gradient.  The bias gradient
// (column sums of dZ) is accumulated from the staged dZ tiles by the otherwise idle epilogue warps.
//
// Folded variant (default, WgradTcParams::fold): the three HORIZONTAL taps go into the MMA N dimension.  With
// p' = p + (s-1):  dW[(r,s)][ci][co] = sum_p' X[p' + (r-1)(W+2)][ci] * dZ[p' - (s-1)][co], so for one pixel block the
// B operand is dZ at three pixel shifts - three 64-wide N blocks of the same [130 pixels][64 co] slab, one pixel
// (128 B) apart through the descriptor's leading-byte-offset - and the A operand is X at the vertical taps only:
// two MMAs of M128 x N192 per 16-pixel K step (rows r = 0,1 | r = 2 and a discarded half) instead of five of
// M128 x N64: 18 KB of operand reads and 192 tensor cycles per K step instead of 30 KB and 240 port cycles.
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

#include <cstdlib>

namespace srk {

using namespace tc;

int* tc_err_flag();

constexpr int TM = 128, KC = 64, NT = 64, TAPS = 9, NACC = 5;
constexpr int SLAB_BOX_ROWS = 32;
constexpr int DZ_TILE_BYTES = TM * NT * 2;
constexpr int kThreads = 192;
constexpr int MAX_STAGES = 6;

constexpr int DZ_FOLD_ROWS = TM + 2;                        // dZ slab of the folded variant: pixels m0-1 .. m0+128
constexpr int DZ_FOLD_BYTES = (DZ_FOLD_ROWS * NT * 2 + 1023) / 1024 * 1024;
constexpr int FOLD_N = 3 * NT;                              // 192 accumulator columns per MMA

struct WgradTcParams {
  int fold;                // 1: horizontal taps folded into N (see the header comment)
  int P, Wp, num_tiles;
  int x_col0, dz_col0;     // channel offsets of this (ci chunk, co chunk) pass
  int slab_rows, stages, stage_bytes;
  float* ws;               // [gridDim.x][9][64][64] fp32 per-CTA partial sums
  float* db;               // [64] slice of the bias gradient (accumulated) or null
  int* err;
};

struct __align__(8) WgradBarriers {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], done;
  uint32_t tmem_base;
  float red[16][NT];
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDz,
                   const WgradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  WgradBarriers* bars = reinterpret_cast<WgradBarriers*>(smem_al + p.stages * p.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int slab_bytes = p.slab_rows * KC * 2;   // dZ tile sits behind the slab inside a stage

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(smem_u32(&bars->full[i]), 1);
      mbar_init(smem_u32(&bars->empty[i]), p.db ? 5 : 1);
    }
    mbar_init(smem_u32(&bars->done), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // barrier init and the TMEM allocation above overlap the previous kernel's tail (launch_dep)
  pdl_trigger();
  const uint32_t tmem_base = bars->tmem_base;
  const int my_tiles = blockIdx.x < p.num_tiles ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    // TMA producer: whole warp (warp-uniform control flow), one elected lane issues
    if (elect_one()) { prefetch_tmap(&tmX); prefetch_tmap(&tmDz); }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int m0 = (blockIdx.x + i * gridDim.x) * TM;
      if (!mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 11)) break;
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->full[s]);
        const uint32_t st0 = smem_base + s * p.stage_bytes;
        mbar_arrive_expect_tx(fb, slab_bytes + (p.fold ? DZ_FOLD_ROWS * NT * 2 : DZ_TILE_BYTES));
        const int row0 = p.fold ? m0 - p.Wp : m0 - p.Wp - 1;   // folded: vertical shifts only on the X side
        for (int j = 0; j < p.slab_rows / SLAB_BOX_ROWS; ++j)
          tma_load_2d(st0 + j * SLAB_BOX_ROWS * KC * 2, &tmX, fb, p.x_col0, row0 + j * SLAB_BOX_ROWS);
        tma_load_2d(st0 + slab_bytes, &tmDz, fb, p.dz_col0, p.fold ? m0 - 1 : m0);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp, one elected lane issues tcgen05.mma / commit
    constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);
    // constant descriptor halves (see srk_conv_tc.cu): per MMA only the lower words are advanced
    const uint64_t hi = make_smem_desc(0, 0, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_base = (uint32_t)(make_smem_desc(0, 0, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t wp_u = (uint32_t)p.Wp * (KC * 2 / 16), px_u = KC * 2 / 16;  // image row / pixel in 16-B units
    // tap pairs (0,1) (2,3) (4,5) (6,7) (8,8): start offset of the first tap | (distance to the second) << 16
    const uint32_t a_off[NACC] = {0u + (px_u << 16), 2 * px_u + ((wp_u - 2 * px_u) << 16),
                                  wp_u + px_u + (px_u << 16), 2 * wp_u + (px_u << 16), 2 * wp_u + 2 * px_u};
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int i = 0; i < my_tiles && ok; ++i) {
      ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 12);
      if (!ok) break;
      tc_fence_after();
      const uint32_t slab_lo = lo_base + ((smem_base + s * p.stage_bytes) >> 4);
      const uint32_t dz_lo = slab_lo + (slab_bytes >> 4);
      if (elect_one()) {
        if (p.fold) {
          constexpr uint32_t idesc_f = make_idesc_bf16(128, FOLD_N, 1, 1);
          // A: rows (r = 0 | r = 1) and (r = 1 again, discarded | r = 2) - every read stays inside the slab;
          // B: dZ slab rows j .. (N block j <-> s = 2 - j)
          const uint32_t af[2] = {0u + (wp_u << 16), wp_u + (wp_u << 16)};
          const uint32_t bf = dz_lo + (px_u << 16);
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int ks = 0; ks < TM / 16; ++ks)
              umma_bf16(tmem_base + a * FOLD_N, hi | (slab_lo + af[a] + ks * (16 * KC * 2 / 16)),
                        hi | (bf + ks * (16 * NT * 2 / 16)), idesc_f, (i | ks) != 0);
        } else {
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
#pragma unroll
            for (int ks = 0; ks < TM / 16; ++ks)
              umma_bf16(tmem_base + a * NT, hi | (slab_lo + a_off[a] + ks * (16 * KC * 2 / 16)),
                        hi | (dz_lo + ks * (16 * NT * 2 / 16)), idesc, (i | ks) != 0);
          }
        }
        umma_commit(smem_u32(&bars->empty[s]));
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(smem_u32(&bars->done));
    __syncwarp();
  } else {
    // ---- epilogue warps: bias-gradient partial sums while the MMAs run, then the accumulator flush ----
    const int et = threadIdx.x - 64;          // 0..127
    if (p.db) {
      const int c8 = et & 7, r0 = et >> 3;    // 16-byte channel chunk, first row
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        if (!mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 13)) break;
        const uint8_t* dz = smem_al + s * p.stage_bytes + slab_bytes;
        const int roff = p.fold ? 1 : 0;      // the folded dZ slab starts one pixel before the tile
#pragma unroll
        for (int rr = 0; rr < TM / 16; ++rr) {
          const int r = r0 + rr * 16 + roff;
          const uint4 q = *reinterpret_cast<const uint4*>(dz + r * 128 + ((c8 ^ (r & 7)) << 4));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 t = __bfloat1622float2(h[e]); acc[2 * e] += t.x; acc[2 * e + 1] += t.y; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->empty[s]));
        if (++s == S) { s = 0; ph ^= 1; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) bars->red[r0][c8 * 8 + e] = acc[e];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < NT) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) t += bars->red[r][et];
        p.ws[(size_t)gridDim.x * (TAPS * KC * NT) + (size_t)blockIdx.x * NT + et] = t;  // per-CTA partial
      }
    }
    if (my_tiles > 0 && mbar_wait(smem_u32(&bars->done), 0, p.err, 14)) {
      tc_fence_after();
      const int lg = warp & 3;
      // Flush: a thread holds one accumulator row (64 fp32 = 256 B), so storing it directly would write 16-byte pieces
      // at a 256-byte lane stride (32 L1 wavefronts per instruction).  Each warp transposes its 32 rows through a
      // padded tile in the (now idle) operand ring instead - conflict-free 16-byte stores, then two full rows per
      // coalesced 512-byte global store.  Rows (tap 2a, ci) and (tap 2a+1, ci) of accumulator a are 128
      // consecutive rows of this CTA's partial block.
      constexpr int ROW_PITCH = NT * 4 + 16;   // 272 B: consecutive rows start in different 16-byte bank groups
      uint8_t* tbuf = smem_al + lg * 32 * ROW_PITCH;
      // blocks of 64 accumulator columns: per-tap  a = 0..4 (taps 2a, 2a+1 in the two 64-row halves);
      // folded  b = 0..5: accumulator b / 3 (rows r = 0,1 | r = (1),2), N block j = b % 3 <-> horizontal tap s = 2 - j
      const int nblocks = p.fold ? 6 : NACC;
#pragma unroll 1
      for (int a = 0; a < nblocks; ++a) {
        uint32_t v[NT];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + a * NT;   // folded: (a / 3) * 192 + (a % 3) * 64
        tmem_ld_32x32(taddr, v);
        tmem_ld_32x32(taddr + 32, v + 32);
        tmem_ld_wait();
        __syncwarp();   // the previous accumulator's rows have been read back
#pragma unroll
        for (int j = 0; j < NT / 4; ++j)
          *reinterpret_cast<uint4*>(tbuf + lane * ROW_PITCH + j * 16) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        int tap0;            // tap of this warp's 64-row half; rows (lg & 1) * 32 .. of that tap's [ci][co] block
        if (p.fold) {
          const int acc = a / 3, sx = 2 - a % 3, r = acc == 0 ? (lg >> 1) : 2;
          if (acc == 1 && lg < 2) continue;       // first half of the second accumulator is a discarded copy of r = 1
          tap0 = r * 3 + sx;
        } else {
          if (a == NACC - 1 && lg >= 2) continue;   // the fifth accumulator holds tap 8 only (rows 0..63)
          tap0 = 2 * a + (lg >> 1);
        }
        // this CTA's partial sum, stored plainly (148 CTAs hammering the same 147 KB with atomics cost ~20 us);
        // wgrad_fold_kernel adds the partials up
        float* dst = p.ws + ((size_t)blockIdx.x * TAPS + tap0) * KC * NT + (size_t)((lg & 1) * 32) * NT;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int r = 2 * j + (lane >> 4), c = lane & 15;
          const uint4 q = *reinterpret_cast<const uint4*>(tbuf + r * ROW_PITCH + c * 16);
          *reinterpret_cast<uint4*>(dst + (size_t)r * NT + c * 4) = q;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dw[co0+co][ci0+ci][tap] += sum over CTAs of ws[cta][tap][ci][co];  db[co0+co] += sum over CTAs of the tail.
// The partials are L2 resident and the sum is latency bound, so it is spread over the partials as well: a block
// owns 32 consecutive outputs (one coalesced 128-byte row of every partial) and its 8 warps each sum every 8th
// partial with all their loads in flight, then fold through shared memory.
constexpr int FOLD_OUT = 32, FOLD_WARPS = 8, FOLD_MAX_PER_WARP = 24;   // up to 192 partials
__global__ void __launch_bounds__(FOLD_OUT * FOLD_WARPS)
wgrad_fold_kernel(const float* __restrict__ ws, int parts, float* __restrict__ dw, int cin_total,
                  int cout_total, int ci0, int co0, float* __restrict__ db, int accumulate, int perm_c4) {
  __shared__ float red[FOLD_WARPS][FOLD_OUT];
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * FOLD_OUT + lane;   // over [tap][ci][co], then the 64 bias columns
  const bool bias_part = i >= NT * KC * TAPS;
  const float* src = bias_part ? ws + (size_t)parts * (TAPS * KC * NT) + (i - NT * KC * TAPS) : ws + i;
  const size_t stride = bias_part ? NT : (size_t)TAPS * KC * NT;
  float v[FOLD_MAX_PER_WARP];
  if (!bias_part || db != nullptr) {
#pragma unroll
    for (int u = 0; u < FOLD_MAX_PER_WARP; ++u) {
      const int k = w + u * FOLD_WARPS;
      v[u] = k < parts ? __ldg(src + (size_t)k * stride) : 0.f;
    }
  } else {
#pragma unroll
    for (int u = 0; u < FOLD_MAX_PER_WARP; ++u) v[u] = 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < FOLD_MAX_PER_WARP; ++u) s += v[u];
  red[w][lane] = s;
  __syncthreads();
  if (w != 0) return;
#pragma unroll
  for (int r = 1; r < FOLD_WARPS; ++r) s += red[r][lane];
  // perm_c4 != 0: dZ arrived with its channels sub-pixel-major (cop = sub * C4 + c, C4 = Cout / 4): reference 4c + sub
  auto ref_co = [&](int cop) { return perm_c4 ? 4 * (cop % perm_c4) + cop / perm_c4 : cop; };
  if (bias_part) {
    const int c = i - NT * KC * TAPS;
    if (db != nullptr && co0 + c < cout_total) {
      const int cr = ref_co(co0 + c);
      db[cr] = accumulate ? db[cr] + s : s;
    }
    return;
  }
  const int co = i % NT, t = i / NT, ci = t % KC, tap = t / KC;
  if (ci0 + ci >= cin_total || co0 + co >= cout_total) return;  // zero-filled tail of a 96-channel tensor
  float* o = &dw[((size_t)ref_co(co0 + co) * cin_total + (ci0 + ci)) * TAPS + tap];
  *o = accumulate ? *o + s : s;
}

bool conv_wgrad_tc_shape_ok(const srk_tensor* x, const srk_tensor* dy, int r, int s) {
  if (r != 3 || s != 3) return false;
  if (x->layout != SRK_LAYOUT_ACT || dy->layout != SRK_LAYOUT_ACT) return false;
  if (x->dtype != SRK_BF16 || dy->dtype != SRK_BF16) return false;
  // channel counts that are not multiples of 64 (96) are covered by zero-filled TMA boxes; the fold drops the tail
  if (x->c % 32 != 0 || dy->c % 32 != 0 || x->c < 64 || dy->c < 64) return false;
  const int Wp = x->w + 2;
  const int slab_rows = ((TM + 2 * Wp + 2) + SLAB_BOX_ROWS - 1) / SLAB_BOX_ROWS * SLAB_BOX_ROWS;
  return 2 * (slab_rows * KC * 2 + DZ_FOLD_BYTES) + 8192 <= 227 * 1024;
}

int64_t conv_wgrad_tc_workspace(const srk_tensor* x, const srk_tensor* dy, int, int) {
  (void)x; (void)dy;
  return (int64_t)kNumSMs * (TAPS * KC * NT + NT) * sizeof(float);  // one partial (+ bias partial) per CTA
}

int conv_wgrad_tc_launch(const srk_tensor* x, const srk_tensor* dy, float* dw, float* db, int r, int s,
                         void* workspace, int accumulate, int perm_shuffle, cudaStream_t st) {
  SRK_REQUIRE(!perm_shuffle || dy->c % 4 == 0, "wgrad_tc: sub-pixel-major dZ needs Cout %% 4 == 0");
  const int Hp = x->h + 2, Wp = x->w + 2;
  const long long P = (long long)x->n * Hp * Wp;
  SRK_REQUIRE(P < (1LL << 31) - 4096, "wgrad_tc: too many pixels");
  static int smem_max = 0;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(wgrad3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  }
  static int fold = -1;
  if (fold < 0) { const char* e = getenv("SRK_WGRAD_FOLD"); fold = e ? atoi(e) != 0 : 1; }
  WgradTcParams p;
  p.fold = fold;
  p.P = (int)P; p.Wp = Wp;
  p.num_tiles = (int)((P + TM - 1) / TM);
  p.slab_rows = ((TM + 2 * Wp + (fold ? 0 : 2)) + SLAB_BOX_ROWS - 1) / SLAB_BOX_ROWS * SLAB_BOX_ROWS;
  p.stage_bytes = p.slab_rows * KC * 2 + (fold ? DZ_FOLD_BYTES : DZ_TILE_BYTES);
  const int fixed = 1024 + (int)sizeof(WgradBarriers);
  p.stages = (smem_max - fixed) / p.stage_bytes;
  // Three stages hide the TMA latency (a stage is ~2000 cycles of MMA work) and leave ~60 KB of the SM's shared
  // memory free: this kernel runs on the side stream NEXT TO the BatchNorm-backward reductions of the following
  // layer (srk/ops.py run_on_side_stream), whose blocks need 8 KB each to become resident beside it.
  if (p.stages > 3) p.stages = 3;
  SRK_REQUIRE(p.stages >= 2, "wgrad_tc: image too wide for the shared-memory slab");
  p.err = tc_err_flag();
  const int smem_bytes = fixed + p.stages * p.stage_bytes;
  CUtensorMap tmX, tmDz;
  if (make_tmap_2d_bf16(&tmX, x->data, (uint64_t)P, (uint64_t)x->c, (uint64_t)x->c, SLAB_BOX_ROWS, KC, 128)) return 1;
  if (make_tmap_2d_bf16(&tmDz, dy->data, (uint64_t)P, (uint64_t)dy->c, (uint64_t)dy->c, fold ? DZ_FOLD_ROWS : TM, NT, 128))
    return 1;
  const int kchunks = (x->c + KC - 1) / KC, nchunks = (dy->c + NT - 1) / NT;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const int pdl_cls = kchunks * nchunks > 1 ? PDL_WGRAD : PDL_WGRAD_SINGLE;
  for (int nc = 0; nc < nchunks; ++nc)
    for (int kc = 0; kc < kchunks; ++kc) {
      p.x_col0 = kc * KC;
      p.dz_col0 = nc * NT;
      p.ws = (float*)workspace;  // reused by every (ci, co) chunk pass: the fold runs right behind on the stream
      p.db = (db != nullptr && kc == 0) ? db + nc * NT : nullptr;
      launch_dep(pdl_cls, wgrad3x3_tc_kernel, dim3(grid), dim3(kThreads), smem_bytes, st, tmX, tmDz, p);
      SRK_CUDA_LAUNCH_CHECK("wgrad3x3_tc");
      launch_dep(pdl_cls, wgrad_fold_kernel, dim3((NT * KC * TAPS + NT) / FOLD_OUT), dim3(FOLD_OUT * FOLD_WARPS), 0, st, p.ws, grid,
                 dw, x->c, dy->c, kc * KC, nc * NT, p.db ? db : nullptr, accumulate, perm_shuffle ? dy->c / 4 : 0);
      SRK_CUDA_LAUNCH_CHECK("wgrad_fold");
    }
  return 0;
}

}  // namespace srk
