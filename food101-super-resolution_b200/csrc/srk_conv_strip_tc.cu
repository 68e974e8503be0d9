// The 3x3 64 -> 64 trunk convolution (fprop, and dgrad through rotated weights) as COLUMN STRIPS: the tensor-bound
// formulation of the layer that carries 55 % of the model's FLOPs (reference models.py:46,49,113: 33 such convs per
// ResNet-SR forward, as many data gradients per backward).
//
// Why another formulation.  With pixels as the MMA M dimension and Cout = 64 as N (srk_conv_fold_tc.cu, kFold = 0) an
// M128 N64 K16 MMA reads 6 KB of operands for 32 tensor cycles: the 128 B/clk shared-memory port caps the layer at
// 67 % of the tensor pipe and 42.8 % was measured.  Folding the three HORIZONTAL taps into N (N = 192,
// 10 KB per 96 tensor cycles: tensor-bound) needs Y[p] = D0[p-1] + D1[p] + D2[p+1], a shift along M = across TMEM
// lanes, which only shuffles / shared memory can do (kFold = 1: slower).
//
// Here the M tile is a vertical strip: 128 consecutive ROWS of the tall image (all N images stacked, N * (H+2) rows;
// the zero border rows between images make vertical taps exact across image boundaries) at ONE column x.  Then
//   * a vertical tap is a row shift of the A operand (row-shifted descriptor into the [130 rows][64 ch] slab),
//   * a horizontal tap s sends input column x to output column x - (s - 1): a DIFFERENT TILE with the same lanes.
// So the MMA of input column x (N = 192 = [s0 | s1 | s2] x 64 Cout, three vertical taps x four K steps = 12 MMAs)
// accumulates straight into the accumulators of output columns x+1, x, x-1, which sit in adjacent 64-column blocks
// of a ring of eight blocks that fills all 512 TMEM columns: the tap shift is a TMEM column offset and costs nothing.
// An output column is complete once its right neighbour has been processed; 16 epilogue warps drain it (bias,
// activation, BatchNorm statistics / backward reduction, residual) while the MMA warp is up to five columns ahead.
// 12 MMAs x 96 cycles = 1152 tensor cycles per 128 x 64 outputs: no wasted MMA work, half the slab traffic of the
// halo formulation (the horizontal halo is shared through the accumulators instead of being re-loaded).
//
// Work split: the (strip, output column) pairs are dealt out in contiguous runs to the persistent CTAs; a run
// re-processes one input column on either side (the neighbours' contributions to its first and last output).
//
// 608 threads: warp 0 TMA producer, warp 1 MMA issuer, warp 2 output TMA store / residual load, warps 3-18 epilogue.
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

#include <cstring>
#include <type_traits>

namespace srk {

int* tc_err_flag();
extern long long* g_tc_trace;

namespace strip {

using namespace tc;

constexpr int TM = 128;                     // rows of the tall image per strip (UMMA M)
constexpr int NT = 64;                      // output channels
constexpr int KC = 64;                      // input channels: one 128-byte swizzle row
constexpr int W_BYTES = 9 * NT * KC * 2;    // 72 KB: [r][s][co][ci]
constexpr int SLAB_USED = TM + 2;           // rows y0-1 .. y0+128
constexpr int SLAB_ROWS = 136;              // stage pitch in rows (multiple of 8: 1024-byte aligned stages)
constexpr int STAGE_BYTES = SLAB_ROWS * KC * 2;
constexpr int STAGES = 4;
constexpr int O_TILE_BYTES = TM * NT * 2;   // bf16 output tile staged for the TMA store
constexpr int Z_TILE_BYTES = TM * NT * 2;   // a tile of zeros: the two border columns of y are stored from it
constexpr int RED_BYTES = 4096;             // statistics fold scratch
constexpr int BIAS_BYTES = 1024;            // bias[64] | pad | BN scale[64] | BN shift[64]
constexpr int RING = 8;                     // accumulator blocks of 64 TMEM columns
constexpr int kEpiWarp0 = 3, kEpiWarps = 16, kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;   // 608
constexpr int CPT = 16;                     // accumulator columns per epilogue thread
constexpr int ACT_RUNTIME = -1;

struct Params {
  int R, Hp, Wp, W;        // tall-image rows (N * Hp), padded image height / width, interior width
  int T;                   // work items: strips * W output columns
  int act;
  const float* bias;
  const float* alpha;
  int has_residual;
  float* stats_sum;
  float* stats_sumsq;
  int bn_red, bn_mask;     // see srk_conv_fold_tc.cu: BatchNorm-backward reduction against the tile on the residual path
  const float* bn_mean; const float* bn_invstd; const float* bn_gamma; const float* bn_beta;
  float* bn_dalpha;
  unsigned* red_ticket;
  float* red_part;
  __nv_bfloat16* zsave;    // PReLU with a slope <= 0: copy of the pre-activation (y geometry) or null
  int* err;
  long long* trace;        // bring-up: clock64 stamps of CTA 0 ([16 roles][32 tiles], srk_tc_probe 100 / 102) or null
};

struct __align__(8) Barriers {
  uint64_t full[STAGES], empty[STAGES], wfull, tfull[RING], tempty[RING];
  uint64_t oready[2], ofree[2], rfull[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// One run of consecutive output columns [oa, ob] (padded x coordinates, 1 .. W) of strip k, and the input columns
// [xlo, xlo + nin) it needs.  Accumulator block j of the run belongs to output column xlo - 1 + j, j = 0 .. nin + 1;
// input column xlo + i adds into blocks i + 2 (s = 0), i + 1 (s = 1), i (s = 2).
struct Run { int k, oa, ob, xlo, nin; };
__device__ __forceinline__ bool next_run(int& t, int t1, int W, Run& r) {
  if (t >= t1) return false;
  r.k = t / W;
  r.oa = t - r.k * W + 1;
  const int tend = (r.k + 1) * W < t1 ? (r.k + 1) * W : t1;
  r.ob = r.oa + (tend - t) - 1;
  r.xlo = r.oa > 1 ? r.oa - 1 : 1;
  const int xhi = r.ob < W ? r.ob + 1 : W;
  r.nin = xhi - r.xlo + 1;
  t = tend;
  return true;
}

template <bool kStats, int kAct>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_strip_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                        const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                        const Params p) {
  const int act = kAct == ACT_RUNTIME ? p.act : kAct;
  float* const stats_sum = kStats ? p.stats_sum : nullptr;
  float* const stats_sumsq = kStats ? p.stats_sumsq : nullptr;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem: [weights 72 KB][A ring][2 output tiles][zero tile][fold scratch][bias][barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t wsm = smem_base;
  const uint32_t asm0 = smem_base + W_BYTES;
  const uint32_t osm = asm0 + STAGES * STAGE_BYTES;
  uint8_t* optr = smem_al + W_BYTES + STAGES * STAGE_BYTES;
  const uint32_t zsm = osm + 2 * O_TILE_BYTES;
  float* redp = reinterpret_cast<float*>(optr + 2 * O_TILE_BYTES + Z_TILE_BYTES);
  float* bias_s = reinterpret_cast<float*>(optr + 2 * O_TILE_BYTES + Z_TILE_BYTES + RED_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(optr + 2 * O_TILE_BYTES + Z_TILE_BYTES + RED_BYTES + BIAS_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->wfull), 1);
    for (int i = 0; i < RING; ++i) { mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), kEpiWarps); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->oready[i]), kEpiWarps); mbar_init(smem_u32(&bars->ofree[i]), 1);
      mbar_init(smem_u32(&bars->rfull[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(&bars->tmem_base), 512); tmem_relinquish(); }
  for (int z = threadIdx.x; z < Z_TILE_BYTES / 16; z += kThreads)
    reinterpret_cast<uint4*>(optr + 2 * O_TILE_BYTES)[z] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();   // the zero tile is read by TMA stores
  pdl_wait();      // from here on global memory is read (bias, BN constants, operands)
  pdl_trigger();
  if (threadIdx.x >= 96 && threadIdx.x < 96 + NT) {
    const int c = threadIdx.x - 96;
    bias_s[c] = p.bias ? __ldg(p.bias + c) : 0.f;
    if (kStats && p.bn_red) {
      const float sc = __ldg(p.bn_gamma + c) * __ldg(p.bn_invstd + c);
      bias_s[128 + c] = sc;
      bias_s[192 + c] = __ldg(p.bn_beta + c) - __ldg(p.bn_mean + c) * sc;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // this CTA's run of work items
  const int t0 = (int)(((long long)blockIdx.x * p.T) / gridDim.x), t1 = (int)(((long long)(blockIdx.x + 1) * p.T) / gridDim.x);

  if (warp == 0) {
    // ================= TMA producer: weights once, then one [130 rows][64 ch] slab per input column =================
    if (elect_one()) {
      prefetch_tmap(&tmA);
      prefetch_tmap(&tmW);
      mbar_arrive_expect_tx(smem_u32(&bars->wfull), W_BYTES);
      for (int r = 0; r < 3; ++r)   // smem [r][s][co][ci]: the three s of a kernel row are one N = 192 operand
        tma_load_2d(wsm + r * 3 * NT * KC * 2, &tmW, smem_u32(&bars->wfull), 0, r * 3 * NT);
    }
    __syncwarp();
    int s = 0, ntile = 0;
    uint32_t ph = 0;
    bool ok = true;
    Run run;
    for (int t = t0; ok && next_run(t, t1, p.W, run);) {
      for (int i = 0; i < run.nin && ok; ++i) {
        ok = mbar_wait_relaxed(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 1);
        if (!ok) break;
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[s]);
          mbar_arrive_expect_tx(fb, SLAB_USED * KC * 2);
          tma_load_3d(asm0 + s * STAGE_BYTES, &tmA, fb, 0, run.xlo + i, run.k * TM - 1);
          if (p.trace && blockIdx.x == 0 && ntile < 32) p.trace[0 * 32 + ntile] = clock64();
        }
        ++ntile;
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // tcgen05.mma issue blocks once the tensor queue is full, so the issuing thread cannot run ahead of the pipe: every
    // cycle it spends between two columns (barrier waits, commits) is a cycle the pipe idles.  The waits for column
    // i + 1 are therefore taken in the MIDDLE of column i's MMAs, while the queued ones execute.
    const uint32_t idesc64 = make_idesc_bf16(TM, 64, 0, 0), idesc128 = make_idesc_bf16(TM, 128, 0, 0),
                   idesc192 = make_idesc_bf16(TM, 192, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_base = (uint32_t)(make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t w_lo = lo_base + (wsm >> 4), a_lo0 = lo_base + (asm0 >> 4);
    constexpr uint32_t kRow = KC * 2 / 16;              // 16-byte units per operand row
    constexpr uint32_t kWr = 3 * NT * kRow, kWs = NT * kRow;   // one kernel row / one horizontal tap of weights
    // Control flow of this warp depends on kernel parameters and blockIdx only (a timed-out wait raises the error
    // flag but steers nothing), so the compiler keeps the stage / ring bookkeeping and the descriptor words in uniform
    // registers: with per-thread loop state every tcgen05.mma cost ~8 R2UR / UMOV instructions (60 - 75 cycles of
    // issue per MMA), which made the ISSUING THREAD, not the tensor pipe, the bottleneck of the tile.
    auto wait_hot = [&](uint32_t bar, uint32_t parity, int code) {
      if (mbar_try_wait(bar, parity)) return;
      const long long tw = clock64();
      while (!mbar_try_wait(bar, parity)) {
        if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, code); return; }
      }
    };
    wait_hot(smem_u32(&bars->wfull), 0, 2);
    struct Col { int t; Run run; int i, base; bool valid; };
    auto advance = [&](Col& c) {
      if (++c.i >= c.run.nin) { c.base += c.run.nin + 2; c.i = 0; c.valid = next_run(c.t, t1, p.W, c.run); }
    };
    // blocks a column opens must have been drained by the epilogue (their previous use, 8 blocks ago); its slab landed
    auto wait_col = [&](const Col& c, int s, uint32_t ph) {
      const int g0 = c.base + c.i;
      if (c.i == 0) {
        wait_hot(smem_u32(&bars->tempty[g0 & 7]), ((g0 >> 3) & 1) ^ 1, 3);
        wait_hot(smem_u32(&bars->tempty[(g0 + 1) & 7]), (((g0 + 1) >> 3) & 1) ^ 1, 3);
      }
      wait_hot(smem_u32(&bars->tempty[(g0 + 2) & 7]), (((g0 + 2) >> 3) & 1) ^ 1, 3);
      wait_hot(smem_u32(&bars->full[s]), ph, 4);
    };
    Col cur;
    cur.t = t0; cur.i = 0; cur.base = 0;
    cur.valid = next_run(cur.t, t1, p.W, cur.run);
    int s = 0, ntile = 0;
    uint32_t ph = 0;
    if (cur.valid) wait_col(cur, s, ph);
    while (cur.valid) {
      Col nxt = cur;
      advance(nxt);
      const int s2 = s + 1 == STAGES ? 0 : s + 1;
      const uint32_t ph2 = s2 == 0 ? ph ^ 1 : ph;
      const bool tr = p.trace && blockIdx.x == 0 && ntile < 32;
      tc_fence_after();
      const int g0 = cur.base + cur.i;          // blocks g0 (s = 2), g0 + 1 (s = 1), g0 + 2 (s = 0)
      const int p0 = g0 & 7, p1 = (g0 + 1) & 7, p2 = (g0 + 2) & 7;
      const uint32_t c0 = tmem_base + (uint32_t)(7 - p0) * NT, c1 = tmem_base + (uint32_t)(7 - p1) * NT,
                     c2 = tmem_base + (uint32_t)(7 - p2) * NT;
      const uint32_t a_lo = a_lo0 + s * (STAGE_BYTES >> 4);
      const uint32_t live = cur.i == 0 ? 0u : 1u;   // blocks g0, g0 + 1 hold earlier columns' contributions
      // kWrap 0: blocks p0+2, p0+1, p0 contiguous (descending columns) -> one N = 192 MMA per (r, ks);
      //       1: ring wrap between s = 0 and s = 1 (p0 == 6);  2: between s = 1 and s = 2 (p0 == 7)
      auto issue_rows = [&](auto wrap_tag, int r_begin, int r_end) {
        constexpr int kWrap = decltype(wrap_tag)::value;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          if (r < r_begin || r >= r_end) continue;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t ad = desc_hi | (a_lo + r * kRow + 2 * ks);
            const uint32_t b0 = w_lo + r * kWr + 2 * ks;
            if (r == 0 && ks == 0) {
              // the block of output column x + 1 is opened here (overwrite); the other two accumulate
              umma_bf16(c2, ad, desc_hi | b0, idesc64, 0u);
              umma_bf16(c1, ad, desc_hi | (b0 + kWs), idesc64, live);
              umma_bf16(c0, ad, desc_hi | (b0 + 2 * kWs), idesc64, live);
            } else if (kWrap == 0) {
              umma_bf16(c2, ad, desc_hi | b0, idesc192, 1u);
            } else if (kWrap == 1) {
              umma_bf16(c2, ad, desc_hi | b0, idesc64, 1u);
              umma_bf16(c1, ad, desc_hi | (b0 + kWs), idesc128, 1u);
            } else {
              umma_bf16(c2, ad, desc_hi | b0, idesc128, 1u);
              umma_bf16(c0, ad, desc_hi | (b0 + 2 * kWs), idesc64, 1u);
            }
          }
        }
      };
      auto issue = [&](int r_begin, int r_end) {
        if (p0 <= 5) issue_rows(std::integral_constant<int, 0>{}, r_begin, r_end);
        else if (p0 == 6) issue_rows(std::integral_constant<int, 1>{}, r_begin, r_end);
        else issue_rows(std::integral_constant<int, 2>{}, r_begin, r_end);
      };
      if (tr && lane == 0) p.trace[1 * 32 + ntile] = clock64();
      if (elect_one()) issue(0, 2);
      __syncwarp();
      if (tr && lane == 0) p.trace[2 * 32 + ntile] = clock64();
      if (nxt.valid) wait_col(nxt, s2, ph2);     // overlaps the queued MMAs
      if (tr && lane == 0) p.trace[3 * 32 + ntile] = clock64();
      if (elect_one()) {
        issue(2, 3);
        umma_commit(smem_u32(&bars->empty[s]));
        umma_commit(smem_u32(&bars->tfull[p0]));          // output column x - 1 is complete
        if (cur.i == cur.run.nin - 1) {                   // end of the run: the two blocks no later column feeds
          umma_commit(smem_u32(&bars->tfull[p1]));
          umma_commit(smem_u32(&bars->tfull[p2]));
        }
        if (tr) p.trace[4 * 32 + ntile] = clock64();
      }
      __syncwarp();
      cur = nxt; s = s2; ph = ph2; ++ntile;
    }
  } else if (warp == 2) {
    // ================= output store / residual load warp: stored output `it` <-> work item t0 + it =================
    const int n_out = t1 - t0;
    if (p.has_residual && elect_one()) {
      prefetch_tmap(&tmR);
      for (int it = 0; it < 2 && it < n_out; ++it) {
        const int t = t0 + it, k = t / p.W, xo = t - k * p.W + 1;
        const uint32_t rb = smem_u32(&bars->rfull[it]);
        mbar_arrive_expect_tx(rb, O_TILE_BYTES);
        tma_load_3d(osm + it * O_TILE_BYTES, &tmR, rb, 0, xo, k * TM);
      }
    }
    __syncwarp();
    for (int it = 0; it < n_out; ++it) {
      const int b = it & 1;
      if (!mbar_wait_relaxed(smem_u32(&bars->oready[b]), (it >> 1) & 1, p.err, 6)) break;
      if (elect_one()) {
        const int t = t0 + it, k = t / p.W, xo = t - k * p.W + 1;
        if (p.trace && blockIdx.x == 0 && it < 32) p.trace[8 * 32 + it] = clock64();
        tma_store_3d(&tmY, osm + b * O_TILE_BYTES, 0, xo, k * TM);
        // the zero border columns of y (layout invariant) ride along with their interior neighbours
        if (xo == 1) tma_store_3d(&tmY, zsm, 0, 0, k * TM);
        if (xo == p.W) tma_store_3d(&tmY, zsm, 0, p.W + 1, k * TM);
        tma_store_commit();
        tma_store_wait_read0();
        if (p.trace && blockIdx.x == 0 && it < 32) p.trace[9 * 32 + it] = clock64();
        if (p.has_residual && it + 2 < n_out) {
          const int t2 = t + 2, k2 = t2 / p.W, xo2 = t2 - k2 * p.W + 1;
          const uint32_t rb = smem_u32(&bars->rfull[b]);
          mbar_arrive_expect_tx(rb, O_TILE_BYTES);
          tma_load_3d(osm + b * O_TILE_BYTES, &tmR, rb, 0, xo2, k2 * TM);
        }
        mbar_arrive(smem_u32(&bars->ofree[b]));
      }
      __syncwarp();
    }
    if (elect_one()) tma_store_wait_all();
    __syncwarp();
  } else {
    // ====== epilogue: 16 warps; a warp owns TMEM lanes 32*(warp&3).. (rows of the strip) and 16 of the 64 channels ======
    const int lg = warp & 3, cq = (warp - kEpiWarp0) >> 2;
    const int c0 = cq * CPT;
    const float alpha = (act == SRK_ACT_PRELU) ? __ldg(p.alpha) : 0.f;
    const bool save_z = act == SRK_ACT_PRELU && p.zsave != nullptr && !(alpha > 0.f);
    const int row = lg * 32 + lane;
    float bn_da = 0.f;
    const float bn_alpha = (kStats && p.bn_red && p.bn_mask) ? __ldg(p.alpha) : 1.f;
    float s1[CPT], s2[CPT];
    if (stats_sum) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    }
    // after a protocol error every wait is skipped, but all warps keep running the same sequence so that the named
    // barriers below stay matched
    bool ok = true;
    int it = 0, base = 0;
    Run run;
    for (int t = t0; next_run(t, t1, p.W, run);) {
      const int y = run.k * TM + row;               // row of the tall image
      const int yi = y % p.Hp;
      const bool interior = y < p.R && yi >= 1 && yi <= p.Hp - 2;
      for (int j = 0; j < run.nin + 2; ++j) {
        const int g = base + j, slot = g & 7;
        const int xo = run.xlo - 1 + j;
        const bool stored = xo >= run.oa && xo <= run.ob;
        const bool tr = p.trace && blockIdx.x == 0 && it < 32 && threadIdx.x == kEpiWarp0 * 32 && stored;
        if (tr) p.trace[5 * 32 + it] = clock64();
        if (ok) ok = mbar_wait_relaxed(smem_u32(&bars->tfull[slot]), (g >> 3) & 1, p.err, 5);
        tc_fence_after();
        if (tr) p.trace[6 * 32 + it] = clock64();
        uint32_t v1[CPT];
        if (stored) {
          tmem_ld_32x16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(7 - slot) * NT + c0, v1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->tempty[slot]));   // the block may be reopened
        if (!stored) continue;
        const int acc = it & 1;
        uint8_t* orow = optr + acc * O_TILE_BYTES + row * 128;
        float f[CPT];
#pragma unroll
        for (int q = 0; q < CPT / 4; ++q) {
          const float4 b4 = reinterpret_cast<const float4*>(bias_s + c0)[q];
          f[4 * q] = __uint_as_float(v1[4 * q]) + b4.x; f[4 * q + 1] = __uint_as_float(v1[4 * q + 1]) + b4.y;
          f[4 * q + 2] = __uint_as_float(v1[4 * q + 2]) + b4.z; f[4 * q + 3] = __uint_as_float(v1[4 * q + 3]) + b4.w;
        }
        if (save_z && interior) {   // rare path (see act_bwd_kernel): the backward needs the pre-activation itself
          uint4* zd = reinterpret_cast<uint4*>(p.zsave + ((long long)y * p.Wp + xo) * NT + c0);
#pragma unroll
          for (int q = 0; q < CPT / 8; ++q)
            zd[q] = make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                               pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
        }
        if (act == SRK_ACT_RELU) {
#pragma unroll
          for (int q = 0; q < CPT; ++q) f[q] = fmaxf(f[q], 0.f);
        } else if (act == SRK_ACT_PRELU) {
#pragma unroll
          for (int q = 0; q < CPT; ++q) f[q] = fmaf(alpha, fminf(f[q], 0.f), fmaxf(f[q], 0.f));
        }
        const bool bn_red = kStats && p.bn_red;
        if (stats_sum && !bn_red && interior) {
#pragma unroll
          for (int q = 0; q < CPT; ++q) { s1[q] += f[q]; s2[q] = fmaf(f[q], f[q], s2[q]); }
        }
        if (ok) ok = mbar_wait_relaxed(smem_u32(&bars->ofree[acc]), ((it >> 1) & 1) ^ 1, p.err, 7);
        if (bn_red) {
          // the "residual" tile is Z, the saved input of the BatchNorm layer below: reduce against it, do not add
          if (ok) ok = mbar_wait_relaxed(smem_u32(&bars->rfull[acc]), (it >> 1) & 1, p.err, 8);
          if (interior) {
#pragma unroll
            for (int q = 0; q < CPT / 8; ++q) {
              const uint4 rr = *reinterpret_cast<const uint4*>(orow + (((cq * (CPT / 8) + q) ^ (row & 7)) << 4));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
              const float4 sc0 = reinterpret_cast<const float4*>(bias_s + 128 + c0 + 8 * q)[0];
              const float4 sc1 = reinterpret_cast<const float4*>(bias_s + 128 + c0 + 8 * q)[1];
              const float4 sh0 = reinterpret_cast<const float4*>(bias_s + 192 + c0 + 8 * q)[0];
              const float4 sh1 = reinterpret_cast<const float4*>(bias_s + 192 + c0 + 8 * q)[1];
              const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
              const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 zf = __bfloat1622float2(h[u]);
                const float zz[2] = {zf.x, zf.y};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int kk = 8 * q + 2 * u + e;
                  float gd = f[kk];
                  if (p.bn_mask) {
                    const float bb = fmaf(zz[e], scv[2 * u + e], shv[2 * u + e]);
                    if (bb < 0.f) { bn_da = fmaf(gd, bb, bn_da); gd *= bn_alpha; }
                  }
                  s1[kk] += gd;
                  s2[kk] = fmaf(gd, zz[e], s2[kk]);
                }
              }
            }
          }
        } else if (p.has_residual) {
          if (ok) ok = mbar_wait_relaxed(smem_u32(&bars->rfull[acc]), (it >> 1) & 1, p.err, 8);
#pragma unroll
          for (int q = 0; q < CPT / 8; ++q) {
            const uint4 rr = *reinterpret_cast<const uint4*>(orow + (((cq * (CPT / 8) + q) ^ (row & 7)) << 4));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
            for (int u = 0; u < 4; ++u) { float2 zf = __bfloat1622float2(h[u]); f[8 * q + 2 * u] += zf.x; f[8 * q + 2 * u + 1] += zf.y; }
          }
        }
        // bf16 tile staged in shared memory ([128 rows][128 B], SWIZZLE_128B) for one TMA store (box [128 y][1 x][64 c]);
        // border rows are stored as zeros (layout invariant), rows past the tensor are clipped by TMA
#pragma unroll
        for (int q = 0; q < CPT / 8; ++q) {
          uint4 o = make_uint4(0, 0, 0, 0);
          if (interior)
            o = make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                           pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
          *reinterpret_cast<uint4*>(orow + (((cq * (CPT / 8) + q) ^ (row & 7)) << 4)) = o;
        }
        // every writer fences its own stores towards the async proxy; one arrival per warp (512 per-thread arrivals
        // on one shared-memory barrier cost more than the tile's MMAs)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->oready[acc]));
        if (tr) p.trace[7 * 32 + it] = clock64();
        ++it;
      }
      base += run.nin + 2;
    }
    if (stats_sum) {
      // per-thread partial sums over this CTA's pixels -> per-channel totals in a fixed order (see
      // srk_conv_fold_tc.cu): half-warps, transposed butterfly, the four lane groups through shared memory, then one
      // ordered fold over the CTAs of the grid
#pragma unroll
      for (int q = 0; q < CPT; ++q) {
        s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], 16);
        s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], 16);
      }
#pragma unroll
      for (int half = CPT / 2; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int q = 0; q < half; ++q) {
          float k1 = up ? s1[q + half] : s1[q], o1 = up ? s1[q] : s1[q + half];
          float k2 = up ? s2[q + half] : s2[q], o2 = up ? s2[q] : s2[q + half];
          s1[q] = k1 + __shfl_xor_sync(0xffffffffu, o1, half);
          s2[q] = k2 + __shfl_xor_sync(0xffffffffu, o2, half);
        }
      }
      float* red = redp;                      // [4 lane groups][sum 64 | sumsq 64] | [16 warps] dalpha
      float* vals = redp + 528;
      const int et = threadIdx.x - kEpiWarp0 * 32;
      if (lane < CPT) {
        red[lg * 128 + c0 + lane] = s1[0];
        red[lg * 128 + 64 + c0 + lane] = s2[0];
      }
      {
        const float tt = warp_sum(bn_da);
        if (lane == 0) red[512 + (warp - kEpiWarp0)] = tt;
      }
      named_bar_sync(7, kEpiThreads);
      if (et < 128) {
        vals[et] = ((red[et] + red[128 + et]) + red[256 + et]) + red[384 + et];
      } else if (et == 128) {
        float tt = 0.f;
        for (int w = 0; w < kEpiWarps; ++w) tt += red[512 + w];
        vals[128] = tt;
      }
      named_bar_sync(7, kEpiThreads);
      if (et < 256) {
        const bool want_da = kStats && p.bn_red && p.bn_mask && p.bn_dalpha != nullptr;
        ordered_fold(vals, 129, p.red_ticket, (int)gridDim.x, (int)blockIdx.x, p.red_part, reinterpret_cast<float4*>(redp),
                     et, 256, [] { named_bar_sync(8, 256); },
                     [&](int i, float v) {
                       if (i < 64) stats_sum[i] = v;
                       else if (i < 128) stats_sumsq[i - 64] = v;
                       else if (want_da) p.bn_dalpha[0] = v;
                     });
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

static int make_tmap_strip(CUtensorMap* out, const void* base, uint64_t R, uint64_t Wp, uint32_t box_rows) {
  // the activation tensor as (channel, padded column, tall-image row); box = [box_rows rows][1 column][64 channels]
  PFN_encodeTiled enc = get_encode_tiled();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {NT, Wp, R};
  cuuint64_t strides[2] = {NT * 2, Wp * NT * 2};
  cuuint32_t box[3] = {NT, 1, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(strip) failed (%d) R=%llu Wp=%llu", (int)r, (unsigned long long)R,
              (unsigned long long)Wp);
  return 0;
}

template <bool kStats, int kAct>
static cudaError_t launch_one(int grid, int smem_bytes, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmW,
                              const CUtensorMap& tmY, const CUtensorMap& tmR, const Params& p) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv3x3_strip_tc_kernel<kStats, kAct>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled(PDL_CONV) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv3x3_strip_tc_kernel<kStats, kAct>, tmA, tmW, tmY, tmR, p);
}

}  // namespace strip

// Returns 0 ok, 1 error, -1 "not applicable" (the caller uses the halo-slab kernels).  Covers the plain 3x3
// 64 -> 64 pass on bf16 ACT tensors with bias / activation / residual / BN statistics / BN-backward reduction.
int conv_fprop_strip_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout, const float* bias,
                            int act, const float* alpha, const srk_tensor* residual, int shuffle, float* stats_sum,
                            float* stats_sumsq, cudaStream_t st, const BnRedArgs* br, void* reduce_ws, void* zsave) {
  using namespace strip;
  if (x->c != KC || cout != NT || shuffle != 0 || x->w < 1) return -1;
  SRK_REQUIRE((stats_sum == nullptr && br == nullptr) || reduce_ws != nullptr,
              "conv_strip: fused statistics need the reduce workspace");
  const int Hp = x->h + 2, Wp = x->w + 2;
  const long long R = (long long)x->n * Hp;
  SRK_REQUIRE(R * Wp < (1LL << 31) - 4096, "conv_strip: too many pixels");
  const int nstrips = (int)((R + TM - 1) / TM);
  const long long T = (long long)nstrips * x->w;
  SRK_REQUIRE(T < (1LL << 30), "conv_strip: too many work items");
  const int smem_bytes = 1024 + W_BYTES + STAGES * STAGE_BYTES + 2 * O_TILE_BYTES + Z_TILE_BYTES + RED_BYTES + BIAS_BYTES +
                         (int)sizeof(Barriers);

  CUtensorMap tmA, tmW, tmY, tmR;
  if (make_tmap_strip(&tmA, x->data, (uint64_t)R, (uint64_t)Wp, SLAB_USED)) return 1;
  if (make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)9 * cout, (uint64_t)KC, (uint64_t)KC, 3 * NT, KC, 128)) return 1;
  if (make_tmap_strip(&tmY, y->data, (uint64_t)R, (uint64_t)Wp, TM)) return 1;
  tmR = tmY;
  if (residual && make_tmap_strip(&tmR, residual->data, (uint64_t)R, (uint64_t)Wp, TM)) return 1;

  Params p;
  memset(&p, 0, sizeof(p));
  p.R = (int)R; p.Hp = Hp; p.Wp = Wp; p.W = x->w; p.T = (int)T;
  p.act = act; p.bias = bias; p.alpha = alpha;
  p.has_residual = residual ? 1 : 0;
  p.stats_sum = stats_sum; p.stats_sumsq = stats_sumsq;
  p.red_ticket = reduce_ws ? red_tickets(reduce_ws) : nullptr;
  p.red_part = reduce_ws ? red_partials(reduce_ws) : nullptr;
  p.zsave = (__nv_bfloat16*)zsave;
  p.err = tc_err_flag();
  p.trace = g_tc_trace;
  if (br) {
    SRK_REQUIRE(act == SRK_ACT_NONE && residual == nullptr && stats_sum == nullptr,
                "conv_strip: the fused BN-backward reduction covers the plain dgrad");
    SRK_REQUIRE(same_geometry(br->z, y) && br->z->dtype == SRK_BF16 && br->z->layout == SRK_LAYOUT_ACT,
                "conv_strip: Z must match the dgrad output geometry (bf16 ACT)");
    if (make_tmap_strip(&tmR, br->z->data, (uint64_t)R, (uint64_t)Wp, TM)) return 1;
    p.has_residual = 1;
    p.stats_sum = stats_sum = br->sum_g; p.stats_sumsq = stats_sumsq = br->sum_gz;
    p.bn_red = 1; p.bn_mask = br->alpha != nullptr; p.alpha = br->alpha;
    p.bn_mean = br->mean; p.bn_invstd = br->invstd; p.bn_gamma = br->gamma; p.bn_beta = br->beta;
    p.bn_dalpha = br->dalpha;
  }
  SRK_REQUIRE(stats_sum == nullptr || br != nullptr || (act == SRK_ACT_NONE && residual == nullptr),
              "conv_strip: fused BN statistics are taken of a plain conv output");
  const int grid = T < kNumSMs ? (int)T : kNumSMs;
  cudaError_t le;
  if (stats_sum) le = launch_one<true, SRK_ACT_NONE>(grid, smem_bytes, st, tmA, tmW, tmY, tmR, p);
  else if (act == SRK_ACT_NONE) le = launch_one<false, SRK_ACT_NONE>(grid, smem_bytes, st, tmA, tmW, tmY, tmR, p);
  else le = launch_one<false, ACT_RUNTIME>(grid, smem_bytes, st, tmA, tmW, tmY, tmR, p);
  SRK_REQUIRE(le == cudaSuccess, "conv3x3_strip_tc: launch failed: %s", cudaGetErrorString(le));
  SRK_CUDA_LAUNCH_CHECK("conv3x3_strip_tc");
  return 0;
}

}  // namespace srk
