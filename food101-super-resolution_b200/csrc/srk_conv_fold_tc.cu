// The 3x3 trunk convolution on tcgen05 / TMEM / TMA (fprop, and dgrad through rotated weights) on the zero-bordered
// channels-last bf16 layout: ONE pipeline, three formulations of the MMA loop (template parameters of the kernel).
//
// Pipeline - one persistent CTA per SM, 640 threads: warp 0 TMA producer (halo slab [128 + 2(W+2) (+2) rows][64 ch] per
// tile, weights once), warps 1 and 19 MMA issuers (even / odd tiles, one accumulator buffer each), warp 2 output TMA
// store / residual (+ Z) TMA loads, warps 3-18 epilogue (four per TMEM lane group, 16 output channels each; activation
// compiled in, incremental pixel walker, one accumulator-free arrival per warp), accumulators double buffered in TMEM.
// Optional epilogue fusions: BatchNorm forward statistics (kStats), BatchNorm BACKWARD reduction against a second
// input tile, with or without a residual added first (Params::bn_red, srk_conv_dgrad_bnred); the per-CTA sums leave
// through the ordered fold or, without a serial tail, as integer digits into an accumulator (Params::acc).
//   kCPT = 32 on single CTAs (Params::direct_out, SRK_TC_WIDE=1): 64 < Cout <= 128 in one pass, epilogue threads move
//     their own rows (no staging tiles next to 110-147 KB of weights).  Parity-green, slower on config C3: off by default.
//
//   kFold = 0 (DEFAULT): one MMA group per tap, N = 64: Y[p, co] = sum_{tap, ci} X[p + d(tap), ci] W[tap][co][ci] with
//     the 9 taps as row-shifted descriptors into the slab.  Bound by the shared-memory port.
//   kFold = 1: the three HORIZONTAL taps folded into N (N = 192, 12 MMAs of 96 cycles = the tensor floor, the A operand
//     read once per kernel ROW):
//       D_s[q, co] = sum_{r, ci} X[q + (r-1)*(W+2), ci] * W[r][s][co][ci]        (accumulator slot s, TMEM columns)
//       Y[p, co]   = D_0[p-1, co] + D_1[p, co] + D_2[p+1, co]                    (epilogue: shift across TMEM lanes)
//     A tile is 128 consecutive padded pixels q0..q0+127 and produces the 126 outputs q0+1..q0+126.  The +-1 lane
//     shift is a rotating warp shuffle per slot; the value that crosses a warp boundary is first swapped in through a
//     4 KB shared-memory exchange buffer.  Parity-green, but the shuffles travel through the same shared-memory port
//     the MMA operands use: slower than kFold = 0 (25.9 vs 23.2 us per C2 layer).
//   kPair = 1: kFold = 0 on CTA pairs (cta_group::2, M = 256), optionally with 128 output channels per pass
//     (kCPT = 32).  Parity-green, slower (the N = 64 MMA is bound by the A operand, which pairing does not reduce).
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

#include <cstdlib>

namespace srk {

int* tc_err_flag();
int tc_dbg();
extern long long* g_tc_trace;
int zero_border(const srk_tensor* t, cudaStream_t st);

namespace fold {

using namespace tc;

constexpr int TM = 128;              // pixels per tile (UMMA M)
// output pixels per tile: 126 with the folded taps (rows 0 and 127 of the tile are halo rows), 128 per-tap
template <bool kFold> struct Tile { static constexpr int TMO = kFold ? 126 : 128, ROW0 = kFold ? 1 : 0; };
constexpr int NT = 64;               // output channels per pass (one accumulator slot)
constexpr int KC = 64;               // contraction channels per pass: one 128-byte swizzle row
constexpr int W_BYTES = 9 * NT * KC * 2;   // 72 KB: [r][s][co][ci]
constexpr int SLAB_BOX_ROWS = 32;
constexpr int kEpiWarp0 = 3;         // first epilogue warp
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kMma2Warp = kEpiWarp0 + kEpiWarps;         // second MMA issuer (odd tiles), behind the epilogue warps
constexpr int kThreads = (kMma2Warp + 1) * 32;           // 640
constexpr int CPT = 16;              // accumulator columns (output channels) per epilogue thread
constexpr int O_TILE_BYTES = TM * NT * 2;  // bf16 output tile staged for the TMA store (126 rows used)
constexpr int XCH_BYTES = 2 * 4 * 4 * 2 * CPT * 4;  // [acc][column quarter][lane group][slot 0 / slot 2][16 floats]
constexpr int BIAS_BYTES = 1024;      // bias[64] | BN scale[64] | BN shift[64]
constexpr int MAX_STAGES = 4;
constexpr int ACC_COLS = 256;        // TMEM column stride between the two accumulator buffers
constexpr int ACT_RUNTIME = -1;

struct Params {
  int P, Hp, Wp, num_tiles;
  int k_col0;          // first contraction channel of this pass
  int w_row_per_tap;   // rows per tap in the packed weight matrix (= total output channels)
  int w_row0;          // first weight row of this pass
  int cout_total;      // channels per pixel of y
  int cout_off;        // channel offset inside a y row
  int act, shuffle;
  int slab_rows, stages, stage_bytes;
  int n_cols;          // output channels of this pass: 64 or 32
  int ksteps;          // 16-channel K steps of this pass: 4 or 2
  int step_n, step_y, step_x;   // (gridDim.x * TMO) pixels decomposed as n * Hp*Wp + y * Wp + x: per-tile walker
  float* partial_out;  // chunked contraction: fp32 partial sums [P][64] written instead of y ...
  const float* partial_in;  // ... and added back (before the activation) by the next chunk
  const float* bias;
  int bias_off;
  const float* alpha;
  int has_residual;
  __nv_bfloat16* y;
  int Hp2, Wp2;        // padded sizes of the pixel-shuffled output
  float* stats_sum;
  float* stats_sumsq;
  // BatchNorm-backward reduction fused into a dgrad (stats instantiation, bn_red = 1): the tile that arrives through
  // the residual path is Z, the saved pre-BN activation of the BN layer this gradient flows into; instead of being
  // added it gives  stats_sum[c] += sum g,  stats_sumsq[c] += sum g * z,  bn_dalpha += sum_{b<0} g * b  with
  // b = z * sc + sh the BN output and g the PReLU-masked gradient (bn_mask = 1) or the gradient itself.
  // bn_red = 2: the dgrad also carries a true residual (the skip gradient, added first: the sums are taken of the
  // total) and Z arrives through its own pair of staging tiles (tmZ) - srk_conv_dgrad_bnred with a residual.
  int bn_red, bn_mask;
  const float* bn_mean; const float* bn_invstd; const float* bn_gamma; const float* bn_beta;
  float* bn_dalpha;
  // cross-CTA stage of the statistics (ordered_fold, srk_common.cuh): ticket + partial rows [grid][132]
  unsigned* red_ticket;
  float* red_part;
  // ... or, when set, the exact integer accumulator (srk_common.cuh "acc"): value i of [sum 64 | sumsq 64 | dalpha]
  // is added to acc slot i with fire-and-forget reductions and the kernel ends without a serial tail
  unsigned long long* acc;
  // wide passes (kCPT = 32 on single CTAs, up to 128 output channels per pass: the 96-channel convs of AttentionSR in
  // ONE output-channel pass): w_bytes of resident weights instead of 72 KB, no staging tiles and no store warp - the
  // epilogue threads store their 64 bytes of y themselves (direct_out) and read the residual row (res) the same way
  int mma_warps;       // 2: warps 1 and kMma2Warp issue alternate tiles; 1: warp 1 issues every tile
  int w_bytes;
  int direct_out;
  int part_stride;     // floats per pixel of partial_out / partial_in
  const __nv_bfloat16* res;
  // PReLU epilogues with a slope <= 0 also store the pre-activation (same geometry as y) for the backward pass
  __nv_bfloat16* zsave;
  int* err;
  long long* trace;    // bring-up: per-tile clock64 stamps of CTA 0 ([16][32]) or null (general instantiation)
  int dbg;             // bring-up knobs (general instantiation only): 1 no stores, 2 no MMAs, 4 no A loads
};

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], wfull, tfull[2], tempty[2];
  uint64_t oready[2], ofree[2], rfull[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
  if (N == 32) tmem_ld_32x32(taddr, v);
  else tmem_ld_32x16(taddr, v);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// kFast: the common single-chunk 64 -> 64 pass (no partial sums, N = 3 x 64, four K steps, no PixelShuffle) with
// those choices and the activation (kAct) compiled in; the general instantiation (kFast = false, kAct = ACT_RUNTIME)
// covers 32-wide tails, chunked contractions and the PixelShuffle store.
// kFold = false: the same pipeline with one MMA group per tap (N = 64, 36 MMAs per tile, A operand re-read per tap
// from the row-shifted slab) and a plain one-slot epilogue - the per-tap formulation of srk_conv_tc.cu on the
// 16-warp epilogue of this file.
// kPair = true (per-tap only): the CTAs of a 2-CTA cluster process the two halves of a 256-pixel tile with
// tcgen05.mma.cta_group::2 (M = 256).  Each CTA stages its own slab and only HALF of the weight rows of every tap, so
// the weight operand costs half the shared-memory bandwidth per SM.  Rank 0 issues the MMAs; all TMA loads signal
// rank 0's barriers; MMA completion is multicast to both CTAs; the accumulator-free arrivals of rank 1's epilogue go
// to rank 0 through shared::cluster.
// kCPT: accumulator columns per epilogue thread.  16 covers passes of up to 64 output channels; 32 (CTA pairs only:
// each CTA then holds 64 of the 128 weight rows of a tap, the same 72 KB as a 64-channel pass) runs N = 128 MMAs at
// the tensor floor - used for the 64 -> 256 PixelShuffle convs, whose threads then store 16 bytes per sub-pixel.
template <bool kFold, bool kFast, bool kStats, int kAct, bool kPair = false, int kCPT = 16>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_fold_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                       const __grid_constant__ CUtensorMap tmZ, const Params p) {
  static_assert(!(kPair && kFold), "the CTA-pair variant is per-tap");
  static_assert(kCPT == 16 || (kCPT == 32 && !kFast && !kStats), "32 columns per thread: general instantiation only");
  constexpr int CPT = kCPT;
  constexpr int TMO = Tile<kFold>::TMO, ROW0 = Tile<kFold>::ROW0;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int n_cols = kFast ? NT : p.n_cols;
  const int w_rows = kPair ? n_cols / 2 : n_cols;   // weight rows per tap held by this CTA
  const int ksteps = kFast ? KC / 16 : p.ksteps;
  const int shuffle = kFast ? 0 : p.shuffle;
  const int act = kAct == ACT_RUNTIME ? p.act : kAct;
  float* const partial_out = kFast ? nullptr : p.partial_out;
  const float* const partial_in = kFast ? nullptr : p.partial_in;
  float* const stats_sum = kStats ? p.stats_sum : nullptr;
  float* const stats_sumsq = kStats ? p.stats_sumsq : nullptr;
  const int dbg = kFast ? 0 : p.dbg;
  long long* const trace = kFast ? nullptr : p.trace;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem: [weights 72 KB][A ring][2 output tiles][2 Z tiles iff bn_red == 2][exchange][bias][barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const bool direct_out = !kFast && p.direct_out != 0;
  const int w_bytes = kFast ? W_BYTES : p.w_bytes;
  const uint32_t wsm = smem_base;
  const uint32_t asm0 = smem_base + w_bytes;
  const uint32_t osm = asm0 + p.stages * p.stage_bytes;
  uint8_t* optr = smem_al + w_bytes + p.stages * p.stage_bytes;
  const bool two_in = kStats && p.bn_red == 2;   // residual tile AND Z tile per output tile
  const int o_bytes = direct_out ? 0 : 2 * O_TILE_BYTES;
  const int z_bytes = two_in ? 2 * O_TILE_BYTES : 0;
  const uint32_t zsm = osm + 2 * O_TILE_BYTES;
  float* xch = reinterpret_cast<float*>(optr + o_bytes + z_bytes);
  float* bias_s = reinterpret_cast<float*>(optr + o_bytes + z_bytes + XCH_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(optr + o_bytes + z_bytes + XCH_BYTES + BIAS_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->wfull), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), (kPair ? 2 : 1) * kEpiWarps);
      mbar_init(smem_u32(&bars->oready[i]), kEpiWarps); mbar_init(smem_u32(&bars->ofree[i]), 1);
      mbar_init(smem_u32(&bars->rfull[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) { tmem_alloc_pair(smem_u32(&bars->tmem_base), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(&bars->tmem_base), 512); tmem_relinquish(); }
  }
  // Everything above touches only this CTA's shared memory and TMEM: under programmatic dependent launch it overlaps
  // the tail of the previous kernel.  From here on global memory is read (bias, BN constants, operands).
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x >= 96 && threadIdx.x < 96 + 4 * CPT) {
    const int c = threadIdx.x - 96;
    int bi = p.bias_off + c;
    if (shuffle == 2) {   // sub-pixel-major rows: packed row cop = sub * C + ch  <->  reference channel 4 ch + sub
      const int cop = p.cout_off + c, sub = cop / p.cout_total;
      bi = 4 * (cop - sub * p.cout_total) + sub;
    }
    bias_s[c] = (p.bias && c < n_cols) ? __ldg(p.bias + bi) : 0.f;
    if (kStats && p.bn_red) {
      const float sc = __ldg(p.bn_gamma + c) * __ldg(p.bn_invstd + c);
      bias_s[128 + c] = sc;
      bias_s[192 + c] = __ldg(p.bn_beta + c) - __ldg(p.bn_mean + c) * sc;
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // barriers initialised in both CTAs before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // a pair walks its tiles in lockstep: both CTAs run an iteration iff the pair's first tile exists
  auto tile_ok = [&](int tile) { return (kPair ? tile - (int)rank : tile) < p.num_tiles; };

  if (warp == 0) {
    // ================= TMA producer =================
    const uint32_t wbar = kPair ? mapa_shared(smem_u32(&bars->wfull), 0) : smem_u32(&bars->wfull);
    if (elect_one()) {
      prefetch_tmap(&tmA);
      prefetch_tmap(&tmW);
      if (!kPair || rank == 0) mbar_arrive_expect_tx(smem_u32(&bars->wfull), 9 * n_cols * KC * 2);
      for (int t = 0; t < 9; ++t) {   // smem order [r][s][rows]: the three s of a row are one N = 3 n_cols tile
        const uint32_t dst = wsm + t * w_rows * KC * 2;
        const int wrow = t * p.w_row_per_tap + p.w_row0 + (int)rank * w_rows;
        if (kPair) tma_load_2d_pair(dst, &tmW, wbar, p.k_col0, wrow);
        else tma_load_2d(dst, &tmW, wbar, p.k_col0, wrow);
      }
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile_ok(tile) && ok; tile += gridDim.x) {
      ok = mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 1);
      if (!ok) break;
      if (elect_one()) {
        const uint32_t fb_local = smem_u32(&bars->full[s]);
        const uint32_t fb = kPair ? mapa_shared(fb_local, 0) : fb_local;
        if (dbg & 4) {
          mbar_arrive(fb_local);
        } else {
          // pair: rank 0 expects the bytes of both slabs on its barrier; rank 1 only issues its loads
          if (!kPair || rank == 0) mbar_arrive_expect_tx(fb_local, (kPair ? 2 : 1) * p.slab_rows * KC * 2);
          const int row0 = tile * TMO - p.Wp - 1;   // folded: pixel (tile*126 - 1) - Wp; per-tap: tile*128 - Wp - 1
          for (int j = 0; j < p.slab_rows / SLAB_BOX_ROWS; ++j) {
            const uint32_t dst = asm0 + s * p.stage_bytes + j * SLAB_BOX_ROWS * KC * 2;
            if (kPair) tma_load_2d_pair(dst, &tmA, fb, p.k_col0, row0 + j * SLAB_BOX_ROWS);
            else tma_load_2d(dst, &tmA, fb, p.k_col0, row0 + j * SLAB_BOX_ROWS);
          }
        }
        if (trace && blockIdx.x == 0 && tile / (int)gridDim.x < 32) trace[0 * 32 + tile / gridDim.x] = clock64();
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 || warp == kMma2Warp) {
    if (rank == 0) {   // rank 1 of a pair has no MMA work: its operands are consumed by rank 0's instructions
    // ================= MMA issuers (rank 0 of a pair issues for both CTAs) =================
    // TWO issuing warps, one per accumulator buffer: warp 1 takes the even tiles of this CTA, warp kMma2Warp the odd
    // ones.  Between two tiles an issuer commits and waits on two mbarriers - ~480 cycles in which at most two queued
    // MMAs (~100 cycles) keep the tensor pipe busy (per-tile trace, profiles/r2_trace_pertap_kernel_*.txt: 36 MMAs
    // issued in ~1 820 cycles, next tile's first MMA ~480 cycles later).  Tiles are independent (different accumulator,
    // read-only operands, per-thread tcgen05.commit), so the second issuer's MMAs fill the first one's gap.
    const uint32_t idesc = make_idesc_bf16(kPair ? 2 * TM : TM, kFold ? 3 * n_cols : n_cols, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_base = (uint32_t)(make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t w_lo = lo_base + (wsm >> 4), a_lo0 = lo_base + (asm0 >> 4);
    const uint32_t row_units = (uint32_t)p.Wp * (KC * 2 / 16);   // one image row of the slab, in 16-byte units
    const uint32_t wrow_units = (uint32_t)(3 * w_rows) * (KC * 2 / 16);   // one kernel row of weights (this CTA's rows)
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
    bool ok = mbar_wait(smem_u32(&bars->wfull), 0, p.err, 2);
    // (with fewer than three slab stages both in-flight tiles would hold every stage and the producer could not run
    // ahead: dgrad + residual + BN-reduce measured 32.5 us with two issuers against 29.6 us with one)
    const int nmw = p.mma_warps;
    const int mw = warp == 1 ? 0 : 1;
    if (mw >= nmw) ok = false;
    for (int tile = blockIdx.x + mw * gridDim.x, it = mw; tile_ok(tile) && ok; tile += nmw * gridDim.x, it += nmw) {
      const int acc = it & 1;                       // = mw
      const int s = it % S;                         // slab stage and its phase follow the CTA's tile counter
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      ok = mbar_wait(smem_u32(&bars->tempty[acc]), ((it >> 1) & 1) ^ 1, p.err, 3);
      if (!ok) break;
      ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 4);
      if (!ok) break;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      const uint32_t a_lo = a_lo0 + s * stage_units;
      if (elect_one()) {
        if (trace && blockIdx.x == 0 && it < 32) trace[1 * 32 + it] = clock64();
        if (kFold) {
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              if (ks < ksteps && !(dbg & 2))
                umma_bf16(d_tmem, desc_hi | (a_lo + r * row_units + 2 * ks), desc_hi | (w_lo + r * wrow_units + 2 * ks),
                          idesc, (r | ks) != 0);
        } else {
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks)
                if (ks < ksteps && !(dbg & 2)) {
                  const uint64_t ad = desc_hi | (a_lo + r * row_units + c * (KC * 2 / 16) + 2 * ks);
                  const uint64_t bd = desc_hi | (w_lo + r * wrow_units + c * (wrow_units / 3) + 2 * ks);
                  if (kPair) umma_bf16_pair(d_tmem, ad, bd, idesc, (r | c | ks) != 0);
                  else umma_bf16(d_tmem, ad, bd, idesc, (r | c | ks) != 0);
                }
        }
        if (kPair) {
          umma_commit_pair(smem_u32(&bars->empty[s]));
          umma_commit_pair(smem_u32(&bars->tfull[acc]));
        } else {
          umma_commit(smem_u32(&bars->empty[s]));
          umma_commit(smem_u32(&bars->tfull[acc]));
        }
        if (trace && blockIdx.x == 0 && it < 32) trace[2 * 32 + it] = clock64();
      }
      __syncwarp();
    }
    }
  } else if (warp == 2) {
    // ================= output store / residual load warp (plain, non-PixelShuffle outputs) =================
    if (shuffle == 0 && partial_out == nullptr && !direct_out && !(dbg & 1)) {
      const int first = kPair ? (int)blockIdx.x - (int)rank : (int)blockIdx.x;   // lockstep with the pair's first tile
      const int my_tiles = first < p.num_tiles ? (p.num_tiles - first + gridDim.x - 1) / gridDim.x : 0;
      if (p.has_residual && elect_one()) {
        prefetch_tmap(&tmR);
        if (two_in) prefetch_tmap(&tmZ);
        for (int it = 0; it < 2 && it < my_tiles; ++it) {
          const uint32_t rb = smem_u32(&bars->rfull[it]);
          mbar_arrive_expect_tx(rb, (two_in ? 2 : 1) * TMO * NT * 2);
          tma_load_2d(osm + it * O_TILE_BYTES, &tmR, rb, p.cout_off, (blockIdx.x + it * gridDim.x) * TMO);
          if (two_in) tma_load_2d(zsm + it * O_TILE_BYTES, &tmZ, rb, p.cout_off, (blockIdx.x + it * gridDim.x) * TMO);
        }
      }
      __syncwarp();
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        if (!mbar_wait(smem_u32(&bars->oready[b]), (it >> 1) & 1, p.err, 6)) break;
        if (elect_one()) {
          if (trace && blockIdx.x == 0 && it < 32) trace[12 * 32 + it] = clock64();
          tma_store_2d(&tmY, osm + b * O_TILE_BYTES, p.cout_off, (blockIdx.x + it * gridDim.x) * TMO);
          tma_store_commit();
          tma_store_wait_read0();
          if (trace && blockIdx.x == 0 && it < 32) trace[13 * 32 + it] = clock64();
          if (p.has_residual && it + 2 < my_tiles) {
            // (the Z tile of this buffer is free as well: every epilogue warp read it before arriving on oready)
            const uint32_t rb = smem_u32(&bars->rfull[b]);
            mbar_arrive_expect_tx(rb, (two_in ? 2 : 1) * TMO * NT * 2);
            tma_load_2d(osm + b * O_TILE_BYTES, &tmR, rb, p.cout_off, (blockIdx.x + (it + 2) * gridDim.x) * TMO);
            if (two_in) tma_load_2d(zsm + b * O_TILE_BYTES, &tmZ, rb, p.cout_off, (blockIdx.x + (it + 2) * gridDim.x) * TMO);
          }
          mbar_arrive(smem_u32(&bars->ofree[b]));
        }
        __syncwarp();
      }
      if (elect_one()) tma_store_wait_all();
      __syncwarp();
    }
  } else {
    // ====== epilogue: 16 warps; a warp owns TMEM lanes 32*(warp&3).. and 16 of the 64 channels ======
    const int lg = warp & 3, cq = (warp - kEpiWarp0) >> 2;
    const int c0 = cq * CPT;
    const float alpha = (act == SRK_ACT_PRELU) ? __ldg(p.alpha) : 0.f;
    const bool active = c0 < n_cols;
    const bool staged = shuffle == 0 && partial_out == nullptr && !direct_out && !(dbg & 1);
    const int part_stride = kFast ? NT : p.part_stride;
    const int row = lg * 32 + lane;
    const bool has_row = row >= ROW0 && row < ROW0 + TMO;
    const int src_up = (lane + 31) & 31, src_dn = (lane + 1) & 31;
    // pixel walker: coordinates (wn, wy, wx) of padded pixel t = tile*TMO + row (= this thread's pixel + ROW0)
    int wn, wy, wx;
    {
      const int t0 = blockIdx.x * TMO + row, img = p.Hp * p.Wp;
      wn = t0 / img;
      const int q = t0 - wn * img;
      wy = q / p.Wp;
      wx = q - wy * p.Wp;
    }
    float bn_da = 0.f;
    const float bn_alpha = (kStats && p.bn_red && p.bn_mask) ? __ldg(p.alpha) : 1.f;
    float s1[CPT], s2[CPT];
    if (stats_sum) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    }
    // after a protocol error every wait is skipped, but all warps keep running the same tile sequence so that
    // the named barriers below stay matched (a hung bar.sync would hang the GPU)
    bool ok = true;
    int it = 0;
    const uint32_t te0 = kPair ? mapa_shared(smem_u32(&bars->tempty[0]), 0) : smem_u32(&bars->tempty[0]);
    const uint32_t te1 = kPair ? mapa_shared(smem_u32(&bars->tempty[1]), 0) : smem_u32(&bars->tempty[1]);
    // one arrival per warp on the accumulator-free barrier (rank 0's, also for rank 1's epilogue)
    auto release_acc = [&](int acc) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(acc ? te1 : te0);
        else mbar_arrive(acc ? te1 : te0);
      }
    };
    for (int tile = blockIdx.x; tile_ok(tile); tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int pix = tile * TMO - ROW0 + row;
      // this thread's pixel is t - ROW0: same row of the image unless that underflows (then it is a border pixel)
      const bool is_out = has_row && pix < p.P;
      const int cn = wn, cy = wy, cx = wx - ROW0;
      const bool interior = is_out && cx >= 1 && cx <= p.Wp - 2 && wy >= 1 && wy <= p.Hp - 2;
      wx += p.step_x;
      if (wx >= p.Wp) { wx -= p.Wp; ++wy; }
      wy += p.step_y;
      if (wy >= p.Hp) { wy -= p.Hp; ++wn; }
      wn += p.step_n;
      const bool tr = trace && blockIdx.x == 0 && it < 32 && threadIdx.x == kEpiWarp0 * 32;
      if (tr) trace[3 * 32 + it] = clock64();
      if (ok) ok = mbar_wait(smem_u32(&bars->tfull[acc]), (it >> 1) & 1, p.err, 5);
      tc_fence_after();
      if (tr) trace[4 * 32 + it] = clock64();
      uint8_t* orow = optr + acc * O_TILE_BYTES + (row - ROW0) * 128;
      if (!active) {
        release_acc(acc);
        if (staged) {
          if (ok) ok = mbar_wait(smem_u32(&bars->ofree[acc]), ((it >> 1) & 1) ^ 1, p.err, 7);
          if (ok && p.has_residual) ok = mbar_wait(smem_u32(&bars->rfull[acc]), (it >> 1) & 1, p.err, 8);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->oready[acc]));
        }
        continue;
      }
      float f[CPT];
      if (!kFold) {
        uint32_t v1[CPT];
        tmem_ld_cols<CPT>(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * ACC_COLS + c0, v1);
        tmem_ld_wait();
        release_acc(acc);   // the accumulator is in registers: MMA may refill it
        if (tr) trace[5 * 32 + it] = clock64();
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float4 b4 = reinterpret_cast<const float4*>(bias_s + c0)[j];
          f[4 * j] = __uint_as_float(v1[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v1[4 * j + 1]) + b4.y;
          f[4 * j + 2] = __uint_as_float(v1[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v1[4 * j + 3]) + b4.w;
        }
      } else if (CPT == 16)
      {
        uint32_t v0[CPT], v1[CPT], v2[CPT];
        const uint32_t tbase = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * ACC_COLS + c0;
        float* xw = xch + (((acc * 4 + cq) * 4 + lg) * 2) * CPT;
        tmem_ld_32x16(tbase + n_cols, v1);        // slot 1: this pixel
        tmem_ld_32x16(tbase, v0);                 // slot 0: belongs to the pixel one lane up
        tmem_ld_32x16(tbase + 2 * n_cols, v2);    // slot 2: belongs to the pixel one lane down
        tmem_ld_wait();
        release_acc(acc);   // the accumulator is in registers: MMA may refill it
        if (tr) trace[5 * 32 + it] = clock64();
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j)
            reinterpret_cast<uint4*>(xw)[j] = make_uint4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j)
            reinterpret_cast<uint4*>(xw + CPT)[j] = make_uint4(v2[4 * j], v2[4 * j + 1], v2[4 * j + 2], v2[4 * j + 3]);
        }
        named_bar_sync(1 + cq, 128);   // the four warps of this column quarter have published their edge values
        if (tr) trace[6 * 32 + it] = clock64();
        // lane 31 / lane 0 swap in the neighbour warp's edge value, so that a ROTATING shuffle delivers the right
        // value to every lane (rows 0 and 127 of the tile are halo rows: whatever they receive is never stored)
        if (lane == 31 && lg > 0) {
          const uint4* src = reinterpret_cast<const uint4*>(xch + (((acc * 4 + cq) * 4 + lg - 1) * 2) * CPT);
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j) {
            const uint4 t = src[j];
            v0[4 * j] = t.x; v0[4 * j + 1] = t.y; v0[4 * j + 2] = t.z; v0[4 * j + 3] = t.w;
          }
        }
        if (lane == 0 && lg < 3) {
          const uint4* src = reinterpret_cast<const uint4*>(xch + (((acc * 4 + cq) * 4 + lg + 1) * 2 + 1) * CPT);
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j) {
            const uint4 t = src[j];
            v2[4 * j] = t.x; v2[4 * j + 1] = t.y; v2[4 * j + 2] = t.z; v2[4 * j + 3] = t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float4 b4 = reinterpret_cast<const float4*>(bias_s + c0)[j];
          f[4 * j] = __uint_as_float(v1[4 * j]) + b4.x; f[4 * j + 1] = __uint_as_float(v1[4 * j + 1]) + b4.y;
          f[4 * j + 2] = __uint_as_float(v1[4 * j + 2]) + b4.z; f[4 * j + 3] = __uint_as_float(v1[4 * j + 3]) + b4.w;
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          f[j] += __uint_as_float(__shfl_sync(0xffffffffu, v0[j], src_up));
          f[j] += __uint_as_float(__shfl_sync(0xffffffffu, v2[j], src_dn));
        }
      }
      if (tr) trace[7 * 32 + it] = clock64();
      if (dbg & 1) continue;
      if (partial_in && is_out) {   // fp32 partial sums of the earlier contraction chunks
        const float4* pin = reinterpret_cast<const float4*>(partial_in + (long long)pix * part_stride + c0);
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float4 q4 = __ldg(pin + j);
          f[4 * j] += q4.x; f[4 * j + 1] += q4.y; f[4 * j + 2] += q4.z; f[4 * j + 3] += q4.w;
        }
      }
      if (partial_out) {              // not the last chunk: keep fp32, no activation, nothing goes to y
        if (is_out) {
          float4* po = reinterpret_cast<float4*>(partial_out + (long long)pix * part_stride + c0);
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j) po[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        continue;
      }
      if (act == SRK_ACT_PRELU && p.zsave != nullptr && !(alpha > 0.f) && interior) {
        // rare path (see act_bwd_kernel): the backward cannot recover sign(z) / z from the output
        long long zo;
        if (shuffle == 2) {
          const int cop0 = p.cout_off + c0, sub = cop0 / p.cout_total, ch = cop0 - sub * p.cout_total;
          zo = (((long long)cn * p.Hp2 + (2 * (cy - 1) + (sub >> 1) + 1)) * p.Wp2 + (2 * (cx - 1) + (sub & 1) + 1)) *
                   p.cout_total + ch;
        } else {
          zo = (long long)pix * p.cout_total + p.cout_off + c0;
        }
        uint4* zd = reinterpret_cast<uint4*>(p.zsave + zo);
#pragma unroll
        for (int j = 0; j < CPT / 8; ++j)
          zd[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                             pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
      }
      if (act == SRK_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) f[j] = fmaxf(f[j], 0.f);
      } else if (act == SRK_ACT_PRELU) {
        // z > 0 ? z : alpha * z, written so that it also holds for alpha <= 0
#pragma unroll
        for (int j = 0; j < CPT; ++j) f[j] = fmaf(alpha, fminf(f[j], 0.f), fmaxf(f[j], 0.f));
      }
      if (shuffle == 2) {
        // PixelShuffle(2) as a store remap (models.py:118,121) over sub-pixel-major weight rows: this thread's CPT
        // columns are packed rows cop = sub * C + ch, i.e. CPT consecutive channels of output pixel
        // (2y + sub/2, 2x + sub%2)
        if (!interior) continue;
        const int cop0 = p.cout_off + c0, sub = cop0 / p.cout_total, ch = cop0 - sub * p.cout_total;
        const long long o2 =
            ((long long)cn * p.Hp2 + (2 * (cy - 1) + (sub >> 1) + 1)) * p.Wp2 + (2 * (cx - 1) + (sub & 1) + 1);
        uint4* dst = reinterpret_cast<uint4*>(p.y + o2 * p.cout_total + ch);
#pragma unroll
        for (int j = 0; j < CPT / 8; ++j)
          dst[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                              pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
        continue;
      }
      if (direct_out) {
        // wide pass: this thread's CPT channels of its pixel, straight from registers (border pixels as zeros)
        if (is_out) {
          const long long go = (long long)pix * p.cout_total + p.cout_off + c0;
          if (p.has_residual && interior) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + go);
#pragma unroll
            for (int j = 0; j < CPT / 8; ++j) {
              const uint4 rr = __ldg(rp + j);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
              for (int t = 0; t < 4; ++t) { const float2 u = __bfloat1622float2(h[t]); f[8 * j + 2 * t] += u.x; f[8 * j + 2 * t + 1] += u.y; }
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(p.y + go);
#pragma unroll
          for (int j = 0; j < CPT / 8; ++j) {
            uint4 o = make_uint4(0, 0, 0, 0);
            if (interior)
              o = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                             pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
            dst[j] = o;
          }
        }
        continue;
      }
      const int bn_red = kStats ? p.bn_red : 0;
      if (stats_sum && !bn_red && interior) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
      }
      if (tr) trace[8 * 32 + it] = clock64();
      if (ok) ok = mbar_wait(smem_u32(&bars->ofree[acc]), ((it >> 1) & 1) ^ 1, p.err, 7);
      if (tr) trace[10 * 32 + it] = clock64();
      if (p.has_residual && ok) ok = mbar_wait(smem_u32(&bars->rfull[acc]), (it >> 1) & 1, p.err, 8);
      if (p.has_residual && bn_red != 1 && has_row) {
#pragma unroll
        for (int j = 0; j < CPT / 8; ++j) {
          const uint4 rr = *reinterpret_cast<const uint4*>(orow + (((cq * (CPT / 8) + j) ^ ((row - ROW0) & 7)) << 4));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
          for (int t = 0; t < 4; ++t) { float2 u = __bfloat1622float2(h[t]); f[8 * j + 2 * t] += u.x; f[8 * j + 2 * t + 1] += u.y; }
        }
      }
      if (bn_red) {
        // bn_red = 1: the "residual" tile is Z (reduce, do not add); 2: Z has its own tile, the residual is already in f
        if (interior) {
          const uint8_t* zrow = bn_red == 2 ? orow + 2 * O_TILE_BYTES : orow;
#pragma unroll
          for (int j = 0; j < CPT / 8; ++j) {
            const uint4 rr = *reinterpret_cast<const uint4*>(zrow + (((cq * (CPT / 8) + j) ^ ((row - ROW0) & 7)) << 4));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
            const float4 sc0 = reinterpret_cast<const float4*>(bias_s + 128 + c0 + 8 * j)[0];
            const float4 sc1 = reinterpret_cast<const float4*>(bias_s + 128 + c0 + 8 * j)[1];
            const float4 sh0 = reinterpret_cast<const float4*>(bias_s + 192 + c0 + 8 * j)[0];
            const float4 sh1 = reinterpret_cast<const float4*>(bias_s + 192 + c0 + 8 * j)[1];
            const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
            const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 u = __bfloat1622float2(h[t]);
              const float zz[2] = {u.x, u.y};
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int k = 8 * j + 2 * t + e;
                float gd = f[k];
                if (p.bn_mask) {
                  const float b = fmaf(zz[e], scv[2 * t + e], shv[2 * t + e]);
                  if (b < 0.f) { bn_da = fmaf(gd, b, bn_da); gd *= bn_alpha; }
                }
                s1[k] += gd;
                s2[k] = fmaf(gd, zz[e], s2[k]);
              }
            }
          }
        }
      }
      if (tr) trace[9 * 32 + it] = clock64();
      if (has_row) {
        // bf16 tile staged in shared memory ([126 rows][128 B], SWIZZLE_128B) for one TMA store; border pixels are
        // stored as zeros (layout invariant), rows past the tensor are clipped by TMA
#pragma unroll
        for (int j = 0; j < CPT / 8; ++j) {
          uint4 o = make_uint4(0, 0, 0, 0);
          if (interior)
            o = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                           pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
          *reinterpret_cast<uint4*>(orow + (((cq * (CPT / 8) + j) ^ ((row - ROW0) & 7)) << 4)) = o;
        }
      }
      // every writer fences its own stores towards the async proxy; one arrival per warp
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->oready[acc]));
      if (tr) trace[11 * 32 + it] = clock64();
    }
    if (stats_sum) {
      // per-thread partial sums over this CTA's pixels -> per-channel totals: fold the two half-warps, then a
      // transposed butterfly (lane l ends up owning column l & 15), then one atomic per column
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
      }
#pragma unroll
      for (int half = CPT / 2; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
          float k1 = up ? s1[j + half] : s1[j], o1 = up ? s1[j] : s1[j + half];
          float k2 = up ? s2[j + half] : s2[j], o2 = up ? s2[j] : s2[j + half];
          s1[j] = k1 + __shfl_xor_sync(0xffffffffu, o1, half);
          s2[j] = k2 + __shfl_xor_sync(0xffffffffu, o2, half);
        }
      }
      // CTA partial in a fixed order: the four lane groups of a column through shared memory (the exchange buffer
      // is free once every epilogue warp has left the tile loop), then one ordered fold over the CTAs of the grid
      float* red = xch;                       // [4 lane groups][sum 64 | sumsq 64] | [16 warps] dalpha
      float* vals = xch + 528;
      const int et = threadIdx.x - kEpiWarp0 * 32;
      named_bar_sync(7, kEpiThreads);
      if (active && lane < CPT) {
        red[lg * 128 + c0 + lane] = s1[0];
        red[lg * 128 + 64 + c0 + lane] = s2[0];
      }
      {
        const float t = warp_sum(bn_da);
        if (lane == 0) red[512 + (warp - kEpiWarp0)] = t;
      }
      named_bar_sync(7, kEpiThreads);
      if (et < 128) {
        vals[et] = ((red[et] + red[128 + et]) + red[256 + et]) + red[384 + et];
      } else if (et == 128) {
        float t = 0.f;
        for (int w = 0; w < kEpiWarps; ++w) t += red[512 + w];
        vals[128] = t;
      }
      named_bar_sync(7, kEpiThreads);
      if (p.acc != nullptr) {
        if (et < 128 || (et == 128 && kStats && p.bn_red && p.bn_mask)) acc_add(p.acc, et, vals[et]);
      } else if (et < 256) {
        const bool want_da = kStats && p.bn_red && p.bn_mask && p.bn_dalpha != nullptr;
        const int n_ok = p.cout_total - p.cout_off;   // columns of this pass that exist in the tensor
        ordered_fold(vals, 129, p.red_ticket, (int)gridDim.x, (int)blockIdx.x, p.red_part, reinterpret_cast<float4*>(xch),
                     et, 256, [] { named_bar_sync(8, 256); },
                     [&](int i, float v) {
                       if (i < 64) { if (i < n_ok) stats_sum[p.cout_off + i] = v; }
                       else if (i < 128) { if (i - 64 < n_ok) stats_sumsq[p.cout_off + i - 64] = v; }
                       else if (want_da) p.bn_dalpha[0] = v;
                     });
      }
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all(); else __syncthreads();   // no CTA of a pair exits while its peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <bool kFold, bool kFast, bool kStats, int kAct, bool kPair, int kCPT = 16>
static void set_smem(int smem_max) {
  cudaFuncSetAttribute(conv3x3_fold_tc_kernel<kFold, kFast, kStats, kAct, kPair, kCPT>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
}
template <bool kFold, bool kPair>
static void set_smem_all(int smem_max) {
  set_smem<kFold, true, false, SRK_ACT_NONE, kPair>(smem_max);
  set_smem<kFold, true, true, SRK_ACT_NONE, kPair>(smem_max);
  set_smem<kFold, true, false, SRK_ACT_RELU, kPair>(smem_max);
  set_smem<kFold, true, false, SRK_ACT_PRELU, kPair>(smem_max);
  set_smem<kFold, false, false, ACT_RUNTIME, kPair>(smem_max);
}

template <bool kFold, bool kFast, bool kStats, int kAct, bool kPair, int kCPT = 16>
static cudaError_t launch_one(int grid, int smem_bytes, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmW,
                              const CUtensorMap& tmY, const CUtensorMap& tmRes, const CUtensorMap& tmZ, const Params& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (!kPair && pdl_enabled(PDL_CONV)) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, conv3x3_fold_tc_kernel<kFold, kFast, kStats, kAct, kPair, kCPT>, tmA, tmW, tmY, tmRes, tmZ, p);
}

template <bool kFold, bool kPair>
static cudaError_t launch_pass(bool fast, bool stats, int act, int grid, int smem_bytes, cudaStream_t st,
                               const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmY,
                               const CUtensorMap& tmRes, const CUtensorMap& tmZ, const Params& p) {
  if (fast && stats) return launch_one<kFold, true, true, SRK_ACT_NONE, kPair>(grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
  if (fast && act == SRK_ACT_NONE) return launch_one<kFold, true, false, SRK_ACT_NONE, kPair>(grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
  if (fast && act == SRK_ACT_RELU) return launch_one<kFold, true, false, SRK_ACT_RELU, kPair>(grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
  if (fast && act == SRK_ACT_PRELU) return launch_one<kFold, true, false, SRK_ACT_PRELU, kPair>(grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
  return launch_one<kFold, false, false, ACT_RUNTIME, kPair>(grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
}

}  // namespace fold

// Returns 0 ok, 1 error, -1 "not applicable" (image too wide for two slab stages: the caller uses the
// per-tap kernel of srk_conv_tc.cu).
int conv_fprop_fold_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout,
                           const float* bias, int act, const float* alpha, const srk_tensor* residual, int shuffle,
                           float* stats_sum, float* stats_sumsq, void* workspace, int variant, cudaStream_t st,
                           const BnRedArgs* br, void* reduce_ws, void* zsave, void* acc) {
  using namespace fold;
  SRK_REQUIRE((stats_sum == nullptr && br == nullptr) || reduce_ws != nullptr || acc != nullptr,
              "conv_fold: fused statistics need the reduce workspace or an accumulator");
  // accumulator mode: the kernel's "statistics on" switch is a non-null stats_sum
  if (acc != nullptr && br == nullptr && stats_sum == nullptr) { stats_sum = (float*)acc; stats_sumsq = (float*)acc; }
  // variant: 0 per-tap, 1 folded taps, 2 per-tap on CTA pairs (cta_group::2), 3 = CTA pairs with 128 output
  // channels per pass (PixelShuffle outputs of a single-chunk contraction: the 64 -> 256 upsample convs)
  // 4 = wide single-CTA passes: all output channels (64 < Cout <= 128, Cout % 32 == 0) in ONE pass per contraction chunk
  const bool folded = variant == 1, pair = variant == 2 || variant == 3, wide = variant == 4;
  const int pass_n = variant == 3 ? 2 * NT : (wide ? cout : NT);
  SRK_REQUIRE(!wide || (shuffle == 0 && cout > NT && cout <= 2 * NT && cout % 32 == 0 && stats_sum == nullptr && br == nullptr),
              "conv_fold: wide passes serve plain convs with 64 < Cout <= 128 output channels");
  SRK_REQUIRE(variant != 3 || (shuffle == 2 && x->c == KC && cout % pass_n == 0 && residual == nullptr && stats_sum == nullptr),
              "conv_fold: 128-channel passes serve single-chunk PixelShuffle convs with Cout %% 128 == 0");
  const int TMO = folded ? Tile<true>::TMO : Tile<false>::TMO;
  const int cin = x->c;
  const int Hp = x->h + 2, Wp = x->w + 2;
  const long long P = (long long)x->n * Hp * Wp;
  SRK_REQUIRE(P < (1LL << 31) - 4096, "conv_fold: too many pixels");
  static int smem_max = 0;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    set_smem_all<true, false>(smem_max);
    set_smem_all<false, false>(smem_max);
    set_smem_all<false, true>(smem_max);
    set_smem<false, false, false, ACT_RUNTIME, true, 32>(smem_max);
    set_smem<false, false, false, ACT_RUNTIME, false, 32>(smem_max);
  }
  const int slab_rows = ((TM + 2 * Wp + (folded ? 0 : 2)) + SLAB_BOX_ROWS - 1) / SLAB_BOX_ROWS * SLAB_BOX_ROWS;
  const int w_bytes = wide ? 9 * cout * KC * 2 : W_BYTES;
  const int fixed = 1024 + w_bytes + (wide ? 0 : 2 * O_TILE_BYTES) + ((br && residual) ? 2 * O_TILE_BYTES : 0) + XCH_BYTES +
                    BIAS_BYTES + (int)sizeof(Barriers);
  const int stage_bytes = slab_rows * KC * 2;
  int stages = (smem_max - fixed) / stage_bytes;
  if (stages < 2) return -1;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const int smem_bytes = fixed + stages * stage_bytes;

  CUtensorMap tmA;
  if (make_tmap_2d_bf16(&tmA, x->data, (uint64_t)P, (uint64_t)cin, (uint64_t)cin, SLAB_BOX_ROWS, KC, 128)) return 1;
  CUtensorMap tmY = tmA, tmR = tmA, tmZ = tmA;  // output store / residual load / Z load maps (plain outputs only)
  if (shuffle == 0) {
    if (make_tmap_2d_bf16(&tmY, y->data, (uint64_t)P, (uint64_t)cout, (uint64_t)cout, TMO, NT, 128)) return 1;
    tmR = tmY;
    if (residual && make_tmap_2d_bf16(&tmR, residual->data, (uint64_t)P, (uint64_t)cout, (uint64_t)cout, TMO, NT, 128))
      return 1;
  }

  Params p;
  p.P = (int)P; p.Hp = Hp; p.Wp = Wp;
  p.num_tiles = (int)((P + TMO - 1) / TMO);
  p.w_row_per_tap = cout;
  p.slab_rows = slab_rows; p.stages = stages; p.stage_bytes = stage_bytes;
  p.alpha = alpha;
  p.y = (__nv_bfloat16*)y->data;
  p.shuffle = shuffle;
  p.Hp2 = y->h + 2; p.Wp2 = y->w + 2;
  p.err = tc_err_flag();
  p.dbg = tc_dbg();
  p.trace = g_tc_trace;
  p.stats_sum = stats_sum; p.stats_sumsq = stats_sumsq;
  p.red_ticket = reduce_ws ? red_tickets(reduce_ws) : nullptr;
  p.red_part = reduce_ws ? red_partials(reduce_ws) : nullptr;
  p.acc = (unsigned long long*)acc;
  p.w_bytes = w_bytes; p.direct_out = wide ? 1 : 0; p.part_stride = wide ? cout : NT;
  {
    static int two = -1;   // SRK_TC_MMA2=0: a single MMA-issuing warp (A/B measurements)
    if (two < 0) { const char* e = getenv("SRK_TC_MMA2"); two = e ? atoi(e) != 0 : 1; }
    p.mma_warps = (two && stages >= 3) ? 2 : 1;
  }
  p.res = residual ? (const __nv_bfloat16*)residual->data : nullptr;
  p.zsave = (__nv_bfloat16*)zsave;
  p.bn_red = 0; p.bn_mask = 0;
  p.bn_mean = p.bn_invstd = p.bn_gamma = p.bn_beta = nullptr; p.bn_dalpha = nullptr;
  if (br) {
    SRK_REQUIRE(cin == KC && cout == NT && shuffle == 0 && act == SRK_ACT_NONE && stats_sum == nullptr && variant == 0,
                "conv_fold: the fused BN-backward reduction covers the plain 64 -> 64 per-tap dgrad");
    SRK_REQUIRE(same_geometry(br->z, y) && br->z->dtype == SRK_BF16 && br->z->layout == SRK_LAYOUT_ACT,
                "conv_fold: Z must match the dgrad output geometry (bf16 ACT)");
    // without a residual Z travels in the residual slot; with one it gets its own map and staging tiles
    if (make_tmap_2d_bf16(residual ? &tmZ : &tmR, br->z->data, (uint64_t)P, (uint64_t)cout, (uint64_t)cout, TMO, NT, 128))
      return 1;
    p.stats_sum = stats_sum = acc ? (float*)acc : br->sum_g; p.stats_sumsq = stats_sumsq = acc ? (float*)acc : br->sum_gz;
    p.bn_red = residual ? 2 : 1; p.bn_mask = br->alpha != nullptr; p.alpha = br->alpha;
    p.bn_mean = br->mean; p.bn_invstd = br->invstd; p.bn_gamma = br->gamma; p.bn_beta = br->beta;
    p.bn_dalpha = br->dalpha;
  }
  const int nchunks = (cout + pass_n - 1) / pass_n, kchunks = (cin + KC - 1) / KC;
  SRK_REQUIRE(stats_sum == nullptr || (kchunks == 1 && shuffle == 0 && act == SRK_ACT_NONE && (residual == nullptr || br)),
              "conv_fold: fused BN statistics need a plain Cin == 64 conv");
  int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  if (pair) grid = (grid + 1) / 2 * 2 <= kNumSMs ? (grid + 1) / 2 * 2 : kNumSMs / 2 * 2;   // whole pairs
  {
    const long long step = (long long)grid * TMO, img = (long long)Hp * Wp;
    p.step_n = (int)(step / img);
    p.step_y = (int)((step % img) / Wp);
    p.step_x = (int)((step % img) % Wp);
  }
  for (int nc = 0; nc < nchunks; ++nc) {
    const int n_cols = cout - nc * pass_n < pass_n ? cout - nc * pass_n : pass_n;
    CUtensorMap tmW;
    if (make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)9 * cout, (uint64_t)cin, (uint64_t)cin, pair ? n_cols / 2 : n_cols, KC,
                          128))
      return 1;
    for (int kc = 0; kc < kchunks; ++kc) {
      const bool first = kc == 0, last = kc == kchunks - 1;
      p.k_col0 = kc * KC;
      p.w_row0 = nc * pass_n;
      p.cout_total = y->c;
      p.cout_off = nc * pass_n;
      p.bias_off = nc * pass_n;
      p.n_cols = n_cols;
      p.ksteps = (cin - kc * KC < KC ? cin - kc * KC : KC) / 16;
      p.bias = first ? bias : nullptr;
      p.act = last ? act : SRK_ACT_NONE;
      // chunked contractions: see srk_conv_tc.cu (bf16 partial sums through y without an activation, fp32 partial
      // sums through the workspace when an activation / PixelShuffle follows)
      // (wide passes have no staged tile to carry a bf16 partial sum: always fp32 through the workspace)
      const bool fp32_partials = kchunks > 1 && (act != SRK_ACT_NONE || shuffle != 0 || wide);
      if (fp32_partials) {
        p.has_residual = (last && residual) ? 1 : 0;
        p.partial_out = last ? nullptr : (float*)workspace;
        p.partial_in = first ? nullptr : (const float*)workspace;
        SRK_REQUIRE(workspace != nullptr, "conv_fold: Cin > 64 with an activation needs the fprop workspace");
      } else {
        p.has_residual = first ? ((residual || br) ? 1 : 0) : 1;
        p.partial_out = nullptr;
        p.partial_in = nullptr;
      }
      const CUtensorMap& tmRes = (fp32_partials || first) ? tmR : tmY;
      const bool fast = kchunks == 1 && p.n_cols == NT && p.ksteps == KC / 16 && shuffle == 0 && p.dbg == 0 &&
                        p.trace == nullptr;
      SRK_REQUIRE(fast || stats_sum == nullptr, "conv_fold: fused BN statistics need the single-chunk 64 -> 64 pass");
      cudaError_t le;
      if (variant == 3)
        le = launch_one<false, false, false, ACT_RUNTIME, true, 32>(grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
      else if (wide)
        le = launch_one<false, false, false, ACT_RUNTIME, false, 32>(grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
      else if (folded) le = launch_pass<true, false>(fast, stats_sum != nullptr, p.act, grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
      else if (pair) le = launch_pass<false, true>(fast, stats_sum != nullptr, p.act, grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
      else le = launch_pass<false, false>(fast, stats_sum != nullptr, p.act, grid, smem_bytes, st, tmA, tmW, tmY, tmRes, tmZ, p);
      SRK_REQUIRE(le == cudaSuccess, "conv3x3_fold_tc: launch failed: %s", cudaGetErrorString(le));
      SRK_CUDA_LAUNCH_CHECK("conv3x3_fold_tc");
    }
  }
  if (shuffle == 2) return zero_border(y, st);
  return 0;
}

}  // namespace srk
