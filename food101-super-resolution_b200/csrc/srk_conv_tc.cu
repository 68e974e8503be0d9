// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution (fprop, and dgrad through rotated weights) on the
// zero-bordered channels-last bf16 activation layout.
//
// GEMM view (reference: nn.Conv2d 3x3 s1 p1 at models.py:46,49,65,67,113,117,120):
//   Y[p, co] = sum_{tap, ci} X[p + d(tap), ci] * W[tap][co][ci],   d(tap) = (r-1)*(W+2) + (s-1)
// where p runs over ALL padded pixels [0, N*(H+2)*(W+2)): because the border is zero, a tap is a constant
// shift of the flat pixel index, so any 128 consecutive padded pixels form a legal M tile and the A operand
// of a tap is the same [rows][64ch] matrix shifted by d(tap) rows.  Border pixels compute garbage that the
// epilogue replaces by zeros (which keeps the zero-border invariant without a second pass).
//
// One persistent CTA per SM, 320 threads:
//   warp 0   TMA producer  (A tiles / halo slabs -> smem ring, weights once)
//   warp 1   MMA issuer    (one thread: tcgen05.mma M128 N64 K16, fp32 accumulators in TMEM, double buffered)
//   warps 2-9 epilogue     (tcgen05.ld -> +bias, ReLU/PReLU, BN statistics, +residual, bf16 -> global,
//                           PixelShuffle remap); two warps share a TMEM lane group and split the 64 columns
//
// A-operand staging modes (runtime `mode`):
//   0  one TMA load of [128 x 64ch] per tap (9 loads per tile; every MMA operand 1024-B aligned)
//   1  one halo slab [128 + 2(W+2) + 2 rows x 64ch] per tile; taps are row-shifted descriptors into it
//   2  as 1, with the descriptor base_offset field set from the (non-1024-B-aligned) start address
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

#include <cstdlib>
#include <mutex>

namespace srk {

using namespace tc;

// ---- host: tensor maps ------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                      uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode_tiled();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box=%ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
  return 0;
}

// ---- device --------------------------------------------------------------------------------------------
constexpr int TM = 128;              // pixels per tile (UMMA M)
constexpr int NT = 64;               // output channels per CTA pass (UMMA N)
constexpr int KC = 64;               // contraction channels per pass: one 128-byte swizzle row
constexpr int TAPS = 9;
constexpr int W_TILE_BYTES = NT * KC * 2;       // 8 KB per tap
constexpr int A_TILE_BYTES = TM * KC * 2;       // 16 KB
constexpr int SLAB_BOX_ROWS = 32;
constexpr int kThreads = 352;           // warps: 0 TMA, 1 MMA, 2-9 epilogue, 10 output store / residual load
constexpr int O_TILE_BYTES = TM * NT * 2;  // bf16 output tile staged for the TMA store
constexpr int MAX_STAGES = 8;
constexpr int RED_BYTES = 4096;

struct TcConvParams {
  int P, Hp, Wp, num_tiles;
  int k_col0;          // first contraction channel of this pass (column coordinate in x and in the weights)
  int w_row_per_tap;   // rows per tap in the packed weight matrix (= total output channels)
  int w_row0;          // first weight row of this pass (= output-channel offset in packed order)
  int cout_total;      // channels per pixel of y (row stride)
  int cout_off;        // channel offset inside a y row
  int act, shuffle, sub;
  int mode, slab_rows, stages, stage_bytes;
  int taps;            // 9 (3x3) or 1 (1x1, staged like mode 0)
  int n_cols;          // output channels of this pass: 64 or 32 (UMMA N)
  int ksteps;          // 16-channel K steps of this pass: 4 (64 channels) or 2 (32 channels)
  int res_first;       // (unused) 
  float* partial_out;  // chunked contraction (Cin > 64): fp32 partial sums [P][64] written instead of y ...
  const float* partial_in;  // ... and added back (before the activation) by the next contraction chunk
  const float* bias;   // indexed [bias_off + c] or null
  int bias_off, bias_stride;   // bias index of column c = bias_off + c * bias_stride
  const float* alpha;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  int Hp2, Wp2;        // padded sizes of the shuffled output
  float* stats_sum;    // fused BatchNorm statistics: per-channel sum / sum of squares of the fp32
  float* stats_sumsq;  // outputs over interior pixels (written), or null
  unsigned* red_ticket;   // cross-CTA stage of the statistics (ordered_fold, srk_common.cuh)
  float* red_part;
  __nv_bfloat16* zsave;   // PReLU with a slope <= 0: copy of the pre-activation (y geometry) for the backward pass
  int* err;
  int dbg;             // bring-up knobs: 1 skip stores, 2 skip MMAs, 4 skip A loads
  long long* trace;    // bring-up: per-tile clock64 stamps of CTA 0 ([6][32]) or null
};

struct __align__(8) TcBarriers {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], wfull, tfull[2], tempty[2];
  uint64_t oready[2], ofree[2], rfull[2];   // output staging tiles: written / drained / residual landed
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// kFast: the common single-chunk 64 -> 64 pass (no partial sums, N = 64, four K steps) with those choices
// compiled in; the general instantiation covers 32-wide tails and chunked contractions.
template <bool kFast, bool kStats>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                  const TcConvParams pp) {
  // in the fast instantiation the chunking parameters are compile-time constants
  const TcConvParams& p = pp;
  const int n_cols = kFast ? NT : pp.n_cols;
  const int ksteps = kFast ? KC / 16 : pp.ksteps;
  float* const partial_out = kFast ? nullptr : pp.partial_out;
  const float* const partial_in = kFast ? nullptr : pp.partial_in;
  const int dbg = kFast ? 0 : pp.dbg;
  long long* const trace = kFast ? nullptr : pp.trace;
  float* const stats_sum = kStats ? pp.stats_sum : nullptr;
  float* const stats_sumsq = kStats ? pp.stats_sumsq : nullptr;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem: [weights 9 x 8 KB][A ring stages x stage_bytes][2 output tiles x 16 KB][barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t wsm = smem_base;
  const uint32_t asm0 = smem_base + TAPS * W_TILE_BYTES;
  const uint32_t osm = asm0 + p.stages * p.stage_bytes;
  uint8_t* optr = smem_al + TAPS * W_TILE_BYTES + p.stages * p.stage_bytes;
  float* redp = reinterpret_cast<float*>(optr + 2 * O_TILE_BYTES);   // RED_BYTES: statistics fold scratch
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(optr + 2 * O_TILE_BYTES + RED_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->wfull), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), 256);
      mbar_init(smem_u32(&bars->oready[i]), 256); mbar_init(smem_u32(&bars->ofree[i]), 1);
      mbar_init(smem_u32(&bars->rfull[i]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bars->tmem_base), 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================= TMA producer =================
    // The whole warp runs the loop (warp-uniform control flow keeps addresses / coordinates in uniform
    // registers); one elected lane issues the TMA instructions.
    if (elect_one()) {
      prefetch_tmap(&tmA);
      prefetch_tmap(&tmW);
      const uint32_t wbar = smem_u32(&bars->wfull);
      mbar_arrive_expect_tx(wbar, p.taps * W_TILE_BYTES);
      for (int t = 0; t < p.taps; ++t)
        tma_load_2d(wsm + t * W_TILE_BYTES, &tmW, wbar, p.k_col0, t * p.w_row_per_tap + p.w_row0);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
      const int m0 = tile * TM;
      if (p.mode == 0) {
        for (int t = 0; t < p.taps && ok; ++t) {
          ok = mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 1);
          if (!ok) break;
          if (elect_one()) {
            const uint32_t fb = smem_u32(&bars->full[s]);
            mbar_arrive_expect_tx(fb, A_TILE_BYTES);
            const int d = p.taps == 1 ? 0 : (t / 3 - 1) * p.Wp + (t % 3 - 1);
            tma_load_2d(asm0 + s * p.stage_bytes, &tmA, fb, p.k_col0, m0 + d);
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
      } else {
        ok = mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 1);
        if (!ok) break;
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->full[s]);
          if (dbg & 4) {
            mbar_arrive(fb);
          } else {
            mbar_arrive_expect_tx(fb, p.slab_rows * KC * 2);
            const int row0 = m0 - p.Wp - 1;
            for (int j = 0; j < p.slab_rows / SLAB_BOX_ROWS; ++j)
              tma_load_2d(asm0 + s * p.stage_bytes + j * SLAB_BOX_ROWS * KC * 2, &tmA, fb, p.k_col0,
                          row0 + j * SLAB_BOX_ROWS);
          }
          if (trace && blockIdx.x == 0 && tile / (int)gridDim.x < 32) trace[0 * 32 + tile / gridDim.x] = clock64();
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // Whole warp, warp-uniform; tcgen05.mma / commit are issued by one elected lane.  Descriptors are built
    // once: the upper word (SBO 1024 B, version, SWIZZLE_128B) is constant, the lower word is
    // (address >> 4) | (LBO >> 4) << 16, so stepping K by 16 elements (32 B) or moving to another tap / stage
    // is a 32-bit add on the lower word.
    const uint32_t idesc = make_idesc_bf16(TM, n_cols, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_base = (uint32_t)(make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t w_lo = lo_base + (wsm >> 4), a_lo0 = lo_base + (asm0 >> 4);
    const uint32_t row_units = (uint32_t)p.Wp * (KC * 2 / 16);   // one image row of the slab, in 16-byte units
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4;
    bool ok = mbar_wait(smem_u32(&bars->wfull), 0, p.err, 2);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      ok = mbar_wait(smem_u32(&bars->tempty[acc]), ((it >> 1) & 1) ^ 1, p.err, 3);
      if (!ok) break;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * NT;
      if (p.mode == 0) {
#pragma unroll 1
        for (int t = 0; t < p.taps; ++t) {
          ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 4);
          if (!ok) break;
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * stage_units, b_lo = w_lo + t * (W_TILE_BYTES / 16);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              if (ks < ksteps)
                umma_bf16(d_tmem, desc_hi | (a_lo + 2 * ks), desc_hi | (b_lo + 2 * ks), idesc, (t | ks) != 0);
            umma_commit(smem_u32(&bars->empty[s]));
          }
          __syncwarp();
          if (++s == S) { s = 0; ph ^= 1; }
        }
        if (!ok) break;
      } else {
        ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 4);
        if (!ok) break;
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * stage_units;
        if (elect_one()) {
          if (trace && blockIdx.x == 0 && it < 32) trace[1 * 32 + it] = clock64();
          if (!(dbg & 2)) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks)
                  if (ks < ksteps)
                    umma_bf16(d_tmem, desc_hi | (a_lo + r * row_units + c * (KC * 2 / 16) + 2 * ks),
                              desc_hi | (w_lo + (r * 3 + c) * (W_TILE_BYTES / 16) + 2 * ks), idesc, (r | c | ks) != 0);
          }
          umma_commit(smem_u32(&bars->empty[s]));
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1; }
      }
      if (elect_one()) {
        umma_commit(smem_u32(&bars->tfull[acc]));
        if (trace && blockIdx.x == 0 && it < 32) trace[2 * 32 + it] = clock64();
      }
      __syncwarp();
    }
  } else if (warp == 10) {
    // ================= output store / residual load warp (plain, non-PixelShuffle outputs) =================
    // tile `it` uses staging buffer b = it & 1: [residual TMA load ->] epilogue writes -> TMA store -> free
    if (p.shuffle == 0 && partial_out == nullptr) {
      const int my_tiles = blockIdx.x < p.num_tiles ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      if (p.residual && elect_one()) {
        prefetch_tmap(&tmR);
        for (int it = 0; it < 2 && it < my_tiles; ++it) {
          const uint32_t rb = smem_u32(&bars->rfull[it]);
          mbar_arrive_expect_tx(rb, O_TILE_BYTES);
          tma_load_2d(osm + it * O_TILE_BYTES, &tmR, rb, p.cout_off, (blockIdx.x + it * gridDim.x) * TM);
        }
      }
      __syncwarp();
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        if (!mbar_wait(smem_u32(&bars->oready[b]), (it >> 1) & 1, p.err, 6)) break;
        if (elect_one()) {
          tma_store_2d(&tmY, osm + b * O_TILE_BYTES, p.cout_off, (blockIdx.x + it * gridDim.x) * TM);
          tma_store_commit();
          tma_store_wait_read0();               // the staging tile has been read out: it may be refilled
          if (p.residual && it + 2 < my_tiles) {
            const uint32_t rb = smem_u32(&bars->rfull[b]);
            mbar_arrive_expect_tx(rb, O_TILE_BYTES);
            tma_load_2d(osm + b * O_TILE_BYTES, &tmR, rb, p.cout_off, (blockIdx.x + (it + 2) * gridDim.x) * TM);
          }
          mbar_arrive(smem_u32(&bars->ofree[b]));
        }
        __syncwarp();
      }
      if (elect_one()) tma_store_wait_all();
      __syncwarp();
    }
  } else {
    // ================= epilogue: 8 warps, warp e owns TMEM lanes 32*(e&3).. and 32 of the 64 columns ====
    const int e = warp - 2, lg = warp & 3, ch = e >> 2;   // warps 2..9 -> lane groups 2,3,0,1,2,3,0,1
    const int c0 = ch * 32;                                // first accumulator column of this thread
    const float alpha = (p.act == SRK_ACT_PRELU) ? __ldg(p.alpha) : 0.f;
    const int img = p.Hp * p.Wp;
    const bool active = c0 < n_cols;   // a 32-column pass leaves the second column half idle (it still signals)
    // PixelShuffle passes run over sub-pixel-major weight rows (SRK_PACK_FPROP_TC with pixel_shuffle = 2): column j of
    // this thread is packed row cop = cout_off + c0 + j = sub * C + c (C = channels of y), i.e. reference channel
    // 4c + sub: one sub-pixel per thread and pass, 32 consecutive output channels
    const int ps_cop0 = p.cout_off + c0, ps_sub = p.shuffle == 2 ? ps_cop0 / p.cout_total : 0;
    const int ps_c = ps_cop0 - ps_sub * p.cout_total;
    float bias[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int bi = p.shuffle == 2 ? 4 * (ps_c + j) + ps_sub : p.bias_off + c0 + j;
      bias[j] = (p.bias && active) ? __ldg(p.bias + bi) : 0.f;
    }
    float s1[32], s2[32];
    if (stats_sum) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    }
    bool ok = true;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int pix = tile * TM + lg * 32 + lane;
      const int pc = pix < p.P ? pix : 0;
      const int n = pc / img, q = pc - n * img;
      const int yy = q / p.Wp, xx = q - yy * p.Wp;
      const bool interior = pix < p.P && yy >= 1 && yy <= p.Hp - 2 && xx >= 1 && xx <= p.Wp - 2;
      ok = mbar_wait(smem_u32(&bars->tfull[acc]), (it >> 1) & 1, p.err, 5);
      if (!ok) break;
      tc_fence_after();
      const bool tr = trace && blockIdx.x == 0 && it < 32 && threadIdx.x == 64;
      if (tr) trace[3 * 32 + it] = clock64();
      uint32_t v[32];
      if (active) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * NT + c0, v);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->tempty[acc]));
      if (tr) trace[4 * 32 + it] = clock64();
      const int row = lg * 32 + lane;
      uint8_t* orow = optr + acc * O_TILE_BYTES + row * 128;
      if (!active) {
        // nothing to compute or store; keep the staging-buffer handshake going
        if (p.shuffle == 0 && partial_out == nullptr) {
          if (!mbar_wait(smem_u32(&bars->ofree[acc]), ((it >> 1) & 1) ^ 1, p.err, 7)) break;
          if (p.residual && !mbar_wait(smem_u32(&bars->rfull[acc]), (it >> 1) & 1, p.err, 8)) break;
          mbar_arrive(smem_u32(&bars->oready[acc]));
        }
        continue;
      }
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + bias[j];
      if (partial_in && pix < p.P) {   // fp32 partial sums of the earlier contraction chunks
        const float4* pin = reinterpret_cast<const float4*>(partial_in + (long long)pix * NT + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 q4 = __ldg(pin + j);
          f[4 * j] += q4.x; f[4 * j + 1] += q4.y; f[4 * j + 2] += q4.z; f[4 * j + 3] += q4.w;
        }
      }
      if (partial_out) {              // not the last chunk: keep fp32, no activation, nothing goes to y
        if (pix < p.P) {
          float4* po = reinterpret_cast<float4*>(partial_out + (long long)pix * NT + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) po[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
        continue;
      }
      if (p.shuffle == 0 && !mbar_wait(smem_u32(&bars->ofree[acc]), ((it >> 1) & 1) ^ 1, p.err, 7)) break;
      if (p.act == SRK_ACT_PRELU && p.zsave != nullptr && !(alpha > 0.f) && interior) {
        // rare path (see act_bwd_kernel): the backward cannot recover sign(z) / z from the output
        long long zo;
        if (p.shuffle == 2)
          zo = (((long long)n * p.Hp2 + (2 * (yy - 1) + (ps_sub >> 1) + 1)) * p.Wp2 + (2 * (xx - 1) + (ps_sub & 1) + 1)) *
                   p.cout_total + ps_c;
        else
          zo = (long long)pix * p.cout_total + p.cout_off + c0;
        uint4* zd = reinterpret_cast<uint4*>(p.zsave + zo);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          zd[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                             pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float a = f[j];
        if (p.act == SRK_ACT_RELU) a = fmaxf(a, 0.f);
        else if (p.act == SRK_ACT_PRELU) a = a > 0.f ? a : alpha * a;
        f[j] = a;
      }
      if (p.shuffle == 2) {
        // PixelShuffle(2) as a store remap (models.py:118,121): this thread's 32 columns are channels ps_c .. ps_c+31
        // of output pixel (2y + sub/2, 2x + sub%2): one 64-byte run
        if (!interior) continue;
        const long long orow =
            ((long long)n * p.Hp2 + (2 * (yy - 1) + (ps_sub >> 1) + 1)) * p.Wp2 + (2 * (xx - 1) + (ps_sub & 1) + 1);
        uint4* dst = reinterpret_cast<uint4*>(p.y + orow * p.cout_total + ps_c);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                              pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
        continue;
      }
      // stage the bf16 tile in shared memory ([128 rows][128 B], SWIZZLE_128B) for one coalesced TMA store;
      // border pixels are stored as zeros (layout invariant), rows past the tensor are clipped by TMA
      if (interior && stats_sum) {
#pragma unroll
        for (int j = 0; j < 32; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
      }
      if (p.residual) {
        if (!mbar_wait(smem_u32(&bars->rfull[acc]), (it >> 1) & 1, p.err, 8)) break;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 rr = *reinterpret_cast<const uint4*>(orow + (((ch * 4 + j) ^ (row & 7)) << 4));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
          for (int t = 0; t < 4; ++t) { float2 u = __bfloat1622float2(h[t]); f[8 * j + 2 * t] += u.x; f[8 * j + 2 * t + 1] += u.y; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o = make_uint4(0, 0, 0, 0);
        if (interior)
          o = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                         pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
        *reinterpret_cast<uint4*>(orow + (((ch * 4 + j) ^ (row & 7)) << 4)) = o;
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->oready[acc]));
      if (tr) trace[5 * 32 + it] = clock64();
    }
    if (stats_sum) {
      // per-thread partial sums over this CTA's pixels -> per-channel totals: transposed butterfly (lane l
      // ends up owning column l), then one atomic per lane
#pragma unroll
      for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
          // keep column block [j] if !up, [j + half] if up; send the other one to the partner lane
          float k1 = up ? s1[j + half] : s1[j], o1 = up ? s1[j] : s1[j + half];
          float k2 = up ? s2[j + half] : s2[j], o2 = up ? s2[j] : s2[j + half];
          s1[j] = k1 + __shfl_xor_sync(0xffffffffu, o1, half);
          s2[j] = k2 + __shfl_xor_sync(0xffffffffu, o2, half);
        }
      }
      // after the butterfly lane l holds the total of column c0 + l over this warp's rows.  CTA partial in a fixed
      // order (the four lane groups of a column through shared memory), then one ordered fold over the grid.
      float* red = redp;               // [4 lane groups][sum 64 | sumsq 64]
      float* vals = redp + 512;
      const int et = threadIdx.x - 64;
      if (active) {
        red[lg * 128 + c0 + lane] = s1[0];
        red[lg * 128 + 64 + c0 + lane] = s2[0];
      }
      asm volatile("bar.sync 7, 256;" ::: "memory");
      if (et < 128) vals[et] = ((red[et] + red[128 + et]) + red[256 + et]) + red[384 + et];
      asm volatile("bar.sync 7, 256;" ::: "memory");
      const int n_ok = p.cout_total - p.cout_off;
      ordered_fold(vals, 128, p.red_ticket, (int)gridDim.x, (int)blockIdx.x, p.red_part, reinterpret_cast<float4*>(redp),
                   et, 256, [] { asm volatile("bar.sync 7, 256;" ::: "memory"); },
                   [&](int i, float v) {
                     if (i < 64) { if (i < n_ok) stats_sum[p.cout_off + i] = v; }
                     else if (i - 64 < n_ok) stats_sumsq[p.cout_off + i - 64] = v;
                   });
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ---- host launcher ---------------------------------------------------------------------------------------
static int g_tc_mode = -1;
long long* g_tc_trace = nullptr;
void tc_set_trace(long long* t) { g_tc_trace = t; }
int tc_mode() {
  if (g_tc_mode < 0) {
    const char* e = getenv("SRK_TC_MODE");
    g_tc_mode = e ? atoi(e) : 1;  // slab staging: validated against the oracle on B200 (mode 2 is wrong)
    if (g_tc_mode < 0 || g_tc_mode > 2) g_tc_mode = 1;
  }
  return g_tc_mode;
}
void tc_set_mode(int m) { g_tc_mode = m; }

// Bring-up knobs of the general kernel instantiations (skip stores / MMAs / loads: results are wrong by design).
// Compiled out of release builds; with -DSRK_DEBUG_KNOBS the environment variable SRK_TC_DBG is read once.
int tc_dbg() {
#ifdef SRK_DEBUG_KNOBS
  static int v = -1;
  if (v < 0) { const char* e = getenv("SRK_TC_DBG"); v = e ? atoi(e) : 0; }
  return v;
#else
  return 0;
#endif
}

int* tc_err_flag() {
  // one device int per process per device; checked lazily by the probe (never synchronises the hot path)
  static int* flag[16] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return nullptr;
  if (!flag[dev]) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!flag[dev]) {
      int* f = nullptr;
      if (cudaMalloc(&f, sizeof(int)) != cudaSuccess) return nullptr;
      cudaMemset(f, 0, sizeof(int));
      flag[dev] = f;
    }
  }
  return flag[dev];
}
int tc_read_err_flag() {
  int* f = tc_err_flag();
  int v = 0;
  if (f) cudaMemcpy(&v, f, sizeof(int), cudaMemcpyDeviceToHost);
  return v;
}

// 3x3 kernel choice: 0 = per-tap kernel of this file (8 epilogue warps), 1 = folded taps, 2 = per-tap on the 16-warp
// pipeline of srk_conv_fold_tc.cu, 3 = as 2 on CTA pairs (cta_group::2, M = 256)
static int g_tc_fold = -1;
static int g_up_pair = -1;
static int g_tc_wide = -1;   // 64 < Cout <= 128 in one output-channel pass (SRK_TC_WIDE, srk_tc_probe 40 / 41)
int tc_fold() {
  if (g_tc_fold < 0) {
    const char* e = getenv("SRK_TC_FOLD");
    g_tc_fold = e ? atoi(e) : 2;   // measured (C2 layer): fprop 23.2 / 24.4, fprop+stats 24.8 / 27.8, dgrad+res 25.8 / 27.1 us (2 / 0)
    if (g_tc_fold < 0 || g_tc_fold > 4) g_tc_fold = 2;
  }
  return g_tc_fold;
}

int64_t conv_fprop_tc_workspace(const srk_tensor* x) {
  if (x->c <= KC) return 0;  // (needed when an activation or PixelShuffle follows a chunked contraction)
  // up to 128 columns per pixel: the wide passes of srk_conv_fold_tc.cu keep all output channels of a <= 128-channel conv
  return (int64_t)x->n * (x->h + 2) * (x->w + 2) * 2 * NT * (int64_t)sizeof(float);
}

bool conv_tc_shape_ok(int cin, int cout, int r, int s, int dtype, int shuffle) {
  if (dtype != SRK_BF16 || r != s || (r != 3 && r != 1)) return false;
  // channels are processed in chunks of 64 with a 32-wide tail (96 = 64 + 32 for AttentionSR)
  if (cin % 32 != 0 || cout % 32 != 0 || cin < 64 || cout < 64) return false;
  if (shuffle != 0 && !(shuffle == 2 && cout == 4 * NT)) return false;   // one sub-pixel per 64-row pass
  return true;
}

int conv_fprop_tc_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout, int r, int s,
                         const float* bias, int act, const float* alpha, const srk_tensor* residual, int shuffle,
                         float* stats_sum, float* stats_sumsq, void* workspace, cudaStream_t st, void* reduce_ws,
                         void* zsave, void* acc) {
  SRK_REQUIRE(stats_sum == nullptr || reduce_ws != nullptr, "conv_tc: fused statistics need the reduce workspace");
  if (acc != nullptr) {
    // statistics into an exact integer accumulator: the 16-warp halo-slab kernel's single-pass 64 -> 64 conv only;
    // 2 = not covered, nothing launched (the caller uses the float path)
    if (!(r == 3 && tc_fold() >= 1 && tc_fold() <= 3 && shuffle == 0 && x->c == 64 && cout == 64 && act == SRK_ACT_NONE &&
          residual == nullptr && stats_sum == nullptr))
      return 2;
    const int rc = conv_fprop_fold_launch(x, y, w_packed, cout, bias, act, alpha, residual, shuffle, nullptr, nullptr,
                                          workspace, tc_fold() == 1 ? 1 : (tc_fold() == 3 ? 2 : 0), st, nullptr, nullptr,
                                          zsave, acc);
    return rc < 0 ? 2 : rc;
  }
  // PixelShuffle outputs stay on the 8-warp kernel below: its threads own 32 channels = 16-byte stores per sub-pixel,
  // the 16-warp pipeline would store 8 bytes at a time (measured 64->256 at 128^2: 424 vs 685 us)
  // 64 -> 256 PixelShuffle convs on CTA pairs with N = 128 MMAs (SRK_TC_UP_PAIR=1): parity-green but slower (128^2:
  // 513 vs 421 us) - these convs are bound by their scattered 16-byte PixelShuffle stores, not by the MMAs
  if (g_up_pair < 0) { const char* e = getenv("SRK_TC_UP_PAIR"); g_up_pair = e ? atoi(e) != 0 : 0; }
  if (r == 3 && g_up_pair && shuffle == 2 && x->c == 64 && cout % 128 == 0 && residual == nullptr &&
      stats_sum == nullptr) {
    const int rc = conv_fprop_fold_launch(x, y, w_packed, cout, bias, act, alpha, residual, shuffle, nullptr, nullptr,
                                          workspace, 3, st, nullptr, nullptr, zsave);
    if (rc >= 0) return rc;
  }
  // 64 < Cout <= 128 (the 96-channel trunk of AttentionSR): every output channel in ONE pass per contraction chunk -
  // two launches per conv instead of four.  Parity-green but OFF by default (SRK_TC_WIDE=1 selects it): measured on
  // config C3 17.5 vs 16.2 ms per step - its epilogue threads store 64-byte rows and read fp32 partial sums / residual
  // rows themselves (no room for staging tiles next to 110 KB of weights), and that costs more than the two launches
  // and the narrower MMAs of the 64 + 32 column passes save.
  if (g_tc_wide < 0) { const char* e = getenv("SRK_TC_WIDE"); g_tc_wide = e ? atoi(e) != 0 : 0; }
  if (r == 3 && g_tc_wide && tc_fold() >= 2 && shuffle == 0 && cout > NT && cout <= 2 * NT && cout % 32 == 0 &&
      stats_sum == nullptr && (x->c <= KC || workspace != nullptr)) {
    const int rc = conv_fprop_fold_launch(x, y, w_packed, cout, bias, act, alpha, residual, shuffle, nullptr, nullptr,
                                          workspace, 4, st, nullptr, nullptr, zsave);
    if (rc >= 0) return rc;   // -1: slab does not fit
  }
  if (r == 3 && tc_fold() == 4) {
    const int rc = conv_fprop_strip_launch(x, y, w_packed, cout, bias, act, alpha, residual, shuffle, stats_sum,
                                           stats_sumsq, st, nullptr, reduce_ws, zsave);
    if (rc >= 0) return rc;   // -1: not a plain 64 -> 64 pass -> the halo-slab kernels below
  }
  if (r == 3 && tc_fold() && !(tc_fold() >= 2 && shuffle != 0)) {
    const int rc = conv_fprop_fold_launch(x, y, w_packed, cout, bias, act, alpha, residual, shuffle, stats_sum,
                                          stats_sumsq, workspace, tc_fold() == 1 ? 1 : (tc_fold() == 3 ? 2 : 0), st, nullptr,
                                          reduce_ws, zsave);
    if (rc >= 0) return rc;   // -1: slab does not fit (very wide images) -> per-tap kernel below
  }
  const int cin = x->c;
  const int Hp = x->h + 2, Wp = x->w + 2;
  const long long P = (long long)x->n * Hp * Wp;
  SRK_REQUIRE(P < (1LL << 31) - 4096, "conv_tc: too many pixels");
  static int smem_max = 0;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(conv3x3_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    cudaFuncSetAttribute(conv3x3_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    cudaFuncSetAttribute(conv3x3_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  }
  int mode = r == 1 ? 0 : tc_mode();
  int slab_rows = ((TM + 2 * Wp + 2) + SLAB_BOX_ROWS - 1) / SLAB_BOX_ROWS * SLAB_BOX_ROWS;
  const int fixed = 1024 + TAPS * W_TILE_BYTES + 2 * O_TILE_BYTES + RED_BYTES + (int)sizeof(TcBarriers);
  int stage_bytes, stages;
  if (mode != 0) {
    stage_bytes = slab_rows * KC * 2;
    stages = (smem_max - fixed) / stage_bytes;
    if (stages < 2) mode = 0;  // image too wide for a double-buffered slab: fall back to per-tap loads
    if (stages > 4) stages = 4;
  }
  if (mode == 0) {
    stage_bytes = A_TILE_BYTES;
    stages = (smem_max - fixed) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    SRK_REQUIRE(stages >= 2, "conv_tc: not enough shared memory");
  }
  const int smem_bytes = fixed + stages * stage_bytes;

  CUtensorMap tmA, tmW;
  if (make_tmap_2d_bf16(&tmA, x->data, (uint64_t)P, (uint64_t)cin, (uint64_t)cin, mode == 0 ? TM : SLAB_BOX_ROWS, KC, 128))
    return 1;
  if (make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)(r * s) * cout, (uint64_t)cin, (uint64_t)cin, NT, KC, 128)) return 1;
  CUtensorMap tmY = tmA, tmR = tmA;  // output store / residual load maps (plain outputs only)
  if (shuffle == 0) {
    if (make_tmap_2d_bf16(&tmY, y->data, (uint64_t)P, (uint64_t)cout, (uint64_t)cout, TM, NT, 128)) return 1;
    tmR = tmY;
    if (residual && make_tmap_2d_bf16(&tmR, residual->data, (uint64_t)P, (uint64_t)cout, (uint64_t)cout, TM, NT, 128))
      return 1;
  }

  TcConvParams p;
  p.P = (int)P; p.Hp = Hp; p.Wp = Wp;
  p.num_tiles = (int)((P + TM - 1) / TM);
  p.w_row_per_tap = cout;
  p.taps = r * s;
  p.mode = mode; p.slab_rows = slab_rows; p.stages = stages; p.stage_bytes = stage_bytes;
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
  p.zsave = (__nv_bfloat16*)zsave;
  const int nchunks = (cout + NT - 1) / NT, kchunks = (cin + KC - 1) / KC;
  SRK_REQUIRE(stats_sum == nullptr || (kchunks == 1 && shuffle == 0 && act == SRK_ACT_NONE && residual == nullptr),
              "conv_tc: fused BN statistics need a plain Cin == 64 conv");
  int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  for (int nc = 0; nc < nchunks; ++nc) {
    for (int kc = 0; kc < kchunks; ++kc) {
      const bool first = kc == 0, last = kc == kchunks - 1;
      p.k_col0 = kc * KC;
      p.w_row0 = nc * NT;
      p.sub = 0;
      p.cout_total = y->c;     // channels per output pixel (Cout, or Cout/4 after PixelShuffle)
      p.cout_off = nc * NT;    // first conv output channel of this pass (reference order)
      p.bias_off = nc * NT; p.bias_stride = 1;
      p.n_cols = cout - nc * NT < NT ? cout - nc * NT : NT;
      p.ksteps = (cin - kc * KC < KC ? cin - kc * KC : KC) / 16;
      p.bias = first ? bias : nullptr;
      p.act = last ? act : SRK_ACT_NONE;
      // A contraction over more than 64 channels runs as one pass per chunk.  Without an activation the partial
      // sums ride through y itself in bf16 (TMA store, then re-read as the residual of the next pass; a caller
      // residual is added by the first pass).  With an activation the sign of the pre-activation matters for the
      // backward, so the partial sums travel in fp32 through the caller's workspace ([P][64] floats) and only the
      // last pass activates, adds the residual and stores.
      const bool fp32_partials = kchunks > 1 && (act != SRK_ACT_NONE || shuffle != 0);
      if (fp32_partials) {
        p.residual = (last && residual) ? (const __nv_bfloat16*)residual->data : nullptr;
        p.partial_out = last ? nullptr : (float*)workspace;
        p.partial_in = first ? nullptr : (const float*)workspace;
        SRK_REQUIRE(workspace != nullptr, "conv_tc: Cin > 64 with an activation needs the fprop workspace");
      } else {
        p.residual = first ? (residual ? (const __nv_bfloat16*)residual->data : nullptr) : p.y;
        p.partial_out = nullptr;
        p.partial_in = nullptr;
      }
      p.res_first = 0;
      const CUtensorMap& tmRes = (fp32_partials || first) ? tmR : tmY;
      const bool fast = kchunks == 1 && p.n_cols == NT && p.ksteps == KC / 16 && p.dbg == 0 && p.trace == nullptr;
      SRK_REQUIRE(fast || stats_sum == nullptr, "conv_tc: fused BN statistics need the single-chunk 64 -> 64 pass");
      if (fast && stats_sum) conv3x3_tc_kernel<true, true><<<grid, kThreads, smem_bytes, st>>>(tmA, tmW, tmY, tmRes, p);
      else if (fast) conv3x3_tc_kernel<true, false><<<grid, kThreads, smem_bytes, st>>>(tmA, tmW, tmY, tmRes, p);
      else conv3x3_tc_kernel<false, false><<<grid, kThreads, smem_bytes, st>>>(tmA, tmW, tmY, tmRes, p);
      SRK_CUDA_LAUNCH_CHECK("conv3x3_tc");
    }
  }
  if (shuffle == 2) {
    extern int zero_border(const srk_tensor* t, cudaStream_t st);
    return zero_border(y, st);
  }
  return 0;
}

}  // namespace srk

#ifdef SRK_WITH_PROBES   // micro-benchmarks (csrc/srk_probe_mma.cu): `python build.py --probes`, not part of the product library
namespace srk {
int probe_mma_rate(int n, int a_row_off, int mn_major, float* out_host);
int probe_ldtm_rate(int nwarps, int batch, float* out_host);
}
#endif

// Test / bring-up hook: variant 0..2 selects the A-staging mode of the tcgen05 conv (see the header
// comment); any other value leaves it unchanged.  out_host[0] = the device-side protocol-error flag
// (0 = none; it is cleared by the call), out_host[1] = the mode now in effect.  Synchronises the device.
extern "C" int srk_tc_probe(int variant, float* out_host, int out_len) {
  if (variant >= 0 && variant <= 2) srk::tc_set_mode(variant);
  if (variant >= 10 && variant <= 14) { srk::tc_fold(); srk::g_tc_fold = variant - 10; }  // 3x3 kernel choice (see tc_fold)
  if (variant == 30 || variant == 31) srk::g_up_pair = variant - 30;   // N = 128 CTA-pair upsample convs off / on
  if (variant == 40 || variant == 41) srk::g_tc_wide = variant - 40;   // wide (64 < Cout <= 128) single-pass convs off / on
  if (variant == 20 && out_host && out_len >= 2) {   // query: out[1] = 1 when the folded-tap kernel is the default
    out_host[0] = 0.f;
    out_host[1] = (float)srk::tc_fold();
    return 0;
  }
#ifdef SRK_WITH_PROBES
  if (variant >= 2000 && variant < 3000 && out_host && out_len >= 2)   // 2000 + nwarps + 100 * batch: TMEM read rate
    return srk::probe_ldtm_rate((variant - 2000) % 100, (variant - 2000) / 100, out_host);
  if (variant >= 1000 && out_host && out_len >= 2) {
    // 1000 + n/8 + 100 * a_row_off + 10000 * mn_major: sustained MMA rate microbenchmark
    int v = variant - 1000;
    return srk::probe_mma_rate((v % 100) * 8, (v / 100) % 100, v / 10000, out_host);
  }
#else
  if (variant >= 1000) SRK_FAIL("srk_tc_probe: the micro-benchmarks are compiled in with `build.py --probes` only");
#endif
  static long long* trace = nullptr;
  if (variant == 100) {
    if (!trace) cudaMalloc(&trace, 16 * 32 * sizeof(long long));
    cudaMemset(trace, 0, 16 * 32 * sizeof(long long));
    srk::tc_set_trace(trace);
  }
  if (variant == 102 && trace && out_host && out_len >= 16 * 32) {   // 16-row trace of the folded-tap kernel
    long long h[16 * 32];
    cudaDeviceSynchronize();
    cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    long long mn = 0;
    for (int i = 0; i < 16 * 32; ++i) if (h[i] && (mn == 0 || h[i] < mn)) mn = h[i];
    for (int i = 0; i < 16 * 32; ++i) out_host[i] = h[i] ? (float)(h[i] - mn) : -1.f;
    srk::tc_set_trace(nullptr);
    return 0;
  }
  if (variant == 101 && trace && out_host && out_len >= 6 * 32) {
    long long h[6 * 32];
    cudaDeviceSynchronize();
    cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    long long mn = 0;
    for (int i = 0; i < 6 * 32; ++i) if (h[i] && (mn == 0 || h[i] < mn)) mn = h[i];
    for (int i = 0; i < 6 * 32; ++i) out_host[i] = h[i] ? (float)(h[i] - mn) : -1.f;
    srk::tc_set_trace(nullptr);
    return 0;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) SRK_FAIL("srk_tc_probe: device error: %s", cudaGetErrorString(e));
  int flag = srk::tc_read_err_flag();
  int* f = srk::tc_err_flag();
  if (f) cudaMemset(f, 0, sizeof(int));
  if (out_host && out_len > 0) out_host[0] = (float)flag;
  if (out_host && out_len > 1) out_host[1] = (float)srk::tc_mode();
  return 0;
}
