// tcgen05 convolution for the "many channels in, RGB out" layers: output_conv 9x9 64->3 (models.py:125,167)
// and SRCNN conv3 5x5 64->3 (models.py:86).
//
//   out[n][co][y][x] = b[co] + sum_{r,s,ci} X[n][y+r-pad][x+s-pad][ci] * W[co][ci][r][s]
//
// With only 3 output channels a direct implicit GEMM wastes the MMA's N dimension and re-reads the A operand
// from shared memory once per tap (81 times).  Here the horizontal taps are folded into N instead:
//
//   Z[p][(s, co)] = sum_r sum_ci X[p + (r - pad, 0)][ci] * W[co][ci][r][s]        N = 3K (27 -> 32, 15 -> 16)
//   out[q][co]    = b[co] + sum_s Z[q + (0, s - pad)][(s, co)]
//
// i.e. the vertical tap shift is applied to the INPUT (a row-shifted UMMA descriptor into the halo tile, K
// descriptors instead of K*K), the horizontal tap shift to the OUTPUT (a shifted sum across neighbouring
// pixels = neighbouring TMEM lanes, done with warp shuffles in the epilogue).  M tile = 8 rows x 16 pixels;
// the 16 - (K-1) centre columns of a tile are complete, tiles overlap horizontally by K-1 pixels.
// One 4-D TMA box brings the (TY + K-1) x 32 pixel halo (zero outside the image) as [rows][32 px][128 B]
// SWIZZLE_128B, so every row shift is a whole number of 1024-byte swizzle atoms.
// A CTA tile is TG = 3 such M tiles stacked vertically (12 rows x 32 pixels, three accumulators) on ONE halo of
// 12 + K-1 rows: with a single 4-row M tile the 9x9 conv fetched 12 halo rows per 4 output rows - 512 bytes of
// L2 -> shared-memory traffic per output pixel, 2.2 GB per launch at config C2, which bounded the kernel (336 us);
// three stacked tiles share the vertical halo (284 bytes per output pixel).
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

namespace srk {

using namespace tc;
int* tc_err_flag();

constexpr int TY = 4, TXW = 32, KC = 64;   // M tile = 4 rows x 32 pixels: 32 - (K-1) of 32 columns are complete
constexpr int TG = 3;                      // M tiles per CTA tile (stacked vertically, one accumulator each)
constexpr int ACC_STRIDE = 32 * TG;        // TMEM columns of one accumulator buffer
constexpr int TMEM_COLS = 256;             // 2 buffers x TG x 32 columns, rounded up to a power of two
constexpr int kThreads = 192;
constexpr int MAX_STAGES = 4;

struct SmallNParams {
  int N, H, W, K, pad, cout, npad, xv;   // xv = valid output columns per tile
  int tiles_x, tiles_y, num_tiles;
  int halo_rows, stage_bytes, stages, w_bytes;
  const float* bias;
  float* out;  // NCHW fp32
  int* err;
};

struct __align__(8) SmallNBarriers {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], wfull, tfull[2], tempty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
template <int K>
__global__ void __launch_bounds__(kThreads, 1)
conv_smalln_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const SmallNParams p) {
  constexpr int PAD = K / 2, NP = (K * 3 > 16) ? 32 : 16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t hsm = smem_base;                                  // halo ring
  const uint32_t wsm = smem_base + p.stages * p.stage_bytes;       // weights [K][NP][64]
  SmallNBarriers* bars = reinterpret_cast<SmallNBarriers*>(smem_al + p.stages * p.stage_bytes + p.w_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->wfull), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), 128); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bars->tmem_base), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // TMA producer: whole warp, one elected lane issues
    if (elect_one()) {
      prefetch_tmap(&tmX);
      prefetch_tmap(&tmW);
      const uint32_t wbar = smem_u32(&bars->wfull);
      mbar_arrive_expect_tx(wbar, K * NP * 128);
      for (int r = 0; r < K; ++r) tma_load_2d(wsm + r * NP * 128, &tmW, wbar, 0, r * NP);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
      const int y0 = (t2 / p.tiles_x) * (TY * TG), x0 = (t2 % p.tiles_x) * p.xv;
      if (!mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 21)) break;
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->full[s]);
        mbar_arrive_expect_tx(fb, p.halo_rows * TXW * 128);
        tma_load_4d(hsm + s * p.stage_bytes, &tmX, fb, 0, x0 - PAD, y0 - PAD, n);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp, one elected lane issues.  A(r) = halo + r * (32 px * 128 B): 1024-byte aligned.
    constexpr uint32_t idesc = make_idesc_bf16(128, NP, 0, 0);
    const uint64_t hi = make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_base = (uint32_t)(make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t w_lo = lo_base + (wsm >> 4);
    bool ok = mbar_wait(smem_u32(&bars->wfull), 0, p.err, 22);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      ok = mbar_wait(smem_u32(&bars->tempty[acc]), ((it >> 1) & 1) ^ 1, p.err, 23);
      if (!ok) break;
      ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 24);
      if (!ok) break;
      tc_fence_after();
      const uint32_t halo_lo = lo_base + ((hsm + s * p.stage_bytes) >> 4), d_tmem = tmem_base + acc * ACC_STRIDE;
      if (elect_one()) {
#pragma unroll
        for (int g = 0; g < TG; ++g)   // M tile g: output rows 4g .. 4g+3 of the CTA tile = halo rows 4g + r ..
#pragma unroll
          for (int r = 0; r < K; ++r)
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              umma_bf16(d_tmem + g * 32, hi | (halo_lo + (g * TY + r) * (TXW * 128 / 16) + 2 * ks),
                        hi | (w_lo + r * (NP * 128 / 16) + 2 * ks), idesc, (r | ks) != 0);
        umma_commit(smem_u32(&bars->empty[s]));
        umma_commit(smem_u32(&bars->tfull[acc]));
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    // epilogue: warp lg owns TMEM lanes 32 lg .. = tile row lg (32 pixels)
    const int lg = warp & 3;
    const int tx = lane, ty = lg;
    float b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) b[c] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      if (!mbar_wait(smem_u32(&bars->tfull[acc]), (it >> 1) & 1, p.err, 25)) break;
      tc_fence_after();
      const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
      const int x = (t2 % p.tiles_x) * p.xv + tx - PAD;
#pragma unroll 1
      for (int g = 0; g < TG; ++g) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * ACC_STRIDE + g * 32;
        tmem_ld_32x16(taddr, v);
        if (NP == 32) tmem_ld_32x16(taddr + 16, v + 16);
        tmem_ld_wait();
        if (g == TG - 1) {   // the whole buffer is in registers (or stored): MMA may refill it
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->tempty[acc]));
        }
        // out[q][co] = sum_s Z[q + (s - PAD)][(s, co)]: the neighbour's value comes by shuffle
        float o[3] = {b[0], b[1], b[2]};
#pragma unroll
        for (int s = 0; s < K; ++s) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float z = __shfl_sync(0xffffffffu, __uint_as_float(v[s * 3 + c]), (lane + s - PAD) & 31);
            o[c] += z;
          }
        }
        const int y = (t2 / p.tiles_x) * (TY * TG) + g * TY + ty;
        if (tx >= PAD && tx < TXW - PAD && y < p.H && x < p.W) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < p.cout) p.out[(((size_t)n * p.cout + c) * p.H + y) * p.W + x] = o[c];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

bool conv_smalln_tc_ok(const srk_tensor* x, const srk_tensor* y, int cout, int r, int s) {
  if (r != s || (r != 5 && r != 9)) return false;
  if (x->layout != SRK_LAYOUT_ACT || x->dtype != SRK_BF16 || x->c != KC) return false;
  if (y->layout != SRK_LAYOUT_IMAGE || cout > 3) return false;
  return true;
}

static int make_tmap_act_4d(CUtensorMap* out, const srk_tensor* x, int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode_tiled();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const uint64_t C = x->c, Wp = x->w + 2, Hp = x->h + 2;
  cuuint64_t dims[4] = {C, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
  cuuint64_t strides[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* base = (char*)x->data + (Wp + 1) * C * 2;  // first interior pixel
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed (%d)", (int)r);
  return 0;
}

// w_packed: SRK_PACK_FPROP_TC_N8 = bf16 [R][NP][64], row n = s * 3 + co, NP = 32 (9x9) or 16 (5x5)
int conv_smalln_tc_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout, int r,
                          const float* bias, cudaStream_t st) {
  static int smem_max = 0;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(conv_smalln_tc_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    cudaFuncSetAttribute(conv_smalln_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  }
  SmallNParams p;
  p.N = x->n; p.H = x->h; p.W = x->w; p.K = r; p.pad = r / 2; p.cout = cout;
  p.npad = r * 3 > 16 ? 32 : 16;
  p.xv = TXW - (r - 1);
  p.tiles_x = (p.W + p.xv - 1) / p.xv; p.tiles_y = (p.H + TY * TG - 1) / (TY * TG);
  const long long nt = (long long)p.N * p.tiles_x * p.tiles_y;
  SRK_REQUIRE(nt < (1LL << 31), "conv_smalln: too many tiles");
  p.num_tiles = (int)nt;
  p.halo_rows = TY * TG + r - 1;
  p.stage_bytes = p.halo_rows * TXW * 128;                 // multiple of 2048
  p.w_bytes = (r * p.npad * 128 + 1023) / 1024 * 1024;
  const int fixed = 1024 + p.w_bytes + (int)sizeof(SmallNBarriers);
  p.stages = (smem_max - fixed) / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  SRK_REQUIRE(p.stages >= 1, "conv_smalln: not enough shared memory");
  p.bias = bias; p.out = (float*)y->data; p.err = tc_err_flag();
  CUtensorMap tmX, tmW;
  if (make_tmap_act_4d(&tmX, x, TXW, p.halo_rows)) return 1;
  if (make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)r * p.npad, KC, KC, p.npad, KC, 128)) return 1;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const int smem = fixed + p.stages * p.stage_bytes;
  if (r == 9) conv_smalln_tc_kernel<9><<<grid, kThreads, smem, st>>>(tmX, tmW, p);
  else conv_smalln_tc_kernel<5><<<grid, kThreads, smem, st>>>(tmX, tmW, p);
  SRK_CUDA_LAUNCH_CHECK("conv_smalln_tc");
  return 0;
}

}  // namespace srk
