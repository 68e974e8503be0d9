// tcgen05 convolution for the "many channels in, RGB out" layers: output_conv 9x9 64->3 (models.py:125,167)
// and SRCNN conv3 5x5 64->3 (models.py:86).  Direct (no im2col) implicit GEMM:
//
//   out[n][co][y][x] = b[co] + sum_{r,s,ci} X[n][y+r-pad][x+s-pad][ci] * W[co][ci][r][s]
//
// M tile = 16 rows x 8 columns of output pixels.  One 4-D TMA box brings the (16+K-1) x (8+K-1) halo of
// 64-channel pixels (zero-filled outside the image by TMA bounds checking, so the 1-pixel activation border
// is irrelevant here) into shared memory as [halo_y][halo_x][128 B] SWIZZLE_128B.  The A operand of tap
// (r, s) is that same buffer viewed through a descriptor that starts at pixel (r, s) with
// stride-byte-offset = halo row pitch: eight consecutive x pixels are the 8 rows of a core matrix, the 16
// tile rows are the 16 row groups.  N is padded to the minimum UMMA N=16; weights are packed
// [tap][8 co][64 ci] so the upper 8 columns read the next tap's rows and are simply never stored.
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

namespace srk {

using namespace tc;
int* tc_err_flag();

constexpr int TY = 16, TX = 8, KC = 64, NPAD = 8;
constexpr int kThreads = 192;
constexpr int MAX_STAGES = 4;

struct SmallNParams {
  int N, H, W, K, pad, cout;
  int tiles_x, tiles_y, num_tiles;
  int box_w, box_h, stage_bytes, stages, w_bytes;
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
__device__ __forceinline__ void tmem_ld_32x4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv_smalln_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const SmallNParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t hsm = smem_base;                                  // halo ring
  const uint32_t wsm = smem_base + p.stages * p.stage_bytes;       // weights [taps][8][64] (+ slack)
  SmallNBarriers* bars = reinterpret_cast<SmallNBarriers*>(smem_al + p.stages * p.stage_bytes + p.w_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages, taps = p.K * p.K;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->wfull), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), 128); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bars->tmem_base), 32);
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
      const int wrows = taps * NPAD;
      mbar_arrive_expect_tx(wbar, wrows * 128);
      for (int r0 = 0; r0 < wrows; r0 += NPAD) tma_load_2d(wsm + r0 * 128, &tmW, wbar, 0, r0);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
      const int y0 = (t2 / p.tiles_x) * TY, x0 = (t2 % p.tiles_x) * TX;
      if (!mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 21)) break;
      if (elect_one()) {
        const uint32_t fb = smem_u32(&bars->full[s]);
        mbar_arrive_expect_tx(fb, p.box_w * p.box_h * 128);
        tma_load_4d(hsm + s * p.stage_bytes, &tmX, fb, 0, x0 - p.pad, y0 - p.pad, n);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp, one elected lane issues
    constexpr uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
    const uint32_t sbo = p.box_w * 128;
    const uint64_t a_hi = make_smem_desc(0, 16, sbo, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
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
      const uint32_t halo_lo = lo_base + ((hsm + s * p.stage_bytes) >> 4), d_tmem = tmem_base + acc * 16;
      if (elect_one()) {
        uint32_t row_lo = halo_lo, b_lo = w_lo;
        for (int r = 0; r < p.K; ++r, row_lo += p.box_w * 8) {
          uint32_t a_lo = row_lo;
          for (int c = 0; c < p.K; ++c, a_lo += 8, b_lo += NPAD * 128 / 16) {
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              umma_bf16(d_tmem, a_hi | (a_lo + 2 * ks), b_hi | (b_lo + 2 * ks), idesc, (r | c | ks) != 0);
          }
        }
        umma_commit(smem_u32(&bars->empty[s]));
        umma_commit(smem_u32(&bars->tfull[acc]));
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    const int lg = warp & 3;
    const int i = lg * 32 + lane, ty = i >> 3, tx = i & 7;
    float b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) b[c] = (p.bias && c < p.cout) ? p.bias[c] : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      if (!mbar_wait(smem_u32(&bars->tfull[acc]), (it >> 1) & 1, p.err, 25)) break;
      tc_fence_after();
      uint32_t v[4];
      tmem_ld_32x4(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * 16, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->tempty[acc]));
      const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
      const int y = (t2 / p.tiles_x) * TY + ty, x = (t2 % p.tiles_x) * TX + tx;
      if (y < p.H && x < p.W) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < p.cout) p.out[(((size_t)n * p.cout + c) * p.H + y) * p.W + x] = __uint_as_float(v[c]) + b[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

bool conv_smalln_tc_ok(const srk_tensor* x, const srk_tensor* y, int cout, int r, int s) {
  if (r != s || (r != 5 && r != 9 && r != 3 && r != 7)) return false;
  if (x->layout != SRK_LAYOUT_ACT || x->dtype != SRK_BF16 || x->c != KC) return false;
  if (y->layout != SRK_LAYOUT_IMAGE || cout > 4) return false;
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

// w_packed: SRK_PACK_FPROP_TC_N8 = bf16 [tap][8][64]
int conv_smalln_tc_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout, int r,
                          const float* bias, cudaStream_t st) {
  static int smem_max = 0;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(conv_smalln_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  }
  SmallNParams p;
  p.N = x->n; p.H = x->h; p.W = x->w; p.K = r; p.pad = r / 2; p.cout = cout;
  p.tiles_x = (p.W + TX - 1) / TX; p.tiles_y = (p.H + TY - 1) / TY;
  const long long nt = (long long)p.N * p.tiles_x * p.tiles_y;
  SRK_REQUIRE(nt < (1LL << 31), "conv_smalln: too many tiles");
  p.num_tiles = (int)nt;
  p.box_w = TX + r - 1; p.box_h = TY + r - 1;
  p.stage_bytes = (p.box_w * p.box_h * 128 + 1023) / 1024 * 1024;
  p.w_bytes = (r * r * NPAD * 128 + NPAD * 128 + 1023) / 1024 * 1024;  // + one tap of slack for the N=16 read
  const int fixed = 1024 + p.w_bytes + (int)sizeof(SmallNBarriers);
  p.stages = (smem_max - fixed) / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  SRK_REQUIRE(p.stages >= 1, "conv_smalln: not enough shared memory");
  p.bias = bias; p.out = (float*)y->data; p.err = tc_err_flag();
  CUtensorMap tmX, tmW;
  if (make_tmap_act_4d(&tmX, x, p.box_w, p.box_h)) return 1;
  if (make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)r * r * NPAD, KC, KC, NPAD, KC, 128)) return 1;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  conv_smalln_tc_kernel<<<grid, kThreads, fixed + p.stages * p.stage_bytes, st>>>(tmX, tmW, p);
  SRK_CUDA_LAUNCH_CHECK("conv_smalln_tc");
  return 0;
}

}  // namespace srk
