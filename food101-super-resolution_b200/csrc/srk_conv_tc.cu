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
// One persistent CTA per SM, 192 threads:
//   warp 0   TMA producer  (A tiles / halo slabs -> smem ring, weights once)
//   warp 1   MMA issuer    (one thread: tcgen05.mma M128 N64 K16, fp32 accumulators in TMEM, double buffered)
//   warps 2-5 epilogue     (tcgen05.ld -> +bias, ReLU/PReLU, +residual, bf16 -> global, PixelShuffle remap)
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
constexpr int kThreads = 192;
constexpr int MAX_STAGES = 8;

struct TcConvParams {
  int P, Hp, Wp, num_tiles;
  int k_col0;          // first contraction channel of this pass (column coordinate in x and in the weights)
  int w_row_per_tap;   // rows per tap in the packed weight matrix (= total output channels)
  int w_row0;          // first weight row of this pass (= output-channel offset in packed order)
  int cout_total;      // channels per pixel of y (row stride)
  int cout_off;        // channel offset inside a y row
  int act, shuffle, sub;
  int mode, slab_rows, stages, stage_bytes;
  const float* bias;   // indexed [bias_off + c] or null
  int bias_off, bias_stride;   // bias index of column c = bias_off + c * bias_stride
  const float* alpha;
  const __nv_bfloat16* residual;
  __nv_bfloat16* y;
  int Hp2, Wp2;        // padded sizes of the shuffled output
  int* err;
};

struct __align__(8) TcBarriers {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], wfull, tfull[2], tempty[2];
  uint32_t tmem_base;
  float bias[NT];
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const TcConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem: [weights 9 x 8 KB][A ring stages x stage_bytes][barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t wsm = smem_base;
  const uint32_t asm0 = smem_base + TAPS * W_TILE_BYTES;
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem_al + TAPS * W_TILE_BYTES + p.stages * p.stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->wfull), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->tfull[i]), 1); mbar_init(smem_u32(&bars->tempty[i]), 128); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&bars->tmem_base), 128);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + NT) {
    int c = threadIdx.x - 64;
    bars->bias[c] = p.bias ? p.bias[p.bias_off + c * p.bias_stride] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      prefetch_tmap(&tmA);
      prefetch_tmap(&tmW);
      const uint32_t wbar = smem_u32(&bars->wfull);
      mbar_arrive_expect_tx(wbar, TAPS * W_TILE_BYTES);
      for (int t = 0; t < TAPS; ++t)
        tma_load_2d(wsm + t * W_TILE_BYTES, &tmW, wbar, p.k_col0, t * p.w_row_per_tap + p.w_row0);
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
        const int m0 = tile * TM;
        if (p.mode == 0) {
          for (int t = 0; t < TAPS && ok; ++t) {
            ok = mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 1);
            if (!ok) break;
            const uint32_t fb = smem_u32(&bars->full[s]);
            mbar_arrive_expect_tx(fb, A_TILE_BYTES);
            const int d = (t / 3 - 1) * p.Wp + (t % 3 - 1);
            tma_load_2d(asm0 + s * p.stage_bytes, &tmA, fb, p.k_col0, m0 + d);
            if (++s == S) { s = 0; ph ^= 1; }
          }
        } else {
          ok = mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1, p.err, 1);
          if (!ok) break;
          const uint32_t fb = smem_u32(&bars->full[s]);
          mbar_arrive_expect_tx(fb, p.slab_rows * KC * 2);
          const int row0 = m0 - p.Wp - 1;
          for (int j = 0; j < p.slab_rows / SLAB_BOX_ROWS; ++j)
            tma_load_2d(asm0 + s * p.stage_bytes + j * SLAB_BOX_ROWS * KC * 2, &tmA, fb, p.k_col0,
                        row0 + j * SLAB_BOX_ROWS);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TM, NT, 0, 0);
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
          for (int t = 0; t < TAPS && ok; ++t) {
            ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 4);
            if (!ok) break;
            tc_fence_after();
            const uint32_t a0 = asm0 + s * p.stage_bytes, b0 = wsm + t * W_TILE_BYTES;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              umma_bf16(d_tmem, make_smem_desc(a0 + ks * 32, 16, 1024, kLayoutSW128, 0),
                        make_smem_desc(b0 + ks * 32, 16, 1024, kLayoutSW128, 0), idesc, (t | ks) != 0);
            umma_commit(smem_u32(&bars->empty[s]));
            if (++s == S) { s = 0; ph ^= 1; }
          }
        } else {
          ok = mbar_wait(smem_u32(&bars->full[s]), ph, p.err, 4);
          if (!ok) break;
          tc_fence_after();
          const uint32_t slab = asm0 + s * p.stage_bytes;
          for (int t = 0; t < TAPS; ++t) {
            const uint32_t a0 = slab + ((t / 3) * p.Wp + (t % 3)) * (KC * 2), b0 = wsm + t * W_TILE_BYTES;
            const uint32_t bo = p.mode == 2 ? ((a0 >> 7) & 7) : 0;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks)
              umma_bf16(d_tmem, make_smem_desc(a0 + ks * 32, 16, 1024, kLayoutSW128, bo),
                        make_smem_desc(b0 + ks * 32, 16, 1024, kLayoutSW128, 0), idesc, (t | ks) != 0);
          }
          umma_commit(smem_u32(&bars->empty[s]));
          if (++s == S) { s = 0; ph ^= 1; }
        }
        if (ok) umma_commit(smem_u32(&bars->tfull[acc]));
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue =================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    const float alpha = (p.act == SRK_ACT_PRELU) ? p.alpha[0] : 0.f;
    const int img = p.Hp * p.Wp;
    bool ok = true;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      ok = mbar_wait(smem_u32(&bars->tfull[acc]), (it >> 1) & 1, p.err, 5);
      if (!ok) break;
      tc_fence_after();
      uint32_t v[NT];
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * NT;
      tmem_ld_32x32(taddr, v);
      tmem_ld_32x32(taddr + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bars->tempty[acc]));

      const int pix = tile * TM + lg * 32 + lane;
      if (pix >= p.P) continue;
      const int n = pix / img, q = pix - n * img;
      const int yy = q / p.Wp, xx = q - yy * p.Wp;
      const bool interior = yy >= 1 && yy <= p.Hp - 2 && xx >= 1 && xx <= p.Wp - 2;
      if (p.shuffle == 2) {
        // PixelShuffle(2) as a store remap (models.py:118,121): column j of this pass is reference channel
        // co = cout_off + j = 4c + sub  ->  output pixel (2y + sub/2, 2x + sub%2), channel c.
        if (!interior) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          float a = __uint_as_float(v[j]) + bars->bias[j];
          if (p.act == SRK_ACT_RELU) a = fmaxf(a, 0.f);
          else if (p.act == SRK_ACT_PRELU) a = a > 0.f ? a : alpha * a;
          v[j] = __float_as_uint(a);
        }
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          const long long orow =
              ((long long)n * p.Hp2 + (2 * (yy - 1) + (sub >> 1) + 1)) * p.Wp2 + (2 * (xx - 1) + (sub & 1) + 1);
          uint4* dst = reinterpret_cast<uint4*>(p.y + orow * p.cout_total + p.cout_off / 4);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = h * 8;
            dst[h] = make_uint4(
                pack_bf16x2(__uint_as_float(v[4 * (c0 + 0) + sub]), __uint_as_float(v[4 * (c0 + 1) + sub])),
                pack_bf16x2(__uint_as_float(v[4 * (c0 + 2) + sub]), __uint_as_float(v[4 * (c0 + 3) + sub])),
                pack_bf16x2(__uint_as_float(v[4 * (c0 + 4) + sub]), __uint_as_float(v[4 * (c0 + 5) + sub])),
                pack_bf16x2(__uint_as_float(v[4 * (c0 + 6) + sub]), __uint_as_float(v[4 * (c0 + 7) + sub])));
          }
        }
        continue;
      }
      uint4* dst = reinterpret_cast<uint4*>(p.y + (long long)pix * p.cout_total + p.cout_off);
      if (!interior) {
#pragma unroll
        for (int j = 0; j < NT / 8; ++j) dst[j] = make_uint4(0, 0, 0, 0);
        continue;
      }
      const uint4* res =
          p.residual ? reinterpret_cast<const uint4*>(p.residual + (long long)pix * p.cout_total + p.cout_off) : nullptr;
#pragma unroll
      for (int j = 0; j < NT / 8; ++j) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a = __uint_as_float(v[j * 8 + e]) + bars->bias[j * 8 + e];
          if (p.act == SRK_ACT_RELU) a = fmaxf(a, 0.f);
          else if (p.act == SRK_ACT_PRELU) a = a > 0.f ? a : alpha * a;
          f[e] = a;
        }
        if (res) {
          uint4 r = res[j];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
          for (int e = 0; e < 4; ++e) { float2 t = __bfloat1622float2(h[e]); f[2 * e] += t.x; f[2 * e + 1] += t.y; }
        }
        dst[j] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                            pack_bf16x2(f[6], f[7]));
      }
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
int tc_mode() {
  if (g_tc_mode < 0) {
    const char* e = getenv("SRK_TC_MODE");
    g_tc_mode = e ? atoi(e) : 1;  // slab staging: validated against the oracle on B200 (mode 2 is wrong)
    if (g_tc_mode < 0 || g_tc_mode > 2) g_tc_mode = 1;
  }
  return g_tc_mode;
}
void tc_set_mode(int m) { g_tc_mode = m; }

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

bool conv_tc_shape_ok(int cin, int cout, int r, int s, int dtype, int shuffle) {
  if (dtype != SRK_BF16 || r != 3 || s != 3) return false;
  if (cin % KC != 0 || cout % NT != 0) return false;
  if (shuffle != 0 && !(shuffle == 2 && cout % NT == 0 && cin == KC)) return false;
  return true;
}

int conv_fprop_tc_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout, int r, int s,
                         const float* bias, int act, const float* alpha, const srk_tensor* residual, int shuffle,
                         cudaStream_t st) {
  const int cin = x->c;
  const int Hp = x->h + 2, Wp = x->w + 2;
  const long long P = (long long)x->n * Hp * Wp;
  SRK_REQUIRE(P < (1LL << 31) - 4096, "conv_tc: too many pixels");
  static int smem_max = 0;
  if (!smem_max) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  }
  int mode = tc_mode();
  int slab_rows = ((TM + 2 * Wp + 2) + SLAB_BOX_ROWS - 1) / SLAB_BOX_ROWS * SLAB_BOX_ROWS;
  const int fixed = 1024 + TAPS * W_TILE_BYTES + (int)sizeof(TcBarriers);
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
  if (make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)TAPS * cout, (uint64_t)cin, (uint64_t)cin, NT, KC, 128)) return 1;

  TcConvParams p;
  p.P = (int)P; p.Hp = Hp; p.Wp = Wp;
  p.num_tiles = (int)((P + TM - 1) / TM);
  p.w_row_per_tap = cout;
  p.mode = mode; p.slab_rows = slab_rows; p.stages = stages; p.stage_bytes = stage_bytes;
  p.alpha = alpha;
  p.y = (__nv_bfloat16*)y->data;
  p.shuffle = shuffle;
  p.Hp2 = y->h + 2; p.Wp2 = y->w + 2;
  p.err = tc_err_flag();
  const int nchunks = cout / NT, kchunks = cin / KC;
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
      p.bias = first ? bias : nullptr;
      p.act = last ? act : SRK_ACT_NONE;
      // partial sums over contraction chunks ride through y itself (bf16) when Cin > 64
      p.residual = first ? (residual ? (const __nv_bfloat16*)residual->data : nullptr) : p.y;
      SRK_REQUIRE(kchunks == 1 || (act == SRK_ACT_NONE && shuffle == 0),
                  "conv_tc: activation / pixel-shuffle epilogues need Cin == 64");
      conv3x3_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(tmA, tmW, p);
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

// Test / bring-up hook: variant 0..2 selects the A-staging mode of the tcgen05 conv (see the header
// comment); any other value leaves it unchanged.  out_host[0] = the device-side protocol-error flag
// (0 = none; it is cleared by the call), out_host[1] = the mode now in effect.  Synchronises the device.
extern "C" int srk_tc_probe(int variant, float* out_host, int out_len) {
  if (variant >= 0 && variant <= 2) srk::tc_set_mode(variant);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) SRK_FAIL("srk_tc_probe: device error: %s", cudaGetErrorString(e));
  int flag = srk::tc_read_err_flag();
  int* f = srk::tc_err_flag();
  if (f) cudaMemset(f, 0, sizeof(int));
  if (out_host && out_len > 0) out_host[0] = (float)flag;
  if (out_host && out_len > 1) out_host[1] = (float)srk::tc_mode();
  return 0;
}
