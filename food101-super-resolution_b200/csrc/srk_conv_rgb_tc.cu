// tcgen05 kernels for the convolutions whose OTHER side has only 3 (RGB) channels:
//   input_conv 9x9 3->64 (models.py:107,150), SRCNN conv1 9x9 3->64 (models.py:84)       : forward + wgrad
//   output_conv 9x9 64->3 (models.py:125,167), SRCNN conv3 5x5 64->3 (models.py:86)       : dgrad + wgrad
//
// All of them contract over k = (kernel row r, column s, rgb channel c).  The K axis is laid out in SEGMENTS of
// SEGW = 28 / 16 elements per kernel row: k = r * SEGW + s * 3 + c (the 27 / 15 real entries of a row, then zero-weight
// padding), the ones column at k = K * SEGW, padded to KP = 256 / 128.  With the halo of the 3-channel image staged
// pixel-interleaved in shared memory ([row][x][rgb] bf16), a segment of A is then 56 / 32 CONTIGUOUS bytes of the
// halo (9 / 5 neighbouring pixels), so the im2col builders copy 32-bit words instead of gathering 2-byte elements
// (a second copy of the halo, shifted by one element, serves the pixels whose segment starts on an odd element).
// The kernel materialises, per 16x8-pixel tile, the im2col matrix A[128 pixels][KP] of the 3-channel fp32
// NCHW image T3 in shared memory (bf16, four/two [128 x 128 B] SWIZZLE_128B sub-tiles) and feeds it to the
// tensor cores twice:
//   (y)  Y[pixel][n]  = sum_k A[pixel][k] * Wk[n][k]          A K-major,  Wk K-major  -> ACT bf16 output
//   (g)  G[k][n]     += sum_pixel A[pixel][k] * T64[pixel][n] A MN-major, T64 MN-major -> fp32 weight gradient
// with T64 the 64-channel activation-layout tensor at the same pixels (TMA 4-D box, zero outside the image).
// The ones column of A is 1 for in-image pixels, so that row of G is the column sum of T64 (a bias gradient for
// free).  G stays in TMEM for the CTA's whole tile range and is flushed once with vector fp32 reductions.
//
// 18 warps: 0 TMA, 1 MMA issue, 2-9 im2col builders (two threads per pixel row), 10-17 epilogue.
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

#include <cstdlib>
#include <cstring>

namespace srk {

using namespace tc;
int* tc_err_flag();
int zero_border(const srk_tensor* t, cudaStream_t st);

namespace rgb {

constexpr int TY = 16, TX = 8, TM = 128, NT = 64;
constexpr int kThreads = 576;   // warps: 0 TMA, 1 MMA, 2-9 im2col builders, 10-17 epilogue
constexpr int kBuilders = 256;
constexpr int SUB_BYTES = TM * 128;  // one [128 x 64 k] sub-tile of A

struct Params {
  int N, H, W, tiles_x, tiles_y, num_tiles;
  int do_y, do_g, act;
  int y_stride, y_col0, n_valid;   // y row pitch in channels, first channel and channel count of this pass
  int w_row0, t_col0;              // first packed-weight row / first T64 channel of this pass
  const float* t3;       // [N][3][H][W]
  const float* bias;     // [64] or null (y)
  const float* alpha;
  __nv_bfloat16* y;      // ACT [N][H+2][W+2][64]
  float* ws;             // [grid][KP][64] fp32: one partial of G per CTA (summed in CTA order by fold_kernel)
  float* ws_small;       // [grid][4] fp32: per-CTA partials of db3[0..2] and ps_dalpha
  int want_db3;          // sum over pixels of T3 (bias gradient of the 64 -> 3 conv)
  __nv_bfloat16* zsave;  // fprop with PReLU and a slope <= 0: copy of the pre-activation (y geometry), or null
  const __nv_bfloat16* ps_zsave;   // unshuffle: pre-activation of the layer below (read when its slope is <= 0), or null
  // rgb_out backward fused with the PReLU + PixelShuffle(2) backward of the layer below (models.py:120-122): T64 is that
  // layer's output, the dgrad result is masked by its sign and stored through tmY as dz of the 64 -> 256 conv,
  // channels sub-pixel-major (row (y/2, x/2), channel sub * 64 + c), the PReLU-slope gradient goes to ps_dalpha
  int unshuffle;
  const float* ps_alpha;
  int want_dalpha;
  int* err;
  int dbg;               // bring-up knobs: 1 builders skip the A rows, 2 no MMAs, 4 no T64 loads, 8 no y stores, 16 no halo
};

struct __align__(8) Barriers {
  uint64_t wfull, afull[2], aempty[2], tfull[2], tempty[2], yfull[2], yempty[2], done;
  uint32_t tmem_base;
  float red3[8][3];
  float redda[8];
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void tma_store_5d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

template <int K>
struct Geo {
  static constexpr int KK = K * K * 3;
  static constexpr int SEGW = (K * 3 + 1) / 2 * 2;      // elements per kernel-row segment (even: whole 32-bit words)
  static constexpr int KONE = K * SEGW;                 // the ones column
  static constexpr int KP = (KONE + 1 + 63) / 64 * 64;
  static constexpr int KCH = KP / 64;
  static constexpr int HH = TY + K - 1, HW = TX + K - 1;
  static constexpr int HALO = 3 * HH * HW;              // bf16 elements of one pixel-interleaved halo copy
  static constexpr int HALO_PITCH = (HALO * 2 + 8 + 15) / 16 * 16;   // bytes per copy (+ the over-read of the last segment)
  static constexpr int HALO_BYTES = (2 * 2 * HALO_PITCH + 1023) / 1024 * 1024;   // 2 copies, double buffered
  static constexpr int A_BYTES = KCH * SUB_BYTES;
  static constexpr int W_BYTES = KCH * NT * 128;
  static constexpr int T_BYTES = TM * 128;
  static constexpr int Y_BYTES = TM * 128;   // bf16 output tile staged for the TMA store
  // smem: [A x2][W][T64 x2][Y][halo][barriers]
  static constexpr int SMEM = 1024 + 2 * A_BYTES + W_BYTES + 2 * T_BYTES + Y_BYTES + HALO_BYTES + (int)sizeof(Barriers);
};

template <int K>
__global__ void __launch_bounds__(kThreads, 1)
conv_rgb_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmT,
                   const __grid_constant__ CUtensorMap tmY, const Params p) {
  using G = Geo<K>;
  constexpr int PAD = K / 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t a_sm = smem_base, w_sm = a_sm + 2 * G::A_BYTES, t_sm = w_sm + G::W_BYTES;
  uint8_t* a_ptr = smem_al;
  uint8_t* ystage = smem_al + 2 * G::A_BYTES + G::W_BYTES + 2 * G::T_BYTES;
  const uint32_t y_sm = t_sm + 2 * G::T_BYTES;
  uint8_t* halo = ystage + G::Y_BYTES;
  Barriers* bars = reinterpret_cast<Barriers*>(halo + G::HALO_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NG = G::KP / 128;        // G accumulators (M = 128 each)
  constexpr int TMEM_COLS = 256;         // Y: 2 x 64, G: up to 2 x 64

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->wfull), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->afull[i]), kBuilders);
      mbar_init(smem_u32(&bars->aempty[i]), 1);
      mbar_init(smem_u32(&bars->tfull[i]), 1);
      mbar_init(smem_u32(&bars->tempty[i]), p.unshuffle ? 257 : 1);   // + the epilogue threads that read the T64 tile
      mbar_init(smem_u32(&bars->yfull[i]), 1);
      mbar_init(smem_u32(&bars->yempty[i]), 256);
    }
    mbar_init(smem_u32(&bars->done), 1);
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
  const int my_tiles = blockIdx.x < p.num_tiles ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    // ================= TMA: weights once, T64 tiles (whole warp, one elected lane issues) =================
    if (p.do_y) {
      if (elect_one()) {
        prefetch_tmap(&tmW);
        const uint32_t wb = smem_u32(&bars->wfull);
        mbar_arrive_expect_tx(wb, G::W_BYTES);
        for (int q = 0; q < G::KCH; ++q) tma_load_2d(w_sm + q * NT * 128, &tmW, wb, q * 64, p.w_row0);
      }
      __syncwarp();
    }
    if (p.do_g) {
      if (elect_one()) prefetch_tmap(&tmT);
      __syncwarp();
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
        const int y0 = (t2 / p.tiles_x) * TY, x0 = (t2 % p.tiles_x) * TX;
        const int s = i & 1;
        if (!mbar_wait(smem_u32(&bars->tempty[s]), ((i >> 1) & 1) ^ 1, p.err, 31)) break;
        if (elect_one()) {
          const uint32_t fb = smem_u32(&bars->tfull[s]);
          if (p.dbg & 4) mbar_arrive(fb);
          else {
            mbar_arrive_expect_tx(fb, G::T_BYTES);
            tma_load_4d(t_sm + s * G::T_BYTES, &tmT, fb, p.t_col0, x0, y0, n);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp, one elected lane issues) =================
    constexpr uint32_t idesc_y = make_idesc_bf16(128, NT, 0, 0);
    constexpr uint32_t idesc_g = make_idesc_bf16(128, NT, 1, 1);
    const uint64_t hi_k = make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFF00000000ull;
    const uint32_t lo_k = (uint32_t)(make_smem_desc(0, 16, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t lo_ga = (uint32_t)(make_smem_desc(0, SUB_BYTES, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    const uint32_t lo_gb = (uint32_t)(make_smem_desc(0, 0, 1024, kLayoutSW128, 0) & 0xFFFFFFFFull);
    bool ok = true;
    if (p.do_y) ok = mbar_wait(smem_u32(&bars->wfull), 0, p.err, 32);
    for (int i = 0; i < my_tiles && ok; ++i) {
      const int buf = i & 1;
      const uint32_t par = (i >> 1) & 1;
      ok = mbar_wait(smem_u32(&bars->afull[buf]), par, p.err, 33);
      if (!ok) break;
      tc_fence_after();
      const uint32_t a0 = (a_sm + buf * G::A_BYTES) >> 4;
      if (p.do_y) {
        ok = mbar_wait(smem_u32(&bars->yempty[buf]), par ^ 1, p.err, 34);
        if (!ok) break;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + buf * NT;
#pragma unroll
          for (int q = 0; q < G::KCH; ++q)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (!(p.dbg & 2)) umma_bf16(d, hi_k | (lo_k + a0 + q * (SUB_BYTES / 16) + 2 * ks),
                        hi_k | (lo_k + (w_sm >> 4) + q * (NT * 128 / 16) + 2 * ks), idesc_y, (q | ks) != 0);
          umma_commit(smem_u32(&bars->yfull[buf]));
        }
        __syncwarp();
      }
      if (p.do_g) {
        ok = mbar_wait(smem_u32(&bars->tfull[buf]), par, p.err, 35);
        if (!ok) break;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t0 = (t_sm + buf * G::T_BYTES) >> 4;
#pragma unroll
          for (int mh = 0; mh < NG; ++mh)
#pragma unroll
            for (int ks = 0; ks < TM / 16; ++ks)
              if (!(p.dbg & 2)) umma_bf16(tmem_base + 2 * NT + mh * NT,
                        hi_k | (lo_ga + a0 + (2 * mh) * (SUB_BYTES / 16) + ks * (16 * 128 / 16)),
                        hi_k | (lo_gb + t0 + ks * (16 * 128 / 16)), idesc_g, (i | ks) != 0);
          umma_commit(smem_u32(&bars->tempty[buf]));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&bars->aempty[buf]));
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&bars->done));
    __syncwarp();
  } else if (warp < 10) {
    // ================= im2col builders: thread (row, half) builds half of the K range of one pixel row ===========
    const int bt = threadIdx.x - 64, bi = bt & 127, half = bt >> 7, ty = bi >> 3, tx = bi & 7;
    constexpr int PER = (G::HALO + kBuilders - 1) / kBuilders;   // halo elements staged per thread
    float s3[3] = {0.f, 0.f, 0.f};
    const __nv_bfloat16 one = __float2bfloat16_rn(1.f);
    float preA[PER], preB[PER];
    auto fetch = [&](int i, float (&pre)[PER]) {   // global -> registers for tile i (issued TWO tiles ahead of its use)
      const int tile = blockIdx.x + i * gridDim.x;
      const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
      const int y0 = (t2 / p.tiles_x) * TY, x0 = (t2 % p.tiles_x) * TX;
      const float* src = p.t3 + (size_t)n * 3 * p.H * p.W;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int idx = bt + u * kBuilders;
        float v = 0.f;
        if (idx < G::HALO) {
          const int c = idx / (G::HH * G::HW), rem = idx - c * (G::HH * G::HW);
          const int hy = rem / G::HW, hx = rem - hy * G::HW;
          const int gy = y0 + hy - PAD, gx = x0 + hx - PAD;
          if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(src + ((size_t)c * p.H + gy) * p.W + gx);
        }
        pre[u] = v;
      }
    };
    // the padding element of a row's last segment is read from just behind the halo (and multiplied by a zero
    // weight): it must never be a NaN / Inf bit pattern
    for (int z = bt; z < G::HALO_BYTES / 16; z += kBuilders) reinterpret_cast<uint4*>(halo)[z] = make_uint4(0, 0, 0, 0);
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (my_tiles > 0) fetch(0, preA);
    if (my_tiles > 1) fetch(1, preB);
    // one tile: stage its halo (fetched two tiles earlier: a global load takes longer than one tile's build), refill
    // the registers for tile i + 2, build the A rows
    auto build_tile = [&](int i, float (&pre)[PER]) -> bool {
      const int tile = blockIdx.x + i * gridDim.x;
      const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
      const int y0 = (t2 / p.tiles_x) * TY, x0 = (t2 % p.tiles_x) * TX;
      (void)n;
      const int buf = i & 1;
      // two pixel-interleaved copies of the halo: copy 0 at element e -> byte 2e, copy 1 shifted down by one element
      // (element e -> byte 2e - 2), so that a segment starting on an odd element is word-aligned in copy 1
      uint8_t* h0 = halo + (buf * 2) * G::HALO_PITCH;
      uint8_t* h1 = h0 + G::HALO_PITCH;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int idx = bt + u * kBuilders;
        if (idx < G::HALO && !(p.dbg & 16)) {
          const int c = idx / (G::HH * G::HW), rem = idx - c * (G::HH * G::HW);   // fetch order: channel planes
          const int e = rem * 3 + c;                                               // interleaved element index
          const __nv_bfloat16 v = __float2bfloat16_rn(pre[u]);
          *reinterpret_cast<__nv_bfloat16*>(h0 + 2 * e) = v;
          if (e > 0) *reinterpret_cast<__nv_bfloat16*>(h1 + 2 * e - 2) = v;
        }
      }
      // one barrier per tile: the other halo buffer is only rewritten after everybody passed this point again
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (i + 2 < my_tiles) fetch(i + 2, pre);
      if (!mbar_wait(smem_u32(&bars->aempty[buf]), ((i >> 1) & 1) ^ 1, p.err, 36)) return false;
      const bool inside = (y0 + ty < p.H) && (x0 + tx < p.W);
      // segment r of this pixel = SEGW elements from element ((ty + r) * HW + tx) * 3
      const int e0 = (ty * G::HW + tx) * 3;
      const uint8_t* hsrc = (e0 & 1) ? h1 + 2 * e0 - 2 : h0 + 2 * e0;
      constexpr int ROW_STEP = G::HW * 3 * 2;        // bytes between consecutive halo rows
      constexpr int WSEG = G::SEGW / 2;              // words per segment
      constexpr int WROW = G::KP / 2;                // words per A row
      constexpr int WONE = G::KONE / 2;              // word holding the ones column (KONE is even)
      uint8_t* arow = a_ptr + buf * G::A_BYTES + bi * 128;
      const uint32_t one_word = inside ? (uint32_t)__bfloat16_as_ushort(one) : 0u;
#pragma unroll
      for (int hsel = 0; hsel < 2; ++hsel) {
        if (hsel != half || (p.dbg & 1)) continue;
#pragma unroll
        for (int cj = 0; cj < WROW / 8; ++cj) {      // 16-byte chunks of this thread's half row
          const int chunk = hsel * (WROW / 8) + cj;
          uint32_t wv[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int w = chunk * 4 + t;             // word index inside the A row (compile-time)
            if (w < K * WSEG) {
              const int r = w / WSEG, j = w - r * WSEG;
              wv[t] = *reinterpret_cast<const uint32_t*>(hsrc + r * ROW_STEP + 4 * j);
            } else if (w == WONE) {
              wv[t] = one_word;
            } else {
              wv[t] = 0u;
            }
          }
          const int sub = chunk >> 3, jj = chunk & 7;
          *reinterpret_cast<uint4*>(arow + sub * SUB_BYTES + ((jj ^ (bi & 7)) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
      }
      if (p.want_db3 && inside && half == 0) {
        const __nv_bfloat16* hc = reinterpret_cast<const __nv_bfloat16*>(h0) + ((ty + PAD) * G::HW + tx + PAD) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) s3[c] += __bfloat162float(hc[c]);
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->afull[buf]));
      return true;
    };
    for (int i = 0; i < my_tiles; i += 2) {
      if (!build_tile(i, preA)) break;
      if (i + 1 < my_tiles && !build_tile(i + 1, preB)) break;
    }
    if (p.want_db3) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = warp_sum(s3[c]);
        if (lane == 0) bars->red3[warp - 2][c] = t;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (bt < 3) {
        float t = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += bars->red3[w8][bt];
        p.ws_small[blockIdx.x * 4 + bt] = t;     // CTA partial; fold_kernel sums the CTAs in order
      }
    }
  } else {
    // ================= epilogue: 8 warps; two warps share a TMEM lane group and split the 64 columns ==========
    const int lg = warp & 3, ch = (warp - 10) >> 2, c0 = ch * 32;
    const int ei = lg * 32 + lane, ty = ei >> 3, tx = ei & 7;
    if (p.do_y) {
      const float alpha = (p.act == SRK_ACT_PRELU) ? p.alpha[0] : 0.f;
      const float ps_alpha = p.unshuffle ? __ldg(p.ps_alpha) : 0.f;
      const float ps_inv_alpha = (p.unshuffle && ps_alpha != 0.f) ? 1.f / ps_alpha : 0.f;
      // slope <= 0: the sign / value of the pre-activation does not follow from the layer's output (see act_bwd_kernel)
      const bool ps_use_z = p.unshuffle && p.ps_zsave != nullptr && !(ps_alpha > 0.f);
      const bool save_z = p.act == SRK_ACT_PRELU && p.zsave != nullptr && !(alpha > 0.f);
      float ps_da = 0.f;
      float bias[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) bias[j] = (p.bias && c0 + j < p.n_valid) ? __ldg(p.bias + p.y_col0 + c0 + j) : 0.f;
      for (int i = 0; i < my_tiles; ++i) {
        const int buf = i & 1;
        if (!mbar_wait(smem_u32(&bars->yfull[buf]), (i >> 1) & 1, p.err, 37)) break;
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + buf * NT + c0, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->yempty[buf]));
        const int tile = blockIdx.x + i * gridDim.x;
        const int n = tile / tiles_per_img, t2 = tile - n * tiles_per_img;
        const int y0 = (t2 / p.tiles_x) * TY, x0 = (t2 % p.tiles_x) * TX;
        if (p.unshuffle && !mbar_wait(smem_u32(&bars->tfull[buf]), (i >> 1) & 1, p.err, 39)) break;
        const uint8_t* trow = smem_al + 2 * G::A_BYTES + G::W_BYTES + buf * G::T_BYTES + ei * 128;
        if (save_z && y0 + ty < p.H && x0 + tx < p.W) {   // rare: PReLU slope <= 0 (see act_bwd_kernel)
          __nv_bfloat16* zrow = p.zsave + (((size_t)n * (p.H + 2) + y0 + ty + 1) * (p.W + 2) + x0 + tx + 1) * p.y_stride +
                                p.y_col0 + c0;
          for (int j = 0; j < 4; ++j) {
            if (c0 + j * 8 >= p.n_valid) break;
            uint32_t w4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 z = __floats2bfloat162_rn(__uint_as_float(v[j * 8 + 2 * e]) + bias[j * 8 + 2 * e],
                                                       __uint_as_float(v[j * 8 + 2 * e + 1]) + bias[j * 8 + 2 * e + 1]);
              w4[e] = *reinterpret_cast<uint32_t*>(&z);
            }
            *reinterpret_cast<uint4*>(zrow + j * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = __uint_as_float(v[j * 8 + e]) + bias[j * 8 + e];
            if (p.act == SRK_ACT_RELU) a = fmaxf(a, 0.f);
            else if (p.act == SRK_ACT_PRELU) a = a > 0.f ? a : alpha * a;
            f[e] = a;
          }
          if (p.unshuffle && !ps_use_z) {   // PReLU backward of the layer below: mask by the sign of its OUTPUT (the T64 tile)
            const uint4 tq = *reinterpret_cast<const uint4*>(trow + (((ch * 4 + j) ^ (ei & 7)) << 4));
            const __nv_bfloat162* th = reinterpret_cast<const __nv_bfloat162*>(&tq);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 tv = __bfloat1622float2(th[e]);
              if (!(tv.x > 0.f)) { ps_da = fmaf(f[2 * e], tv.x * ps_inv_alpha, ps_da); f[2 * e] *= ps_alpha; }
              if (!(tv.y > 0.f)) { ps_da = fmaf(f[2 * e + 1], tv.y * ps_inv_alpha, ps_da); f[2 * e + 1] *= ps_alpha; }
            }
          } else if (p.unshuffle) {         // slope <= 0: mask by the sign of its saved PRE-activation (global memory)
            uint4 tq = make_uint4(0, 0, 0, 0);
            if (y0 + ty < p.H && x0 + tx < p.W)
              tq = *reinterpret_cast<const uint4*>(p.ps_zsave + (((size_t)n * (p.H + 2) + y0 + ty + 1) * (p.W + 2) + x0 + tx + 1) * 64 +
                                                   c0 + j * 8);
            const __nv_bfloat162* th = reinterpret_cast<const __nv_bfloat162*>(&tq);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 tv = __bfloat1622float2(th[e]);
              if (tv.x < 0.f) { ps_da = fmaf(f[2 * e], tv.x, ps_da); f[2 * e] *= ps_alpha; }
              if (tv.y < 0.f) { ps_da = fmaf(f[2 * e + 1], tv.y, ps_da); f[2 * e + 1] *= ps_alpha; }
            }
          }
          __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
          o[j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                            *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
        if (p.unshuffle) mbar_arrive(smem_u32(&bars->tempty[buf]));   // the T64 tile has been read
        // The [16 x 8 pixels][64 ch] tile goes out as ONE TMA store from a swizzled staging tile (row = pixel):
        // per-thread 16-byte stores at a 128-byte lane stride cost ~1000 L1 wavefronts per tile (measured 176 us of
        // the 612 us output-conv backward).  Pixels / channels outside the tensor are clipped by the TMA unit.
        // unshuffle: staging row = (coarse pixel (ty/2, tx/2), sub-pixel (ty%2)*2 + tx%2) - the 5-D box
        // [8 y2][4 x2][4 sub][64 ch] of the sub-pixel-major dz tensor
        const int srow = p.unshuffle ? (((ty >> 1) * (TX / 2) + (tx >> 1)) * 4 + (ty & 1) * 2 + (tx & 1)) : ei;
        if (threadIdx.x == 320) tma_store_wait_read0();      // the previous tile has left the staging buffer
        asm volatile("bar.sync 3, 256;" ::: "memory");
        uint8_t* orow = ystage + srow * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(orow + (((ch * 4 + j) ^ (srow & 7)) << 4)) = o[j];
        fence_proxy_async();
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (threadIdx.x == 320 && !(p.dbg & 8)) {
          if (p.unshuffle) tma_store_5d(&tmY, y_sm, 0, 0, x0 / 2, y0 / 2, n);
          else tma_store_4d(&tmY, y_sm, p.y_col0, x0, y0, n);
          tma_store_commit();
        }
      }
      if (p.unshuffle && p.want_dalpha) {
        const float t = warp_sum(ps_da);
        if (lane == 0) bars->redda[warp - 10] = t;
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (threadIdx.x == 320) {
          float u = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) u += bars->redda[w8];
          p.ws_small[blockIdx.x * 4 + 3] = u;    // CTA partial; fold_kernel sums the CTAs in order
        }
      }
      if (threadIdx.x == 320) tma_store_wait_all();
    }
    if (p.do_g && my_tiles > 0 && mbar_wait(smem_u32(&bars->done), 0, p.err, 38)) {
      tc_fence_after();
#pragma unroll 1
      for (int mh = 0; mh < NG; ++mh) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + 2 * NT + mh * NT + c0, v);
        tmem_ld_wait();
        // this CTA's partial of G: plain stores; fold_kernel adds the partials in CTA order (float atomics here
        // would make the weight gradient depend on the order in which the CTAs retire)
        float4* dst = reinterpret_cast<float4*>(p.ws + ((size_t)blockIdx.x * G::KP + mh * 128 + ei) * NT + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
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

// sum over the CTA partials ws[b][k][n] (in CTA order) -> OIHW gradient, k = r * SEGW + s * 3 + c.
//   rgb_out = 0: dW[n][c][tap] (3 -> 64 conv), db[n] = column sums (the ones row of G)
//   rgb_out = 1: dW[c][n][taps-1-tap] (64 -> 3 conv; im2col over dY); db3 / dalpha from the per-CTA small partials
// Every output element is WRITTEN exactly once over the 64-channel passes of a call.
// Block = 32 consecutive outputs x 8 row sets: row set g adds partials g, g + 8, ... (148 dependent-free loads per
// output were latency-bound on 64 blocks: 30 us), then the 8 sets are added in set order.
__global__ void __launch_bounds__(256) fold_kernel(const float* __restrict__ ws, const float* __restrict__ ws_small,
                                                   int nblk, int kp, float* __restrict__ dw, float* __restrict__ db,
                                                   float* __restrict__ db3, float* __restrict__ dalpha, int K, int rgb_out,
                                                   int n0, int n_total) {
  __shared__ float part[8][33];
  const int taps = K * K, segw = (K * 3 + 1) / 2 * 2, kone = K * segw;
  const int il = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + il;
  const size_t stride = (size_t)kp * NT;
  const int total = (kone + 1) * NT;
  float v = 0.f;
  if (i < total) {
#pragma unroll 4
    for (int b = g; b < nblk; b += 8) v += __ldg(ws + b * stride + i);
  } else if (i < total + 4) {
    for (int b = g; b < nblk; b += 8) v += __ldg(ws_small + b * 4 + (i - total));
  }
  part[g][il] = v;
  __syncthreads();
  if (g != 0 || i >= total + 4) return;
  v = ((part[0][il] + part[1][il]) + (part[2][il] + part[3][il])) + ((part[4][il] + part[5][il]) + (part[6][il] + part[7][il]));
  if (i >= total) {
    const int c = i - total;
    float* dst = c < 3 ? (db3 ? db3 + c : nullptr) : dalpha;
    if (dst != nullptr) *dst = v;
    return;
  }
  const int n = i % NT, k = i / NT;
  if (n0 + n >= n_total) return;
  if (k == kone) {
    if (db != nullptr && !rgb_out) db[n0 + n] = v;
    return;
  }
  const int r = k / segw, t = k - r * segw;
  if (t >= K * 3) return;
  const int sx = t / 3, c = t - sx * 3, tap = r * K + sx;
  if (!rgb_out) dw[((size_t)(n0 + n) * 3 + c) * taps + tap] = v;
  else dw[((size_t)c * n_total + n0 + n) * taps + (taps - 1 - tap)] = v;
}

static int make_tmap_act_4d_tile(CUtensorMap* out, const srk_tensor* x) {
  PFN_encodeTiled enc = get_encode_tiled();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const uint64_t C = x->c, Wp = x->w + 2, Hp = x->h + 2;
  cuuint64_t dims[4] = {C, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->n};
  cuuint64_t strides[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
  cuuint32_t box[4] = {64, TX, TY, 1};  // 64 channels per pass; a 96-channel tensor's tail is zero-filled
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* base = (char*)x->data + (Wp + 1) * C * 2;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d tile) failed (%d)", (int)r);
  return 0;
}

// dz of the 64 -> 256 PixelShuffle conv below the output conv: bf16 ACT [N][256][H/2][W/2] with channels sub-pixel-major
// (sub * 64 + c), viewed as (c = 64, sub = 4, x2, y2, n); a 16 x 8 tile of fine pixels is the box [8 y2][4 x2][4 sub][64 c]
static int make_tmap_unshuffle_5d(CUtensorMap* out, const srk_tensor* dz) {
  PFN_encodeTiled enc = get_encode_tiled();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const uint64_t C = dz->c, Wp = dz->w + 2, Hp = dz->h + 2;
  cuuint64_t dims[5] = {64, 4, (cuuint64_t)dz->w, (cuuint64_t)dz->h, (cuuint64_t)dz->n};
  cuuint64_t strides[4] = {64 * 2, C * 2, Wp * C * 2, Hp * Wp * C * 2};
  cuuint32_t box[5] = {64, 4, TX / 2, TY / 2, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  void* base = (char*)dz->data + (Wp + 1) * C * 2;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(5d unshuffle) failed (%d)", (int)r);
  return 0;
}

template <int K>
static int launch(const CUtensorMap& tmW, const CUtensorMap& tmT, const CUtensorMap& tmY, const Params& p,
                  cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_rgb_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<K>::SMEM);
    attr = true;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  conv_rgb_tc_kernel<K><<<grid, kThreads, Geo<K>::SMEM, st>>>(tmW, tmT, tmY, p);
  SRK_CUDA_LAUNCH_CHECK("conv_rgb_tc");
  return 0;
}

}  // namespace rgb

// one partial of G ([KP][64] fp32) plus four small partials per CTA
int64_t conv_rgb_workspace_bytes(int k) {
  return (int64_t)kNumSMs * ((int64_t)(k == 9 ? rgb::Geo<9>::KP : rgb::Geo<5>::KP) * 64 + 4) * 4;
}

// t3: IMAGE [N,3,H,W]; y (optional): ACT bf16 [N,64,H,W] = act(conv + bias) with Wk = w_packed (bf16 [64][KP]);
// t64 (optional): ACT bf16 [N,64,H,W] -> dw (OIHW fp32, accumulated), db / db3 (accumulated).
// dz_ps (optional, rgb_out only): instead of y = dgrad(dY) the kernel stores the gradient of the 64 -> 256 PixelShuffle +
// PReLU layer below (see Params::unshuffle); ps_alpha / ps_dalpha: that PReLU's slope and its gradient (accumulated).
int conv_rgb_tc_run(const srk_tensor* t3, const srk_tensor* y, const void* w_packed, const float* bias, int act,
                    const float* alpha, const srk_tensor* t64, float* dw, float* db, float* db3, int rgb_out, int k,
                    void* workspace, cudaStream_t st, const srk_tensor* dz_ps, const float* ps_alpha,
                    float* ps_dalpha, void* zsave, const void* ps_zsave) {
  SRK_REQUIRE(k == 9 || k == 5, "conv_rgb: kernel size must be 9 or 5");
  SRK_REQUIRE(t3 && t3->layout == SRK_LAYOUT_IMAGE && t3->c == 3, "conv_rgb: needs a 3-channel IMAGE tensor");
  rgb::Params p;
  p.N = t3->n; p.H = t3->h; p.W = t3->w;
  p.tiles_x = (p.W + rgb::TX - 1) / rgb::TX; p.tiles_y = (p.H + rgb::TY - 1) / rgb::TY;
  const long long nt = (long long)p.N * p.tiles_x * p.tiles_y;
  SRK_REQUIRE(nt < (1LL << 31), "conv_rgb: too many tiles");
  p.num_tiles = (int)nt;
  p.do_y = y != nullptr || dz_ps != nullptr; p.do_g = t64 != nullptr; p.act = act;
  p.unshuffle = dz_ps != nullptr; p.ps_alpha = ps_alpha; p.want_dalpha = ps_dalpha != nullptr;
  p.zsave = (__nv_bfloat16*)zsave; p.ps_zsave = (const __nv_bfloat16*)ps_zsave;
  if (dz_ps) {
    SRK_REQUIRE(rgb_out && y == nullptr && t64 != nullptr && t64->c == 64 && ps_alpha != nullptr && w_packed != nullptr,
                "conv_rgb: the fused unshuffle needs the 64 -> 3 backward with its 64-channel input");
    SRK_REQUIRE(dz_ps->layout == SRK_LAYOUT_ACT && dz_ps->dtype == SRK_BF16 && dz_ps->c == 256 && dz_ps->n == t3->n &&
                    2 * dz_ps->h == t3->h && 2 * dz_ps->w == t3->w,
                "conv_rgb: dz must be a bf16 ACT tensor [N, 256, H/2, W/2]");
  }
  p.t3 = (const float*)t3->data; p.bias = bias; p.alpha = alpha;
  p.y = y ? (__nv_bfloat16*)y->data : nullptr;
  const int KP = k == 9 ? rgb::Geo<9>::KP : rgb::Geo<5>::KP;
  p.ws = (float*)workspace; p.ws_small = p.ws ? p.ws + (size_t)kNumSMs * KP * 64 : nullptr;
  p.want_db3 = db3 != nullptr; p.err = tc_err_flag();
#ifdef SRK_DEBUG_KNOBS
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("SRK_RGB_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
#else
  p.dbg = 0;
#endif
  CUtensorMap tmW, tmT, tmY;
  memset(&tmW, 0, sizeof(tmW)); memset(&tmT, 0, sizeof(tmT)); memset(&tmY, 0, sizeof(tmY));
  // the 64-channel side may have 64 or 96 channels: one pass per 64-channel chunk (tails are zero-filled by TMA)
  const int c64 = y ? y->c : t64->c;
  if (dz_ps) {
    if (make_tmap_2d_bf16(&tmW, w_packed, 64, (uint64_t)KP, (uint64_t)KP, 64, 64, 128)) return 1;
    if (rgb::make_tmap_unshuffle_5d(&tmY, dz_ps)) return 1;
  } else if (p.do_y) {
    SRK_REQUIRE(y->layout == SRK_LAYOUT_ACT && y->dtype == SRK_BF16 && y->c % 32 == 0 && y->c >= 64 && y->n == p.N &&
                    y->h == p.H && y->w == p.W,
                "conv_rgb: y must be a bf16 ACT tensor with 64 or 96 channels");
    if (make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)(rgb_out ? 64 : y->c), (uint64_t)KP, (uint64_t)KP, 64, 64, 128)) return 1;
    if (rgb::make_tmap_act_4d_tile(&tmY, y)) return 1;   // output tiles leave through a 4-D TMA store
  }
  if (p.do_g) {
    SRK_REQUIRE(t64->layout == SRK_LAYOUT_ACT && t64->dtype == SRK_BF16 && t64->c % 32 == 0 && t64->c >= 64 &&
                    t64->n == p.N && t64->h == p.H && t64->w == p.W,
                "conv_rgb: t64 must be a bf16 ACT tensor with 64 or 96 channels");
    SRK_REQUIRE(workspace != nullptr && dw != nullptr, "conv_rgb: workspace / dw required");
    SRK_REQUIRE(!(y && t64->c != y->c), "conv_rgb: y and t64 channel counts differ");
    if (rgb::make_tmap_act_4d_tile(&tmT, t64)) return 1;
  }
  SRK_REQUIRE(!rgb_out || c64 == 64, "conv_rgb: the 64 -> 3 backward takes a 64-channel input");
  for (int n0 = 0; n0 < c64; n0 += 64) {
    p.y_stride = y ? y->c : 64;
    p.y_col0 = n0;
    p.n_valid = c64 - n0 < 64 ? c64 - n0 : 64;
    p.w_row0 = rgb_out ? 0 : n0;
    p.t_col0 = n0;
    p.want_db3 = n0 == 0 && db3 != nullptr;
    p.want_dalpha = n0 == 0 && ps_dalpha != nullptr;
    int rc = k == 9 ? rgb::launch<9>(tmW, tmT, tmY, p, st) : rgb::launch<5>(tmW, tmT, tmY, p, st);
    if (rc) return rc;
    if (p.do_g) {
      const int total = (k * ((k * 3 + 1) / 2 * 2) + 1) * 64 + 4;
      const int nblk = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;   // = the grid of rgb::launch
      rgb::fold_kernel<<<(total + 31) / 32, 256, 0, st>>>(p.ws, p.ws_small, nblk, KP, dw, db, p.want_db3 ? db3 : nullptr,
                                                            p.want_dalpha ? ps_dalpha : nullptr, k, rgb_out, n0, c64);
      SRK_CUDA_LAUNCH_CHECK("conv_rgb_fold");
    }
  }
  if (dz_ps) return zero_border(dz_ps, st);
  if (y) return zero_border(y, st);
  return 0;
}

}  // namespace srk
