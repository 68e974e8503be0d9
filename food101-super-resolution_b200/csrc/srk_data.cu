// The training-sample pipeline of the reference's FoodSRDataset.__getitem__ (reference src/dataset.py:27-41) as one
// batched kernel: RandomCrop / CenterCrop(crop) + RandomHorizontalFlip + ToTensor (uint8 -> float / 255) give the HR
// crop, transforms.Resize((crop / s, crop / s), BICUBIC) on that TENSOR - i.e. F.interpolate(mode='bicubic',
// antialias=True, align_corners=False), ATen's _upsample_bicubic2d_aa - gives the LR image (not clamped, as in the
// reference).  The host ships the decoded uint8 image (4x fewer bytes than the fp32 pair) and draws the crop offsets
// and flip flags with torch's generator, exactly where torchvision draws them; everything else happens here.
//
// Antialiased bicubic (ATen aten/src/ATen/native/cpu/UpSampleKernel.cpp, _compute_indices_min_size_weights_aa, cubic
// filter a = -0.5): for scale = in / out >= 1 the filter is stretched by scale: support = 2 * scale, centre =
// scale * (i + 0.5), taps j in [xmin, xmin + xsize) with weights filter((j - centre + 0.5) / scale) normalised to
// sum 1.  Separable: horizontal pass into shared memory, then the vertical pass.
//
// One block = LT x LT LR pixels of one image (all three channels): it stages the (LT * s + 2 * support) ^ 2 HR patch
// from uint8 in shared memory (flipped / cropped addressing), writes the HR pixels it owns (coalesced fp32 rows) and
// produces its LR pixels.  HBM traffic: the crop is read once as uint8, HR and LR are written once.
#include "srk_common.cuh"

namespace srk {
namespace data {

constexpr int LT = 8;            // LR pixels per block edge
constexpr int MAX_SCALE = 4;
constexpr int MAX_TAPS = 4 * MAX_SCALE + 2;

__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
  return 0.f;
}

struct Params {
  const uint8_t* src;      // [N][Hs][Ws][3] (hwc != 0) or [N][3][Hs][Ws]
  int hwc, N, Hs, Ws;
  const int* offsets;      // [N][2]: top, left of the crop inside the source image
  const uint8_t* flips;    // [N]: 1 = horizontal flip of the crop
  int crop, scale, lr;     // HR edge, downscale factor, LR edge (crop / scale)
  float* hr;               // [N][3][crop][crop]
  float* lrimg;            // [N][3][lr][lr]
};

template <int S>
__global__ void __launch_bounds__(256) crop_flip_downsample_kernel(const Params p) {
  constexpr int SUP = 2 * S;                       // filter support in HR pixels
  constexpr int PW = LT * S + 2 * SUP + 2;         // staged patch edge (+2: the integer window rounding of ATen)
  extern __shared__ float sm[];
  float* patch = sm;                               // [3][PW][PW] HR values (fp32, already / 255)
  float* hpass = sm + 3 * PW * PW;                 // [3][PW][LT] after the horizontal pass
  const int n = blockIdx.z;
  const int ly0 = blockIdx.y * LT, lx0 = blockIdx.x * LT;
  const int top = p.offsets[2 * n], left = p.offsets[2 * n + 1];
  const bool flip = p.flips[n] != 0;
  // patch origin in HR coordinates (may be negative / run past the crop: such pixels are never given weight)
  const int py0 = ly0 * S - SUP - 1, px0 = lx0 * S - SUP - 1;
  const size_t plane = (size_t)p.Hs * p.Ws;
  for (int i = threadIdx.x; i < PW * PW; i += blockDim.x) {
    const int yy = i / PW, xx = i - yy * PW;
    const int hy = py0 + yy, hx = px0 + xx;
    float v[3] = {0.f, 0.f, 0.f};
    if (hy >= 0 && hy < p.crop && hx >= 0 && hx < p.crop) {
      const int sy = top + hy, sx = left + (flip ? p.crop - 1 - hx : hx);
      if (p.hwc) {
        const uint8_t* s = p.src + ((size_t)n * plane + (size_t)sy * p.Ws + sx) * 3;
        v[0] = (float)s[0] / 255.f; v[1] = (float)s[1] / 255.f; v[2] = (float)s[2] / 255.f;
      } else {
        const uint8_t* s = p.src + (size_t)n * 3 * plane + (size_t)sy * p.Ws + sx;
        v[0] = (float)s[0] / 255.f; v[1] = (float)s[plane] / 255.f; v[2] = (float)s[2 * plane] / 255.f;
      }
      // this block owns the HR pixels under its LR tile
      if (hy >= ly0 * S && hy < (ly0 + LT) * S && hx >= lx0 * S && hx < (lx0 + LT) * S) {
        const size_t o = ((size_t)n * 3 * p.crop + hy) * p.crop + hx;
        p.hr[o] = v[0];
        p.hr[o + (size_t)p.crop * p.crop] = v[1];
        p.hr[o + 2 * (size_t)p.crop * p.crop] = v[2];
      }
    }
    patch[i] = v[0]; patch[PW * PW + i] = v[1]; patch[2 * PW * PW + i] = v[2];
  }
  __syncthreads();
  const float scale = (float)p.crop / (float)p.lr, invscale = 1.f / scale, support = 2.f * scale;
  // horizontal pass: (channel, patch row, LR column)
  for (int i = threadIdx.x; i < 3 * PW * LT; i += blockDim.x) {
    const int lxl = i % LT, yy = (i / LT) % PW, c = i / (LT * PW);
    const int lx = lx0 + lxl;
    float acc = 0.f;
    if (lx < p.lr) {
      const float center = scale * (lx + 0.5f);
      const int xmin = max(0, (int)(center - support + 0.5f));
      const int xsize = min(p.crop, (int)(center + support + 0.5f)) - xmin;
      float wsum = 0.f, w[MAX_TAPS];
#pragma unroll
      for (int j = 0; j < MAX_TAPS; ++j) {
        w[j] = j < xsize ? cubic_aa((j + xmin - center + 0.5f) * invscale) : 0.f;
        wsum += w[j];
      }
      const float* row = patch + (c * PW + yy) * PW + (xmin - px0);
#pragma unroll
      for (int j = 0; j < MAX_TAPS; ++j)
        if (j < xsize) acc = fmaf(w[j] / wsum, row[j], acc);
    }
    hpass[(c * PW + yy) * LT + lxl] = acc;
  }
  __syncthreads();
  // vertical pass: (channel, LR row, LR column)
  for (int i = threadIdx.x; i < 3 * LT * LT; i += blockDim.x) {
    const int lxl = i % LT, lyl = (i / LT) % LT, c = i / (LT * LT);
    const int lx = lx0 + lxl, ly = ly0 + lyl;
    if (lx >= p.lr || ly >= p.lr) continue;
    const float center = scale * (ly + 0.5f);
    const int ymin = max(0, (int)(center - support + 0.5f));
    const int ysize = min(p.crop, (int)(center + support + 0.5f)) - ymin;
    float wsum = 0.f, w[MAX_TAPS];
#pragma unroll
    for (int j = 0; j < MAX_TAPS; ++j) {
      w[j] = j < ysize ? cubic_aa((j + ymin - center + 0.5f) * invscale) : 0.f;
      wsum += w[j];
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_TAPS; ++j)
      if (j < ysize) acc = fmaf(w[j] / wsum, hpass[(c * PW + (ymin - py0) + j) * LT + lxl], acc);
    p.lrimg[((size_t)(n * 3 + c) * p.lr + ly) * p.lr + lx] = acc;
  }
}

}  // namespace data
}  // namespace srk

using namespace srk;

extern "C" int srk_sr_make_batch(const void* src_u8, int hwc, int n, int hs, int ws, const int32_t* offsets,
                                 const uint8_t* flips, int crop, int scale, float* hr, float* lr, void* stream) {
  SRK_REQUIRE(src_u8 && offsets && flips && hr && lr, "srk_sr_make_batch: null argument");
  SRK_REQUIRE(n > 0 && crop > 0 && hs >= crop && ws >= crop, "srk_sr_make_batch: the source image must contain the crop");
  SRK_REQUIRE(scale == 2 || scale == 3 || scale == 4, "srk_sr_make_batch: scale must be 2, 3 or 4");
  SRK_REQUIRE(crop % scale == 0, "srk_sr_make_batch: crop size must be divisible by the scale factor");
  data::Params p;
  p.src = (const uint8_t*)src_u8; p.hwc = hwc; p.N = n; p.Hs = hs; p.Ws = ws;
  p.offsets = offsets; p.flips = flips; p.crop = crop; p.scale = scale; p.lr = crop / scale;
  p.hr = hr; p.lrimg = lr;
  const int tiles = (p.lr + data::LT - 1) / data::LT;
  SRK_REQUIRE(n <= 65535 && tiles <= 65535, "srk_sr_make_batch: batch or image too large for one launch");
  dim3 grid(tiles, tiles, n);
  cudaStream_t st = (cudaStream_t)stream;
#define SRK_LAUNCH_DATA(S)                                                                         \
  do {                                                                                             \
    constexpr int PW = data::LT * S + 4 * S + 2;                                                   \
    const size_t smem = (size_t)(3 * PW * PW + 3 * PW * data::LT) * sizeof(float);                 \
    static bool attr = false;                                                                      \
    if (!attr) {                                                                                   \
      cudaFuncSetAttribute(data::crop_flip_downsample_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      attr = true;                                                                                 \
    }                                                                                              \
    data::crop_flip_downsample_kernel<S><<<grid, 256, smem, st>>>(p);                              \
  } while (0)
  if (scale == 2) SRK_LAUNCH_DATA(2);
  else if (scale == 3) SRK_LAUNCH_DATA(3);
  else SRK_LAUNCH_DATA(4);
#undef SRK_LAUNCH_DATA
  SRK_CUDA_LAUNCH_CHECK("sr_make_batch");
  return 0;
}
