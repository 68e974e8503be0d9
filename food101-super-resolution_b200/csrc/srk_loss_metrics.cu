// Pixel losses (nn.L1Loss / nn.MSELoss, loss.py:81-86), NLPD loss (loss.py:31-79) forward+backward,
// PSNR / SSIM reductions (metrics.py:14-31 via torchmetrics 1.8.2) and a flat Adam step
// (train.py:55,120).  All inputs are NCHW fp32 images; kernels are coalesced along W, vectorised
// where alignment allows, and reduce with warp shuffles -> one atomic per block.
#include "srk_common.cuh"

namespace srk {

constexpr int kLossPartials = 148 * 8;    // upper bound of every reducing grid in this file (red_blocks)

// ---- L1 / MSE -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_loss_fwd_kernel(const float* __restrict__ a,
    const float* __restrict__ b, long long n, int mode, double* __restrict__ acc) {
  __shared__ double red[32];
  float s = 0.f;
  long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  long long stride = (long long)gridDim.x * blockDim.x * 4;
  const bool aligned = ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0;
  for (; i < n; i += stride) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (aligned && i + 3 < n) {
      float4 x = *reinterpret_cast<const float4*>(a + i), y = *reinterpret_cast<const float4*>(b + i);
      d[0] = x.x - y.x; d[1] = x.y - y.y; d[2] = x.z - y.z; d[3] = x.w - y.w;
    } else {
      for (int j = 0; j < 4 && i + j < n; ++j) d[j] = a[i + j] - b[i + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) s += mode == 0 ? fabsf(d[j]) : d[j] * d[j];
  }
  double t = block_sum_d((double)s, red);
  if (threadIdx.x == 0) acc[blockIdx.x] = t;      // per-block partial: summed in block order by the finishing kernel
}
// Fixed-order sum of `n` per-block partials by one 256-thread block (thread t takes partials t, t + 256, ...; then the
// block tree of block_sum_d): the loss value does not depend on the order in which blocks retire.
__device__ __forceinline__ double ordered_partial_sum(const double* part, int n, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  return block_sum_d(s, red);
}
__global__ void __launch_bounds__(256) scale_to_float_kernel(const double* part, int n, double scale, float* out) {
  __shared__ double red[32];
  const double t = ordered_partial_sum(part, n, red);
  if (threadIdx.x == 0) out[0] = (float)(t * scale);
}

__global__ void __launch_bounds__(256) pixel_loss_bwd_kernel(const float* __restrict__ a,
    const float* __restrict__ b, long long n, int mode, const float* __restrict__ gout,
    float* __restrict__ g) {
  const float go = gout[0] / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - b[i];
    g[i] = mode == 0 ? (d > 0.f ? go : (d < 0.f ? -go : 0.f)) : 2.f * d * go;
  }
}

// ---- NLPD ------------------------------------------------------------------------------------------
// The Laplacian pyramid is linear, so Lap_l(sr) - Lap_l(hr) = Lap_l(sr - hr): one pyramid of the
// difference image d is built (half the reference's work, same value up to fp32 rounding).
// Workspace layout (floats): for l in 0..L: cur_l [NC*h_l*w_l] (cur_0 = d); then per level l<L a
// sign map s_l [NC*h_l*w_l] (int8), and two ping-pong gradient buffers of the level-0/level-1 size.
struct NlpdPlan {
  int L;
  int h[8], w[8];
  long long cur_off[8];   // float offsets
  long long sign_off[8];  // byte offsets from sign base
  long long g_off[2];     // float offsets
  long long floats, sign_bytes, total_bytes;
  long long acc_off;      // byte offset of double acc[8][kLossPartials] (per-block partial sums)
};
static NlpdPlan nlpd_plan(int n, int c, int h, int w, int L) {
  NlpdPlan p; p.L = L;
  long long nc = (long long)n * c, off = 0;
  p.h[0] = h; p.w[0] = w;
  // every region starts on a 16-byte boundary: the exact-2x kernels use float2 / char2 accesses on fine rows
  auto pad4 = [](long long v) { return (v + 3) / 4 * 4; };
  for (int l = 0; l <= L; ++l) {
    if (l > 0) { p.h[l] = (p.h[l - 1] + 1) / 2; p.w[l] = (p.w[l - 1] + 1) / 2; }
    p.cur_off[l] = off; off += pad4(nc * p.h[l] * p.w[l]);
  }
  p.g_off[0] = off; off += pad4(nc * p.h[0] * p.w[0]);
  p.g_off[1] = off; off += pad4(nc * p.h[1] * p.w[1]);
  p.floats = off;
  long long sb = 0;
  for (int l = 0; l < L; ++l) { p.sign_off[l] = sb; sb += (nc * p.h[l] * p.w[l] + 15) / 16 * 16; }
  p.sign_bytes = sb;
  p.acc_off = (p.floats * 4 + 15) / 16 * 16 + sb;
  p.total_bytes = p.acc_off + 8 * (long long)kLossPartials * sizeof(double);
  return p;
}

__global__ void __launch_bounds__(256) nlpd_diff_kernel(const float* __restrict__ a,
    const float* __restrict__ b, long long n, int clamp01, float* __restrict__ d, double* __restrict__ acc) {
  __shared__ double red[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float x = a[i], y = b[i];
    if (clamp01) { x = fminf(fmaxf(x, 0.f), 1.f); y = fminf(fmaxf(y, 0.f), 1.f); }
    float v = x - y;
    d[i] = v; s += fabsf(v);
  }
  double t = block_sum_d((double)s, red);
  if (threadIdx.x == 0) acc[blockIdx.x] = t;
}

// Row-wise iteration over an [NC][H][W] image stack: a block walks whole rows (several per pass when W is
// narrower than the block), so the (nc, y, x) of an element costs one 32-bit division per ROW instead of two
// 64-bit divisions per element.
#define NLPD_LOOP_BEGIN(NC_, H_, W_)                                                                   \
  {                                                                                                    \
    const bool wide__ = (W_) >= (int)blockDim.x;                                                       \
    const int rpi__ = wide__ ? 1 : (int)blockDim.x / (W_);                                             \
    const int tr__ = wide__ ? 0 : (int)threadIdx.x / (W_);                                             \
    const int tx__ = wide__ ? (int)threadIdx.x : (int)threadIdx.x - tr__ * (W_);                       \
    const int xs__ = wide__ ? (int)blockDim.x : (W_);                                                  \
    const int rows__ = (NC_) * (H_);                                                                   \
    for (int row = blockIdx.x * rpi__ + tr__; row < rows__ && tr__ < rpi__; row += gridDim.x * rpi__) { \
      const int nc = row / (H_), y = row - nc * (H_);                                                  \
      for (int x = tx__; x < (W_); x += xs__) {                                                        \
        const long long i = (long long)row * (W_) + x;
#define NLPD_LOOP_END }}}

// down[y][x] = sum_{ky,kx} k[ky][kx] * cur[2y+ky-2][2x+kx-2] (zero padded)   (loss.py:61-62)
__global__ void __launch_bounds__(256) nlpd_blur_down_kernel(const float* __restrict__ cur, int NC, int H,
    int W, int h2, int w2, const float* __restrict__ k25, float* __restrict__ down) {
  __shared__ float k[25];
  if (threadIdx.x < 25) k[threadIdx.x] = k25[threadIdx.x];
  __syncthreads();
  NLPD_LOOP_BEGIN(NC, h2, w2)
    const float* p = cur + (long long)nc * H * W;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      int yy = 2 * y + ky - 2;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        int xx = 2 * x + kx - 2;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(k[ky * 5 + kx], p[(long long)yy * W + xx], acc);
      }
    }
    down[i] = acc;
  NLPD_LOOP_END
}

// Even fine width (W = 2 w2): one thread produces TWO horizontally adjacent outputs from the 5 x 7 fine window they
// share, read as three float2 and one float per row (20 load instructions for two outputs instead of 50).  Taps that
// fall outside the image contribute k * 0 (the generic kernel skips them): the same fmaf chain, identical bits.
__global__ void __launch_bounds__(256) nlpd_blur_down_pair_kernel(const float* __restrict__ cur, int NC, int H,
    int W, int h2, int w2, const float* __restrict__ k25, float* __restrict__ down) {
  __shared__ float k[25];
  if (threadIdx.x < 25) k[threadIdx.x] = k25[threadIdx.x];
  __syncthreads();
  const int pairs = (w2 + 1) / 2;
  NLPD_LOOP_BEGIN(NC, h2, pairs)
    (void)i;
    const float* p = cur + (long long)nc * H * W;
    const int c0 = 4 * x - 2;   // first fine column of the window; outputs 2x and 2x + 1
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int yy = 2 * y + ky - 2;
      if (yy < 0 || yy >= H) continue;
      const float* row = p + (long long)yy * W + c0;
      float v[7];
      {
        const float2 t0 = c0 >= 0 ? *reinterpret_cast<const float2*>(row) : make_float2(0.f, 0.f);
        const float2 t1 = *reinterpret_cast<const float2*>(row + 2);   // columns 4x, 4x + 1 always exist
        const float2 t2 = c0 + 4 < W ? *reinterpret_cast<const float2*>(row + 4) : make_float2(0.f, 0.f);
        v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y; v[4] = t2.x; v[5] = t2.y;
        v[6] = c0 + 6 < W ? row[6] : 0.f;
      }
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        a0 = fmaf(k[ky * 5 + kx], v[kx], a0);
        a1 = fmaf(k[ky * 5 + kx], v[kx + 2], a1);
      }
    }
    float* o = down + ((long long)nc * h2 + y) * w2 + 2 * x;
    if ((w2 & 1) == 0) {
      *reinterpret_cast<float2*>(o) = make_float2(a0, a1);
    } else {
      o[0] = a0;
      if (2 * x + 1 < w2) o[1] = a1;
    }
  NLPD_LOOP_END
}

// bilinear, align_corners=False, explicit output size (loss.py:63): src = max(scale*(o+.5)-.5, 0)
__device__ __forceinline__ void bilin_src(int o, float scale, int in, int& i0, int& i1, float& lam) {
  float src = fmaxf(scale * (o + 0.5f) - 0.5f, 0.f);
  i0 = (int)src; if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  lam = src - (float)i0;
  lam = fminf(fmaxf(lam, 0.f), 1.f);
}

// The generic and the exact-2x kernels share these two expressions with explicit roundings, so that a level gives the
// same bits on either path (the compiler is otherwise free to contract a*b + c*d either way).
__device__ __forceinline__ float bilerp(float ly, float lx, float a00, float a01, float a10, float a11) {
  const float top = __fmaf_rn(lx, a01, __fmul_rn(1.f - lx, a00));
  const float bot = __fmaf_rn(lx, a11, __fmul_rn(1.f - lx, a10));
  return __fmaf_rn(ly, bot, __fmul_rn(1.f - ly, top));
}
__device__ __forceinline__ float gdown_of(float g_next, float c_l, float acc) { return __fmaf_rn(-c_l, acc, g_next); }
__device__ __forceinline__ signed char sign_of(float d) { return d > 0.f ? 1 : (d < 0.f ? -1 : 0); }

// diff = cur - up(down); accumulates sum|diff|; optionally stores sign(diff) as int8
__global__ void __launch_bounds__(256) nlpd_lap_abs_kernel(const float* __restrict__ cur,
    const float* __restrict__ down, int NC, int H, int W, int h2, int w2, float sy, float sx,
    signed char* __restrict__ sign, double* __restrict__ acc) {
  __shared__ double red[32];
  float s = 0.f;
  NLPD_LOOP_BEGIN(NC, H, W)
    int y0, y1, x0, x1; float ly, lx;
    bilin_src(y, sy, h2, y0, y1, ly);
    bilin_src(x, sx, w2, x0, x1, lx);
    const float* p = down + (long long)nc * h2 * w2;
    float up = bilerp(ly, lx, p[y0 * w2 + x0], p[y0 * w2 + x1], p[y1 * w2 + x0], p[y1 * w2 + x1]);
    float d = cur[i] - up;
    s += fabsf(d);
    if (sign) sign[i] = sign_of(d);
  NLPD_LOOP_END
  double t = block_sum_d((double)s, red);
  if (threadIdx.x == 0) acc[blockIdx.x] = t;
}

// Exact-2x levels (H = 2 h2, W = 2 w2; every level of a power-of-two crop): one thread per COARSE pixel produces its
// 2x2 block of fine outputs.  align_corners=False at scale 1/2 puts fine row 2y at coarse y - 0.25 (rows y-1, y with
// lambda 0.75; row 0 alone when y = 0) and fine row 2y+1 at y + 0.25 (rows y, min(y+1, h2-1) with lambda 0.25): the
// 3x3 coarse neighbourhood is loaded once (9 loads for 4 outputs instead of 16), bilin_src disappears, the fine row
// goes out as float2 / char2.  Same bilerp() on the same operands as the generic kernel: identical signs.
__global__ void __launch_bounds__(256) nlpd_lap_abs_2x_kernel(const float* __restrict__ cur,
    const float* __restrict__ down, int NC, int H, int W, int h2, int w2, signed char* __restrict__ sign,
    double* __restrict__ acc) {
  __shared__ double red[32];
  float s = 0.f;
  NLPD_LOOP_BEGIN(NC, h2, w2)
    (void)i;
    const float* p = down + (long long)nc * h2 * w2;
    const int yr[3] = {max(y - 1, 0), y, min(y + 1, h2 - 1)}, xr[3] = {max(x - 1, 0), x, min(x + 1, w2 - 1)};
    float g[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) g[a][b] = p[yr[a] * w2 + xr[b]];
    const float lys[2] = {y == 0 ? 0.f : 0.75f, 0.25f}, lxs[2] = {x == 0 ? 0.f : 0.75f, 0.25f};
    const long long base = ((long long)nc * H + 2 * y) * W + 2 * x;
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const long long o = base + (long long)py * W;
      const float2 c = *reinterpret_cast<const float2*>(cur + o);
      const float d0 = c.x - bilerp(lys[py], lxs[0], g[py][0], g[py][1], g[py + 1][0], g[py + 1][1]);
      const float d1 = c.y - bilerp(lys[py], lxs[1], g[py][1], g[py][2], g[py + 1][1], g[py + 1][2]);
      s += fabsf(d0);
      s += fabsf(d1);
      if (sign) *reinterpret_cast<char2*>(sign + o) = make_char2(sign_of(d0), sign_of(d1));
    }
  NLPD_LOOP_END
  double t = block_sum_d((double)s, red);
  if (threadIdx.x == 0) acc[blockIdx.x] = t;
}

struct NlpdWeights { double w[8]; int nblk[8]; };
// acc: [8][kLossPartials] per-block partials of the L1 term (row 0) and of the pyramid levels (rows 1..)
__global__ void __launch_bounds__(256) nlpd_combine_kernel(const double* acc, NlpdWeights wt, float* loss) {
  __shared__ double red[32];
  double v = 0.0;
  for (int i = 0; i < 8; ++i) {
    if (wt.nblk[i] == 0) continue;
    const double t = ordered_partial_sum(acc + (size_t)i * kLossPartials, wt.nblk[i], red);
    v += t * wt.w[i];     // valid in thread 0
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)v;
}

// g_down[y][x] = G_next[y][x] - c_l * sum_{Y,X} Uy[Y][y] Ux[X][x] sign_l[Y][X]    (bilinear^T)
__global__ void __launch_bounds__(256) nlpd_bwd_down_kernel(const signed char* __restrict__ sign,
    const float* __restrict__ g_next, int NC, int H, int W, int h2, int w2, float sy, float sx, float c_l,
    float* __restrict__ g_down) {
  NLPD_LOOP_BEGIN(NC, h2, w2)
    const signed char* sp = sign + (long long)nc * H * W;
    float acc = 0.f;
    // an exact 2x level touches fine rows / columns 2y-1 .. 2y+2 only; other ratios get the wider safe window
    const int ry = (H == 2 * h2) ? 1 : 3, rx = (W == 2 * w2) ? 1 : 3;
    // the column weights do not depend on the row: compute them once (up to 8 columns)
    const int X0 = max(2 * x - rx, 0), X1 = min(2 * x + rx + 1, W - 1);
    float wxs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int X = X0 + k;
      float wv = 0.f;
      if (X <= X1) {
        int x0, x1; float lx;
        bilin_src(X, sx, w2, x0, x1, lx);
        wv = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
      }
      wxs[k] = wv;
    }
    for (int Y = max(2 * y - ry, 0); Y <= min(2 * y + ry + 1, H - 1); ++Y) {
      int y0, y1; float ly;
      bilin_src(Y, sy, h2, y0, y1, ly);
      float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      if (wy == 0.f) continue;
      const signed char* row = sp + (long long)Y * W + X0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (X0 + k > X1 || wxs[k] == 0.f) continue;
        acc = fmaf(wy * wxs[k], (float)row[k], acc);
      }
    }
    g_down[i] = gdown_of(g_next ? g_next[i] : 0.f, c_l, acc);
  NLPD_LOOP_END
}

// Exact-2x levels: coarse pixel (y, x) receives fine rows 2y-1 .. 2y+2 with weights 0.25, 0.75, 0.75, 0.25 - the
// transposed rule of nlpd_lap_abs_2x_kernel; at the borders the missing outer row drops out and the outermost fine row
// counts fully (weight 1: both of its bilinear taps are the border pixel) - and the same in x.  No bilin_src, a
// char / char2 / char per fine row, same products in the same order as the generic kernel: identical bits.
__global__ void __launch_bounds__(256) nlpd_bwd_down_2x_kernel(const signed char* __restrict__ sign,
    const float* __restrict__ g_next, int NC, int H, int W, int h2, int w2, float c_l,
    float* __restrict__ g_down) {
  NLPD_LOOP_BEGIN(NC, h2, w2)
    const signed char* sp = sign + ((long long)nc * H + 2 * y) * W + 2 * x;   // fine pixel (2y, 2x)
    const float wys[4] = {0.25f, y == 0 ? 1.f : 0.75f, y == h2 - 1 ? 1.f : 0.75f, 0.25f};
    const float wxs[4] = {0.25f, x == 0 ? 1.f : 0.75f, x == w2 - 1 ? 1.f : 0.75f, 0.25f};
    const bool left = x > 0, right = x < w2 - 1;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if ((r == 0 && y == 0) || (r == 3 && y == h2 - 1)) continue;
      const signed char* row = sp + (long long)(r - 1) * W;
      const char2 mid = *reinterpret_cast<const char2*>(row);
      if (left) acc = fmaf(wys[r] * wxs[0], (float)row[-1], acc);
      acc = fmaf(wys[r] * wxs[1], (float)mid.x, acc);
      acc = fmaf(wys[r] * wxs[2], (float)mid.y, acc);
      if (right) acc = fmaf(wys[r] * wxs[3], (float)row[2], acc);
    }
    g_down[i] = gdown_of(g_next ? g_next[i] : 0.f, c_l, acc);
  NLPD_LOOP_END
}

// G_l[Y][X] = c_l*sign_l[Y][X] + sum_{ky,kx: (Y+2-ky),(X+2-kx) even} k[ky][kx] g_down[(Y+2-ky)/2][(X+2-kx)/2]
// level 0 additionally adds alpha*sign(d)/numel, multiplies by gout and writes grad_sr.
__global__ void __launch_bounds__(256) nlpd_bwd_up_kernel(const signed char* __restrict__ sign,
    const float* __restrict__ g_down, int NC, int H, int W, int h2, int w2, const float* __restrict__ k25,
    float c_l, const float* __restrict__ d0, float c_mae, const float* __restrict__ gout,
    float* __restrict__ G) {
  __shared__ float k[25];
  if (threadIdx.x < 25) k[threadIdx.x] = k25[threadIdx.x];
  __syncthreads();
  const float go = gout ? gout[0] : 1.f;
  NLPD_LOOP_BEGIN(NC, H, W)
    const int X = x, Y = y;
    const float* gp = g_down + (long long)nc * h2 * w2;
    float acc = c_l * (float)sign[i];
    // blurred[yy][xx] reads cur[yy+ky-2][xx+kx-2]; only even (yy,xx) are kept as down[yy/2][xx/2]
    // only taps with Y + 2 - ky even (and likewise in x) land on a kept sample: step the taps by 2
    for (int ky = Y & 1; ky < 5; ky += 2) {
      int yy = Y + 2 - ky;
      if (yy < 0 || (yy >> 1) >= h2) continue;
      for (int kx = X & 1; kx < 5; kx += 2) {
        int xx = X + 2 - kx;
        if (xx < 0 || (xx >> 1) >= w2) continue;
        acc = fmaf(k[ky * 5 + kx], gp[(long long)(yy >> 1) * w2 + (xx >> 1)], acc);
      }
    }
    if (d0) {
      float d = d0[i];
      acc += d > 0.f ? c_mae : (d < 0.f ? -c_mae : 0.f);
      acc *= go;
    }
    G[i] = acc;
  NLPD_LOOP_END
}

// Exact-2x levels (H = 2 h2, W = 2 w2): one thread per COARSE pixel produces the 2x2 block of fine outputs from the
// 3x3 neighbourhood of g_down it shares (9 loads for 4 outputs instead of ~6 per output), same tap order as the
// generic kernel, so the results are bit-identical.
__global__ void __launch_bounds__(256) nlpd_bwd_up_2x_kernel(const signed char* __restrict__ sign,
    const float* __restrict__ g_down, int NC, int H, int W, int h2, int w2, const float* __restrict__ k25,
    float c_l, const float* __restrict__ d0, float c_mae, const float* __restrict__ gout,
    float* __restrict__ G) {
  __shared__ float k[25];
  if (threadIdx.x < 25) k[threadIdx.x] = k25[threadIdx.x];
  __syncthreads();
  const float go = gout ? gout[0] : 1.f;
  NLPD_LOOP_BEGIN(NC, h2, w2)
    (void)i;
    const float* gp = g_down + (long long)nc * h2 * w2;
    float gv[3][3];   // gv[a + 1][b + 1] = g_down[y + a][x + b], 0 outside
#pragma unroll
    for (int a = -1; a <= 1; ++a)
#pragma unroll
      for (int b = -1; b <= 1; ++b) {
        const int yy = y + a, xx = x + b;
        gv[a + 1][b + 1] = (yy >= 0 && yy < h2 && xx >= 0 && xx < w2) ? gp[(long long)yy * w2 + xx] : 0.f;
      }
    const bool ya[3] = {y - 1 >= 0, true, y + 1 < h2}, xa[3] = {x - 1 >= 0, true, x + 1 < w2};
    const long long base = ((long long)nc * H + 2 * y) * W + 2 * x;
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const char2 sg = *reinterpret_cast<const char2*>(sign + base + (long long)py * W);
      float acc[2] = {c_l * (float)sg.x, c_l * (float)sg.y};
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        // generic order: ky = Y&1, +2, ... (i.e. a = +1 first), kx likewise
#pragma unroll
        for (int a = 1; a >= (py ? 0 : -1); --a) {
          const int ky = (py ? 3 : 2) - 2 * a;
          if (!ya[a + 1]) continue;
#pragma unroll
          for (int b = 1; b >= (px ? 0 : -1); --b) {
            const int kx = (px ? 3 : 2) - 2 * b;
            if (!xa[b + 1]) continue;
            acc[px] = fmaf(k[ky * 5 + kx], gv[a + 1][b + 1], acc[px]);
          }
        }
      }
      const long long o = base + (long long)py * W;
      if (d0) {
        const float2 d = *reinterpret_cast<const float2*>(d0 + o);
        acc[0] += d.x > 0.f ? c_mae : (d.x < 0.f ? -c_mae : 0.f);
        acc[1] += d.y > 0.f ? c_mae : (d.y < 0.f ? -c_mae : 0.f);
        acc[0] *= go; acc[1] *= go;
      }
      *reinterpret_cast<float2*>(G + o) = make_float2(acc[0], acc[1]);
    }
  NLPD_LOOP_END
}

// ---- PSNR / SSIM ------------------------------------------------------------------------------------
// Per-image sum of squared errors: 16-byte loads (four pixels of both images per thread and step, two steps in
// flight), fp32 partial per thread, fp64 from the warp level up.  The per-image totals are accumulated with fp64
// atomics (a handful of blocks per image): order-dependent only at the 1e-16 relative level, far inside the
// 0.01 dB the metric is specified to.
__global__ void __launch_bounds__(256) psnr_sse_kernel(const float* __restrict__ a, const float* __restrict__ b,
    long long per_image, int clamp01, double* __restrict__ sse) {
  __shared__ double red[32];
  const int n = blockIdx.y;
  const float* pa = a + (long long)n * per_image;
  const float* pb = b + (long long)n * per_image;
  float s = 0.f;
  auto term = [&](float x, float y) {
    if (clamp01) { x = fminf(fmaxf(x, 0.f), 1.f); y = fminf(fmaxf(y, 0.f), 1.f); }
    const float d = x - y;
    s = fmaf(d, d, s);
  };
  const bool vec = (per_image & 3) == 0 && ((((uintptr_t)pa) | ((uintptr_t)pb)) & 15) == 0;
  if (vec) {
    const long long n4 = per_image >> 2, stride = (long long)gridDim.x * blockDim.x;
    const float4* qa = reinterpret_cast<const float4*>(pa);
    const float4* qb = reinterpret_cast<const float4*>(pb);
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {
      const float4 x0 = __ldg(qa + i), y0 = __ldg(qb + i), x1 = __ldg(qa + i + stride), y1 = __ldg(qb + i + stride);
      term(x0.x, y0.x); term(x0.y, y0.y); term(x0.z, y0.z); term(x0.w, y0.w);
      term(x1.x, y1.x); term(x1.y, y1.y); term(x1.z, y1.z); term(x1.w, y1.w);
    }
    if (i < n4) {
      const float4 x0 = __ldg(qa + i), y0 = __ldg(qb + i);
      term(x0.x, y0.x); term(x0.y, y0.y); term(x0.z, y0.z); term(x0.w, y0.w);
    }
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_image;
         i += (long long)gridDim.x * blockDim.x)
      term(pa[i], pb[i]);
  }
  double t = block_sum_d((double)s, red);
  if (threadIdx.x == 0) atomicAdd(&sse[n], t);
}

// SSIM (gaussian 11x11, sigma 1.5, k1=.01, k2=.03, data_range 1): the reflect padding of
// torchmetrics is cropped away again, so only the (H-10)x(W-10) valid windows contribute.
// Tile: 32x16 outputs from a 42x26 input patch; separable: horizontal pass of the 5 maps into shared
// memory, vertical pass per output.
constexpr int ST_W = 32, ST_H = 16, SK = 11;
__global__ void __launch_bounds__(ST_W * ST_H) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b,
    int C, int H, int W, int clamp01, double* __restrict__ out) {
  __shared__ float pa[ST_H + SK - 1][ST_W + SK - 1 + 1];
  __shared__ float pb[ST_H + SK - 1][ST_W + SK - 1 + 1];
  __shared__ float hm[5][ST_H + SK - 1][ST_W + 1];
  __shared__ float g[SK];
  __shared__ double red[32];
  const int tid = threadIdx.y * ST_W + threadIdx.x;
  if (tid == 0) {
    double s = 0.0, t[SK];
    for (int i = 0; i < SK; ++i) { double d = (i - 5) / 1.5; t[i] = exp(-d * d / 2.0); s += t[i]; }
    for (int i = 0; i < SK; ++i) g[i] = (float)(t[i] / s);
  }
  const int nc = blockIdx.z, n = nc / C;
  const int OH = H - (SK - 1), OW = W - (SK - 1);
  const int ox0 = blockIdx.x * ST_W, oy0 = blockIdx.y * ST_H;
  const float* A = a + (long long)nc * H * W;
  const float* B = b + (long long)nc * H * W;
  for (int i = tid; i < (ST_H + SK - 1) * (ST_W + SK - 1); i += ST_W * ST_H) {
    int r = i / (ST_W + SK - 1), cc = i - r * (ST_W + SK - 1);
    int y = oy0 + r, x = ox0 + cc;
    float va = 0.f, vb = 0.f;
    if (y < H && x < W) {
      va = A[(long long)y * W + x]; vb = B[(long long)y * W + x];
      if (clamp01) { va = fminf(fmaxf(va, 0.f), 1.f); vb = fminf(fmaxf(vb, 0.f), 1.f); }
    }
    pa[r][cc] = va; pb[r][cc] = vb;
  }
  __syncthreads();
  for (int i = tid; i < (ST_H + SK - 1) * ST_W; i += ST_W * ST_H) {
    int r = i / ST_W, cc = i - r * ST_W;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float w = g[k], x = pa[r][cc + k], y = pb[r][cc + k];
      m0 = fmaf(w, x, m0); m1 = fmaf(w, y, m1);
      m2 = fmaf(w, x * x, m2); m3 = fmaf(w, y * y, m3); m4 = fmaf(w, x * y, m4);
    }
    hm[0][r][cc] = m0; hm[1][r][cc] = m1; hm[2][r][cc] = m2; hm[3][r][cc] = m3; hm[4][r][cc] = m4;
  }
  __syncthreads();
  float val = 0.f;
  {
    int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
    if (oy < OH && ox < OW) {
      float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
      for (int k = 0; k < SK; ++k) {
        float w = g[k];
        m0 = fmaf(w, hm[0][threadIdx.y + k][threadIdx.x], m0);
        m1 = fmaf(w, hm[1][threadIdx.y + k][threadIdx.x], m1);
        m2 = fmaf(w, hm[2][threadIdx.y + k][threadIdx.x], m2);
        m3 = fmaf(w, hm[3][threadIdx.y + k][threadIdx.x], m3);
        m4 = fmaf(w, hm[4][threadIdx.y + k][threadIdx.x], m4);
      }
      const float c1 = 1e-4f, c2 = 9e-4f;
      float mu_pp = m0 * m0, mu_tt = m1 * m1, mu_pt = m0 * m1;
      float s_pp = fmaxf(m2 - mu_pp, 0.f), s_tt = fmaxf(m3 - mu_tt, 0.f), s_pt = m4 - mu_pt;
      val = ((2.f * mu_pt + c1) * (2.f * s_pt + c2)) / ((mu_pp + mu_tt + c1) * (s_pp + s_tt + c2));
    }
  }
  // block reduce (2-D block -> linear tid)
  double v = warp_sum_d((double)val);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    double r = tid < (ST_W * ST_H / 32) ? red[tid] : 0.0;
    r = warp_sum_d(r);
    if (tid == 0) atomicAdd(&out[n], r);
  }
}

// ---- Adam -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
    float* __restrict__ m, float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
    const long long* __restrict__ step, float grad_scale) {
  const double t = (double)step[0];
  const float bc1 = (float)(1.0 - pow((double)b1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    float mi = m[i] + (1.f - b1) * (gi - m[i]);      // torch: exp_avg.lerp_(grad, 1-beta1)
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

// Multi-tensor Adam: one launch updates up to ADAM_MT tensors; the pointer table travels by value in the
// kernel parameters (so the launch is CUDA-graph capturable without a host-side table upload).
constexpr int ADAM_MT = 48;
struct AdamTable {
  float* p[ADAM_MT];
  const float* g[ADAM_MT];
  float* m[ADAM_MT];
  float* v[ADAM_MT];
  int n[ADAM_MT];
};
__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamTable tb, float lr, float b1, float b2, float eps,
    const long long* __restrict__ step, float grad_scale, const float* __restrict__ lr_dev,
    const float* __restrict__ scale_dev) {
  // device-resident learning rate / gradient scale: a captured graph keeps working when a scheduler changes the
  // rate, and a gradient-norm clip can feed its factor without a host round trip
  if (lr_dev) lr = lr_dev[0];
  if (scale_dev) grad_scale *= scale_dev[0];
  const int t = blockIdx.y;
  const int n = tb.n[t];
  const int i0 = blockIdx.x * 1024;
  if (i0 >= n) return;
  const double tt = (double)step[0];
  const float bc1 = (float)(1.0 - pow((double)b1, tt));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, tt));
  const float step_size = lr / bc1;
  float* __restrict__ p = tb.p[t];
  const float* __restrict__ g = tb.g[t];
  float* __restrict__ m = tb.m[t];
  float* __restrict__ v = tb.v[t];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + k * 256 + threadIdx.x;
    if (i < n) {
      float gi = g[i] * grad_scale;
      float mi = m[i] + (1.f - b1) * (gi - m[i]);
      float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi; v[i] = vi;
      p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
  }
}

}  // namespace srk

using namespace srk;

// SRK_NLPD_2X=0 keeps exact-2x pyramid levels on the generic kernels (read per call: the tests compare both paths).
static inline bool nlpd_2x() {
  const char* e = getenv("SRK_NLPD_2X");
  return e == nullptr || atoi(e) != 0;
}

// The NLPD kernels are row-stride loops: more blocks than fit on the GPU at once only add a partial second wave
// (lap_abs_2x: 40 registers -> 6 blocks / SM, so 1184 blocks ran as 1.33 waves).  Cap the grid at what is resident.
template <typename K>
static int resident_cap(K kernel, int blocks) {
  int occ = 0;   // queried per launch (microseconds, and nothing at all when a captured graph is replayed)
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) occ = 1;
  const int cap = kNumSMs * occ;
  return blocks < cap ? blocks : cap;
}

static inline int red_blocks(long long n, int per_thread) {
  long long b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int64_t srk_pixel_loss_scratch_bytes(void) { return (int64_t)kLossPartials * sizeof(double); }

extern "C" int srk_pixel_loss_fwd(const float* sr, const float* hr, int64_t numel, int mode, float* loss,
                                  double* scratch, void* stream) {
  SRK_REQUIRE(numel > 0 && (mode == 0 || mode == 1), "srk_pixel_loss_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = red_blocks(numel, 16);   // <= kLossPartials: scratch holds one double per block
  pixel_loss_fwd_kernel<<<blocks, 256, 0, st>>>(sr, hr, numel, mode, scratch);
  scale_to_float_kernel<<<1, 256, 0, st>>>(scratch, blocks, 1.0 / (double)numel, loss);
  SRK_CUDA_LAUNCH_CHECK("pixel_loss_fwd");
  return 0;
}

extern "C" int srk_pixel_loss_bwd(const float* sr, const float* hr, int64_t numel, int mode,
                                  const float* gout, float* grad_sr, void* stream) {
  SRK_REQUIRE(numel > 0 && (mode == 0 || mode == 1), "srk_pixel_loss_bwd: bad arguments");
  pixel_loss_bwd_kernel<<<red_blocks(numel, 4), 256, 0, (cudaStream_t)stream>>>(sr, hr, numel, mode, gout, grad_sr);
  SRK_CUDA_LAUNCH_CHECK("pixel_loss_bwd");
  return 0;
}

extern "C" int64_t srk_nlpd_workspace_bytes(int n, int c, int h, int w, int levels) {
  if (levels < 1 || levels > 6) return -1;
  return nlpd_plan(n, c, h, w, levels).total_bytes;
}

extern "C" int srk_nlpd_fwd(const float* sr, const float* hr, int n, int c, int h, int w, int levels,
                            float alpha, const float* kernel25, int clamp01, void* workspace, float* loss,
                            void* stream) {
  SRK_REQUIRE(levels >= 1 && levels <= 6, "srk_nlpd_fwd: levels must be in [1,6]");
  SRK_REQUIRE(((uintptr_t)workspace & 15) == 0, "srk_nlpd_fwd: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  NlpdPlan p = nlpd_plan(n, c, h, w, levels);
  float* wsf = (float*)workspace;
  signed char* sbase = (signed char*)workspace + (p.floats * 4 + 15) / 16 * 16;
  double* acc = (double*)((char*)workspace + p.acc_off);
  const int NC = n * c;
  long long n0 = (long long)NC * h * w;
  NlpdWeights wt;
  for (int i = 0; i < 8; ++i) { wt.w[i] = 0.0; wt.nblk[i] = 0; }
  wt.nblk[0] = resident_cap(nlpd_diff_kernel, red_blocks(n0, 8));
  nlpd_diff_kernel<<<wt.nblk[0], 256, 0, st>>>(sr, hr, n0, clamp01, wsf + p.cur_off[0], acc);
  wt.w[0] = (double)alpha / (double)n0;
  for (int l = 0; l < levels; ++l) {
    int H = p.h[l], W = p.w[l], h2 = p.h[l + 1], w2 = p.w[l + 1];
    long long nd = (long long)NC * h2 * w2, nu = (long long)NC * H * W;
    if (nlpd_2x() && W == 2 * w2)
      nlpd_blur_down_pair_kernel<<<resident_cap(nlpd_blur_down_pair_kernel, red_blocks((long long)NC * h2 * ((w2 + 1) / 2), 1)), 256, 0, st>>>(wsf + p.cur_off[l], NC, H, W, h2, w2, kernel25, wsf + p.cur_off[l + 1]);
    else
      nlpd_blur_down_kernel<<<resident_cap(nlpd_blur_down_kernel, red_blocks(nd, 2)), 256, 0, st>>>(wsf + p.cur_off[l], NC, H, W, h2, w2, kernel25, wsf + p.cur_off[l + 1]);
    float sy = (float)((double)h2 / H), sx = (float)((double)w2 / W);
    double* acc_l = acc + (size_t)(1 + l) * kLossPartials;
    if (nlpd_2x() && H == 2 * h2 && W == 2 * w2) {
      wt.nblk[1 + l] = resident_cap(nlpd_lap_abs_2x_kernel, red_blocks(nd, 1));
      nlpd_lap_abs_2x_kernel<<<wt.nblk[1 + l], 256, 0, st>>>(wsf + p.cur_off[l], wsf + p.cur_off[l + 1], NC, H, W, h2, w2, sbase + p.sign_off[l], acc_l);
    } else {
      wt.nblk[1 + l] = resident_cap(nlpd_lap_abs_kernel, red_blocks(nu, 4));
      nlpd_lap_abs_kernel<<<wt.nblk[1 + l], 256, 0, st>>>(wsf + p.cur_off[l], wsf + p.cur_off[l + 1], NC, H, W, h2, w2, sy, sx, sbase + p.sign_off[l], acc_l);
    }
    wt.w[1 + l] = (1.0 - (double)alpha) / (double)nu;
  }
  nlpd_combine_kernel<<<1, 256, 0, st>>>(acc, wt, loss);
  SRK_CUDA_LAUNCH_CHECK("nlpd_fwd");
  return 0;
}

extern "C" int srk_nlpd_bwd(int n, int c, int h, int w, int levels, float alpha, const float* kernel25,
                            void* workspace, const float* gout, float* grad_sr, void* stream) {
  SRK_REQUIRE(levels >= 1 && levels <= 6, "srk_nlpd_bwd: levels must be in [1,6]");
  cudaStream_t st = (cudaStream_t)stream;
  NlpdPlan p = nlpd_plan(n, c, h, w, levels);
  float* wsf = (float*)workspace;
  signed char* sbase = (signed char*)workspace + (p.floats * 4 + 15) / 16 * 16;
  const int NC = n * c;
  // coarse -> fine.  G_{l+1} (gradient w.r.t. cur_{l+1}) lives in g buffer (l+1)&1 ... sizes shrink
  // with l, so buffer 0 (level-0 sized) and buffer 1 (level-1 sized) alternate safely: G_l for odd l
  // in buffer 1, even l >= 2 in buffer 0; g_down_l reuses the cur_{l+1} slot (no longer needed).
  const float* g_next = nullptr;  // G_{L} = 0 : nothing consumes cur_L except the top-level upsample
  for (int l = levels - 1; l >= 0; --l) {
    int H = p.h[l], W = p.w[l], h2 = p.h[l + 1], w2 = p.w[l + 1];
    long long nd = (long long)NC * h2 * w2, nu = (long long)NC * H * W;
    float c_l = (float)((1.0 - (double)alpha) / (double)nu);
    float sy = (float)((double)h2 / H), sx = (float)((double)w2 / W);
    float* g_down = wsf + p.cur_off[l + 1];
    if (nlpd_2x() && H == 2 * h2 && W == 2 * w2)
      nlpd_bwd_down_2x_kernel<<<resident_cap(nlpd_bwd_down_2x_kernel, red_blocks(nd, 1)), 256, 0, st>>>(sbase + p.sign_off[l], g_next, NC, H, W, h2, w2, c_l, g_down);
    else
      nlpd_bwd_down_kernel<<<resident_cap(nlpd_bwd_down_kernel, red_blocks(nd, 1)), 256, 0, st>>>(sbase + p.sign_off[l], g_next, NC, H, W, h2, w2, sy, sx, c_l, g_down);
    if (l == 0) {
      float c_mae = (float)((double)alpha / (double)nu);
      if (H == 2 * h2 && W == 2 * w2)
        nlpd_bwd_up_2x_kernel<<<resident_cap(nlpd_bwd_up_2x_kernel, red_blocks(nd, 1)), 256, 0, st>>>(sbase + p.sign_off[0], g_down, NC, H, W, h2, w2, kernel25, c_l, wsf + p.cur_off[0], c_mae, gout, grad_sr);
      else
        nlpd_bwd_up_kernel<<<resident_cap(nlpd_bwd_up_kernel, red_blocks(nu, 2)), 256, 0, st>>>(sbase + p.sign_off[0], g_down, NC, H, W, h2, w2, kernel25, c_l, wsf + p.cur_off[0], c_mae, gout, grad_sr);
    } else {
      float* G = wsf + p.g_off[l & 1];
      if (H == 2 * h2 && W == 2 * w2)
        nlpd_bwd_up_2x_kernel<<<resident_cap(nlpd_bwd_up_2x_kernel, red_blocks(nd, 1)), 256, 0, st>>>(sbase + p.sign_off[l], g_down, NC, H, W, h2, w2, kernel25, c_l, nullptr, 0.f, nullptr, G);
      else
        nlpd_bwd_up_kernel<<<resident_cap(nlpd_bwd_up_kernel, red_blocks(nu, 2)), 256, 0, st>>>(sbase + p.sign_off[l], g_down, NC, H, W, h2, w2, kernel25, c_l, nullptr, 0.f, nullptr, G);
      g_next = G;
    }
  }
  SRK_CUDA_LAUNCH_CHECK("nlpd_bwd");
  return 0;
}

extern "C" int srk_psnr_sse(const float* sr, const float* hr, int n, int64_t per_image, int clamp01,
                            double* sse, void* stream) {
  SRK_REQUIRE(n > 0 && per_image > 0, "srk_psnr_sse: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(sse, 0, sizeof(double) * n, st);
  int bx = red_blocks(per_image, 8);
  int cap = (148 * 8 + n - 1) / n; if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  psnr_sse_kernel<<<dim3(bx, n), 256, 0, st>>>(sr, hr, per_image, clamp01, sse);
  SRK_CUDA_LAUNCH_CHECK("psnr_sse");
  return 0;
}

extern "C" int srk_ssim(const float* sr, const float* hr, int n, int c, int h, int w, int clamp01,
                        double* ssim_sum, void* stream) {
  SRK_REQUIRE(n > 0 && c > 0, "srk_ssim: empty input");
  SRK_REQUIRE(h > 10 && w > 10, "srk_ssim: image smaller than the 11x11 window");
  SRK_REQUIRE((long long)n * c <= 65535, "srk_ssim: too many planes in one call");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ssim_sum, 0, sizeof(double) * n, st);
  int OH = h - 10, OW = w - 10;
  dim3 grid((OW + ST_W - 1) / ST_W, (OH + ST_H - 1) / ST_H, n * c);
  ssim_kernel<<<grid, dim3(ST_W, ST_H), 0, st>>>(sr, hr, c, h, w, clamp01, ssim_sum);
  SRK_CUDA_LAUNCH_CHECK("ssim");
  return 0;
}

extern "C" int srk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             int64_t numel, float lr, float beta1, float beta2, float eps,
                             const int64_t* step_count, float grad_scale, void* stream) {
  SRK_REQUIRE(numel > 0, "srk_adam_step: empty parameter buffer");
  adam_kernel<<<red_blocks(numel, 4), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, (const long long*)step_count, grad_scale);
  SRK_CUDA_LAUNCH_CHECK("adam");
  return 0;
}

extern "C" int srk_adam_multi(int count, float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2,
                              float eps, const int64_t* step_count, float grad_scale, const float* lr_dev,
                              const float* grad_scale_dev, void* stream) {
  SRK_REQUIRE(count >= 0, "srk_adam_multi: negative count");
  for (int base = 0; base < count; base += ADAM_MT) {
    AdamTable tb;
    const int c = count - base < ADAM_MT ? count - base : ADAM_MT;
    int64_t mx = 0;
    for (int i = 0; i < ADAM_MT; ++i) {
      if (i < c) {
        SRK_REQUIRE(numel[base + i] > 0 && numel[base + i] < (1LL << 31), "srk_adam_multi: bad tensor size");
        tb.p[i] = params[base + i]; tb.g[i] = grads[base + i]; tb.m[i] = exp_avg[base + i];
        tb.v[i] = exp_avg_sq[base + i]; tb.n[i] = (int)numel[base + i];
        if (numel[base + i] > mx) mx = numel[base + i];
      } else {
        tb.p[i] = nullptr; tb.g[i] = nullptr; tb.m[i] = nullptr; tb.v[i] = nullptr; tb.n[i] = 0;
      }
    }
    dim3 grid((unsigned)((mx + 1023) / 1024), (unsigned)c);
    adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tb, lr, beta1, beta2, eps, (const long long*)step_count,
                                                              grad_scale, lr_dev, grad_scale_dev);
    SRK_CUDA_LAUNCH_CHECK("adam_multi");
  }
  return 0;
}
