// CUDA-core (fp32-accurate) implicit-GEMM convolution: fprop (also dgrad via rotated weights) and
// wgrad for every conv shape on the path.  This is the fp32 parity path and the fallback for the
// shapes the tcgen05 kernel does not take (9x9 / 5x5 / 1x1, Cin=3, Cout=3).
//
// GEMM view (reference: F.conv2d at models.py:46..167, stride 1, pad R/2):
//   fprop:  Y[p, co] = sum_k A[p, k] * Wp[k, co],  p = (n,y,x), k = (r,s,ci), A gathered on the fly
//   wgrad:  dW[k, co] = sum_p A[p, k] * dY[p, co]
#include "srk_common.cuh"

namespace srk {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename TI>
__device__ __forceinline__ float load_in(const View& v, int n, int y, int x, int c) {
  const TI* p = (const TI*)v.p;
  return to_f<TI>(p[v.off + n * v.sn + y * v.sh + x * v.sw + c * v.sc]);
}

struct FpropParams {
  View x, y, res;
  const float* w;  // [R][S][Cin][Cout]
  const float* bias;
  const float* alpha;
  void* zsave;      // PReLU with a slope <= 0: copy of the pre-activation in y's geometry / dtype, or null
  int Cin, Cout, R, S, pad;
  int act, shuffle, has_res;
  int K;            // R*S*Cin
  long long M;      // N*H*W
};

// KFAST: consecutive threads walk k (channels-last input), else consecutive threads walk pixels.
template <typename TI, typename TO, bool KFAST>
__global__ void __launch_bounds__(256) conv_fprop_simt(FpropParams P) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int H = P.x.H, W = P.x.W;

  // A-load assignment
  int a_m[4], a_k[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (KFAST) { a_k[j] = tid & 15; a_m[j] = (tid >> 4) + 16 * j; }
    else       { a_m[j] = tid & 63; a_k[j] = (tid >> 6) + 4 * j; }
  }
  int pn[4], py[4], px[4];
  bool pv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    long long m = m0 + a_m[j];
    pv[j] = m < P.M;
    long long mm = pv[j] ? m : 0;
    pn[j] = (int)(mm / ((long long)H * W));
    int rem = (int)(mm - (long long)pn[j] * H * W);
    py[j] = rem / W; px[j] = rem - py[j] * W;
  }
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < P.K; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + a_k[j];
      float v = 0.f;
      if (pv[j] && k < P.K) {
        int tap = k / P.Cin, ci = k - tap * P.Cin;
        int r = tap / P.S, s = tap - r * P.S;
        int yy = py[j] + r - P.pad, xx = px[j] + s - P.pad;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = load_in<TI>(P.x, pn[j], yy, xx, ci);
      }
      As[a_k[j]][a_m[j]] = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = (tid >> 6) + 4 * j, nn = tid & 63;
      int k = k0 + kk, co = n0 + nn;
      Bs[kk][nn] = (k < P.K && co < P.Cout) ? P.w[(long long)k * P.Cout + co] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float alpha = (P.act == SRK_ACT_PRELU) ? P.alpha[0] : 0.f;
  TO* yp = (TO*)P.y.p;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= P.M) continue;
    int n = (int)(m / ((long long)H * W));
    int rem = (int)(m - (long long)n * H * W);
    int y = rem / W, x = rem - y * W;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= P.Cout) continue;
      float v = acc[i][j] + (P.bias ? P.bias[co] : 0.f);
      int oc = co, oy = y, ox = x;
      if (P.shuffle == 2) { oc = co >> 2; oy = 2 * y + ((co >> 1) & 1); ox = 2 * x + (co & 1); }
      long long oidx = P.y.off + n * P.y.sn + oy * P.y.sh + ox * P.y.sw + oc * P.y.sc;
      if (P.act == SRK_ACT_PRELU && P.zsave != nullptr && !(alpha > 0.f)) ((TO*)P.zsave)[oidx] = from_f<TO>(v);
      if (P.act == SRK_ACT_RELU) v = fmaxf(v, 0.f);
      else if (P.act == SRK_ACT_PRELU) v = v > 0.f ? v : alpha * v;
      if (P.has_res) {
        // residual shares the output geometry
        long long ridx = P.res.off + n * P.res.sn + oy * P.res.sh + ox * P.res.sw + oc * P.res.sc;
        v += (P.res.dtype == SRK_BF16) ? to_f(((const __nv_bfloat16*)P.res.p)[ridx])
                                       : ((const float*)P.res.p)[ridx];
      }
      yp[oidx] = from_f<TO>(v);
    }
  }
}

// Zero the 1-pixel border ring of an ACT tensor (interior untouched).
template <typename T>
__global__ void zero_border_kernel(T* p, int N, int Hp, int Wp, int C) {
  int ring = 2 * Wp + 2 * (Hp - 2);
  long long total = (long long)N * ring * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int rp = (int)(t % ring);
    int n = (int)(t / ring);
    int y, x;
    if (rp < Wp) { y = 0; x = rp; }
    else if (rp < 2 * Wp) { y = Hp - 1; x = rp - Wp; }
    else { int q = rp - 2 * Wp; y = 1 + (q >> 1); x = (q & 1) ? Wp - 1 : 0; }
    p[(((long long)n * Hp + y) * Wp + x) * C + c] = from_f<T>(0.f);
  }
}

// the same with 16-byte stores: one thread per 16-byte chunk of a ring pixel (CB = chunks per pixel)
__global__ void zero_border_vec_kernel(uint4* p, int N, int Hp, int Wp, int CB) {
  const int ring = 2 * Wp + 2 * (Hp - 2);
  const long long total = (long long)N * ring * CB;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int per_img = ring * CB;
    const int n = (int)(i / per_img);
    const int t = (int)(i - (long long)n * per_img);
    const int rp = t / CB, cb = t - rp * CB;
    int y, x;
    if (rp < Wp) { y = 0; x = rp; }
    else if (rp < 2 * Wp) { y = Hp - 1; x = rp - Wp; }
    else { const int q = rp - 2 * Wp; y = 1 + (q >> 1); x = (q & 1) ? Wp - 1 : 0; }
    p[(((long long)n * Hp + y) * Wp + x) * CB + cb] = make_uint4(0, 0, 0, 0);
  }
}

int zero_border(const srk_tensor* t, cudaStream_t st) {
  if (t->layout != SRK_LAYOUT_ACT) return 0;
  int Hp = t->h + 2, Wp = t->w + 2;
  const int pixel_bytes = t->c * (t->dtype == SRK_BF16 ? 2 : 4);
  if (pixel_bytes % 16 == 0 && ((uintptr_t)t->data & 15) == 0) {
    const int CB = pixel_bytes / 16;
    const long long chunks = (long long)t->n * (2 * Wp + 2 * (Hp - 2)) * CB;
    int blocks = (int)((chunks + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    zero_border_vec_kernel<<<blocks, 256, 0, st>>>((uint4*)t->data, t->n, Hp, Wp, CB);
    SRK_CUDA_LAUNCH_CHECK("zero_border");
    return 0;
  }
  long long total = (long long)t->n * (2 * Wp + 2 * (Hp - 2)) * t->c;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (t->dtype == SRK_BF16)
    zero_border_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)t->data, t->n, Hp, Wp, t->c);
  else
    zero_border_kernel<float><<<blocks, 256, 0, st>>>((float*)t->data, t->n, Hp, Wp, t->c);
  SRK_CUDA_LAUNCH_CHECK("zero_border");
  return 0;
}

int conv_fprop_simt_launch(const srk_tensor* x, const srk_tensor* y, const float* w, int cout, int r,
                           int s, const float* bias, int act, const float* alpha,
                           const srk_tensor* residual, int shuffle, cudaStream_t st, void* zsave) {
  FpropParams P;
  P.zsave = zsave;
  P.x = make_view(x); P.y = make_view(y);
  P.has_res = residual != nullptr;
  P.res = residual ? make_view(residual) : P.y;
  P.w = w; P.bias = bias; P.alpha = alpha;
  P.Cin = x->c; P.Cout = cout; P.R = r; P.S = s; P.pad = r / 2;
  P.act = act; P.shuffle = shuffle;
  P.K = r * s * x->c;
  P.M = (long long)x->n * x->h * x->w;
  dim3 grid((unsigned)((P.M + BM - 1) / BM), (unsigned)((cout + BN - 1) / BN));
  bool in_bf = x->layout == SRK_LAYOUT_ACT && x->dtype == SRK_BF16;
  bool out_bf = y->layout == SRK_LAYOUT_ACT && y->dtype == SRK_BF16;
  bool kfast = x->layout == SRK_LAYOUT_ACT;
#define LAUNCH(TI, TO, KF) conv_fprop_simt<TI, TO, KF><<<grid, 256, 0, st>>>(P)
  if (in_bf && out_bf) LAUNCH(__nv_bfloat16, __nv_bfloat16, true);
  else if (in_bf && !out_bf) LAUNCH(__nv_bfloat16, float, true);
  else if (!in_bf && out_bf) { if (kfast) LAUNCH(float, __nv_bfloat16, true); else LAUNCH(float, __nv_bfloat16, false); }
  else { if (kfast) LAUNCH(float, float, true); else LAUNCH(float, float, false); }
#undef LAUNCH
  SRK_CUDA_LAUNCH_CHECK("conv_fprop_simt");
  if (y->layout == SRK_LAYOUT_ACT) return zero_border(y, st);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// wgrad: dW[k][co] = sum_p A[p][k] dY[p][co], db[co] = sum_p dY[p][co].  The pixel range is split over blockIdx.z;
// every slice stores its partial [K][Cout] (+ [Cout]) into the workspace and conv_wgrad_simt_fold adds the slices in
// slice order into the OIHW gradient (no float atomics: the result does not depend on block scheduling).
struct WgradParams {
  View x, dy;
  float* part;      // [gz][K * Cout + Cout]
  int has_bias;
  int Cin, Cout, R, S, pad, K;
  long long M;
  long long chunk;  // pixels per z-slice
};

template <typename TI, typename TG>
__global__ void __launch_bounds__(256) conv_wgrad_simt(WgradParams P) {
  __shared__ float As[BK][BM + 4];  // [pixel][k]
  __shared__ float Bs[BK][BN + 4];  // [pixel][co]
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int H = P.x.H, W = P.x.W;
  const long long p_begin = (long long)blockIdx.z * P.chunk;
  long long p_end = p_begin + P.chunk;
  if (p_end > P.M) p_end = P.M;

  const int kl = tid & 63;
  const int k = k0 + kl;
  const bool kvalid = k < P.K;
  int ci = 0, r = 0, s = 0;
  if (kvalid) { int tap = k / P.Cin; ci = k - tap * P.Cin; r = tap / P.S; s = tap - r * P.S; }
  const int col = tid & 63, co = n0 + col;
  const bool covalid = co < P.Cout;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  float* slice = P.part + (long long)blockIdx.z * ((long long)P.K * P.Cout + P.Cout);

  for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int pl = (tid >> 6) + 4 * j;
      long long p = p0 + pl;
      float a = 0.f, g = 0.f;
      if (p < p_end) {
        int n = (int)(p / ((long long)H * W));
        int rem = (int)(p - (long long)n * H * W);
        int y = rem / W, x = rem - y * W;
        if (kvalid) {
          int yy = y + r - P.pad, xx = x + s - P.pad;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) a = load_in<TI>(P.x, n, yy, xx, ci);
        }
        if (covalid) g = load_in<TG>(P.dy, n, y, x, co);
      }
      As[pl][kl] = a;
      Bs[pl][col] = g;
      if (blockIdx.x == 0) bsum += g;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < BK; ++pp) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[pp][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[pp][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int kk = k0 + ty * 4 + i;
    if (kk >= P.K) continue;
    int tap = kk / P.Cin, cii = kk - tap * P.Cin;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c2 = n0 + tx * 4 + j;
      if (c2 >= P.Cout) continue;
      (void)tap; (void)cii;
      slice[(long long)kk * P.Cout + c2] = acc[i][j];
    }
  }
  if (P.has_bias && blockIdx.x == 0) {
    // threads tid, tid+64, tid+128, tid+192 share a column: reduce through smem
    __shared__ float red[256];
    red[tid] = bsum;
    __syncthreads();
    if (tid < 64 && covalid) slice[(long long)P.K * P.Cout + co] = red[tid] + red[tid + 64] + red[tid + 128] + red[tid + 192];
  }
}

__global__ void conv_wgrad_simt_fold(const float* __restrict__ part, int gz, int K, int Cin, int Cout, int taps,
                                     float* __restrict__ dw, float* __restrict__ db, int accumulate) {
  const long long stride = (long long)K * Cout + Cout;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= stride || (i >= (long long)K * Cout && db == nullptr)) return;
  float v = 0.f;
  for (int z = 0; z < gz; ++z) v += part[z * stride + i];
  if (i < (long long)K * Cout) {
    const int kk = (int)(i / Cout), c2 = (int)(i - (long long)kk * Cout);
    const int tap = kk / Cin, cii = kk - tap * Cin;
    float* d = dw + ((long long)c2 * Cin + cii) * taps + tap;
    *d = accumulate ? *d + v : v;
  } else {
    float* d = db + (i - (long long)K * Cout);
    *d = accumulate ? *d + v : v;
  }
}

struct WgradGrid { int gx, gy, gz; long long chunk; };
static WgradGrid wgrad_simt_grid(const srk_tensor* x, const srk_tensor* dy, int r, int s) {
  WgradGrid g;
  const int K = r * s * x->c;
  const long long M = (long long)x->n * x->h * x->w;
  g.gx = (K + BM - 1) / BM; g.gy = (dy->c + BN - 1) / BN;
  long long want = (148LL * 4 + (long long)g.gx * g.gy - 1) / ((long long)g.gx * g.gy);
  long long max_splits = (M + 1023) / 1024;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  long long chunk = (M + want - 1) / want;
  chunk = (chunk + BK - 1) / BK * BK;
  g.gz = (int)((M + chunk - 1) / chunk);
  g.chunk = chunk;
  return g;
}
int64_t conv_wgrad_simt_workspace(const srk_tensor* x, const srk_tensor* dy, int r, int s) {
  const WgradGrid g = wgrad_simt_grid(x, dy, r, s);
  return (int64_t)g.gz * ((int64_t)r * s * x->c * dy->c + dy->c) * (int64_t)sizeof(float);
}

int conv_wgrad_simt_launch(const srk_tensor* x, const srk_tensor* dy, float* dw, float* db, int r,
                           int s, void* workspace, int accumulate, cudaStream_t st) {
  SRK_REQUIRE(workspace != nullptr, "conv_wgrad_simt: workspace required (srk_conv_wgrad_workspace_bytes)");
  WgradParams P;
  P.x = make_view(x); P.dy = make_view(dy);
  P.part = (float*)workspace; P.has_bias = db != nullptr;
  P.Cin = x->c; P.Cout = dy->c; P.R = r; P.S = s; P.pad = r / 2;
  P.K = r * s * x->c;
  P.M = (long long)x->n * x->h * x->w;
  const WgradGrid g = wgrad_simt_grid(x, dy, r, s);
  P.chunk = g.chunk;
  dim3 grid(g.gx, g.gy, g.gz);
  bool x_bf = x->layout == SRK_LAYOUT_ACT && x->dtype == SRK_BF16;
  bool g_bf = dy->layout == SRK_LAYOUT_ACT && dy->dtype == SRK_BF16;
  if (x_bf && g_bf) conv_wgrad_simt<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(P);
  else if (x_bf) conv_wgrad_simt<__nv_bfloat16, float><<<grid, 256, 0, st>>>(P);
  else if (g_bf) conv_wgrad_simt<float, __nv_bfloat16><<<grid, 256, 0, st>>>(P);
  else conv_wgrad_simt<float, float><<<grid, 256, 0, st>>>(P);
  SRK_CUDA_LAUNCH_CHECK("conv_wgrad_simt");
  const long long total = (long long)P.K * P.Cout + P.Cout;
  conv_wgrad_simt_fold<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P.part, g.gz, P.K, P.Cin, P.Cout, r * s, dw, db,
                                                                        accumulate);
  SRK_CUDA_LAUNCH_CHECK("conv_wgrad_simt_fold");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weight packing from OIHW fp32 master weights
__device__ __forceinline__ void pack_weights_body(const float* __restrict__ w, void* __restrict__ out, int Cout,
                                                  int Cin, int R, int S, int kind, int shuffle, long long bx,
                                                  long long gx) {
  long long total = (long long)Cout * Cin * R * S;
  const int NPn8 = S * 3 > 16 ? 32 : 16;
  if (kind == SRK_PACK_FPROP_TC_N8) total = (long long)R * NPn8 * Cin;
  const int SEGr = (S * 3 + 1) / 2 * 2;                 // K layout of the RGB-side tcgen05 kernels: k = r * SEG + s * 3 + c
  const int KPr = (R * SEGr + 1 + 63) / 64 * 64;
  if (kind == SRK_PACK_RGBOUT_DGRAD_TC) total = 64LL * KPr;
  if (kind == SRK_PACK_RGBIN_TC) total = (long long)Cout * KPr;
  for (long long i = bx * (long long)blockDim.x + threadIdx.x; i < total; i += gx * blockDim.x) {
    // i indexes the OUTPUT linearly
    if (kind == SRK_PACK_RGBIN_TC || kind == SRK_PACK_RGBOUT_DGRAD_TC) {  // bf16 [64][KP]
      int k = (int)(i % KPr), n = (int)(i / KPr);
      float v = 0.f;
      const int r = k / SEGr, tk = k - r * SEGr;
      if (r < R && tk < S * 3) {
        int s = tk / 3, c = tk - s * 3;
        if (kind == SRK_PACK_RGBIN_TC) {            // n = co (Cout == 64), c = ci (Cin == 3)
          if (n < Cout && c < Cin) v = w[(((long long)n * Cin + c) * R + r) * S + s];
        } else {                                      // n = ci (Cin == 64), c = co (Cout <= 3), rot180
          if (n < Cin && c < Cout) v = w[(((long long)c * Cin + n) * R + (R - 1 - r)) * S + (S - 1 - s)];
        }
      }
      ((__nv_bfloat16*)out)[i] = __float2bfloat16_rn(v);
    } else if (kind == SRK_PACK_FPROP_TC_N8) {  // bf16 [R][NP][Cin], row n = s * 3 + co
      int ci = (int)(i % Cin); long long t = i / Cin;
      int n = (int)(t % NPn8); int r = (int)(t / NPn8);
      int s = n / 3, co = n - s * 3;
      float v = (s < S && co < Cout) ? w[(((long long)co * Cin + ci) * R + r) * S + s] : 0.f;
      ((__nv_bfloat16*)out)[i] = __float2bfloat16_rn(v);
    } else if (kind == SRK_PACK_FPROP_SIMT) {  // [R][S][Cin][Cout]
      int co = (int)(i % Cout); long long t = i / Cout;
      int ci = (int)(t % Cin); t /= Cin;
      int s = (int)(t % S); int r = (int)(t / S);
      ((float*)out)[i] = w[(((long long)co * Cin + ci) * R + r) * S + s];
    } else if (kind == SRK_PACK_DGRAD_SIMT) {  // [R][S][Cout][Cin] rot180
      int ci = (int)(i % Cin); long long t = i / Cin;
      int co = (int)(t % Cout); t /= Cout;
      int s = (int)(t % S); int r = (int)(t / S);
      ((float*)out)[i] = w[(((long long)co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s)];
    } else if (kind == SRK_PACK_FPROP_TC) {  // bf16 [tap][Cout'][Cin]
      int ci = (int)(i % Cin); long long t = i / Cin;
      int cop = (int)(t % Cout); int tap = (int)(t / Cout);
      int co = cop;
      if (shuffle == 2) { int C4 = Cout / 4; int sub = cop / C4, c = cop - sub * C4; co = 4 * c + sub; }
      int r = tap / S, s = tap - r * S;
      ((__nv_bfloat16*)out)[i] = __float2bfloat16_rn(w[(((long long)co * Cin + ci) * R + r) * S + s]);
    } else {  // SRK_PACK_DGRAD_TC: bf16 [tap rot180][Cin][Cout']
      int cop = (int)(i % Cout); long long t = i / Cout;
      int ci = (int)(t % Cin); int tap = (int)(t / Cin);
      int co = cop;
      if (shuffle == 2) { int C4 = Cout / 4; int sub = cop / C4, c = cop - sub * C4; co = 4 * c + sub; }
      int r = tap / S, s = tap - r * S;
      ((__nv_bfloat16*)out)[i] =
          __float2bfloat16_rn(w[(((long long)co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s)]);
    }
  }
}

__global__ void pack_weights_kernel(const float* __restrict__ w, void* __restrict__ out, int Cout, int Cin, int R,
                                    int S, int kind, int shuffle) {
  pack_weights_body(w, out, Cout, Cin, R, S, kind, shuffle, blockIdx.x, gridDim.x);
}

// Many packs in one launch: the table travels by value (CUDA-graph capturable); blockIdx.y selects the entry.
constexpr int PACK_MT = 64;
struct PackTable {
  const float* w[PACK_MT];
  void* out[PACK_MT];
  short cout[PACK_MT], cin[PACK_MT];
  signed char r[PACK_MT], kind[PACK_MT], shuffle[PACK_MT];
};
__global__ void pack_weights_multi_kernel(const PackTable tb) {
  const int t = blockIdx.y;
  pack_weights_body(tb.w[t], tb.out[t], tb.cout[t], tb.cin[t], tb.r[t], tb.r[t], tb.kind[t], tb.shuffle[t], blockIdx.x,
                    gridDim.x);
}

}  // namespace srk

using namespace srk;

extern "C" int64_t srk_weight_pack_bytes(int cout, int cin, int r, int s, int kind) {
  int64_t n = (int64_t)cout * cin * r * s;
  if (kind == SRK_PACK_FPROP_TC_N8) return (int64_t)r * (s * 3 > 16 ? 32 : 16) * cin * 2;
  const int64_t kp_rgb = (r * ((s * 3 + 1) / 2 * 2) + 1 + 63) / 64 * 64;
  if (kind == SRK_PACK_RGBOUT_DGRAD_TC) return 64LL * kp_rgb * 2;
  if (kind == SRK_PACK_RGBIN_TC) return (int64_t)cout * kp_rgb * 2;
  return (kind == SRK_PACK_FPROP_SIMT || kind == SRK_PACK_DGRAD_SIMT) ? n * 4 : n * 2;
}

extern "C" int srk_weight_pack(const float* w_oihw, void* out, int cout, int cin, int r, int s,
                               int kind, int pixel_shuffle, void* stream) {
  SRK_REQUIRE(kind >= 0 && kind <= 6, "srk_weight_pack: bad kind %d", kind);
  SRK_REQUIRE(kind != SRK_PACK_RGBIN_TC || (cin == 3 && cout % 32 == 0 && cout >= 64),
              "srk_weight_pack: RGBIN pack needs a 3 -> 64 / 96 conv");
  SRK_REQUIRE(kind != SRK_PACK_RGBOUT_DGRAD_TC || (cin == 64 && cout <= 3), "srk_weight_pack: RGBOUT pack needs a 64 -> 3 conv");
  SRK_REQUIRE(kind != SRK_PACK_FPROP_TC_N8 || cout <= 3, "srk_weight_pack: the RGB-output pack needs Cout <= 3");
  SRK_REQUIRE(pixel_shuffle == 0 || (pixel_shuffle == 2 && cout % 4 == 0), "srk_weight_pack: bad pixel_shuffle");
  long long total = (long long)cout * cin * r * s;
  if (kind == SRK_PACK_FPROP_TC_N8) total = (long long)r * (s * 3 > 16 ? 32 : 16) * cin;
  const long long kp_rgb = (r * ((s * 3 + 1) / 2 * 2) + 1 + 63) / 64 * 64;
  if (kind == SRK_PACK_RGBOUT_DGRAD_TC) total = 64LL * kp_rgb;
  if (kind == SRK_PACK_RGBIN_TC) total = (long long)cout * kp_rgb;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, out, cout, cin, r, s, kind, pixel_shuffle);
  SRK_CUDA_LAUNCH_CHECK("pack_weights");
  return 0;
}

extern "C" int srk_weight_pack_multi(int count, const float* const* w_oihw, void* const* out, const int32_t* cout,
                                     const int32_t* cin, const int32_t* r, const int32_t* kind,
                                     const int32_t* pixel_shuffle, void* stream) {
  SRK_REQUIRE(count >= 0, "srk_weight_pack_multi: negative count");
  for (int base = 0; base < count; base += PACK_MT) {
    PackTable tb;
    const int c = count - base < PACK_MT ? count - base : PACK_MT;
    for (int i = 0; i < PACK_MT; ++i) {
      const int j = base + (i < c ? i : 0);
      SRK_REQUIRE(kind[j] >= 0 && kind[j] <= 6 && cout[j] > 0 && cout[j] < 32768 && cin[j] > 0 && cin[j] < 32768 &&
                      r[j] >= 1 && r[j] <= 11,
                  "srk_weight_pack_multi: bad entry %d", j);
      tb.w[i] = w_oihw[j]; tb.out[i] = out[j];
      tb.cout[i] = (short)cout[j]; tb.cin[i] = (short)cin[j]; tb.r[i] = (signed char)r[j]; tb.kind[i] = (signed char)kind[j];
      tb.shuffle[i] = (signed char)(pixel_shuffle ? pixel_shuffle[j] : 0);
    }
    pack_weights_multi_kernel<<<dim3(64, c), 256, 0, (cudaStream_t)stream>>>(tb);
    SRK_CUDA_LAUNCH_CHECK("pack_weights_multi");
  }
  return 0;
}
