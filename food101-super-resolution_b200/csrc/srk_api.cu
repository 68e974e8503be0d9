// C-ABI front door of libsrk: error channel, conv dispatch (CUDA-core fp32 path vs tcgen05 path),
// version / capability queries.  Declarations: include/srk.h.
#include "srk_common.cuh"
#include "srk_tc_common.cuh"

#include <cstring>

namespace srk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// srk_conv_simt.cu
int conv_fprop_simt_launch(const srk_tensor* x, const srk_tensor* y, const float* w, int cout, int r,
                           int s, const float* bias, int act, const float* alpha,
                           const srk_tensor* residual, int shuffle, cudaStream_t st, void* zsave);
int conv_wgrad_simt_launch(const srk_tensor* x, const srk_tensor* dy, float* dw, float* db, int r,
                           int s, void* workspace, int accumulate, cudaStream_t st);
int64_t conv_wgrad_simt_workspace(const srk_tensor* x, const srk_tensor* dy, int r, int s);
// srk_conv_tc.cu
bool conv_tc_shape_ok(int cin, int cout, int r, int s, int dtype, int shuffle);
int conv_fprop_tc_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout,
                         int r, int s, const float* bias, int act, const float* alpha,
                         const srk_tensor* residual, int shuffle, float* stats_sum, float* stats_sumsq,
                         void* workspace, cudaStream_t st, void* reduce_ws, void* zsave, void* acc = nullptr);
int64_t conv_fprop_tc_workspace(const srk_tensor* x);
bool conv_smalln_tc_ok(const srk_tensor* x, const srk_tensor* y, int cout, int r, int s);
int conv_smalln_tc_launch(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int cout, int r,
                          const float* bias, cudaStream_t st);
bool conv_wgrad_tc_shape_ok(const srk_tensor* x, const srk_tensor* dy, int r, int s);
int64_t conv_wgrad_tc_workspace(const srk_tensor* x, const srk_tensor* dy, int r, int s);
int conv_wgrad_tc_launch(const srk_tensor* x, const srk_tensor* dy, float* dw, float* db, int r, int s,
                         void* workspace, int accumulate, int perm_shuffle, cudaStream_t st);

int64_t conv_rgb_workspace_bytes(int k);
int conv_rgb_tc_run(const srk_tensor* t3, const srk_tensor* y, const void* w_packed, const float* bias, int act,
                    const float* alpha, const srk_tensor* t64, float* dw, float* db, float* db3, int rgb_out, int k,
                    void* workspace, cudaStream_t st, const srk_tensor* dz_ps, const float* ps_alpha, float* ps_dalpha,
                    void* zsave, const void* ps_zsave);

static bool tensor_ok(const srk_tensor* t) {
  if (t == nullptr || t->data == nullptr) return false;
  if (t->n <= 0 || t->c <= 0 || t->h <= 0 || t->w <= 0) return false;
  if (t->layout == SRK_LAYOUT_IMAGE) return t->dtype == SRK_F32;
  if (t->layout == SRK_LAYOUT_ACT) return t->dtype == SRK_F32 || t->dtype == SRK_BF16;
  return false;
}

}  // namespace srk

using namespace srk;

extern "C" const char* srk_last_error(void) { return g_err; }
extern "C" int srk_version(void) { return 100; }

extern "C" int srk_conv_tc_supported(int cin, int cout, int r, int s, int dtype, int pixel_shuffle) {
  /* 2 = the RGB-output variant (bf16 ACT in, IMAGE out, SRK_PACK_FPROP_TC_N8 weights) */
  if (dtype == SRK_BF16 && cin == 64 && cout <= 3 && r == s && (r == 5 || r == 9) && pixel_shuffle == 0) return 2;
  return conv_tc_shape_ok(cin, cout, r, s, dtype, pixel_shuffle) ? 1 : 0;
}

extern "C" int64_t srk_conv_fprop_workspace_bytes(const srk_tensor* x, int pack_kind) {
  if (x == nullptr) return -1;
  const bool tc_kind = pack_kind == SRK_PACK_FPROP_TC || pack_kind == SRK_PACK_DGRAD_TC;
  return (tc_kind && x->layout == SRK_LAYOUT_ACT) ? conv_fprop_tc_workspace(x) : 0;
}

extern "C" int srk_conv_fprop(const srk_tensor* x, const srk_tensor* y, const void* w_packed,
                              int pack_kind, int cout, int r, int s, const float* bias, int act,
                              const float* alpha, const srk_tensor* residual, int pixel_shuffle,
                              int impl, float* bn_sums, void* reduce_ws, void* bn_acc, const srk_tensor* prelu_z,
                              void* workspace, void* stream) {
  if (bn_acc != nullptr) {   // 2 = the accumulator path does not cover this conv: nothing launched
    SRK_REQUIRE(bn_sums == nullptr, "srk_conv_fprop: bn_sums and bn_acc exclude each other");
    if (!(tensor_ok(x) && tensor_ok(y) && w_packed && pack_kind == SRK_PACK_FPROP_TC && (impl == SRK_IMPL_AUTO || impl == SRK_IMPL_TC) &&
          x->layout == SRK_LAYOUT_ACT && y->layout == SRK_LAYOUT_ACT && x->dtype == SRK_BF16 && y->dtype == SRK_BF16 &&
          r == 3 && s == 3 && x->c == 64 && cout == 64 && same_geometry(x, y) && act == SRK_ACT_NONE && residual == nullptr &&
          pixel_shuffle == 0 && prelu_z == nullptr))
      return 2;
    return conv_fprop_tc_launch(x, y, w_packed, cout, r, s, bias, act, alpha, nullptr, 0, nullptr, nullptr, workspace,
                                (cudaStream_t)stream, nullptr, nullptr, bn_acc);
  }
  float* bn_sum = bn_sums;
  float* bn_sumsq = bn_sums ? bn_sums + cout : nullptr;
  SRK_REQUIRE(tensor_ok(x) && tensor_ok(y), "srk_conv_fprop: bad x / y tensor");
  SRK_REQUIRE(bn_sum == nullptr || (act == SRK_ACT_NONE && residual == nullptr && pixel_shuffle == 0 &&
                                    y->layout == SRK_LAYOUT_ACT),
              "srk_conv_fprop: BN statistics are taken of a plain conv output in the ACT layout");
  SRK_REQUIRE(bn_sum == nullptr || reduce_ws != nullptr, "srk_conv_fprop: bn_sums needs reduce_ws");
  SRK_REQUIRE(w_packed != nullptr, "srk_conv_fprop: null weights");
  SRK_REQUIRE(r == s && (r & 1) == 1 && r >= 1 && r <= 11, "srk_conv_fprop: odd square kernels only (got %dx%d)", r, s);
  SRK_REQUIRE(act == SRK_ACT_NONE || act == SRK_ACT_RELU || act == SRK_ACT_PRELU, "srk_conv_fprop: bad act %d", act);
  SRK_REQUIRE(act != SRK_ACT_PRELU || alpha != nullptr, "srk_conv_fprop: PReLU needs alpha");
  SRK_REQUIRE(pixel_shuffle == 0 || pixel_shuffle == 2, "srk_conv_fprop: pixel_shuffle must be 0 or 2");
  if (pixel_shuffle == 2) {
    SRK_REQUIRE(cout % 4 == 0 && y->c == cout / 4 && y->h == 2 * x->h && y->w == 2 * x->w && y->n == x->n,
                "srk_conv_fprop: pixel-shuffle output geometry mismatch");
  } else {
    SRK_REQUIRE(y->c == cout && y->h == x->h && y->w == x->w && y->n == x->n,
                "srk_conv_fprop: output geometry mismatch");
  }
  if (residual) {
    SRK_REQUIRE(tensor_ok(residual) && same_geometry(residual, y) && residual->layout == y->layout,
                "srk_conv_fprop: residual must match the output geometry and layout");
  }
  void* zsave = nullptr;
  if (prelu_z) {
    SRK_REQUIRE(act == SRK_ACT_PRELU && tensor_ok(prelu_z) && same_geometry(prelu_z, y) && prelu_z->layout == y->layout &&
                    prelu_z->dtype == y->dtype && y->layout == SRK_LAYOUT_ACT,
                "srk_conv_fprop: prelu_z must match an ACT output of a PReLU conv");
    zsave = prelu_z->data;
  }
  if (pack_kind == SRK_PACK_FPROP_TC_N8) {
    SRK_REQUIRE(conv_smalln_tc_ok(x, y, cout, r, s) && act == SRK_ACT_NONE && residual == nullptr && pixel_shuffle == 0 &&
                    bn_sum == nullptr,
                "srk_conv_fprop: the RGB-output tcgen05 path takes bf16 ACT input with 64 channels, an IMAGE "
                "output with <= 4 channels and no activation / residual");
    return conv_smalln_tc_launch(x, y, w_packed, cout, r, bias, (cudaStream_t)stream);
  }
  const bool tc_kind = pack_kind == SRK_PACK_FPROP_TC || pack_kind == SRK_PACK_DGRAD_TC;
  const bool simt_kind = pack_kind == SRK_PACK_FPROP_SIMT || pack_kind == SRK_PACK_DGRAD_SIMT;
  SRK_REQUIRE(tc_kind || simt_kind, "srk_conv_fprop: bad pack kind %d", pack_kind);
  if (impl == SRK_IMPL_AUTO) impl = tc_kind ? SRK_IMPL_TC : SRK_IMPL_SIMT;
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == SRK_IMPL_TC) {
    SRK_REQUIRE(tc_kind, "srk_conv_fprop: tcgen05 path needs SRK_PACK_*_TC weights");
    SRK_REQUIRE(x->layout == SRK_LAYOUT_ACT && y->layout == SRK_LAYOUT_ACT && x->dtype == SRK_BF16 &&
                    y->dtype == SRK_BF16,
                "srk_conv_fprop: tcgen05 path needs bf16 ACT tensors");
    SRK_REQUIRE(conv_tc_shape_ok(x->c, cout, r, s, SRK_BF16, pixel_shuffle),
                "srk_conv_fprop: shape Cin=%d Cout=%d %dx%d not supported by the tcgen05 path", x->c, cout, r, s);
    if (bn_sum != nullptr && (x->c != 64 || cout != 64)) {  // the fused statistics cover the single-pass 64 -> 64 conv only
      if (conv_fprop_tc_launch(x, y, w_packed, cout, r, s, bias, act, alpha, residual, pixel_shuffle, nullptr,
                               nullptr, workspace, st, nullptr, zsave))
        return 1;
      return srk_bn_stats(y, bn_sums, reduce_ws, stream);
    }
    return conv_fprop_tc_launch(x, y, w_packed, cout, r, s, bias, act, alpha, residual, pixel_shuffle, bn_sum,
                                bn_sumsq, workspace, st, reduce_ws, zsave);
  }
  SRK_REQUIRE(impl == SRK_IMPL_SIMT && simt_kind, "srk_conv_fprop: CUDA-core path needs SRK_PACK_*_SIMT weights");
  if (conv_fprop_simt_launch(x, y, (const float*)w_packed, cout, r, s, bias, act, alpha, residual,
                             pixel_shuffle, st, zsave))
    return 1;
  return bn_sum ? srk_bn_stats(y, bn_sums, reduce_ws, stream) : 0;
}

// Data gradient of a 3x3 64 -> 64 conv with the BatchNorm-backward reduction of the layer BELOW fused into the
// epilogue: dx = dgrad(dz), and over interior pixels  sum_g[c] += sum g,  sum_gz[c] += sum g * z,
// dalpha += sum_{b<0} g * b, with z the saved pre-BN activation, b = BN(z) and g = dx masked by the PReLU that sits
// between that BN and this conv (alpha != NULL) or dx itself; with a residual, dx = dgrad(dz) + residual.  Returns 2 (and launches nothing) when the shape is
// outside the fused kernel: the caller then runs srk_conv_fprop + srk_bn_bwd_reduce.
extern "C" int srk_conv_dgrad_bnred(const srk_tensor* dz, const srk_tensor* dx, const void* w_packed_dgrad,
                                    const srk_tensor* z, const float* mean, const float* invstd, const float* gamma,
                                    const float* beta, const float* alpha, float* sum_g, float* sum_gz, float* dalpha,
                                    const srk_tensor* residual, void* reduce_ws, void* acc, void* stream) {
  SRK_REQUIRE(tensor_ok(dz) && tensor_ok(dx) && tensor_ok(z) && w_packed_dgrad, "srk_conv_dgrad_bnred: bad tensors");
  SRK_REQUIRE(reduce_ws != nullptr || acc != nullptr, "srk_conv_dgrad_bnred: reduce_ws (or acc) is required");
  SRK_REQUIRE(mean && invstd && gamma && beta && ((sum_g && sum_gz) || acc), "srk_conv_dgrad_bnred: null statistics");
  SRK_REQUIRE(alpha == nullptr || dalpha != nullptr || acc != nullptr, "srk_conv_dgrad_bnred: dalpha is required with alpha");
  if (acc != nullptr && !(tc_fold() >= 1 && tc_fold() <= 3)) return 2;
  if (!(dz->layout == SRK_LAYOUT_ACT && dx->layout == SRK_LAYOUT_ACT && dz->dtype == SRK_BF16 && dx->dtype == SRK_BF16 &&
        dz->c == 64 && dx->c == 64 && same_geometry(dz, dx)))
    return 2;
  if (residual && !(tensor_ok(residual) && same_geometry(residual, dx) && residual->layout == SRK_LAYOUT_ACT &&
                    residual->dtype == SRK_BF16))
    return 2;
  BnRedArgs br = {z, mean, invstd, gamma, beta, alpha, sum_g, sum_gz, dalpha};
  if (tc_fold() == 4 && residual == nullptr && acc == nullptr) {
    const int rs = conv_fprop_strip_launch(dz, dx, w_packed_dgrad, 64, nullptr, SRK_ACT_NONE, nullptr, nullptr, 0, nullptr,
                                           nullptr, (cudaStream_t)stream, &br, reduce_ws, nullptr);
    if (rs >= 0) return rs;
  }
  const int rc = conv_fprop_fold_launch(dz, dx, w_packed_dgrad, 64, nullptr, SRK_ACT_NONE, nullptr, residual, 0, nullptr,
                                        nullptr, nullptr, 0, (cudaStream_t)stream, &br, reduce_ws, nullptr, acc);
  return rc < 0 ? 2 : rc;
}

extern "C" int64_t srk_conv_wgrad_workspace_bytes(const srk_tensor* x, const srk_tensor* dy, int r, int s,
                                                  int impl) {
  if (x == nullptr || dy == nullptr) return -1;
  if (impl == SRK_IMPL_TC || (impl == SRK_IMPL_AUTO && conv_wgrad_tc_shape_ok(x, dy, r, s)))
    return conv_wgrad_tc_workspace(x, dy, r, s);
  return conv_wgrad_simt_workspace(x, dy, r, s);
}

extern "C" int srk_conv_wgrad(const srk_tensor* x, const srk_tensor* dy, float* dw, float* db, int r, int s,
                              int impl, int accumulate, int perm_shuffle, void* workspace, void* stream) {
  SRK_REQUIRE(tensor_ok(x) && tensor_ok(dy), "srk_conv_wgrad: bad x / dy tensor");
  SRK_REQUIRE(dw != nullptr, "srk_conv_wgrad: null dw");
  SRK_REQUIRE(r == s && (r & 1) == 1 && r >= 1 && r <= 11, "srk_conv_wgrad: odd square kernels only");
  SRK_REQUIRE(x->n == dy->n && x->h == dy->h && x->w == dy->w, "srk_conv_wgrad: x / dy geometry mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == SRK_IMPL_AUTO) impl = conv_wgrad_tc_shape_ok(x, dy, r, s) ? SRK_IMPL_TC : SRK_IMPL_SIMT;
  if (impl == SRK_IMPL_TC) {
    SRK_REQUIRE(conv_wgrad_tc_shape_ok(x, dy, r, s), "srk_conv_wgrad: shape not supported by the tcgen05 path");
    SRK_REQUIRE(workspace != nullptr || conv_wgrad_tc_workspace(x, dy, r, s) == 0, "srk_conv_wgrad: workspace required");
    return conv_wgrad_tc_launch(x, dy, dw, db, r, s, workspace, accumulate, perm_shuffle, st);
  }
  SRK_REQUIRE(!perm_shuffle, "srk_conv_wgrad: sub-pixel-major dY is a tcgen05-path layout");
  return conv_wgrad_simt_launch(x, dy, dw, db, r, s, workspace, accumulate, st);
}

extern "C" int64_t srk_conv_rgb_workspace_bytes(int k) { return (k == 9 || k == 5) ? conv_rgb_workspace_bytes(k) : -1; }

extern "C" int srk_conv_rgb_fprop(const srk_tensor* img3, const srk_tensor* y, const void* w_packed, int k,
                                  const float* bias, int act, const float* alpha, const srk_tensor* prelu_z,
                                  void* stream) {
  SRK_REQUIRE(tensor_ok(img3) && tensor_ok(y) && w_packed != nullptr, "srk_conv_rgb_fprop: bad arguments");
  SRK_REQUIRE(act != SRK_ACT_PRELU || alpha != nullptr, "srk_conv_rgb_fprop: PReLU needs alpha");
  SRK_REQUIRE(prelu_z == nullptr || (act == SRK_ACT_PRELU && tensor_ok(prelu_z) && same_geometry(prelu_z, y) &&
                                     prelu_z->layout == SRK_LAYOUT_ACT && prelu_z->dtype == SRK_BF16),
              "srk_conv_rgb_fprop: prelu_z must match the bf16 ACT output of a PReLU conv");
  return conv_rgb_tc_run(img3, y, w_packed, bias, act, alpha, nullptr, nullptr, nullptr, nullptr, 0, k, nullptr,
                         (cudaStream_t)stream, nullptr, nullptr, nullptr, prelu_z ? prelu_z->data : nullptr, nullptr);
}

extern "C" int srk_conv_rgb_bwd(const srk_tensor* img3, const srk_tensor* t64, const void* w_packed,
                                const srk_tensor* dx, float* dw, float* db, int k, int rgb_out, void* workspace,
                                void* stream) {
  SRK_REQUIRE(tensor_ok(img3) && tensor_ok(t64) && dw != nullptr && workspace != nullptr, "srk_conv_rgb_bwd: bad arguments");
  SRK_REQUIRE(dx == nullptr || (tensor_ok(dx) && w_packed != nullptr && rgb_out == 1),
              "srk_conv_rgb_bwd: dx needs rgb_out = 1 and SRK_PACK_RGBOUT_DGRAD_TC weights");
  return conv_rgb_tc_run(img3, dx, w_packed, nullptr, SRK_ACT_NONE, nullptr, t64, dw, rgb_out ? nullptr : db,
                         rgb_out ? db : nullptr, rgb_out, k, workspace, (cudaStream_t)stream, nullptr, nullptr, nullptr,
                         nullptr, nullptr);
}

// Backward of the 64 -> 3 output conv fused with the PReLU + PixelShuffle(2) backward of the upsample stage below it
// (models.py:120-125): dY = gradient of the RGB image, t64 = the upsample stage's output (= the conv input), w_packed =
// SRK_PACK_RGBOUT_DGRAD_TC weights.  Produces dw [3][64][K][K], db [3] (accumulated), dalpha (accumulated) and
// dz_ps = gradient of the 64 -> 256 conv output, bf16 ACT [N, 256, H/2, W/2] with channels SUB-PIXEL-MAJOR
// (sub * 64 + c instead of 4c + sub): feed it to srk_conv_fprop with SRK_PACK_DGRAD_TC weights packed with
// pixel_shuffle = 2 and to srk_conv_wgrad with perm_shuffle = 1.
extern "C" int srk_conv_rgbout_bwd_unshuffle(const srk_tensor* dy_img, const srk_tensor* t64, const srk_tensor* t64_z,
                                             const void* w_packed, const srk_tensor* dz_ps, float* dw, float* db,
                                             const float* alpha, float* dalpha, int k, void* workspace, void* stream) {
  SRK_REQUIRE(tensor_ok(dy_img) && tensor_ok(t64) && tensor_ok(dz_ps) && dw && workspace && w_packed && alpha,
              "srk_conv_rgbout_bwd_unshuffle: bad arguments");
  SRK_REQUIRE(t64_z == nullptr || (tensor_ok(t64_z) && same_geometry(t64_z, t64) && t64_z->layout == SRK_LAYOUT_ACT &&
                                   t64_z->dtype == SRK_BF16),
              "srk_conv_rgbout_bwd_unshuffle: t64_z must match t64 (bf16 ACT)");
  return conv_rgb_tc_run(dy_img, nullptr, w_packed, nullptr, SRK_ACT_NONE, nullptr, t64, dw, nullptr, db, 1, k, workspace,
                         (cudaStream_t)stream, dz_ps, alpha, dalpha, nullptr, t64_z ? t64_z->data : nullptr);
}
