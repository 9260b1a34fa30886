// upsample.cu -- nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) + F.pad to the skip size
// (reference models/unet_parts_depthwise_separable.py:64,78-81), forward.
//
// Write-bound (output 4x the input).  One CTA produces a 128 x 16 output tile of one (b, c) plane: the
// <= 66 x 10 source pixels it depends on are staged once in shared memory with coalesced loads, then
// every thread computes 2 rows x 4 consecutive pixels from smem taps and issues 128-bit stores.  Index
// math follows torch's area_pixel_compute_source_index for align_corners=True: src = dst*(in-1)/(out-1).
#include "common.cuh"

namespace smaat {

constexpr int UP_TW = 128, UP_TH = 16;      // output tile
constexpr int UP_SW = UP_TW / 2 + 4, UP_SH = UP_TH / 2 + 4;  // source tile bound (scale < 0.5, +2 taps, +slack)

template <bool VEC>
__global__ void __launch_bounds__(256) upsample2x_pad_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             int64_t y_bstride, int C, int H, int W, int Ho, int Wo,
                                                             int pad_t, int pad_l, float ry, float rx, int tiles_x) {
  __shared__ float src[UP_SH][UP_SW + 1];
  const int tx0 = (blockIdx.x % tiles_x) * UP_TW;
  const int ty0 = (blockIdx.x / tiles_x) * UP_TH;
  const int c = blockIdx.y, b = blockIdx.z;
  const float* xp = x + ((int64_t)b * C + c) * H * W;
  // source window of this output tile (clamped to the image)
  const int ux_lo = max(tx0 - pad_l, 0), uy_lo = max(ty0 - pad_t, 0);
  const int sx0 = min((int)(rx * ux_lo), W - 1), sy0 = min((int)(ry * uy_lo), H - 1);
  for (int i = threadIdx.x; i < UP_SH * UP_SW; i += 256) {
    const int r = i / UP_SW, cc = i - r * UP_SW;
    const int gy = min(sy0 + r, H - 1), gx = min(sx0 + cc, W - 1);  // clamp == the x1/y1 = min(.+1, in-1) rule
    src[r][cc] = __ldg(xp + (int64_t)gy * W + gx);
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, tyq = threadIdx.x >> 5;
  const int ox0 = tx0 + 4 * tx;
  if (ox0 >= Wo) return;
  float* yp = y + (int64_t)b * y_bstride + (int64_t)c * Ho * Wo;
  // per-column taps are shared by the two rows this thread produces
  int cx0[4], cx1[4];
  float lxv[4];
  bool xin[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ux = ox0 + j - pad_l;
    xin[j] = (ux >= 0) && (ux < 2 * W) && (ox0 + j < Wo);
    const float sx = rx * (float)max(ux, 0);
    const int x0 = min((int)sx, W - 1);
    lxv[j] = sx - (float)x0;
    cx0[j] = min(x0 - sx0, UP_SW - 1);
    cx1[j] = min(min(x0 + 1, W - 1) - sx0, UP_SW - 1);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int oy = ty0 + tyq + half * 8;
    if (oy >= Ho) continue;
    const int uy = oy - pad_t;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (uy >= 0 && uy < 2 * H) {
      const float sy = ry * (float)uy;
      const int y0 = min((int)sy, H - 1);
      const float ly = sy - (float)y0;
      const int r0 = min(y0 - sy0, UP_SH - 1), r1 = min(min(y0 + 1, H - 1) - sy0, UP_SH - 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (xin[j]) {
          const float lx = lxv[j];
          // same association as torch's upsample_bilinear2d: w_y0*(w_x0*v00 + w_x1*v01) + w_y1*(...)
          o[j] = (1.f - ly) * ((1.f - lx) * src[r0][cx0[j]] + lx * src[r0][cx1[j]]) +
                 ly * ((1.f - lx) * src[r1][cx0[j]] + lx * src[r1][cx1[j]]);
        }
      }
    }
    float* dst = yp + (int64_t)oy * Wo + ox0;
    if (VEC) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ox0 + j < Wo) dst[j] = o[j];
    }
  }
}

}  // namespace smaat

using namespace smaat;

extern "C" int smaat_upsample2x_pad_fwd(const float* x, float* y, int64_t y_bstride, int B, int C, int H, int W, int Ho, int Wo,
                                        void* stream) {
  SMAAT_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, "upsample2x: bad arguments");
  SMAAT_REQUIRE(Ho >= 2 * H && Wo >= 2 * W, "upsample2x: target %dx%d smaller than 2x source %dx%d (negative pad = crop unsupported)",
                Ho, Wo, H, W);
  SMAAT_REQUIRE(y_bstride >= (int64_t)C * Ho * Wo, "upsample2x: y batch stride too small");
  SMAAT_REQUIRE(C <= 65535 && B <= 65535, "upsample2x: C/B too large for grid.y/z");
  const int pad_t = (Ho - 2 * H) / 2, pad_l = (Wo - 2 * W) / 2;
  const float ry = (2 * H > 1) ? (float)(H - 1) / (float)(2 * H - 1) : 0.f;
  const float rx = (2 * W > 1) ? (float)(W - 1) / (float)(2 * W - 1) : 0.f;
  const bool vec = (Wo % 4 == 0) && aligned16(y) && (y_bstride % 4 == 0);
  const int tiles_x = ceil_div(Wo, UP_TW), tiles_y = ceil_div(Ho, UP_TH);
  dim3 grid(tiles_x * tiles_y, C, B);
  if (vec)
    upsample2x_pad_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, y_bstride, C, H, W, Ho, Wo, pad_t, pad_l, ry, rx, tiles_x);
  else
    upsample2x_pad_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, y_bstride, C, H, W, Ho, Wo, pad_t, pad_l, ry, rx, tiles_x);
  SMAAT_LAUNCH_CHECK("smaat_upsample2x_pad_fwd");
  return SMAAT_OK;
}
