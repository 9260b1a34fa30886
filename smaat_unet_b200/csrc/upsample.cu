// upsample.cu -- nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) + F.pad to the skip size
// (reference models/unet_parts_depthwise_separable.py:64,78-81), forward.
//
// Write-bound (output 4x the input).  One thread produces a 2-row x 4-pixel output block: with a source
// step < 0.5 per output pixel those 8 outputs depend on at most 3 source rows x 4 source columns, which
// are loaded once (12 loads per 8 outputs instead of 32) from a 4x smaller plane that stays in L1/L2;
// two 128-bit stores.  grid = (row-pair x quad blocks of one plane, C, B): 32-bit index math only.
// Index math follows torch's area_pixel_compute_source_index for align_corners=True: src = dst*(in-1)/(out-1).
#include "common.cuh"

namespace smaat {

__device__ __forceinline__ float sel3(int i, float a, float b, float c, float d) {
  return i == 0 ? a : (i == 1 ? b : (i == 2 ? c : d));
}

template <bool VEC>
__global__ void __launch_bounds__(256) upsample2x_pad_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             int64_t y_bstride, int C, int H, int W, int Ho, int Wo,
                                                             int pad_t, int pad_l, float ry, float rx) {
  const int wq = (Wo + 3) >> 2, hp = (Ho + 1) >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= hp * wq) return;
  const int op = idx / wq;
  const int oy0 = op << 1;
  const int ox0 = (idx - op * wq) << 2;
  const int c = blockIdx.y, b = blockIdx.z;
  const float* src = x + ((int64_t)b * C + c) * H * W;
  float* dst = y + (int64_t)b * y_bstride + ((int64_t)c * Ho + oy0) * Wo + ox0;

  // source columns: xa .. xa+3 (clamped) cover every tap of the 4 output pixels
  const int uxa = max(ox0 - pad_l, 0);
  const int xa = min((int)(rx * (float)uxa), W - 1);
  int ci[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) ci[i] = min(xa + i, W - 1);
  int ix0[4];
  float lx[4];
  bool xin[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ux = ox0 + j - pad_l;
    xin[j] = (ux >= 0) && (ux < 2 * W);
    const float sx = rx * (float)max(ux, 0);
    const int x0 = min((int)sx, W - 1);
    lx[j] = sx - (float)x0;
    ix0[j] = min(max(x0 - xa, 0), 2);
  }
  // source rows: ya .. ya+2 (clamped) cover both output rows
  const int uya = max(oy0 - pad_t, 0);
  const int ya = min((int)(ry * (float)uya), H - 1);
  float v[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* rp = src + (int64_t)min(ya + r, H - 1) * W;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[r][i] = __ldg(rp + ci[i]);
  }
  // horizontal pass once per source row (same association as torch: w_x0*v0 + w_x1*v1), then the vertical blend
  float hrow[3][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = sel3(ix0[j], v[r][0], v[r][1], v[r][2], v[r][3]);
      const float a1 = sel3(ix0[j] + 1, v[r][0], v[r][1], v[r][2], v[r][3]);
      hrow[r][j] = (1.f - lx[j]) * a0 + lx[j] * a1;
    }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int oy = oy0 + rr;
    if (oy >= Ho) break;
    const int uy = oy - pad_t;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (uy >= 0 && uy < 2 * H) {
      const float sy = ry * (float)uy;
      const int y0 = min((int)sy, H - 1);
      const float ly = sy - (float)y0;
      const bool r0 = (y0 - ya) >= 1;  // y0 - ya is 0 or 1; the row below holds min(y0+1, H-1)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (xin[j]) o[j] = (1.f - ly) * (r0 ? hrow[1][j] : hrow[0][j]) + ly * (r0 ? hrow[2][j] : hrow[1][j]);
    }
    float* d = dst + (int64_t)rr * Wo;
    if (VEC) {
      *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ox0 + j < Wo) d[j] = o[j];
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
  dim3 grid(ceil_div(ceil_div(Ho, 2) * ceil_div(Wo, 4), 256), C, B);
  if (vec)
    upsample2x_pad_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, y_bstride, C, H, W, Ho, Wo, pad_t, pad_l, ry, rx);
  else
    upsample2x_pad_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, y_bstride, C, H, W, Ho, Wo, pad_t, pad_l, ry, rx);
  SMAAT_LAUNCH_CHECK("smaat_upsample2x_pad_fwd");
  return SMAAT_OK;
}
