// dw3x3_bwd.cu -- depthwise 3x3 backward, smem-tiled (replaces the naive per-element versions).
//
// Backward of DepthwiseSeparableConv.depthwise (reference models/layers.py:38-44,48):
//   input : dx[b,c,i,j]  = sum_kk sum_{dy,dx} w[c*k+kk][dy][dx] * dd[b, c*k+kk, i-dy+1, j-dx+1]
//   weight: dW[o][dy][dx] += sum_{b,i,j} dd[b,o,i,j] * in[b,o/k,i+dy-1,j+dx-1],  db[o] += sum dd
// One CTA per (plane, TH x TW tile): the halo tile(s) are staged in shared memory once (coalesced loads,
// zero fill = padding), so every global element is read once instead of 9-18 times.  The weight kernel
// keeps 10*k partial sums per thread, block-reduces them and merges tiles with fp32 atomics.
#include "common.cuh"

namespace smaat {

constexpr int DB_TH = 32;     // tile rows
constexpr int DB_KMAX = 4;    // kernels_per_layer supported by the tiled kernels

static int pick_tw(int W) {
  if (W % 4 == 0)
    for (int c = 96; c >= 16; c -= 4)
      if (W % c == 0) return c;
  return W >= 64 ? 64 : ((W + 3) / 4) * 4;
}

__global__ void __launch_bounds__(256) dw3x3_bwd_input_tiled(const float* __restrict__ dd, const float* __restrict__ w,
                                                             float* __restrict__ dx0, int C0, int64_t bs0,
                                                             float* __restrict__ dx1, int C1, int64_t bs1, int H, int W, int k,
                                                             int TW, int tiles_x, int tiles_y) {
  extern __shared__ float sm[];  // [k][DB_TH+2][TW+2]
  const int Cin = C0 + C1;
  const int tiles = tiles_x * tiles_y;
  const int plane = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int b = plane / Cin, c = plane - b * Cin;
  const int y0 = (tile / tiles_x) * DB_TH, x0 = (tile % tiles_x) * TW;
  const int SW = TW + 2, SH = DB_TH + 2;
  const int P = H * W;
  const float* g = dd + ((int64_t)b * Cin + c) * k * P;
  for (int i = threadIdx.x; i < k * SH * SW; i += blockDim.x) {
    const int kk = i / (SH * SW), r = (i / SW) % SH, cc = i % SW;
    const int gy = y0 - 1 + r, gx = x0 - 1 + cc;
    sm[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(g + (int64_t)kk * P + (int64_t)gy * W + gx) : 0.f;
  }
  __shared__ float ws[DB_KMAX * 9];
  if (threadIdx.x < k * 9) ws[threadIdx.x] = __ldg(w + (int64_t)c * k * 9 + threadIdx.x);
  __syncthreads();
  float* dst = (c < C0) ? dx0 + (int64_t)b * bs0 + (int64_t)c * P : dx1 + (int64_t)b * bs1 + (int64_t)(c - C0) * P;
  for (int i = threadIdx.x; i < DB_TH * TW; i += blockDim.x) {
    const int y = i / TW, x = i - y * TW;
    const int gy = y0 + y, gx = x0 + x;
    if (gy >= H || gx >= W) continue;
    float acc = 0.f;
    for (int kk = 0; kk < k; ++kk) {
      const float* t = sm + (kk * SH + y + 1) * SW + x + 1;   // centre
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dxx = 0; dxx < 3; ++dxx) acc = fmaf(ws[kk * 9 + dy * 3 + dxx], t[(1 - dy) * SW + (1 - dxx)], acc);
    }
    dst[(int64_t)gy * W + gx] = acc;
  }
}

__global__ void __launch_bounds__(256) dw3x3_bwd_weight_tiled(const float* __restrict__ dd, const float* __restrict__ x0p, int C0,
                                                              int64_t bs0, const float* __restrict__ x1p, int C1, int64_t bs1,
                                                              const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                                                              float* __restrict__ dw, float* __restrict__ db, int H, int W, int k,
                                                              int TW, int tiles_x, int tiles_y) {
  extern __shared__ float sm[];  // [DB_TH+2][TW+2] input halo tile (activation applied, zero padding)
  const int Cin = C0 + C1;
  const int tiles = tiles_x * tiles_y;
  const int plane = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int b = plane / Cin, c = plane - b * Cin;
  const int y0 = (tile / tiles_x) * DB_TH, x0 = (tile % tiles_x) * TW;
  const int SW = TW + 2, SH = DB_TH + 2;
  const int P = H * W;
  const float* src = (c < C0) ? x0p + (int64_t)b * bs0 + (int64_t)c * P : x1p + (int64_t)b * bs1 + (int64_t)(c - C0) * P;
  const bool pro = in_scale != nullptr;
  const float s = pro ? __ldg(in_scale + c) : 1.f, t = pro ? __ldg(in_shift + c) : 0.f;
  for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) {
    const int r = i / SW, cc = i - r * SW;
    const int gy = y0 - 1 + r, gx = x0 - 1 + cc;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      v = __ldg(src + (int64_t)gy * W + gx);
      if (pro) v = fmaxf(fmaf(v, s, t), 0.f);
    }
    sm[i] = v;
  }
  __syncthreads();
  float acc[DB_KMAX][10];
#pragma unroll
  for (int kk = 0; kk < DB_KMAX; ++kk)
#pragma unroll
    for (int q = 0; q < 10; ++q) acc[kk][q] = 0.f;
  const float* g = dd + ((int64_t)b * Cin + c) * k * P;
  for (int i = threadIdx.x; i < DB_TH * TW; i += blockDim.x) {
    const int y = i / TW, x = i - y * TW;
    const int gy = y0 + y, gx = x0 + x;
    if (gy >= H || gx >= W) continue;
    const float* tl = sm + y * SW + x;  // top-left of the 3x3 window
    float win[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dxx = 0; dxx < 3; ++dxx) win[dy * 3 + dxx] = tl[dy * SW + dxx];
#pragma unroll
    for (int kk = 0; kk < DB_KMAX; ++kk) {
      if (kk < k) {
        const float gv = __ldg(g + (int64_t)kk * P + (int64_t)gy * W + gx);
#pragma unroll
        for (int q = 0; q < 9; ++q) acc[kk][q] = fmaf(gv, win[q], acc[kk][q]);
        acc[kk][9] += gv;
      }
    }
  }
  __shared__ float red[DB_KMAX * 10][8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int kk = 0; kk < DB_KMAX; ++kk) {
    if (kk < k) {
#pragma unroll
      for (int q = 0; q < 10; ++q) {
        const float v = warp_sum(acc[kk][q]);
        if (lane == 0) red[kk * 10 + q][wp] = v;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < k * 10) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[threadIdx.x][i];
    const int kk = threadIdx.x / 10, q = threadIdx.x % 10;
    const int o = c * k + kk;
    if (q < 9) atomicAdd(dw + (int64_t)o * 9 + q, v);
    else if (db) atomicAdd(db + o, v);
  }
}

int dw3x3_bwd_input_tiled_launch(const float* dd, const float* w, float* dx0, int C0, int64_t bs0, float* dx1, int C1, int64_t bs1,
                                 int B, int H, int W, int k, cudaStream_t st) {
  const int TW = pick_tw(W);
  const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, DB_TH);
  const int64_t grid = (int64_t)B * (C0 + C1) * tiles_x * tiles_y;
  SMAAT_REQUIRE(grid < (1ll << 31), "dw3x3_bwd_input: grid too large");
  const size_t smem = (size_t)k * (DB_TH + 2) * (TW + 2) * sizeof(float);
  dw3x3_bwd_input_tiled<<<(unsigned)grid, 256, smem, st>>>(dd, w, dx0, C0, bs0, dx1, C1, bs1, H, W, k, TW, tiles_x, tiles_y);
  SMAAT_LAUNCH_CHECK("smaat_dw3x3_bwd_input");
  return SMAAT_OK;
}

int dw3x3_bwd_weight_tiled_launch(const float* dd, const float* x0, int C0, int64_t bs0, const float* x1, int C1, int64_t bs1,
                                  const float* in_scale, const float* in_shift, float* dw, float* db, int B, int H, int W, int k,
                                  cudaStream_t st) {
  const int TW = pick_tw(W);
  const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, DB_TH);
  const int64_t grid = (int64_t)B * (C0 + C1) * tiles_x * tiles_y;
  SMAAT_REQUIRE(grid < (1ll << 31), "dw3x3_bwd_weight: grid too large");
  const size_t smem = (size_t)(DB_TH + 2) * (TW + 2) * sizeof(float);
  dw3x3_bwd_weight_tiled<<<(unsigned)grid, 256, smem, st>>>(dd, x0, C0, bs0, x1, C1, bs1, in_scale, in_shift, dw, db, H, W, k, TW,
                                                            tiles_x, tiles_y);
  SMAAT_LAUNCH_CHECK("smaat_dw3x3_bwd_weight");
  return SMAAT_OK;
}

}  // namespace smaat
