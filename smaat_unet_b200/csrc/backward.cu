// backward.cu -- gradient kernels of the SmaAt-UNet hot path (training step, BASELINE configs[2]).
//
// First complete set: correctness-first streaming / reduction kernels (fp32 data, fp64 or atomic fp32
// accumulation where sums span the batch).  The two GEMM-shaped gradients reuse the tensor-core path:
// the pointwise input gradient IS a pointwise forward with the transposed weight (smaat_pw1x1_fwd),
// only the pointwise weight gradient (a reduction over B*H*W pixels) runs on the CUDA cores here.
//
// Notation follows functional.py: z = pre-BatchNorm activation, a = act(scale*z + shift),
// dA = dL/da masked by the activation, dz = dL/dz.
#include "common.cuh"

namespace smaat {

// ---------------------------------------------------------------------------------------------
// BatchNorm (+ReLU) backward, 3 steps: per-channel sums -> per-channel coefficients -> apply
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                double* __restrict__ sums, int B, int C, int P, int act, int chunks) {
  const int c = blockIdx.y;
  const int64_t n = (int64_t)B * P;
  const int64_t per = (n + chunks - 1) / chunks;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
  const float s = scale ? __ldg(scale + c) : 1.f, t = shift ? __ldg(shift + c) : 0.f;
  double s1 = 0.0, s2 = 0.0;
  float f1 = 0.f, f2 = 0.f;
  int cnt = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int64_t b = i / P, pp = i - b * P;
    const int64_t idx = (b * C + c) * (int64_t)P + pp;
    const float zv = __ldg(z + idx);
    float g = __ldg(dy + idx);
    if (act == 1 && !(fmaf(zv, s, t) > 0.f)) g = 0.f;
    f1 += g;
    f2 = fmaf(g, zv, f2);
    if (++cnt == 64) { s1 += f1; s2 += f2; f1 = f2 = 0.f; cnt = 0; }
  }
  s1 += f1; s2 += f2;
  __shared__ double r1[256], r2[256];
  r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { atomicAdd(sums + c, r1[0]); atomicAdd(sums + C + c, r2[0]); }
}

// dz = a[c]*dA + b[c]*z + cc[c];  train: a = g*istd, b = -a*c2*istd, cc = -a*c1 + a*c2*mean*istd with
// c1 = S1/n, c2 = dgamma/n; eval: a = g*istd, b = cc = 0.  dgamma = istd*(S2 - mean*S1), dbeta = S1.
__global__ void bn_bwd_coeffs_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                     const float* __restrict__ mean, const float* __restrict__ invstd, int train,
                                     float* __restrict__ a, float* __restrict__ b, float* __restrict__ cc,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dz_sum, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double S1 = sums[c], S2 = sums[C + c];
  const double m = mean[c], is = invstd[c], g = gamma ? gamma[c] : 1.0;
  const double dg = is * (S2 - m * S1);
  if (dgamma) dgamma[c] += (float)dg;
  if (dbeta) dbeta[c] += (float)S1;
  const double sc = g * is;
  // sum_{b,p} dz (= the bias gradient of the conv that produced z): a*S1 + b*n*mean + n*cc, which is identically 0 with
  // batch statistics and a*S1 with running statistics -- no extra pass over dz
  if (dz_sum && !train) dz_sum[c] += (float)(sc * S1);
  a[c] = (float)sc;
  if (train) {
    const double c1 = S1 / count, c2 = dg / count;
    b[c] = (float)(-sc * c2 * is);
    cc[c] = (float)(-sc * c1 + sc * c2 * m * is);
  } else {
    b[c] = 0.f;
    cc[c] = 0.f;
  }
}

__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               const float* __restrict__ a, const float* __restrict__ b,
                                                               const float* __restrict__ cc, float* __restrict__ dz, int C, int P,
                                                               int act) {
  const int plane = blockIdx.x, c = plane % C;
  const float s = scale ? __ldg(scale + c) : 1.f, t = shift ? __ldg(shift + c) : 0.f;
  const float av = __ldg(a + c), bv = __ldg(b + c), cv = __ldg(cc + c);
  const int64_t base = (int64_t)plane * P;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P; i += gridDim.y * blockDim.x) {
    const float zv = __ldg(z + base + i);
    float g = __ldg(dy + base + i);
    if (act == 1 && !(fmaf(zv, s, t) > 0.f)) g = 0.f;
    dz[base + i] = fmaf(av, g, fmaf(bv, zv, cv));
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3 backward
// ---------------------------------------------------------------------------------------------
// dx[b,c,i,j] = sum_kk sum_{dy,dx} w[c*k+kk][dy][dx] * dd[b, c*k+kk, i-dy+1, j-dx+1]; split over the virtual concat
__global__ void __launch_bounds__(256) dw3x3_bwd_input_kernel(const float* __restrict__ dd, const float* __restrict__ w,
                                                              float* __restrict__ dx0, int C0, int64_t bs0,
                                                              float* __restrict__ dx1, int C1, int64_t bs1, int H, int W, int k) {
  const int Cin = C0 + C1;
  const int plane = blockIdx.y;  // b*Cin + c
  const int b = plane / Cin, c = plane - b * Cin;
  float* dst = (c < C0) ? dx0 + (int64_t)b * bs0 + (int64_t)c * H * W : dx1 + (int64_t)b * bs1 + (int64_t)(c - C0) * H * W;
  const int P = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    float acc = 0.f;
    for (int kk = 0; kk < k; ++kk) {
      const int o = c * k + kk;
      const float* g = dd + ((int64_t)b * Cin * k + o) * P;
      const float* wk = w + (int64_t)o * 9;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y - dy + 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xx = x - dx + 1;
          if (xx < 0 || xx >= W) continue;
          acc = fmaf(__ldg(wk + dy * 3 + dx), __ldg(g + (int64_t)yy * W + xx), acc);
        }
      }
    }
    dst[i] = acc;
  }
}

// dW[o][dy][dx] += sum dd[b,o,i,j] * in[b,o/k,i+dy-1,j+dx-1];  db[o] += sum dd.  in = act(in_scale*x+in_shift) if given.
__global__ void __launch_bounds__(256) dw3x3_bwd_weight_kernel(const float* __restrict__ dd, const float* __restrict__ x0, int C0,
                                                               int64_t bs0, const float* __restrict__ x1, int C1, int64_t bs1,
                                                               const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                                                               float* __restrict__ dw, float* __restrict__ db, int B, int H, int W,
                                                               int k, int chunks) {
  const int Cin = C0 + C1;
  const int o = blockIdx.y, c = o / k;
  const int P = H * W;
  const int64_t n = (int64_t)B * P;
  const int64_t per = (n + chunks - 1) / chunks;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
  const bool pro = in_scale != nullptr;
  const float s = pro ? __ldg(in_scale + c) : 1.f, t = pro ? __ldg(in_shift + c) : 0.f;
  float acc[10];
#pragma unroll
  for (int q = 0; q < 10; ++q) acc[q] = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int b = (int)(i / P);
    const int pp = (int)(i - (int64_t)b * P);
    const int y = pp / W, x = pp - y * W;
    const float g = __ldg(dd + ((int64_t)b * Cin * k + o) * P + pp);
    const float* src = (c < C0) ? x0 + (int64_t)b * bs0 + (int64_t)c * P : x1 + (int64_t)b * bs1 + (int64_t)(c - C0) * P;
    acc[9] += g;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= W) continue;
        float v = __ldg(src + (int64_t)yy * W + xx);
        if (pro) v = fmaxf(fmaf(v, s, t), 0.f);
        acc[dy * 3 + dx] = fmaf(g, v, acc[dy * 3 + dx]);
      }
    }
  }
  __shared__ float red[10][8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    const float v = warp_sum(acc[q]);
    if (lane == 0) red[q][wp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[threadIdx.x][i];
    if (threadIdx.x < 9) atomicAdd(dw + (int64_t)o * 9 + threadIdx.x, v);
    else if (db) atomicAdd(db + o, v);
  }
}

// ---------------------------------------------------------------------------------------------
// pointwise 1x1 weight gradient: dW[o][c] += sum_{b,p} dz[b,o,p] * d[b,c,p];  db[o] += sum dz
// CTA: 64 (o) x 64 (c) tile over one pixel chunk; 256 threads, 4x4 micro-tile; fp32 atomics to merge chunks.
// ---------------------------------------------------------------------------------------------
constexpr int WG_T = 64, WG_PX = 32;
__global__ void __launch_bounds__(256) pw1x1_bwd_weight_kernel(const float* __restrict__ dz, const float* __restrict__ d,
                                                               float* __restrict__ dW, float* __restrict__ db, int K, int Cout,
                                                               int P, int px_per_cta) {
  __shared__ float Zs[WG_PX][WG_T + 1];
  __shared__ float Ds[WG_PX][WG_T + 1];
  const int o0 = blockIdx.x * WG_T, c0 = blockIdx.y * WG_T;
  const int chunks_per_img = (P + px_per_cta - 1) / px_per_cta;
  const int b = blockIdx.z / chunks_per_img;
  const int p_lo = (blockIdx.z - b * chunks_per_img) * px_per_cta;
  const int p_hi = min(P, p_lo + px_per_cta);
  const float* zb = dz + (int64_t)b * Cout * P;
  const float* dbp = d + (int64_t)b * K * P;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each 4 (o) x 4 (c)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  for (int p0 = p_lo; p0 < p_hi; p0 += WG_PX) {
    // load 64 channels x 32 px of dz and d (coalesced along px), transposed into [px][ch]
    for (int i = threadIdx.x; i < WG_T * WG_PX; i += 256) {
      const int ch = i / WG_PX, px = i - ch * WG_PX;
      const int pp = p0 + px;
      const bool pv = pp < p_hi;
      Zs[px][ch] = (pv && o0 + ch < Cout) ? __ldg(zb + (int64_t)(o0 + ch) * P + pp) : 0.f;
      Ds[px][ch] = (pv && c0 + ch < K) ? __ldg(dbp + (int64_t)(c0 + ch) * P + pp) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int px = 0; px < WG_PX; ++px) {
      float zv[4], dv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { zv[i] = Zs[px][ty * 4 + i]; dv[i] = Ds[px][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zv[i], dv[j], acc[i][j]);
    }
    if (db && blockIdx.y == 0 && threadIdx.x < WG_T) {
      float sacc = 0.f;
      for (int px = 0; px < WG_PX; ++px) sacc += Zs[px][threadIdx.x];
      bsum += sacc;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (o < Cout && c < K) atomicAdd(dW + (int64_t)o * K + c, acc[i][j]);
    }
  if (db && blockIdx.y == 0 && threadIdx.x < WG_T && o0 + threadIdx.x < Cout) atomicAdd(db + o0 + threadIdx.x, bsum);
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float t[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8)
    if (x < cols && y0 + j < rows) t[j][threadIdx.x] = src[(int64_t)(y0 + j) * cols + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8)
    if (xo < rows && yo0 + j < cols) dst[(int64_t)(yo0 + j) * rows + xo] = t[threadIdx.x][j];
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(2) backward: dy goes to the first maximum of each 2x2 window (row-major order, as torch)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ dx, int64_t N, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const int Hc = (H + 1) / 2, Wc = (W + 1) / 2;  // windows incl. the ragged last row/col (which only get zeros)
  const int64_t total = N * (int64_t)Hc * Wc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wx = (int)(i % Wc);
    const int64_t t = i / Wc;
    const int wy = (int)(t % Hc);
    const int64_t n = t / Hc;
    const float* xp = x + n * H * (int64_t)W;
    float* dp = dx + n * H * (int64_t)W;
    const int y0 = 2 * wy, x0 = 2 * wx;
    const bool full = (wy < Ho) && (wx < Wo);
    int best = 0;
    if (full) {
      float bv = xp[(int64_t)y0 * W + x0];
      const float v1 = xp[(int64_t)y0 * W + x0 + 1], v2 = xp[(int64_t)(y0 + 1) * W + x0], v3 = xp[(int64_t)(y0 + 1) * W + x0 + 1];
      if (v1 > bv) { bv = v1; best = 1; }
      if (v2 > bv) { bv = v2; best = 2; }
      if (v3 > bv) { bv = v3; best = 3; }
    }
    const float g = full ? dy[(n * Ho + wy) * (int64_t)Wo + wx] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = y0 + (q >> 1), xx = x0 + (q & 1);
      if (yy < H && xx < W) dp[(int64_t)yy * W + xx] = (full && q == best) ? g : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear x2 (align_corners) + pad backward, as a gather (no atomics, no memset): source row t receives from
// the upsampled rows u in [2t-2, 2t+3] (src = u*(n-1)/(2n-1) lies in [u/2-1/2, u/2]), likewise for columns.
// Separable: phase 1 reduces each gradient row horizontally into shared memory (6 taps per source column),
// phase 2 reduces vertically (6 taps).  The weights repeat the forward's fp32 index arithmetic (upsample.cu),
// so this is the exact adjoint of smaat_upsample2x_pad_fwd.
// ---------------------------------------------------------------------------------------------
constexpr int UB_TY = 16, UB_TX = 64, UB_ROWS = 2 * UB_TY + 4;

__device__ __forceinline__ float up_adj_weight(int u, int t, int n, float r) {
  if (u < 0 || u >= 2 * n) return 0.f;
  const float s = r * (float)u;
  const int i0 = min((int)s, n - 1), i1 = min(i0 + 1, n - 1);
  const float l = s - (float)i0;
  return (i0 == t ? 1.f - l : 0.f) + (i1 == t ? l : 0.f);
}

__global__ void __launch_bounds__(256) upsample2x_pad_bwd_kernel(const float* __restrict__ dy, int64_t dy_bstride,
                                                                 float* __restrict__ dx, int C, int H, int W, int Ho, int Wo,
                                                                 int pad_t, int pad_l, float ry, float rx, int tiles_x) {
  __shared__ float hs[UB_ROWS][UB_TX + 1];
  __shared__ float wys[UB_TY][6];
  const int c = blockIdx.y, b = blockIdx.z;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int y0 = ty * UB_TY, x0 = tx * UB_TX;
  const float* g = dy + (int64_t)b * dy_bstride + (int64_t)c * Ho * Wo + (int64_t)pad_t * Wo + pad_l;
  const int col = threadIdx.x & (UB_TX - 1), rgrp = threadIdx.x / UB_TX;   // 4 row groups
  const int x = min(x0 + col, W - 1);   // columns past the edge recompute the last one (never stored)
  const int u0 = 2 * x - 2;
  // out-of-range taps get weight 0 and a clamped (valid) address: no predicates in the row loop
  float wx[6];
  int ox[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    wx[i] = up_adj_weight(u0 + i, x, W, rx);
    ox[i] = min(max(u0 + i, 0), 2 * W - 1);
  }
  if (threadIdx.x < UB_TY * 6) {
    const int yy = threadIdx.x / 6, i = threadIdx.x - yy * 6;
    const int y = y0 + yy;
    wys[yy][i] = (y < H) ? up_adj_weight(2 * y - 2 + i, y, H, ry) : 0.f;
  }
  const int r0 = 2 * y0 - 2;   // first upsampled row this tile needs
#pragma unroll 3
  for (int r = rgrp; r < UB_ROWS; r += 256 / UB_TX) {
    const int uy = min(max(r0 + r, 0), 2 * H - 1);   // rows outside the image only meet zero vertical weights
    const float* row = g + (int64_t)uy * Wo;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) a = fmaf(wx[i], __ldg(row + ox[i]), a);
    hs[r][col] = a;
  }
  __syncthreads();
  if (x0 + col >= W) return;
  float* dst = dx + ((int64_t)b * C + c) * H * W;
#pragma unroll
  for (int yy = rgrp; yy < UB_TY; yy += 256 / UB_TX) {
    const int y = y0 + yy;
    if (y >= H) break;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) a = fmaf(wys[yy][i], hs[2 * yy + i][col], a);
    dst[(int64_t)y * W + x] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// OutConv backward (ncls small)
// ---------------------------------------------------------------------------------------------
// 128-bit variants (P % 4 == 0): the input gradient streams dy once per 4 pixels and writes every channel; the weight
// gradient is one CTA per (b, c) plane slice (no per-element index division), merged with fp32 atomics.
__global__ void __launch_bounds__(256) outconv_bwd_input_v4(const float* __restrict__ dy, const float* __restrict__ w,
                                                            float* __restrict__ dx, int Cin, int ncls, int P4) {
  const int b = blockIdx.y;
  const float4* g4 = reinterpret_cast<const float4*>(dy) + (int64_t)b * ncls * P4;
  float4* o4 = reinterpret_cast<float4*>(dx) + (int64_t)b * Cin * P4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P4; i += gridDim.x * blockDim.x) {
    if (ncls == 1) {
      const float4 g = __ldg(g4 + i);
      for (int c = 0; c < Cin; ++c) {
        const float wv = __ldg(w + c);
        o4[(int64_t)c * P4 + i] = make_float4(wv * g.x, wv * g.y, wv * g.z, wv * g.w);
      }
    } else {
      for (int c = 0; c < Cin; ++c) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < ncls; ++j) {
          const float wv = __ldg(w + (int64_t)j * Cin + c);
          const float4 g = __ldg(g4 + (int64_t)j * P4 + i);
          a.x = fmaf(wv, g.x, a.x); a.y = fmaf(wv, g.y, a.y); a.z = fmaf(wv, g.z, a.z); a.w = fmaf(wv, g.w, a.w);
        }
        o4[(int64_t)c * P4 + i] = a;
      }
    }
  }
}

__global__ void __launch_bounds__(256) outconv_bwd_weight_v4(const float* __restrict__ dy, const float* __restrict__ x,
                                                             float* __restrict__ dW, float* __restrict__ db, int Cin, int ncls,
                                                             int P4) {
  const int plane = blockIdx.x, b = plane / Cin, c = plane - b * Cin;
  const float4* x4 = reinterpret_cast<const float4*>(x) + (int64_t)plane * P4;
  __shared__ float red[2][8];
  for (int j = 0; j < ncls; ++j) {
    const float4* g4 = reinterpret_cast<const float4*>(dy) + ((int64_t)b * ncls + j) * P4;
    float a = 0.f, sg = 0.f;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P4; i += gridDim.y * blockDim.x) {
      const float4 g = __ldg(g4 + i), xv = __ldg(x4 + i);
      a += (g.x * xv.x + g.y * xv.y) + (g.z * xv.z + g.w * xv.w);
      sg += (g.x + g.y) + (g.z + g.w);
    }
    a = warp_sum(a);
    sg = warp_sum(sg);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = sg; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f, u = 0.f;
      for (int i = 0; i < 8; ++i) { v += red[0][i]; u += red[1][i]; }
      atomicAdd(dW + (int64_t)j * Cin + c, v);
      if (c == 0 && db) atomicAdd(db + j, u);
    }
  }
}

__global__ void __launch_bounds__(256) outconv_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                float* __restrict__ dx, int Cin, int ncls, int P) {
  const int c = blockIdx.y, b = blockIdx.z;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < ncls; ++j) acc = fmaf(__ldg(w + (int64_t)j * Cin + c), __ldg(dy + ((int64_t)b * ncls + j) * P + i), acc);
    dx[((int64_t)b * Cin + c) * P + i] = acc;
  }
}
// dW[j][c] += sum_{b,p} dy[b,j,p]*x[b,c,p]; db[j] += sum dy   (grid: chunks x Cin; loops over classes)
__global__ void __launch_bounds__(256) outconv_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                 float* __restrict__ dW, float* __restrict__ db, int B, int Cin,
                                                                 int ncls, int P, int chunks) {
  const int c = blockIdx.y;
  const int64_t n = (int64_t)B * P;
  const int64_t per = (n + chunks - 1) / chunks;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
  __shared__ float red[8];
  for (int j = 0; j < ncls; ++j) {
    float a = 0.f, s = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const int64_t bb = i / P, pp = i - bb * P;
      const float g = __ldg(dy + (bb * ncls + j) * (int64_t)P + pp);
      a = fmaf(g, __ldg(x + (bb * Cin + c) * (int64_t)P + pp), a);
      s += g;
    }
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
      for (int i = 0; i < 8; ++i) v += red[i];
      atomicAdd(dW + (int64_t)j * Cin + c, v);
    }
    __syncthreads();
    if (c == 0 && db) {
      s = warp_sum(s);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        float v = 0.f;
        for (int i = 0; i < 8; ++i) v += red[i];
        atomicAdd(db + j, v);
      }
      __syncthreads();
    }
  }
}

}  // namespace smaat

namespace smaat {
bool bn_bwd_v4_ok(const void* dy, const void* z, const void* dz, int P);
int bn_act_bwd_reduce_v4_launch(const float* dy, const float* z, const float* scale, const float* shift, double* sums, int B, int C,
                                int P, int act, cudaStream_t st);
int bn_act_bwd_apply_v4_launch(const float* dy, const float* z, const float* scale, const float* shift, const float* a, const float* b,
                               const float* cc, float* dz, int B, int C, int P, int act, cudaStream_t st);
int dw3x3_bwd_input_tiled_launch(const float* dd, const float* w, float* dx0, int C0, int64_t bs0, float* dx1, int C1, int64_t bs1,
                                 int B, int H, int W, int k, cudaStream_t st);
int dw3x3_bwd_weight_tiled_launch(const float* dd, const float* x0, int C0, int64_t bs0, const float* x1, int C1, int64_t bs1,
                                  const float* in_scale, const float* in_shift, float* dw, float* db, int B, int H, int W, int k,
                                  cudaStream_t st);
}  // namespace smaat

using namespace smaat;

static inline unsigned grid1d(int64_t items, int threads, int cap_mult = 32) {
  int64_t b = ceil_div64(items, threads);
  const int64_t cap = (int64_t)num_sms() * cap_mult;
  return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}
static inline int pick_chunks(int64_t n, int C) {
  int chunks = (int)ceil_div64(n, 256 * 64);
  const int maxc = ceil_div(num_sms() * 8, C);
  if (chunks > maxc) chunks = maxc;
  return chunks < 1 ? 1 : chunks;
}

extern "C" int smaat_bn_act_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift, double* sums, int B,
                                       int C, int P, int act, void* stream) {
  SMAAT_REQUIRE(dy && z && sums && B > 0 && C > 0 && P > 0 && C <= 65535, "bn_act_bwd_reduce: bad arguments");
  if (bn_bwd_v4_ok(dy, z, nullptr, P)) return bn_act_bwd_reduce_v4_launch(dy, z, scale, shift, sums, B, C, P, act, (cudaStream_t)stream);
  const int chunks = pick_chunks((int64_t)B * P, C);
  bn_act_bwd_reduce_kernel<<<dim3(chunks, C), 256, 0, (cudaStream_t)stream>>>(dy, z, scale, shift, sums, B, C, P, act, chunks);
  SMAAT_LAUNCH_CHECK("smaat_bn_act_bwd_reduce");
  return SMAAT_OK;
}

extern "C" int smaat_bn_bwd_coeffs(const double* sums, double count, const float* gamma, const float* mean, const float* invstd,
                                   int train, float* a, float* b, float* cc, float* dgamma, float* dbeta, float* dz_sum, int C,
                                   void* stream) {
  SMAAT_REQUIRE(sums && mean && invstd && a && b && cc && C > 0 && count > 0, "bn_bwd_coeffs: bad arguments");
  bn_bwd_coeffs_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, gamma, mean, invstd, train, a, b, cc, dgamma,
                                                                           dbeta, dz_sum, C);
  SMAAT_LAUNCH_CHECK("smaat_bn_bwd_coeffs");
  return SMAAT_OK;
}

extern "C" int smaat_bn_act_bwd_apply(const float* dy, const float* z, const float* scale, const float* shift, const float* a,
                                      const float* b, const float* cc, float* dz, int B, int C, int P, int act, void* stream) {
  SMAAT_REQUIRE(dy && z && a && b && cc && dz && B > 0 && C > 0 && P > 0, "bn_act_bwd_apply: bad arguments");
  if (bn_bwd_v4_ok(dy, z, dz, P))
    return bn_act_bwd_apply_v4_launch(dy, z, scale, shift, a, b, cc, dz, B, C, P, act, (cudaStream_t)stream);
  unsigned gy = grid1d(P, 256, 4);
  if (gy > 65535u) gy = 65535u;
  dim3 grid(B * C, gy);
  bn_act_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, z, scale, shift, a, b, cc, dz, C, P, act);
  SMAAT_LAUNCH_CHECK("smaat_bn_act_bwd_apply");
  return SMAAT_OK;
}

extern "C" int smaat_dw3x3_bwd_input(const float* dd, const float* w, float* dx0, int C0, int64_t dx0_bstride, float* dx1, int C1,
                                     int64_t dx1_bstride, int B, int H, int W, int k, void* stream) {
  SMAAT_REQUIRE(dd && w && dx0 && B > 0 && C0 > 0 && C1 >= 0 && H > 0 && W > 0 && k > 0, "dw3x3_bwd_input: bad arguments");
  SMAAT_REQUIRE(C1 == 0 || dx1, "dw3x3_bwd_input: C1 > 0 but dx1 null");
  if (k <= 4) return dw3x3_bwd_input_tiled_launch(dd, w, dx0, C0, dx0_bstride, dx1, C1, dx1_bstride, B, H, W, k, (cudaStream_t)stream);
  SMAAT_REQUIRE((int64_t)B * (C0 + C1) <= 65535 * 32767ll, "dw3x3_bwd_input: too many planes");
  const int64_t planes = (int64_t)B * (C0 + C1);
  SMAAT_REQUIRE(planes <= 65535, "dw3x3_bwd_input: B*Cin > 65535 not supported yet");
  dim3 grid(grid1d((int64_t)H * W, 256, 4), (unsigned)planes);
  dw3x3_bwd_input_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dd, w, dx0, C0, dx0_bstride, dx1, C1, dx1_bstride, H, W, k);
  SMAAT_LAUNCH_CHECK("smaat_dw3x3_bwd_input");
  return SMAAT_OK;
}

extern "C" int smaat_dw3x3_bwd_weight(const float* dd, const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1,
                                      int64_t x1_bstride, const float* in_scale, const float* in_shift, float* dw, float* db, int B,
                                      int H, int W, int k, void* stream) {
  SMAAT_REQUIRE(dd && x0 && dw && B > 0 && C0 > 0 && C1 >= 0 && H > 0 && W > 0 && k > 0, "dw3x3_bwd_weight: bad arguments");
  if (k <= 4)
    return dw3x3_bwd_weight_tiled_launch(dd, x0, C0, x0_bstride, x1, C1, x1_bstride, in_scale, in_shift, dw, db, B, H, W, k,
                                         (cudaStream_t)stream);
  const int KC = k * (C0 + C1);
  SMAAT_REQUIRE(KC <= 65535, "dw3x3_bwd_weight: too many channels");
  const int chunks = pick_chunks((int64_t)B * H * W, KC);
  dw3x3_bwd_weight_kernel<<<dim3(chunks, KC), 256, 0, (cudaStream_t)stream>>>(dd, x0, C0, x0_bstride, x1, C1, x1_bstride, in_scale,
                                                                              in_shift, dw, db, B, H, W, k, chunks);
  SMAAT_LAUNCH_CHECK("smaat_dw3x3_bwd_weight");
  return SMAAT_OK;
}

extern "C" int smaat_pw1x1_bwd_weight(const float* dz, const float* d, float* dW, float* db, int B, int K, int Cout, int P,
                                      void* stream) {
  SMAAT_REQUIRE(dz && d && dW && B > 0 && K > 0 && Cout > 0 && P > 0, "pw1x1_bwd_weight: bad arguments");
  // pixel chunk per CTA: enough CTAs to fill the GPU, few enough that the fp32 atomics stay cheap
  const int tiles = ceil_div(Cout, WG_T) * ceil_div(K, WG_T);
  int64_t want = (int64_t)num_sms() * 8 / tiles;
  if (want < 1) want = 1;
  int64_t px = ceil_div64((int64_t)B * P, want);
  px = ((px + WG_PX - 1) / WG_PX) * WG_PX;
  if (px < 256) px = 256;
  if (px > P) px = ((P + WG_PX - 1) / WG_PX) * WG_PX;
  const int chunks_per_img = (int)ceil_div64(P, px);
  SMAAT_REQUIRE((int64_t)B * chunks_per_img <= 65535, "pw1x1_bwd_weight: grid.z too large");
  dim3 grid(ceil_div(Cout, WG_T), ceil_div(K, WG_T), B * chunks_per_img);
  pw1x1_bwd_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dz, d, dW, db, K, Cout, P, (int)px);
  SMAAT_LAUNCH_CHECK("smaat_pw1x1_bwd_weight");
  return SMAAT_OK;
}

extern "C" int smaat_transpose(const float* src, float* dst, int rows, int cols, void* stream) {
  SMAAT_REQUIRE(src && dst && rows > 0 && cols > 0, "transpose: bad arguments");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, rows, cols);
  SMAAT_LAUNCH_CHECK("smaat_transpose");
  return SMAAT_OK;
}

extern "C" int smaat_maxpool2_bwd(const float* x, const float* dy, float* dx, int64_t N, int H, int W, void* stream) {
  SMAAT_REQUIRE(x && dy && dx && N > 0 && H >= 2 && W >= 2, "maxpool2_bwd: bad arguments");
  maxpool2_bwd_kernel<<<grid1d(N * ((H + 1) / 2) * ((W + 1) / 2), 256, 64), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, N, H, W);
  SMAAT_LAUNCH_CHECK("smaat_maxpool2_bwd");
  return SMAAT_OK;
}

extern "C" int smaat_upsample2x_pad_bwd(const float* dy, int64_t dy_bstride, float* dx, int B, int C, int H, int W, int Ho, int Wo,
                                        void* stream) {
  SMAAT_REQUIRE(dy && dx && B > 0 && C > 0 && H > 0 && W > 0 && Ho >= 2 * H && Wo >= 2 * W, "upsample2x_bwd: bad arguments");
  SMAAT_REQUIRE(C <= 65535 && B <= 65535, "upsample2x_bwd: C/B too large");
  const int pad_t = (Ho - 2 * H) / 2, pad_l = (Wo - 2 * W) / 2;
  const float ry = (2 * H > 1) ? (float)(H - 1) / (float)(2 * H - 1) : 0.f;
  const float rx = (2 * W > 1) ? (float)(W - 1) / (float)(2 * W - 1) : 0.f;
  const int tiles_x = ceil_div(W, UB_TX), tiles_y = ceil_div(H, UB_TY);
  dim3 grid(tiles_x * tiles_y, C, B);
  upsample2x_pad_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, dy_bstride, dx, C, H, W, Ho, Wo, pad_t, pad_l, ry, rx, tiles_x);
  SMAAT_LAUNCH_CHECK("smaat_upsample2x_pad_bwd");
  return SMAAT_OK;
}

extern "C" int smaat_outconv_bwd(const float* dy, const float* x, const float* w, float* dx, float* dW, float* db, int B, int Cin,
                                 int ncls, int P, void* stream) {
  SMAAT_REQUIRE(dy && x && w && B > 0 && Cin > 0 && ncls > 0 && P > 0, "outconv_bwd: bad arguments");
  SMAAT_REQUIRE(Cin <= 65535 && B <= 65535, "outconv_bwd: Cin/B too large");
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = (P % 4 == 0) && aligned16(dy) && aligned16(x) && (!dx || aligned16(dx));
  if (v4) {
    const int P4 = P / 4;
    if (dx) {
      unsigned gx = (unsigned)ceil_div(P4, 256);
      outconv_bwd_input_v4<<<dim3(gx, B), 256, 0, st>>>(dy, w, dx, Cin, ncls, P4);
      SMAAT_LAUNCH_CHECK("smaat_outconv_bwd(input)");
    }
    if (dW) {
      int slices = ceil_div(P4, 256 * 8);
      if (slices > 65535) slices = 65535;
      SMAAT_REQUIRE((int64_t)B * Cin < (1ll << 31), "outconv_bwd: too many planes");
      outconv_bwd_weight_v4<<<dim3((unsigned)(B * Cin), (unsigned)slices), 256, 0, st>>>(dy, x, dW, db, Cin, ncls, P4);
      SMAAT_LAUNCH_CHECK("smaat_outconv_bwd(weight)");
    }
    return SMAAT_OK;
  }
  if (dx) {
    dim3 grid(grid1d(P, 256, 4), Cin, B);
    outconv_bwd_input_kernel<<<grid, 256, 0, st>>>(dy, w, dx, Cin, ncls, P);
    SMAAT_LAUNCH_CHECK("smaat_outconv_bwd(input)");
  }
  if (dW) {
    const int chunks = pick_chunks((int64_t)B * P, Cin);
    outconv_bwd_weight_kernel<<<dim3(chunks, Cin), 256, 0, st>>>(dy, x, dW, db, B, Cin, ncls, P, chunks);
    SMAAT_LAUNCH_CHECK("smaat_outconv_bwd(weight)");
  }
  return SMAAT_OK;
}
