// cbam_bwd.cu -- CBAM backward (reference models/layers.py:90-141 differentiated).
//
// Forward (cbam.cu): sc = sigmoid(MLP(avg_p x) + MLP(max_p x)); u = x*sc; pooled = [mean_c u, max_c u];
// raw = conv_kxk(pooled); sa = sigmoid(BN1(raw)); out = u*sa.   Given g = dL/dout, with
//   d_u[b,c,p] = g*sa[p] + d_pooled[0][p]/C + [c == argmax_c u(p)] d_pooled[1][p]:
//   gate_in : d_pre[b,p] = (sum_c g*u) * sa*(1-sa), and the channel argmax of u      (reads g, x)
//             (then BN(1) backward -> d_raw, bn.cu/backward.cu)
//   conv    : d_pooled = conv_transpose(d_raw, w);  dW = corr(pooled, d_raw)          (maps only, smem-tiled)
//   dsc     : d_sc[b,c] = sum_p d_u*x, and the plane argmax of x                     (reads g, x)
//   mlp     : tiny per-image MLP backward -> dW1, db1, dW2, db2, d_avg, d_max
//   dx      : dx = d_u*sc + d_avg/P + [p == argmax_p x] d_max                         (reads g, writes dx)
// HBM traffic 6|x| (5 reads + 1 write); every |x|-sized pass uses 128-bit accesses when P % 4 == 0.
#include "common.cuh"

namespace smaat {

// ---- d_pre[b,p] = (sum_c g*x*sc) * sa*(1-sa);  amax[b,p] = argmax_c x*sc (lowest c on ties) ------------------
__global__ void __launch_bounds__(256) cbam_bwd_gate_in_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                               const float* __restrict__ sc, const float* __restrict__ sa,
                                                               float* __restrict__ dpre, int* __restrict__ amax, int C, int P) {
  __shared__ float rs[8][33];
  __shared__ float mv[8][33];
  __shared__ int mi[8][33];
  const int tx = threadIdx.x, cg = threadIdx.y, b = blockIdx.y;
  const int pp = blockIdx.x * 32 + tx;
  float s = 0.f, best = -INFINITY;
  int bi = 0x7fffffff;
  if (pp < P) {
    const float* gb = g + (int64_t)b * C * P + pp;
    const float* xb = x + (int64_t)b * C * P + pp;
    for (int c = cg; c < C; c += 8) {
      const float u = __ldg(xb + (int64_t)c * P) * __ldg(sc + (int64_t)b * C + c);
      s = fmaf(__ldg(gb + (int64_t)c * P), u, s);
      if (u > best) { best = u; bi = c; }
    }
  }
  rs[cg][tx] = s;
  mv[cg][tx] = best;
  mi[cg][tx] = bi;
  __syncthreads();
  if (cg == 0 && pp < P) {
    for (int i = 1; i < 8; ++i) s += rs[i][tx];
    const float a = __ldg(sa + (int64_t)b * P + pp);
    dpre[(int64_t)b * P + pp] = s * a * (1.f - a);
    if (amax) {
      float bb = mv[0][tx];
      int am = mi[0][tx];
      for (int i = 1; i < 8; ++i) {
        const float v = mv[i][tx];
        const int id = mi[i][tx];
        if (v > bb || (v == bb && id < am)) { bb = v; am = id; }
      }
      amax[(int64_t)b * P + pp] = am;
    }
  }
}

// 128-bit variant for large planes: a thread owns 4 pixels and walks every channel (no cross-thread reduction)
__global__ void __launch_bounds__(256) cbam_bwd_gate_in_v4(const float* __restrict__ g, const float* __restrict__ x,
                                                           const float* __restrict__ sc, const float* __restrict__ sa,
                                                           float* __restrict__ dpre, int* __restrict__ amax, int C, int P4) {
  extern __shared__ float scs[];   // sc[b, :]
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) scs[c] = __ldg(sc + (int64_t)b * C + c);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P4) return;
  const float4* g4 = reinterpret_cast<const float4*>(g) + (int64_t)b * C * P4 + i;
  const float4* x4 = reinterpret_cast<const float4*>(x) + (int64_t)b * C * P4 + i;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int bi[4] = {0, 0, 0, 0};
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float4 gv = __ldg(g4 + (int64_t)c * P4), xv = __ldg(x4 + (int64_t)c * P4);
    const float w = scs[c];
    const float u[4] = {xv.x * w, xv.y * w, xv.z * w, xv.w * w};
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] = fmaf(gg[j], u[j], s[j]);
      if (u[j] > best[j]) { best[j] = u[j]; bi[j] = c; }
    }
  }
  const float4 a = __ldg(reinterpret_cast<const float4*>(sa) + (int64_t)b * P4 + i);
  reinterpret_cast<float4*>(dpre)[(int64_t)b * P4 + i] =
      make_float4(s[0] * a.x * (1.f - a.x), s[1] * a.y * (1.f - a.y), s[2] * a.z * (1.f - a.z), s[3] * a.w * (1.f - a.w));
  if (amax) reinterpret_cast<int4*>(amax)[(int64_t)b * P4 + i] = make_int4(bi[0], bi[1], bi[2], bi[3]);
}

// ---- spatial conv backward, smem-tiled ----------------------------------------------------------------------
constexpr int GB_T = 32;  // tile edge

// d_pooled[b,ch,y,x] = sum_{dy,dx} w[ch][dy][dx] * d_raw[b, y-dy+R, x-dx+R]
template <int KS>
__global__ void __launch_bounds__(256) cbam_gate_bwd_input_kernel(const float* __restrict__ draw, const float* __restrict__ wsp,
                                                                  float* __restrict__ dpooled, int H, int W, int tiles_x) {
  constexpr int R = KS / 2, TP = GB_T + 2 * R + 1;   // odd pitch
  __shared__ float tile[(GB_T + 2 * R) * TP];
  __shared__ float ws[2 * KS * KS];
  const int b = blockIdx.y;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int y0 = ty * GB_T, x0 = tx * GB_T;
  const int P = H * W;
  const float* src = draw + (int64_t)b * P;
  for (int i = threadIdx.x; i < (GB_T + 2 * R) * (GB_T + 2 * R); i += 256) {
    const int r = i / (GB_T + 2 * R), c = i - r * (GB_T + 2 * R);
    const int gy = y0 - R + r, gx = x0 - R + c;
    tile[r * TP + c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(src + (int64_t)gy * W + gx) : 0.f;
  }
  if (threadIdx.x < 2 * KS * KS) ws[threadIdx.x] = __ldg(wsp + threadIdx.x);
  __syncthreads();
  const int row = threadIdx.x >> 3, col = (threadIdx.x & 7) << 2;   // 4 pixels per thread
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < KS; ++dy) {
    // output (row, col+j) reads d_raw at tile row (row + R) - dy + R = row + 2R - dy, col (col+j) + 2R - dx
    const float* tr = tile + (row + 2 * R - dy) * TP + col;
    float v[4 + 2 * R];
#pragma unroll
    for (int i = 0; i < 4 + 2 * R; ++i) v[i] = tr[i];
#pragma unroll
    for (int dx = 0; dx < KS; ++dx) {
      const float w0 = ws[dy * KS + dx], w1 = ws[KS * KS + dy * KS + dx];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a0[j] = fmaf(w0, v[j + 2 * R - dx], a0[j]);
        a1[j] = fmaf(w1, v[j + 2 * R - dx], a1[j]);
      }
    }
  }
  const int gy = y0 + row;
  if (gy < H) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gx = x0 + col + j;
      if (gx < W) {
        dpooled[((int64_t)b * 2) * P + (int64_t)gy * W + gx] = a0[j];
        dpooled[((int64_t)b * 2 + 1) * P + (int64_t)gy * W + gx] = a1[j];
      }
    }
  }
}

// dW[ch][dy][dx] += sum_{b,y,x} d_raw[b,y,x] * pooled[b,ch,y+dy-R,x+dx-R]; persistent CTAs, thread pair per tap
template <int KS>
__global__ void __launch_bounds__(2 * 2 * KS * KS) cbam_gate_bwd_weight_kernel(const float* __restrict__ draw,
                                                                               const float* __restrict__ pooled,
                                                                               float* __restrict__ dW, int B, int H, int W,
                                                                               int tiles_x, int tiles_y) {
  constexpr int R = KS / 2, TE = GB_T + 2 * R, TP = TE + 1, NT = 2 * KS * KS;
  __shared__ float dt[GB_T * GB_T];
  __shared__ float pt[2][TE * TP];
  const int tap = threadIdx.x % NT, half = threadIdx.x / NT;       // blockDim = 2*NT
  const int ch = tap / (KS * KS), r = tap - ch * KS * KS;
  const int dy = r / KS, dx = r - dy * KS;
  const int P = H * W;
  const int tiles = tiles_x * tiles_y;
  float acc = 0.f;
  for (int t = blockIdx.x; t < B * tiles; t += gridDim.x) {
    const int b = t / tiles, tt = t - b * tiles;
    const int ty = tt / tiles_x, tx = tt - ty * tiles_x;
    const int y0 = ty * GB_T, x0 = tx * GB_T;
    __syncthreads();
    for (int i = threadIdx.x; i < GB_T * GB_T; i += blockDim.x) {
      const int rr = i / GB_T, cc = i - rr * GB_T;
      const int gy = y0 + rr, gx = x0 + cc;
      dt[i] = (gy < H && gx < W) ? __ldg(draw + (int64_t)b * P + (int64_t)gy * W + gx) : 0.f;
    }
    for (int i = threadIdx.x; i < 2 * TE * TE; i += blockDim.x) {
      const int c2 = i / (TE * TE), j = i - c2 * TE * TE;
      const int rr = j / TE, cc = j - rr * TE;
      const int gy = y0 - R + rr, gx = x0 - R + cc;
      pt[c2][rr * TP + cc] = (gy >= 0 && gy < H && gx >= 0 && gx < W)
                                 ? __ldg(pooled + ((int64_t)b * 2 + c2) * P + (int64_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    const float* pp = pt[ch] + dy * TP + dx;
    const int r0 = half * (GB_T / 2);
    for (int rr = r0; rr < r0 + GB_T / 2; ++rr) {
#pragma unroll 8
      for (int cc = 0; cc < GB_T; ++cc) acc = fmaf(dt[rr * GB_T + cc], pp[rr * TP + cc], acc);
    }
  }
  atomicAdd(dW + tap, acc);
}

// ---- |x|-sized passes ------------------------------------------------------------------------------------------
constexpr int CB_CH = 8;     // channels per CTA
constexpr int CB_IT = 4;     // pixel-vector iterations per thread

template <int V> struct VecT;
template <> struct VecT<4> { using F = float4; using I = int4; };
template <> struct VecT<1> { using F = float;  using I = int;  };

template <int V> __device__ __forceinline__ void ldv(float* d, const float* p) {
  if (V == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w; }
  else d[0] = __ldg(p);
}
template <int V> __device__ __forceinline__ void ldvi(int* d, const int* p) {
  if (V == 4) { const int4 t = __ldg(reinterpret_cast<const int4*>(p)); d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w; }
  else d[0] = __ldg(p);
}
template <int V> __device__ __forceinline__ void stv(float* p, const float* s) {
  if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(s[0], s[1], s[2], s[3]);
  else p[0] = s[0];
}

// order-preserving float -> uint32 (larger float = larger key)
__device__ __forceinline__ unsigned fkey(float v) {
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// d_sc[b,c] += sum_p d_u*x;  pkey[b,c] = max over p of (fkey(x) << 32 | ~p)   (ties -> lowest p)
template <int V>
__global__ void __launch_bounds__(256) cbam_bwd_dsc_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                           const float* __restrict__ sa, const float* __restrict__ dpooled,
                                                           const int* __restrict__ amax, float* __restrict__ dsc,
                                                           unsigned long long* __restrict__ pkey, int C, int P) {
  const int c0 = blockIdx.x * CB_CH, b = blockIdx.z;
  const int chunk = 256 * V * CB_IT;
  const int p_lo = blockIdx.y * chunk;
  float acc[CB_CH], bv[CB_CH];
  int bi[CB_CH];
#pragma unroll
  for (int q = 0; q < CB_CH; ++q) { acc[q] = 0.f; bv[q] = -INFINITY; bi[q] = 0x7fffffff; }
  const float invC = 1.f / (float)C;
#pragma unroll 1
  for (int it = 0; it < CB_IT; ++it) {
    const int p = p_lo + (it * 256 + threadIdx.x) * V;
    if (p >= P) break;
    float av[V], d0[V], d1[V];
    int am[V];
    ldv<V>(av, sa + (int64_t)b * P + p);
    ldv<V>(d0, dpooled + ((int64_t)b * 2) * P + p);
    ldv<V>(d1, dpooled + ((int64_t)b * 2 + 1) * P + p);
    ldvi<V>(am, amax + (int64_t)b * P + p);
#pragma unroll
    for (int q = 0; q < CB_CH; ++q) {
      const int c = c0 + q;
      if (c < C) {
        float gv[V], xv[V];
        ldv<V>(gv, g + ((int64_t)b * C + c) * P + p);
        ldv<V>(xv, x + ((int64_t)b * C + c) * P + p);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float du = fmaf(gv[j], av[j], d0[j] * invC) + (am[j] == c ? d1[j] : 0.f);
          acc[q] = fmaf(du, xv[j], acc[q]);
          if (xv[j] > bv[q]) { bv[q] = xv[j]; bi[q] = p + j; }
        }
      }
    }
  }
  __shared__ float rs[8][CB_CH];
  __shared__ unsigned long long rk[8][CB_CH];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < CB_CH; ++q) {
    const float s = warp_sum(acc[q]);
    unsigned long long k = bi[q] == 0x7fffffff ? 0ull : (((unsigned long long)fkey(bv[q]) << 32) | (unsigned)(~(unsigned)bi[q]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
      k = other > k ? other : k;
    }
    if (lane == 0) { rs[wp][q] = s; rk[wp][q] = k; }
  }
  __syncthreads();
  if (threadIdx.x < CB_CH && c0 + threadIdx.x < C) {
    float s = 0.f;
    unsigned long long k = 0ull;
    for (int i = 0; i < 8; ++i) { s += rs[i][threadIdx.x]; k = rk[i][threadIdx.x] > k ? rk[i][threadIdx.x] : k; }
    atomicAdd(dsc + (int64_t)b * C + c0 + threadIdx.x, s);
    atomicMax(pkey + (int64_t)b * C + c0 + threadIdx.x, k);
  }
}

// dx = d_u*sc + d_avg/P + [p == argmax_p x] d_max
template <int V>
__global__ void __launch_bounds__(256) cbam_bwd_dx_kernel(const float* __restrict__ g, const float* __restrict__ sc,
                                                          const float* __restrict__ sa, const float* __restrict__ dpooled,
                                                          const int* __restrict__ amax, const float* __restrict__ davg,
                                                          const float* __restrict__ dmx,
                                                          const unsigned long long* __restrict__ pkey, float* __restrict__ dx,
                                                          int C, int P) {
  const int c0 = blockIdx.x * CB_CH, b = blockIdx.z;
  const int chunk = 256 * V * CB_IT;
  const int p_lo = blockIdx.y * chunk;
  float s[CB_CH], ga[CB_CH], gm[CB_CH];
  int pi[CB_CH];
  const float invP = 1.f / (float)P, invC = 1.f / (float)C;
#pragma unroll
  for (int q = 0; q < CB_CH; ++q) {
    const int c = c0 + q;
    if (c < C) {
      s[q] = __ldg(sc + (int64_t)b * C + c);
      ga[q] = __ldg(davg + (int64_t)b * C + c) * invP;
      gm[q] = __ldg(dmx + (int64_t)b * C + c);
      pi[q] = (int)(~(unsigned)(pkey[(int64_t)b * C + c] & 0xffffffffull));
    } else { s[q] = ga[q] = gm[q] = 0.f; pi[q] = -1; }
  }
#pragma unroll 1
  for (int it = 0; it < CB_IT; ++it) {
    const int p = p_lo + (it * 256 + threadIdx.x) * V;
    if (p >= P) break;
    float av[V], d0[V], d1[V];
    int am[V];
    ldv<V>(av, sa + (int64_t)b * P + p);
    ldv<V>(d0, dpooled + ((int64_t)b * 2) * P + p);
    ldv<V>(d1, dpooled + ((int64_t)b * 2 + 1) * P + p);
    ldvi<V>(am, amax + (int64_t)b * P + p);
#pragma unroll
    for (int q = 0; q < CB_CH; ++q) {
      const int c = c0 + q;
      if (c < C) {
        float gv[V], o[V];
        ldv<V>(gv, g + ((int64_t)b * C + c) * P + p);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float du = fmaf(gv[j], av[j], d0[j] * invC) + (am[j] == c ? d1[j] : 0.f);
          o[j] = fmaf(du, s[q], ga[q]) + (p + j == pi[q] ? gm[q] : 0.f);
        }
        stv<V>(dx + ((int64_t)b * C + c) * P + p, o);
      }
    }
  }
}

// ---- channel MLP backward (one CTA per image) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) cbam_mlp_bwd_kernel(const float* __restrict__ avg, const float* __restrict__ mx,
                                                           const float* __restrict__ w1, const float* __restrict__ b1,
                                                           const float* __restrict__ w2, const float* __restrict__ sc,
                                                           const float* __restrict__ dsc, float* __restrict__ dw1,
                                                           float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2,
                                                           float* __restrict__ davg, float* __restrict__ dmx, int C, int hidden) {
  extern __shared__ float sh[];  // avg[C] mx[C] dm[C] ha[h] hm[h] dha[h] dhm[h]
  float* sa_ = sh;
  float* sm_ = sa_ + C;
  float* dm = sm_ + C;
  float* ha = dm + C;
  float* hm = ha + hidden;
  float* dha = hm + hidden;
  float* dhm = dha + hidden;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sa_[c] = avg[(int64_t)b * C + c];
    sm_[c] = mx[(int64_t)b * C + c];
    const float s = sc[(int64_t)b * C + c];
    dm[c] = dsc[(int64_t)b * C + c] * s * (1.f - s);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < hidden; j += nw) {
    float da = 0.f, dmm = 0.f, back = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float wv = __ldg(w1 + (int64_t)j * C + c);
      da = fmaf(wv, sa_[c], da);
      dmm = fmaf(wv, sm_[c], dmm);
      back = fmaf(__ldg(w2 + (int64_t)c * hidden + j), dm[c], back);
    }
    da = warp_sum(da); dmm = warp_sum(dmm); back = warp_sum(back);
    if (lane == 0) {
      const float pa = da + b1[j], pm = dmm + b1[j];
      ha[j] = fmaxf(pa, 0.f);
      hm[j] = fmaxf(pm, 0.f);
      dha[j] = pa > 0.f ? back : 0.f;
      dhm[j] = pm > 0.f ? back : 0.f;
      atomicAdd(db1 + j, dha[j] + dhm[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float d = dm[c];
    atomicAdd(db2 + c, 2.f * d);   // the second-layer bias enters twice (layers.py:109)
    float ga = 0.f, gm = 0.f;
    for (int j = 0; j < hidden; ++j) {
      atomicAdd(dw2 + (int64_t)c * hidden + j, d * (ha[j] + hm[j]));
      const float wv = __ldg(w1 + (int64_t)j * C + c);
      ga = fmaf(wv, dha[j], ga);
      gm = fmaf(wv, dhm[j], gm);
      atomicAdd(dw1 + (int64_t)j * C + c, dha[j] * sa_[c] + dhm[j] * sm_[c]);
    }
    davg[(int64_t)b * C + c] = ga;
    dmx[(int64_t)b * C + c] = gm;
  }
}

}  // namespace smaat

using namespace smaat;

extern "C" int smaat_cbam_bwd_gate_in(const float* g, const float* x, const float* sc, const float* sa, float* dpre, int* amax,
                                      int B, int C, int P, void* stream) {
  SMAAT_REQUIRE(g && x && sc && sa && dpre && B > 0 && C > 0 && P > 0 && B <= 65535, "cbam_bwd_gate_in: bad arguments");
  if (P % 4 == 0 && P >= 8192 && C * sizeof(float) <= 48 * 1024 && aligned16(g) && aligned16(x) && aligned16(sa) && aligned16(dpre) &&
      (!amax || aligned16(amax))) {
    cbam_bwd_gate_in_v4<<<dim3(ceil_div(P / 4, 256), B), 256, C * sizeof(float), (cudaStream_t)stream>>>(g, x, sc, sa, dpre, amax, C,
                                                                                                        P / 4);
    SMAAT_LAUNCH_CHECK("smaat_cbam_bwd_gate_in");
    return SMAAT_OK;
  }
  cbam_bwd_gate_in_kernel<<<dim3(ceil_div(P, 32), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(g, x, sc, sa, dpre, amax, C, P);
  SMAAT_LAUNCH_CHECK("smaat_cbam_bwd_gate_in");
  return SMAAT_OK;
}

template <int KS>
static int gate_bwd_launch(const float* draw, const float* pooled, const float* wsp, float* dpooled, float* dW, int B, int H, int W,
                           cudaStream_t st) {
  const int tiles_x = ceil_div(W, GB_T), tiles_y = ceil_div(H, GB_T);
  cbam_gate_bwd_input_kernel<KS><<<dim3(tiles_x * tiles_y, B), 256, 0, st>>>(draw, wsp, dpooled, H, W, tiles_x);
  SMAAT_LAUNCH_CHECK("smaat_cbam_gate_bwd(input)");
  int64_t grid = (int64_t)B * tiles_x * tiles_y;
  const int64_t cap = (int64_t)num_sms() * 2;
  if (grid > cap) grid = cap;
  cbam_gate_bwd_weight_kernel<KS><<<(unsigned)grid, 4 * KS * KS, 0, st>>>(draw, pooled, dW, B, H, W, tiles_x, tiles_y);
  SMAAT_LAUNCH_CHECK("smaat_cbam_gate_bwd(weight)");
  return SMAAT_OK;
}

extern "C" int smaat_cbam_gate_bwd(const float* draw, const float* pooled, const float* wsp, float* dpooled, float* dW, int B, int H,
                                   int W, int ks, void* stream) {
  SMAAT_REQUIRE(draw && pooled && wsp && dpooled && dW && B > 0 && H > 0 && W > 0, "cbam_gate_bwd: bad arguments");
  SMAAT_REQUIRE(ks == 3 || ks == 7, "cbam_gate_bwd: kernel size must be 3 or 7");
  SMAAT_REQUIRE(B <= 65535, "cbam_gate_bwd: batch too large");
  if (ks == 7) return gate_bwd_launch<7>(draw, pooled, wsp, dpooled, dW, B, H, W, (cudaStream_t)stream);
  return gate_bwd_launch<3>(draw, pooled, wsp, dpooled, dW, B, H, W, (cudaStream_t)stream);
}

static bool cb_vec_ok(int P, const void* a, const void* b, const void* c, const void* d, const void* e) {
  return P % 4 == 0 && aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d) && aligned16(e);
}

extern "C" int smaat_cbam_bwd_dsc(const float* g, const float* x, const float* sa, const float* dpooled, const int* amax, float* dsc,
                                  unsigned long long* pkey, int B, int C, int P, void* stream) {
  SMAAT_REQUIRE(g && x && sa && dpooled && amax && dsc && pkey && B > 0 && C > 0 && P > 0 && B <= 65535,
                "cbam_bwd_dsc: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (cb_vec_ok(P, g, x, sa, dpooled, amax)) {
    const dim3 grid(ceil_div(C, CB_CH), ceil_div(P, 256 * 4 * CB_IT), B);
    SMAAT_REQUIRE(grid.y <= 65535, "cbam_bwd_dsc: plane too large");
    cbam_bwd_dsc_kernel<4><<<grid, 256, 0, st>>>(g, x, sa, dpooled, amax, dsc, pkey, C, P);
  } else {
    const dim3 grid(ceil_div(C, CB_CH), ceil_div(P, 256 * CB_IT), B);
    SMAAT_REQUIRE(grid.y <= 65535, "cbam_bwd_dsc: plane too large");
    cbam_bwd_dsc_kernel<1><<<grid, 256, 0, st>>>(g, x, sa, dpooled, amax, dsc, pkey, C, P);
  }
  SMAAT_LAUNCH_CHECK("smaat_cbam_bwd_dsc");
  return SMAAT_OK;
}

extern "C" int smaat_cbam_bwd_dx(const float* g, const float* sc, const float* sa, const float* dpooled, const int* amax,
                                 const float* davg, const float* dmx, const unsigned long long* pkey, float* dx, int B, int C, int P,
                                 void* stream) {
  SMAAT_REQUIRE(g && sc && sa && dpooled && amax && davg && dmx && pkey && dx && B > 0 && C > 0 && P > 0 && B <= 65535,
                "cbam_bwd_dx: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (cb_vec_ok(P, g, dx, sa, dpooled, amax)) {
    const dim3 grid(ceil_div(C, CB_CH), ceil_div(P, 256 * 4 * CB_IT), B);
    SMAAT_REQUIRE(grid.y <= 65535, "cbam_bwd_dx: plane too large");
    cbam_bwd_dx_kernel<4><<<grid, 256, 0, st>>>(g, sc, sa, dpooled, amax, davg, dmx, pkey, dx, C, P);
  } else {
    const dim3 grid(ceil_div(C, CB_CH), ceil_div(P, 256 * CB_IT), B);
    SMAAT_REQUIRE(grid.y <= 65535, "cbam_bwd_dx: plane too large");
    cbam_bwd_dx_kernel<1><<<grid, 256, 0, st>>>(g, sc, sa, dpooled, amax, davg, dmx, pkey, dx, C, P);
  }
  SMAAT_LAUNCH_CHECK("smaat_cbam_bwd_dx");
  return SMAAT_OK;
}

extern "C" int smaat_cbam_mlp_bwd(const float* avg, const float* mx, const float* w1, const float* b1, const float* w2,
                                  const float* sc, const float* dsc, float* dw1, float* db1, float* dw2, float* db2, float* davg,
                                  float* dmx, int B, int C, int hidden, void* stream) {
  SMAAT_REQUIRE(avg && mx && w1 && b1 && w2 && sc && dsc && dw1 && db1 && dw2 && db2 && davg && dmx && B > 0 && C > 0 && hidden > 0,
                "cbam_mlp_bwd: bad arguments");
  const size_t smem = (size_t)(3 * C + 4 * hidden) * sizeof(float);
  SMAAT_REQUIRE(smem <= 48 * 1024, "cbam_mlp_bwd: C too large");
  cbam_mlp_bwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(avg, mx, w1, b1, w2, sc, dsc, dw1, db1, dw2, db2, davg, dmx, C, hidden);
  SMAAT_LAUNCH_CHECK("smaat_cbam_mlp_bwd");
  return SMAAT_OK;
}
