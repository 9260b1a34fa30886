// cbam_bwd.cu -- CBAM backward (reference models/layers.py:90-141 differentiated).
//
// Forward (cbam.cu): sc = sigmoid(MLP(avg_p x) + MLP(max_p x)); u = x*sc; pooled = [mean_c u, max_c u];
// raw = conv_kxk(pooled); sa = sigmoid(BN1(raw)); out = u*sa.   Given g = dL/dout:
//   gate_in : d_pre[b,p] = (sum_c g*u) * sa*(1-sa)                       (then BN(1) backward -> d_raw, bn.cu/backward.cu)
//   conv    : d_pooled = conv_transpose(d_raw, w);  dW = corr(pooled, d_raw)
//   main    : d_u = g*sa + d_pooled[0]/C + [c == argmax_c u] d_pooled[1];  dx = d_u*sc;  d_sc[b,c] = sum_p d_u*x
//   mlp     : tiny per-image MLP backward -> dW1, db1, dW2, db2, d_avg, d_max
//   pool    : dx += d_avg/P + [p == argmax_p x] d_max
#include "common.cuh"

namespace smaat {

// ---- d_pre[b,p] = (sum_c g*x*sc) * sa*(1-sa) ------------------------------------------------------
__global__ void __launch_bounds__(256) cbam_bwd_gate_in_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                               const float* __restrict__ sc, const float* __restrict__ sa,
                                                               float* __restrict__ dpre, int C, int P) {
  __shared__ float rs[8][33];
  const int tx = threadIdx.x, cg = threadIdx.y, b = blockIdx.y;
  const int pp = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (pp < P) {
    const float* gb = g + (int64_t)b * C * P + pp;
    const float* xb = x + (int64_t)b * C * P + pp;
    for (int c = cg; c < C; c += 8) s = fmaf(__ldg(gb + (int64_t)c * P) * __ldg(xb + (int64_t)c * P), __ldg(sc + (int64_t)b * C + c), s);
  }
  rs[cg][tx] = s;
  __syncthreads();
  if (cg == 0 && pp < P) {
    for (int i = 1; i < 8; ++i) s += rs[i][tx];
    const float a = __ldg(sa + (int64_t)b * P + pp);
    dpre[(int64_t)b * P + pp] = s * a * (1.f - a);
  }
}

// ---- spatial conv backward ------------------------------------------------------------------------------
template <int KS>
__global__ void __launch_bounds__(256) cbam_gate_bwd_input_kernel(const float* __restrict__ draw, const float* __restrict__ wsp,
                                                                  float* __restrict__ dpooled, int H, int W) {
  constexpr int R = KS / 2;
  const int b = blockIdx.z, ch = blockIdx.y;
  const int P = H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
      const int yy = y - dy + R;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        const int xx = x - dx + R;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(__ldg(wsp + (ch * KS + dy) * KS + dx), __ldg(draw + (int64_t)b * P + (int64_t)yy * W + xx), acc);
      }
    }
    dpooled[((int64_t)b * 2 + ch) * P + i] = acc;
  }
}

// dW[ch][dy][dx] += sum_{b,y,x} draw[b,y,x] * pooled[b,ch,y+dy-R,x+dx-R]   (grid: chunks x (2*KS*KS))
__global__ void __launch_bounds__(256) cbam_gate_bwd_weight_kernel(const float* __restrict__ draw, const float* __restrict__ pooled,
                                                                   float* __restrict__ dW, int B, int H, int W, int KS, int chunks) {
  const int tap = blockIdx.y;
  const int ch = tap / (KS * KS), r = tap - ch * KS * KS;
  const int dy = r / KS - KS / 2, dx = r % KS - KS / 2;
  const int P = H * W;
  const int64_t n = (int64_t)B * P;
  const int64_t per = (n + chunks - 1) / chunks;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
  float acc = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int64_t b = i / P;
    const int pp = (int)(i - b * P);
    const int y = pp / W + dy, x = pp % W + dx;
    if (y >= 0 && y < H && x >= 0 && x < W)
      acc = fmaf(__ldg(draw + i), __ldg(pooled + (b * 2 + ch) * (int64_t)P + (int64_t)y * W + x), acc);
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[i];
    atomicAdd(dW + tap, v);
  }
}

// ---- main pass: dx = d_u*sc, d_sc[b,c] += sum_p d_u*x -------------------------------------------------------
__global__ void __launch_bounds__(256) cbam_bwd_main_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                            const float* __restrict__ sc, const float* __restrict__ sa,
                                                            const float* __restrict__ dpooled, float* __restrict__ dx,
                                                            float* __restrict__ dsc, int C, int P) {
  __shared__ float mv[8][33];
  __shared__ int mi[8][33];
  const int tx = threadIdx.x, cg = threadIdx.y, b = blockIdx.y;
  const int pp = blockIdx.x * 32 + tx;
  const bool pv = pp < P;
  const float* xb = x + (int64_t)b * C * P + pp;
  const float* scb = sc + (int64_t)b * C;
  // pass 1: argmax over channels of u = x*sc (first maximum, like torch.max(dim=1))
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (pv)
    for (int c = cg; c < C; c += 8) {
      const float u = __ldg(xb + (int64_t)c * P) * __ldg(scb + c);
      if (u > best) { best = u; bi = c; }
    }
  mv[cg][tx] = best;
  mi[cg][tx] = bi;
  __syncthreads();
  int amax = 0;
  {
    float bb = mv[0][tx];
    amax = mi[0][tx];
    for (int i = 1; i < 8; ++i) {
      const float v = mv[i][tx];
      const int id = mi[i][tx];
      if (v > bb || (v == bb && id < amax)) { bb = v; amax = id; }
    }
  }
  // pass 2
  const float av = pv ? __ldg(sa + (int64_t)b * P + pp) : 0.f;
  const float dp0 = pv ? __ldg(dpooled + ((int64_t)b * 2) * P + pp) / (float)C : 0.f;
  const float dp1 = pv ? __ldg(dpooled + ((int64_t)b * 2 + 1) * P + pp) : 0.f;
  const float* gb = g + (int64_t)b * C * P + pp;
  float* dxb = dx + (int64_t)b * C * P + pp;
  for (int c = cg; c < C; c += 8) {   // warp-uniform trip count: the 32 lanes of a warp share cg
    float du = 0.f, xv = 0.f;
    const float s = __ldg(scb + c);
    if (pv) {
      xv = __ldg(xb + (int64_t)c * P);
      du = fmaf(__ldg(gb + (int64_t)c * P), av, dp0) + (c == amax ? dp1 : 0.f);
      dxb[(int64_t)c * P] = du * s;
    }
    const float part = warp_sum(du * xv);
    if (tx == 0) atomicAdd(dsc + (int64_t)b * C + c, part);
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

// ---- dx[n,p] += d_avg[n]/P + [p == argmax_p x[n,:]] d_max[n]   (one CTA per plane) ----------------------------
__global__ void __launch_bounds__(256) cbam_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ davg,
                                                            const float* __restrict__ dmx, float* __restrict__ dx, int P) {
  const int64_t n = blockIdx.x;
  const float* xp = x + n * P;
  float* dp = dx + n * P;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float v = __ldg(xp + i);
    if (v > best) { best = v; bi = i; }
  }
  __shared__ float sv[256];
  __shared__ int si[256];
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v = sv[threadIdx.x + o];
      const int id = si[threadIdx.x + o];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && id < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = id; }
    }
    __syncthreads();
  }
  const int amax = si[0];
  const float ga = __ldg(davg + n) / (float)P, gm = __ldg(dmx + n);
  for (int i = threadIdx.x; i < P; i += blockDim.x) dp[i] += ga + (i == amax ? gm : 0.f);
}

}  // namespace smaat

using namespace smaat;

extern "C" int smaat_cbam_bwd_gate_in(const float* g, const float* x, const float* sc, const float* sa, float* dpre, int B, int C,
                                      int P, void* stream) {
  SMAAT_REQUIRE(g && x && sc && sa && dpre && B > 0 && C > 0 && P > 0 && B <= 65535, "cbam_bwd_gate_in: bad arguments");
  cbam_bwd_gate_in_kernel<<<dim3(ceil_div(P, 32), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(g, x, sc, sa, dpre, C, P);
  SMAAT_LAUNCH_CHECK("smaat_cbam_bwd_gate_in");
  return SMAAT_OK;
}

extern "C" int smaat_cbam_gate_bwd(const float* draw, const float* pooled, const float* wsp, float* dpooled, float* dW, int B, int H,
                                   int W, int ks, void* stream) {
  SMAAT_REQUIRE(draw && pooled && wsp && dpooled && dW && B > 0 && H > 0 && W > 0, "cbam_gate_bwd: bad arguments");
  SMAAT_REQUIRE(ks == 3 || ks == 7, "cbam_gate_bwd: kernel size must be 3 or 7");
  SMAAT_REQUIRE(B <= 65535, "cbam_gate_bwd: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t bl = ceil_div64((int64_t)H * W, 256);
  const unsigned gx = (unsigned)(bl < 1024 ? bl : 1024);
  if (ks == 7) cbam_gate_bwd_input_kernel<7><<<dim3(gx, 2, B), 256, 0, st>>>(draw, wsp, dpooled, H, W);
  else cbam_gate_bwd_input_kernel<3><<<dim3(gx, 2, B), 256, 0, st>>>(draw, wsp, dpooled, H, W);
  SMAAT_LAUNCH_CHECK("smaat_cbam_gate_bwd(input)");
  int chunks = (int)ceil_div64((int64_t)B * H * W, 256 * 32);
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  cbam_gate_bwd_weight_kernel<<<dim3(chunks, 2 * ks * ks), 256, 0, st>>>(draw, pooled, dW, B, H, W, ks, chunks);
  SMAAT_LAUNCH_CHECK("smaat_cbam_gate_bwd(weight)");
  return SMAAT_OK;
}

extern "C" int smaat_cbam_bwd_main(const float* g, const float* x, const float* sc, const float* sa, const float* dpooled, float* dx,
                                   float* dsc, int B, int C, int P, void* stream) {
  SMAAT_REQUIRE(g && x && sc && sa && dpooled && dx && dsc && B > 0 && C > 0 && P > 0 && B <= 65535, "cbam_bwd_main: bad arguments");
  cbam_bwd_main_kernel<<<dim3(ceil_div(P, 32), B), dim3(32, 8), 0, (cudaStream_t)stream>>>(g, x, sc, sa, dpooled, dx, dsc, C, P);
  SMAAT_LAUNCH_CHECK("smaat_cbam_bwd_main");
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

extern "C" int smaat_cbam_pool_bwd(const float* x, const float* davg, const float* dmx, float* dx, int64_t N, int P, void* stream) {
  SMAAT_REQUIRE(x && davg && dmx && dx && N > 0 && P > 0 && N < (1ll << 31), "cbam_pool_bwd: bad arguments");
  cbam_pool_bwd_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(x, davg, dmx, dx, P);
  SMAAT_LAUNCH_CHECK("smaat_cbam_pool_bwd");
  return SMAAT_OK;
}
