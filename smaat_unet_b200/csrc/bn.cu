// bn.cu -- train-mode BatchNorm2d pieces (reference parts_ds.py:25,34 and layers.py:120,127 in .train()):
//   * per-channel batch statistics are accumulated in fp64 by the producing kernel's epilogue
//     (pw1x1 / dsconv `stats`) or by channel_stats_kernel;
//   * bn_finalize turns them into the (scale, shift) the consumer applies, saves mean / inv-std for
//     the backward pass and updates running_mean / running_var exactly like torch
//     (momentum 0.1, UNBIASED variance into running_var, biased variance for normalisation);
//   * affine_act applies y = act(scale[c] * x + shift[c]) (ReLU for the block output, sigmoid for the
//     spatial gate) -- streaming, 128-bit.
#include "common.cuh"

namespace smaat {

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   long long* __restrict__ num_batches_tracked, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (c >= C) return;
  const double mean = stats[c] / count;
  double var = stats[C + c] / count - mean * mean;  // biased
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  const float s = (float)((double)g * invstd);
  scale[c] = s;
  shift[c] = (float)((double)bt - mean * (double)g * invstd);
  if (mean_out) mean_out[c] = (float)mean;
  if (invstd_out) invstd_out[c] = (float)invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, float* __restrict__ y, int C, int P,
                                                         int act) {
  const int plane = blockIdx.x;  // b * C + c
  const int c = plane % C;
  const float s = scale ? __ldg(scale + c) : 1.f, t = shift ? __ldg(shift + c) : 0.f;
  const float* xp = x + (int64_t)plane * P;
  float* yp = y + (int64_t)plane * P;
  const int i4 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (i4 >= P) return;
  float v[4];
  if (VEC) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(xp + i4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (i4 + j < P) ? __ldg(xp + i4 + j) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float r = fmaf(v[j], s, t);
    if (act == 1) r = fmaxf(r, 0.f);
    else if (act == 2) r = 1.f / (1.f + expf(-r));
    v[j] = r;
  }
  if (VEC) {
    *reinterpret_cast<float4*>(yp + i4) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i4 + j < P) yp[i4 + j] = v[j];
  }
}

// per-channel sum / sum of squares over (B, P) of x[B][C][P]  (+= into fp64 accumulators)
__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int B, int C,
                                                            int P, int chunks) {
  const int c = blockIdx.y;
  const int64_t n = (int64_t)B * P;
  const int64_t per = (n + chunks - 1) / chunks;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n, lo + per);
  double s1 = 0.0, s2 = 0.0;
  float f1 = 0.f, f2 = 0.f;
  int cnt = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int64_t b = i / P, pp = i - b * P;
    const float v = __ldg(x + (b * C + c) * (int64_t)P + pp);
    f1 += v;
    f2 = fmaf(v, v, f2);
    if (++cnt == 64) {  // flush fp32 partials into fp64 regularly
      s1 += f1; s2 += f2; f1 = f2 = 0.f; cnt = 0;
    }
  }
  s1 += f1;
  s2 += f2;
  __shared__ double r1[256], r2[256];
  r1[threadIdx.x] = s1;
  r2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      r1[threadIdx.x] += r1[threadIdx.x + o];
      r2[threadIdx.x] += r2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(stats + c, r1[0]);
    atomicAdd(stats + C + c, r2[0]);
  }
}

}  // namespace smaat

using namespace smaat;

extern "C" int smaat_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float eps,
                                 float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                                 float* mean_out, float* invstd_out, long long* num_batches_tracked, int C, void* stream) {
  SMAAT_REQUIRE(stats && scale && shift && C > 0 && count > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(stats, count, gamma, beta, eps, momentum, running_mean,
                                                                         running_var, scale, shift, mean_out, invstd_out,
                                                                         num_batches_tracked, C);
  SMAAT_LAUNCH_CHECK("smaat_bn_finalize");
  return SMAAT_OK;
}

extern "C" int smaat_affine_act_fwd(const float* x, const float* scale, const float* shift, float* y, int B, int C, int P,
                                    int act, void* stream) {
  SMAAT_REQUIRE(x && y && B > 0 && C > 0 && P > 0 && act >= 0 && act <= 2, "affine_act: bad arguments");
  SMAAT_REQUIRE(ceil_div(ceil_div(P, 4), 256) <= 65535, "affine_act: plane too large for grid.y");
  const bool vec = (P % 4 == 0) && aligned16(x) && aligned16(y);
  dim3 grid(B * C, ceil_div(ceil_div(P, 4), 256));
  if (vec) affine_act_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, scale, shift, y, C, P, act);
  else affine_act_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, scale, shift, y, C, P, act);
  SMAAT_LAUNCH_CHECK("smaat_affine_act_fwd");
  return SMAAT_OK;
}

extern "C" int smaat_channel_stats(const float* x, double* stats, int B, int C, int P, void* stream) {
  SMAAT_REQUIRE(x && stats && B > 0 && C > 0 && P > 0 && C <= 65535, "channel_stats: bad arguments");
  const int64_t n = (int64_t)B * P;
  int chunks = (int)ceil_div64(n, 256 * 64);
  const int maxc = ceil_div(num_sms() * 8, C);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, C);
  channel_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, stats, B, C, P, chunks);
  SMAAT_LAUNCH_CHECK("smaat_channel_stats");
  return SMAAT_OK;
}
