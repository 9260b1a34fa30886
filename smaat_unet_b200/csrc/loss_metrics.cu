// loss_metrics.cu -- the training/validation step's loss and metric bookkeeping in ONE pass, no host sync.
//
// Replaces UNetBase.loss_func (reference models/regression_lightning.py:57-65: mse_loss(reduction="sum") / B)
// and PrecipitationMetrics.update (reference metric/precipitation_metrics.py:37-95): the NaN guard (:46, a
// host-syncing `.any()` in the reference), the normalised and de-normalised squared errors (:60-73), and the
// TN/FP/FN/TP confusion counts of `x * factor * 12 > threshold` (:80-93).  HBM-bound: 8 bytes per pixel read
// (+4 written when the loss gradient 2*(p-y)*grad_scale is requested in the same pass).
//
// batch_acc (double[8], zeroed by the call): [0] sum (p-y)^2   [1] sum (p*f - y*f)^2   [2] #NaN in p or y
//                                            [3] TN  [4] FP  [5] FN  [6] TP            [7] n
// smaat_metrics_commit folds one batch into the running totals ON THE DEVICE and skips the batch when [2] != 0,
// which is what the reference's early `return` does -- without the device->host round trip.
#include "common.cuh"

namespace smaat {

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct LmAcc {
  float sse, sse_d;          // per-thread fp32 partials over <= a few hundred elements, merged in fp64
  unsigned nan, cnt[4];
};

__device__ __forceinline__ void lm_one(float p, float y, float factor, float thr, bool denorm, LmAcc& a, float* g, float gs) {
  const float d = p - y;
  a.sse = fmaf(d, d, a.sse);
  // same fp32 operation order as the reference: (x * factor) * 12 > threshold
  const float pu = denorm ? __fmul_rn(p, factor) : p, yu = denorm ? __fmul_rn(y, factor) : y;
  const float dd = pu - yu;
  a.sse_d = fmaf(dd, dd, a.sse_d);
  a.nan += (unsigned)((p != p) | (y != y));
  const int pm = __fmul_rn(pu, 12.f) > thr, tm = __fmul_rn(yu, 12.f) > thr;
  a.cnt[tm * 2 + pm] += 1u;
  if (g) *g = 2.f * d * gs;
}

template <bool VEC>
__global__ void __launch_bounds__(256) mse_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                                          float factor, float thr, int denorm, double* __restrict__ acc,
                                                          float* __restrict__ dpred, float gs) {
  double tot[7] = {0, 0, 0, 0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  LmAcc a;
  a.sse = a.sse_d = 0.f; a.nan = 0u; a.cnt[0] = a.cnt[1] = a.cnt[2] = a.cnt[3] = 0u;
  if (VEC) {
    const int64_t n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(pred);
    const float4* y4 = reinterpret_cast<const float4*>(target);
    float4* g4 = reinterpret_cast<float4*>(dpred);
    int it = 0;
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 p = __ldg(p4 + i), y = __ldg(y4 + i);
      float4 g;
      lm_one(p.x, y.x, factor, thr, denorm, a, dpred ? &g.x : nullptr, gs);
      lm_one(p.y, y.y, factor, thr, denorm, a, dpred ? &g.y : nullptr, gs);
      lm_one(p.z, y.z, factor, thr, denorm, a, dpred ? &g.z : nullptr, gs);
      lm_one(p.w, y.w, factor, thr, denorm, a, dpred ? &g.w : nullptr, gs);
      if (dpred) g4[i] = g;
      if (++it == 64) {  // bound the fp32 partial sums to 256 terms
        tot[0] += a.sse; tot[1] += a.sse_d; a.sse = a.sse_d = 0.f; it = 0;
      }
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride)
      lm_one(__ldg(pred + i), __ldg(target + i), factor, thr, denorm, a, dpred ? dpred + i : nullptr, gs);
  } else {
    int it = 0;
    for (int64_t i = tid; i < n; i += stride) {
      lm_one(__ldg(pred + i), __ldg(target + i), factor, thr, denorm, a, dpred ? dpred + i : nullptr, gs);
      if (++it == 256) { tot[0] += a.sse; tot[1] += a.sse_d; a.sse = a.sse_d = 0.f; it = 0; }
    }
  }
  tot[0] += a.sse; tot[1] += a.sse_d; tot[2] = a.nan;
  tot[3] = a.cnt[0]; tot[4] = a.cnt[1]; tot[5] = a.cnt[2]; tot[6] = a.cnt[3];
  __shared__ double red[7][8];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const double v = warp_sum_f64(tot[q]);
    if (lane == 0) red[q][wp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0.0;
    for (int i = 0; i < 8; ++i) v += red[threadIdx.x][i];
    if (v != 0.0) atomicAdd(acc + threadIdx.x, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 7) acc[7] = (double)n;
}

// totals (double[9]): [0] total_loss  [1] total_loss_denorm  [2] total_samples  [3] total_pixels
//                     [4] TN [5] FP [6] FN [7] TP   [8] batches skipped by the NaN guard
__global__ void metrics_commit_kernel(const double* __restrict__ b, double* __restrict__ t, int batch_size, int denorm) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (b[2] != 0.0) { t[8] += 1.0; return; }
  t[0] += b[0] / batch_size;
  if (denorm) t[1] += b[1] / batch_size;
  t[2] += batch_size;
  t[3] += b[7];
  t[4] += b[3]; t[5] += b[4]; t[6] += b[5]; t[7] += b[6];
}

}  // namespace smaat

using namespace smaat;

extern "C" int smaat_mse_metrics_fwd(const float* pred, const float* target, int64_t n, float factor, float threshold,
                                     int denormalize, double* batch_acc, float* dpred, float grad_scale, void* stream) {
  SMAAT_REQUIRE(pred && target && batch_acc && n > 0, "mse_metrics: bad arguments (n=%lld)", (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(batch_acc, 0, 8 * sizeof(double), st);
  if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "mse_metrics: memset: %s", cudaGetErrorString(e));
  const bool vec = aligned16(pred) && aligned16(target) && (!dpred || aligned16(dpred));
  int64_t blocks = ceil_div64(n, 256 * 16);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (vec)
    mse_metrics_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(pred, target, n, factor, threshold, denormalize, batch_acc, dpred, grad_scale);
  else
    mse_metrics_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(pred, target, n, factor, threshold, denormalize, batch_acc, dpred, grad_scale);
  SMAAT_LAUNCH_CHECK("smaat_mse_metrics_fwd");
  return SMAAT_OK;
}

extern "C" int smaat_metrics_commit(const double* batch_acc, double* totals, int batch_size, int denormalize, void* stream) {
  SMAAT_REQUIRE(batch_acc && totals && batch_size > 0, "metrics_commit: bad arguments");
  metrics_commit_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(batch_acc, totals, batch_size, denormalize);
  SMAAT_LAUNCH_CHECK("smaat_metrics_commit");
  return SMAAT_OK;
}
