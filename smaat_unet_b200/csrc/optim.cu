// optim.cu -- the optimizer step of the reference's training loop as ONE kernel over a flat parameter bucket.
//
// Replaces torch.optim.Adam(self.parameters(), lr) (reference models/regression_lightning.py:47-48; train_SmaAtUNet.py:25:
// Adam defaults betas = (0.9, 0.999), eps = 1e-8, weight_decay = 0, amsgrad = False) as TrainSession uses it: parameters,
// gradients and both moments live in four flat fp32 buffers with identical layout (train.py), so a step is one
// streaming pass (16 bytes read + 12 written per parameter) instead of ~20 multi-tensor launches over 214 tensors plus a
// gradient gather.  The learning rate and the step count are DEVICE scalars: a CUDA graph that captured the step follows
// later changes of the learning rate (ReduceLROnPlateau, regression_lightning.py:49-55) without re-capture.
//
// Arithmetic = torch's single-tensor Adam, operation for operation:
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g g;  denom = sqrt(v) / sqrt(1 - b2^t) + eps;  p -= (lr / (1 - b1^t)) m / denom
#include "common.cuh"

namespace smaat {

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n4, const float* __restrict__ lr_ptr,
                                                        const float* __restrict__ step_ptr, double b1d, double b2d, float omb1, float omb2,
                                                        float eps) {
  const float b1 = (float)b1d, b2 = (float)b2d;
  const double t = (double)__ldg(step_ptr) + 1.0;             // this step's number (the counter holds completed steps)
  const float bc1 = (float)(1.0 - pow(b1d, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(b2d, t));
  const float step_size = __ldg(lr_ptr) / bc1;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    float* pe = &pp.x; float* me = &mm.x; float* ve = &vv.x;
    const float* ge = &gg.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      me[e] = fmaf(ge[e] - me[e], omb1, me[e]);                // exp_avg.lerp_(grad, 1 - beta1)
      ve[e] = b2 * ve[e] + omb2 * ge[e] * ge[e];
      const float denom = sqrtf(ve[e]) / bc2_sqrt + eps;
      pe[e] -= step_size * (me[e] / denom);
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
}

__global__ void adam_advance_kernel(float* step_ptr) { *step_ptr += 1.f; }

}  // namespace smaat

using namespace smaat;

/* One Adam step over flat buffers of n floats (n % 4 == 0, 16-byte aligned; padding elements must hold zero gradients).
 * lr, step: device scalars (fp32; step = number of completed steps, incremented here). */
extern "C" int smaat_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                               float* step, double beta1, double beta2, double eps, void* stream) {
  SMAAT_REQUIRE(params && grads && exp_avg && exp_avg_sq && lr && step, "adam_step: null pointer");
  SMAAT_REQUIRE(n > 0 && n % 4 == 0, "adam_step: n must be a positive multiple of 4 (pad the bucket)");
  SMAAT_REQUIRE(aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq), "adam_step: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  const int grid = (int)(ceil_div64(n4, 256) < (int64_t)num_sms() * 8 ? ceil_div64(n4, 256) : (int64_t)num_sms() * 8);
  // torch evaluates 1 - beta in double and rounds once: (float)(1 - 0.999) != 1.f - 0.999f
  adam_flat_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n4, lr, step, beta1, beta2, (float)(1.0 - beta1),
                                         (float)(1.0 - beta2), (float)eps);
  SMAAT_LAUNCH_CHECK("smaat_adam_step");
  adam_advance_kernel<<<1, 1, 0, st>>>(step);
  SMAAT_LAUNCH_CHECK("smaat_adam_step(advance)");
  return SMAAT_OK;
}
