// bn_bwd.cu -- vectorised BatchNorm(+ReLU) backward passes (the streaming 2|x| reduce and 3|x| apply).
// One CTA per (plane, slice): 128-bit loads, no per-element index division; fp64 atomics merge the slices.
#include "common.cuh"

namespace smaat {

__global__ void __launch_bounds__(256) bn_act_bwd_reduce_v4(const float* __restrict__ dy, const float* __restrict__ z,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            double* __restrict__ sums, int C, int P4, int act) {
  const int plane = blockIdx.x, c = plane % C;
  const float s = scale ? __ldg(scale + c) : 1.f, t = shift ? __ldg(shift + c) : 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(dy) + (int64_t)plane * P4;
  const float4* z4 = reinterpret_cast<const float4*>(z) + (int64_t)plane * P4;
  float f1 = 0.f, f2 = 0.f;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P4; i += gridDim.y * blockDim.x) {
    const float4 zv = __ldg(z4 + i);
    float4 g = __ldg(g4 + i);
    if (act == 1) {
      if (!(fmaf(zv.x, s, t) > 0.f)) g.x = 0.f;
      if (!(fmaf(zv.y, s, t) > 0.f)) g.y = 0.f;
      if (!(fmaf(zv.z, s, t) > 0.f)) g.z = 0.f;
      if (!(fmaf(zv.w, s, t) > 0.f)) g.w = 0.f;
    }
    f1 += (g.x + g.y) + (g.z + g.w);
    f2 += (g.x * zv.x + g.y * zv.y) + (g.z * zv.z + g.w * zv.w);
  }
  __shared__ double r1[8], r2[8];
  const float w1 = warp_sum(f1), w2 = warp_sum(f2);
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = w1; r2[threadIdx.x >> 5] = w2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += r1[i]; b += r2[i]; }
    atomicAdd(sums + c, a);
    atomicAdd(sums + C + c, b);
  }
}

__global__ void __launch_bounds__(256) bn_act_bwd_apply_v4(const float* __restrict__ dy, const float* __restrict__ z,
                                                           const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ cc, float* __restrict__ dz, int C, int P4, int act) {
  const int plane = blockIdx.x, c = plane % C;
  const float s = scale ? __ldg(scale + c) : 1.f, t = shift ? __ldg(shift + c) : 0.f;
  const float av = __ldg(a + c), bv = __ldg(b + c), cv = __ldg(cc + c);
  const float4* g4 = reinterpret_cast<const float4*>(dy) + (int64_t)plane * P4;
  const float4* z4 = reinterpret_cast<const float4*>(z) + (int64_t)plane * P4;
  float4* o4 = reinterpret_cast<float4*>(dz) + (int64_t)plane * P4;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < P4; i += gridDim.y * blockDim.x) {
    const float4 zv = __ldg(z4 + i);
    float4 g = __ldg(g4 + i);
    if (act == 1) {
      if (!(fmaf(zv.x, s, t) > 0.f)) g.x = 0.f;
      if (!(fmaf(zv.y, s, t) > 0.f)) g.y = 0.f;
      if (!(fmaf(zv.z, s, t) > 0.f)) g.z = 0.f;
      if (!(fmaf(zv.w, s, t) > 0.f)) g.w = 0.f;
    }
    float4 o;
    o.x = fmaf(av, g.x, fmaf(bv, zv.x, cv));
    o.y = fmaf(av, g.y, fmaf(bv, zv.y, cv));
    o.z = fmaf(av, g.z, fmaf(bv, zv.z, cv));
    o.w = fmaf(av, g.w, fmaf(bv, zv.w, cv));
    o4[i] = o;
  }
}

bool bn_bwd_v4_ok(const void* dy, const void* z, const void* dz, int P) {
  return (P % 4 == 0) && aligned16(dy) && aligned16(z) && (dz == nullptr || aligned16(dz));
}

static dim3 plane_grid(int64_t planes, int P4) {
  int slices = ceil_div(P4, 256 * 8);                         // >= 8 float4 per thread
  const int64_t want = (int64_t)num_sms() * 16;
  if (planes * slices > want * 4) slices = (int)((want * 4 + planes - 1) / planes);
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  return dim3((unsigned)planes, (unsigned)slices);
}

int bn_act_bwd_reduce_v4_launch(const float* dy, const float* z, const float* scale, const float* shift, double* sums, int B, int C,
                                int P, int act, cudaStream_t st) {
  bn_act_bwd_reduce_v4<<<plane_grid((int64_t)B * C, P / 4), 256, 0, st>>>(dy, z, scale, shift, sums, C, P / 4, act);
  SMAAT_LAUNCH_CHECK("smaat_bn_act_bwd_reduce");
  return SMAAT_OK;
}

int bn_act_bwd_apply_v4_launch(const float* dy, const float* z, const float* scale, const float* shift, const float* a, const float* b,
                               const float* cc, float* dz, int B, int C, int P, int act, cudaStream_t st) {
  bn_act_bwd_apply_v4<<<plane_grid((int64_t)B * C, P / 4), 256, 0, st>>>(dy, z, scale, shift, a, b, cc, dz, C, P / 4, act);
  SMAAT_LAUNCH_CHECK("smaat_bn_act_bwd_apply");
  return SMAAT_OK;
}

}  // namespace smaat
