// dsconv_fused.cu -- DepthwiseSeparableConv forward as ONE kernel: depthwise 3x3 on the CUDA
// cores feeding the pointwise 1x1 on the tensor cores, with the BN-affine/ReLU epilogue.
//
// Replaces DepthwiseSeparableConv.forward (reference models/layers.py:47-50: depthwise then
// pointwise, nothing in between) + eval BatchNorm2d + ReLU (parts_ds.py:25-26,34-35).  The
// k x -expanded depthwise result never reaches HBM: per B=32 forward the DS blocks move
// 4*B*S^2*(Cin + Cout) bytes instead of 4*B*S^2*(Cin + 2*k*Cin + Cout) (32.7 -> 10.4 GB).
//
// One persistent CTA per SM; a tile is a PH x PW = 128-pixel patch of one image (M = 128 TMEM
// lanes) and ALL Cout <= 128 output channels (N_TILE TMEM columns), so the depthwise work is done
// exactly once.  K = k*Cin is walked in chunks of 32 depthwise channels (CC = 32/k input channels):
//   warp 0      TMA: (PH+2) x (PW+8) x CC input halo box per chunk (OOB zero fill = padding=1; box
//               starts at x0-4: the inner TMA coordinate must be 16-byte aligned) into an IS-deep ring;
//               input may be the virtual concat [x0, x1] of UpDS (parts_ds.py:85)
//   warps 6-17  three depthwise producer groups (128 threads each; group g takes every third chunk): 3x3 stencil
//               from the staged tile with a sliding register window (one LDS.128 per row, edge columns from the
//               neighbouring quads by shuffle), then write the result straight into the UMMA A-operand layout
//               (MN-major tf32, 128B span / 32B-atom swizzle) -- as hi and lo tf32 parts in TF32X3 mode (the split
//               is free here: values are in registers) -- into a 3/4-stage A ring
//   warp 18     prefetches the weight chunks (K-major SW128, [hi rows | lo rows]) into their own ring
//   warp 1      MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma kind::tf32 into TMEM (TF32X3: a wide
//               A_hi x [B_hi | B_lo] MMA + A_lo x B_hi per k-step) and commits
//   warps 2-5   epilogue: tcgen05.ld (lane = pixel, 16 columns per step) -> scale/shift/ReLU -> coalesced NCHW stores
//               (or the fused 1-class OutConv dot product; or BatchNorm batch statistics from fragment-shaped
//               reads); two TMEM accumulator stages overlap it with the next tile's MMAs
// Debug: SMAAT_DSCONV_TIMING=1 makes CTA 0 record per-stage cycle counters (smaat_debug_dsconv_timing).
#include <stdlib.h>

#include "tc_common.cuh"

namespace smaat {

// Stage timing of CTA 0 (clock64 cycles, accumulated over launches until read): see smaat_debug_dsconv_timing.
//  [0] producer group 0: wait input box   [1] wait free A stage   [2] stencil + A-operand writes   [3] chunks
//  [4] MMA lane: wait A   [5] wait B   [6] wait free accumulator   [7] issue   [8] chunks
//  [9] epilogue warp 2: wait accumulator  [10] drain + store   [11] tiles      [12] kernel cycles
__device__ unsigned long long g_ds_timing[16];

struct DsParams {
  const float* dw_w;
  const float* dw_b;
  const float* scale;
  const float* shift;
  float* y;
  int64_t y_bstride;
  double* stats;
  const float* oc_w;   // fused OutConv (1 class): logits = sum_c oc_w[c] * act[c] + oc_b, written instead of y
  const float* oc_b;
  float* oc_y;
  int C0, C1, H, W, Cout, relu, K;
  int tiles_x, tiles_y, total_tiles, nchunks;
  int timing;          // SMAAT_DSCONV_TIMING=1: CTA 0 records stage timers (debug)
};

template <int N_TILE, int KPL, int PW, bool X3>
struct DsCfg {
  static constexpr int PH = TC_BM / PW;
  static constexpr int BW = PW + 8, BH = PH + 2;
  static constexpr int CC = TC_BK / KPL;                       // input channels per chunk
  static constexpr int IN_BYTES = CC * BH * BW * 4;            // multiple of 128 for PW in {16,32}
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;            // 16 KB
  static constexpr int B_BYTES = N_TILE * TC_BK * 4;
  static constexpr int AST_BYTES = (X3 ? 2 : 1) * A_BYTES;     // A ring stage: hi [+ lo]
  static constexpr int BST_BYTES = (X3 ? 2 : 1) * B_BYTES;     // B ring stage: hi [+ lo]
  static constexpr int OFF_ALO = A_BYTES;
  static constexpr int OFF_BLO = B_BYTES;
  // A ring: 3 stages in TF32X3 (32 KB each; the input ring must stay deep enough to cover HBM latency), 4 in TF32
  static constexpr int AS = X3 ? ((KPL == 1 && N_TILE > 64) ? 2 : 3) : 4;
  // depthwise producer groups (128 threads each).  NG <= AS always: the per-stage a_empty barriers are tested by phase parity,
  // which is only unambiguous while a group can never be two hand-backs of a stage behind (see csrc/dsconv_tmem.cu)
  static constexpr int NG = AS < 3 ? 2 : 3;
  static constexpr int BS = X3 ? 2 : 4;                         // weight ring, prefetched by its own warp
  static constexpr int IS_FIT = (218 * 1024 - AS * AST_BYTES - BS * BST_BYTES) / IN_BYTES;
  static constexpr int IS = IS_FIT > 8 ? 8 : IS_FIT;           // input ring: as deep as shared memory allows
  static constexpr int OFF_A = ((IS * IN_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_BR = OFF_A + AS * AST_BYTES;
  static constexpr int OFF_BAR = OFF_BR + BS * BST_BYTES;
  static constexpr int BAR_BYTES = 512;
  static_assert((IS * NG + IS + 2 * AS + 2 * BS + 4) * 8 + 8 <= BAR_BYTES, "barrier block");
  static constexpr int AFF_N = 128;
  static constexpr int TOTAL = OFF_BAR + BAR_BYTES + 3 * AFF_N * 4 + 1024;   // scale | shift | OutConv weights
  static constexpr uint32_t B_TX = BST_BYTES;
  static constexpr int LOADER_WARP = 6 + 4 * NG;               // weight-ring loader
  static constexpr int THREADS = 64 + 128 + 128 * NG + 32;     // TMA, MMA | 4 epilogue warps | producers | loader
  static_assert(IS >= 2, "input ring");
  // TF32X3: the weight stage holds [hi rows | lo rows] contiguously, so ONE N = 2*N_TILE MMA computes A_hi*[B_hi | B_lo]
  // into 2*N_TILE accumulator columns and a second N = N_TILE MMA adds A_lo*B_hi to the first half: 2 instead of 3 MMAs
  // per k-step (each MMA re-reads its 4 KB A slice from shared memory whatever N is); the epilogue adds the two halves.
  static constexpr int ACC_COLS = X3 ? 2 * N_TILE : N_TILE;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;               // two accumulator stages
  static_assert(IN_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
  static_assert(N_TILE <= AFF_N, "epilogue affine staging");
};

template <int N_TILE, int KPL, int PW, bool X3>
__global__ void __launch_bounds__(DsCfg<N_TILE, KPL, PW, X3>::THREADS, 1)
    dsconv_fused_kernel(const __grid_constant__ CUtensorMap map_in0, const __grid_constant__ CUtensorMap map_in1,
                        const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo,
                        const DsParams p) {
  using L = DsCfg<N_TILE, KPL, PW, X3>;
  constexpr int PH = L::PH, BW = L::BW, BH = L::BH, CC = L::CC, IS = L::IS, AS = L::AS, BS = L::BS;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* a_base = smem + L::OFF_A;
  unsigned char* b_base = smem + L::OFF_BR;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  // [IS][NG] TMA input box landed, one barrier per (stage, group that reads the fill): a group meets a stage only at every
  // NG-th of its fills (unless NG divides IS), and TMA loads may complete out of order -- a parity test on a per-stage
  // barrier could then be satisfied by the wrong fill (csrc/dsconv_tmem.cu has the failure this caused there)
  uint64_t* in_full = bars;
  uint64_t* in_empty = in_full + IS * L::NG;      // [IS] producer group finished reading the box (128 arrivals)
  uint64_t* a_full = in_empty + IS;               // [AS] A operand written (128 arrivals)
  uint64_t* a_empty = a_full + AS;                // [AS] MMAs reading the A stage retired (commit)
  uint64_t* b_full = a_empty + AS;                // [BS] weight chunk landed (TMA tx)
  uint64_t* b_empty = b_full + BS;                // [BS] MMAs reading the B stage retired (commit)
  uint64_t* tmem_full = b_empty + BS;             // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]  (128 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* aff = reinterpret_cast<float*>(smem + L::OFF_BAR + L::BAR_BYTES);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  const int nch = p.nchunks;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_in0);
    tma_prefetch_desc(&map_in1);
    tma_prefetch_desc(&map_w);
    if (X3) tma_prefetch_desc(&map_wlo);
    for (int s = 0; s < IS; ++s) {
      for (int g = 0; g < L::NG; ++g) mbar_init(&in_full[s * L::NG + g], 1);
      mbar_init(&in_empty[s], 128);
    }
    for (int s = 0; s < AS; ++s) {
      mbar_init(&a_full[s], 128);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < BS; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, L::TMEM_COLS);
  for (int c = threadIdx.x; c < L::AFF_N; c += blockDim.x) {
    aff[c] = (c < p.Cout && p.scale) ? __ldg(p.scale + c) : 1.f;
    aff[L::AFF_N + c] = (c < p.Cout && p.shift) ? __ldg(p.shift + c) : 0.f;
    aff[2 * L::AFF_N + c] = (c < p.Cout && p.oc_w) ? __ldg(p.oc_w + c) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const long long t_kernel0 = ((p.timing & 1) && blockIdx.x == 0 && threadIdx.x == 0) ? clock64() : 0;

  if (warp == 0) {
    // ===== TMA: input halo boxes, running ahead through the IS-deep ring =====
    if (lane == 0) {
      uint32_t gc = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int t2 = tile - b * tiles_per_img;
        const int ty = t2 / p.tiles_x, tx = t2 - ty * p.tiles_x;
        const int x0 = tx * PW, y0 = ty * PH;
        for (int i = 0; i < nch; ++i, ++gc) {
          const int s = gc % IS;
          mbar_wait(&in_empty[s], ((gc / IS) & 1u) ^ 1u);
          uint64_t* full = &in_full[s * L::NG + (int)(gc % (uint32_t)L::NG)];      // the barrier of the group that reads this chunk
          mbar_arrive_expect_tx(full, L::IN_BYTES);
          const int cb = i * CC;
          const CUtensorMap* m = (cb < p.C0) ? &map_in0 : &map_in1;
          const int cc = (cb < p.C0) ? cb : cb - p.C0;
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                  "r"(smem_u32(smem + s * L::IN_BYTES)),
              "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(full)), "r"(x0 - 4), "r"(y0 - 1), "r"(cc), "r"(b)
              : "memory");
          // pull the same chunk of this CTA's NEXT tile into L2 now: by the time it is TMA-loaded the HBM
          // latency is already paid, so a few 15 KB boxes in flight per SM are enough to stream at HBM speed
          const int ntile = tile + gridDim.x;
          if (ntile < p.total_tiles) {
            const int nb = ntile / tiles_per_img;
            const int nt2 = ntile - nb * tiles_per_img;
            const int nty = nt2 / p.tiles_x, ntx = nt2 - nty * p.tiles_x;
            asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                             reinterpret_cast<uint64_t>(m)),
                         "r"(ntx * PW - 4), "r"(nty * PH - 1), "r"(cc), "r"(nb)
                         : "memory");
          }
        }
      }
    }
  } else if (warp == L::LOADER_WARP) {
    // ===== weight-ring loader: K-major SW128 chunks (hi [+lo]), decoupled from the input ring =====
    if (lane == 0) {
      uint32_t gc = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int i = 0; i < nch; ++i, ++gc) {
          const int sb = gc % BS;
          // rings of equal depth advance in lockstep: one commit (a_empty) releases both the A and the weight stage
          mbar_wait((AS == BS) ? &a_empty[sb] : &b_empty[sb], ((gc / BS) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&b_full[sb], L::B_TX);
          tma_load_2d(b_base + sb * L::BST_BYTES, &map_w, &b_full[sb], i * TC_BK, 0);
          if (X3) tma_load_2d(b_base + sb * L::BST_BYTES + L::OFF_BLO, &map_wlo, &b_full[sb], i * TC_BK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the loop (warp-uniform control flow and descriptors, which the compiler keeps
    // in uniform registers), one elected lane issues.  Descriptors are built once per chunk and advanced by constant adds:
    // the issuing thread's own instruction stream was the limit (~30 SASS instructions, ~115 cycles per MMA before). =====
    constexpr uint32_t idesc = make_idesc_tf32(N_TILE);
    constexpr uint32_t idesc_wide = make_idesc_tf32(X3 ? 2 * N_TILE : N_TILE);
    const bool rec = (p.timing & 1) && (blockIdx.x == 0) && (lane == 0);
    uint32_t gc = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1u;
      const long long te0 = rec ? clock64() : 0;
      mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1u) ^ 1u);
      if (rec) atomicAdd(&g_ds_timing[6], (unsigned long long)(clock64() - te0));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * L::ACC_COLS;
      for (int i = 0; i < nch; ++i, ++gc) {
        const int sa = gc % AS, sb = gc % BS;
        long long tk0 = 0, tk1 = 0, tk2 = 0;
        if (rec) tk0 = clock64();
        mbar_wait(&a_full[sa], (gc / AS) & 1u);
        if (rec) tk1 = clock64();
        mbar_wait(&b_full[sb], (gc / BS) & 1u);
        if (rec) tk2 = clock64();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(a_base + sa * L::AST_BYTES);
          const uint32_t b_addr = smem_u32(b_base + sb * L::BST_BYTES);
          // k-step kk: A advances 8 k-rows = 1 KB, the K-major weights 8 tf32 = 32 B (descriptor address field = bytes >> 4)
          const uint64_t ad0 = make_a_desc(a_addr, TC_BK * 128);
          const uint64_t al0 = make_a_desc(a_addr + L::OFF_ALO, TC_BK * 128);
          const uint64_t bd0 = make_b_desc(b_addr);
          const int kc = min(TC_BK, p.K - i * TC_BK);
          if (kc == TC_BK) {
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk) {
              const uint32_t first = (kk > 0) ? 1u : (i > 0 ? 1u : 0u);
              if (X3) {
                // D[:, 0:N) += A_hi*B_hi and D[:, N:2N) += A_hi*B_lo in one wide MMA over the contiguous hi|lo weight rows,
                // then D[:, 0:N) += A_lo*B_hi
                umma_tf32(d_tmem, ad0 + (uint64_t)(kk * 64), bd0 + (uint64_t)(kk * 2), idesc_wide, first);
                umma_tf32(d_tmem, al0 + (uint64_t)(kk * 64), bd0 + (uint64_t)(kk * 2), idesc, 1u);
              } else {
                umma_tf32(d_tmem, ad0 + (uint64_t)(kk * 64), bd0 + (uint64_t)(kk * 2), idesc, first);
              }
            }
          } else {
            const int nmma = (kc + 7) >> 3;
            for (int kk = 0; kk < nmma; ++kk) {
              const uint32_t first = (i > 0 || kk > 0) ? 1u : 0u;
              if (X3) {
                umma_tf32(d_tmem, ad0 + (uint64_t)(kk * 64), bd0 + (uint64_t)(kk * 2), idesc_wide, first);
                umma_tf32(d_tmem, al0 + (uint64_t)(kk * 64), bd0 + (uint64_t)(kk * 2), idesc, 1u);
              } else {
                umma_tf32(d_tmem, ad0 + (uint64_t)(kk * 64), bd0 + (uint64_t)(kk * 2), idesc, first);
              }
            }
          }
          umma_commit(&a_empty[sa]);  // each commit tracks completion of all MMAs issued so far
          if (AS != BS) umma_commit(&b_empty[sb]);
          if (i == nch - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (rec) {
          const long long tk3 = clock64();
          atomicAdd(&g_ds_timing[4], (unsigned long long)(tk1 - tk0));
          atomicAdd(&g_ds_timing[5], (unsigned long long)(tk2 - tk1));
          atomicAdd(&g_ds_timing[7], (unsigned long long)(tk3 - tk2));
          atomicAdd(&g_ds_timing[8], 1ull);
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue warps 2..5 =====
    const int q = warp & 3;
    const float act_lo = p.relu ? 0.f : -INFINITY;
    const int m = q * 32 + lane;  // pixel of the patch = TMEM lane
    const int pr = m / PW, pc = m % PW;
    const int64_t P = (int64_t)p.H * p.W;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int b = tile / tiles_per_img;
      const int t2 = tile - b * tiles_per_img;
      const int ty = t2 / p.tiles_x, tx = t2 - ty * p.tiles_x;
      const int gy = ty * PH + pr, gx = tx * PW + pc;
      const bool pvalid = (gy < p.H) && (gx < p.W);
      const uint32_t acc = tcount & 1u;
      const bool rec = (p.timing & 1) && (blockIdx.x == 0) && (warp == 2) && (lane == 0);
      long long tq0 = 0, tq1 = 0;
      if (rec) tq0 = clock64();
      mbar_wait(&tmem_full[acc], (tcount >> 1) & 1u);
      if (rec) tq1 = clock64();
      tc_fence_after();
      float* ypix = p.y + (int64_t)b * p.y_bstride + (int64_t)gy * p.W + gx;
      float oc_dot = 0.f;   // fused OutConv: this pixel's dot product over all Cout activations (lane = pixel)
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * L::ACC_COLS;
      if (p.stats) {
        // BatchNorm batch statistics from the RAW accumulators, re-read in fragment layout (tc_common.cuh); patch pixels
        // outside the image are masked (their stencil still sees the image edge), channels past Cout are exact zeros (TMA
        // zero fill of the weight rows); the affine is applied to the sums analytically
        const uint32_t vmask = __ballot_sync(0xffffffffu, pvalid);
        const double npix = (double)__popc(vmask);
#pragma unroll 1
        for (int c0 = 0; c0 < N_TILE; c0 += 32) {
          if (c0 >= p.Cout) break;
          float s1, s2;
          tmem_colsum32<X3 ? N_TILE : 0>(tacc + (uint32_t)c0, lane, s1, s2, vmask);
          const int c = c0 + tmem_colsum32_col(lane);
          if (c < p.Cout) {
            const double sc = (double)aff[c], sh = (double)aff[L::AFF_N + c];
            atomicAdd(p.stats + c, sc * (double)s1 + npix * sh);
            atomicAdd(p.stats + p.Cout + c, sc * sc * (double)s2 + 2.0 * sc * sh * (double)s1 + npix * sh * sh);
          }
        }
      }
      // 16 accumulator columns per step keep the epilogue within the 104-register budget of the 608-thread CTA
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 16) {
        if (c0 >= p.Cout) break;
        uint32_t r[16];
        tmem_ld16(tacc + (uint32_t)c0, r);
        if (X3) {   // second half of the accumulator: the A_hi*B_lo term
          uint32_t r2[16];
          tmem_ld16(tacc + (uint32_t)(N_TILE + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        float scv[16], shv[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 a = *reinterpret_cast<const float4*>(aff + c0 + 4 * j4);
          const float4 t = *reinterpret_cast<const float4*>(aff + L::AFF_N + c0 + 4 * j4);
          scv[4 * j4] = a.x; scv[4 * j4 + 1] = a.y; scv[4 * j4 + 2] = a.z; scv[4 * j4 + 3] = a.w;
          shv[4 * j4] = t.x; shv[4 * j4 + 1] = t.y; shv[4 * j4 + 2] = t.z; shv[4 * j4 + 3] = t.w;
        }
        tmem_ld_wait();
        const int nchn = min(16, p.Cout - c0);
        float* yp = ypix + (int64_t)c0 * P;
        if (p.oc_y) {
          // channels past Cout have zero accumulators, identity affine and zero OutConv weight: no mask needed
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w4 = *reinterpret_cast<const float4*>(aff + 2 * L::AFF_N + c0 + 4 * j4);
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * j4 + e;
              oc_dot = fmaf(fmaxf(fmaf(__uint_as_float(r[j]), scv[j], shv[j]), act_lo), wv[e], oc_dot);
            }
          }
        } else if (nchn == 16) {
          if (pvalid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              *yp = fmaxf(fmaf(__uint_as_float(r[j]), scv[j], shv[j]), act_lo);
              yp += P;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (pvalid && j < nchn) yp[(int64_t)j * P] = fmaxf(fmaf(__uint_as_float(r[j]), scv[j], shv[j]), act_lo);
        }
      }
      if (p.oc_y && pvalid) p.oc_y[(int64_t)b * P + (int64_t)gy * p.W + gx] = oc_dot + (p.oc_b ? __ldg(p.oc_b) : 0.f);
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (rec) {
        atomicAdd(&g_ds_timing[9], (unsigned long long)(tq1 - tq0));
        atomicAdd(&g_ds_timing[10], (unsigned long long)(clock64() - tq1));
        atomicAdd(&g_ds_timing[11], 1ull);
      }
    }
  } else {
    // ===== depthwise producer groups: NG groups of 4 warps (warps 6 ..), group g takes every NG-th chunk =====
    const int g = (warp - 6) >> 2;
    const int t = threadIdx.x - 192 - 128 * g;  // 0..127
    const int Cin = p.C0 + p.C1;
    uint32_t gc = 0, iph = 0;       // iph: phase bit per input stage of this group's fill barriers
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int i = 0; i < nch; ++i, ++gc) {
        if ((int)(gc % (uint32_t)L::NG) != g) continue;
        const int s = gc % IS;
        const bool rec = (p.timing & 1) && (blockIdx.x == 0) && (t == 0) && (g == 0);
        long long tk0 = 0, tk1 = 0, tk2 = 0;
        // depthwise weights of this thread's (first) task: issued before the ring waits so that their latency hides there
        float wr[KPL][9], br[KPL];
        auto task_channel = [&](int task) { return (PW == 32) ? (task >> 3) : (((task >> 4) << 1) | ((task >> 2) & 1)); };
        auto load_weights = [&](int task) {
          const int gch = i * CC + task_channel(task);
          const bool chv = gch < Cin;
#pragma unroll
          for (int kk = 0; kk < KPL; ++kk) {
            const int gk = gch * KPL + kk;
#pragma unroll
            for (int w9 = 0; w9 < 9; ++w9) wr[kk][w9] = chv ? __ldg(p.dw_w + (int64_t)gk * 9 + w9) : 0.f;
            br[kk] = (chv && p.dw_b) ? __ldg(p.dw_b + gk) : 0.f;
          }
        };
        load_weights(t);
        if (rec) tk0 = clock64();
        mbar_wait(&in_full[s * L::NG + g], (iph >> s) & 1u);
        iph ^= 1u << s;
        if (rec) tk1 = clock64();
        const int sa = gc % AS;
        mbar_wait(&a_empty[sa], ((gc / AS) & 1u) ^ 1u);  // MMAs that read this A stage 3 chunks ago retired
        if (rec) tk2 = clock64();
        unsigned char* my_op = a_base + sa * L::AST_BYTES;
        const float* in_stage = reinterpret_cast<const float*>(smem + s * L::IN_BYTES);
#pragma unroll 1
        for (int task = t; task < CC * 8; task += 128) {
          // task -> (input channel ci, column quad qc, row group rg).  A quarter-warp (8 lanes: one 128-bit shared-memory
          // wavefront) must touch 8 different 16-byte bank groups: PW = 32 -> the 8 quads of one channel row; PW = 16 ->
          // the 4 quads of TWO channels (16 words apart in the input tile, and 2 k-rows apart = a different 32-byte-atom
          // swizzle phase in the A operand).  Pairing the two row groups of one channel instead (rows 4 apart: 96 words in
          // the input tile, 4 KB in the A operand) put both halves on the same banks: every LDS.128 / STS.128 2-way.
          int ci, qc, rg;
          if (PW == 32) {
            ci = task >> 3; qc = task & 7; rg = 0;
          } else {
            qc = task & 3; ci = ((task >> 4) << 1) | ((task >> 2) & 1); rg = (task >> 3) & 1;
          }
          constexpr int NQ = PW / 4;
          const int c0 = qc << 2, r0 = rg << 2;
          if (task != t) load_weights(task);   // KPL = 1: a second task per chunk
          // smem column of patch column c (dx = -1..1) is c + 4 + dx: the 4 outputs read cols c0+3 .. c0+8
          const float* trow = in_stage + (ci * BH + r0) * BW + c0 + 3;
          float win[3][6];
          // one LDS.128 per row; the two edge values come from the neighbouring quads' registers (lane -1 / +1 hold columns
          // c0-4..c0-1 / c0+4..c0+7 of the same channel row), only the first / last quad of a patch row reads the halo
          // column -- the 32 lanes' scalar loads would all fall on 8 banks (stride 4 words): a 4-way conflict each
          const bool lb = (qc == 0), rb = (qc == NQ - 1);
          const int edge = lb ? 0 : 5;
          auto load_row = [&](float* wl, int r) {
            const float4 a = *reinterpret_cast<const float4*>(trow + r * BW + 1);
            float left = __shfl_up_sync(0xffffffffu, a.w, 1), right = __shfl_down_sync(0xffffffffu, a.x, 1);
            if (lb | rb) {
              const float e = trow[r * BW + edge];
              if (lb) left = e; else right = e;
            }
            wl[0] = left; wl[1] = a.x; wl[2] = a.y; wl[3] = a.z; wl[4] = a.w; wl[5] = right;
          };
#pragma unroll
          for (int r = 0; r < 2; ++r) load_row(win[r], r);
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            load_row(win[(rr + 2) % 3], rr + 2);
            const float* w0 = win[rr % 3];
            const float* w1 = win[(rr + 1) % 3];
            const float* w2 = win[(rr + 2) % 3];
            const int m = (r0 + rr) * PW + c0;
#pragma unroll
            for (int kk = 0; kk < KPL; ++kk) {
              float o4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float a = br[kk];
                a = fmaf(wr[kk][0], w0[j], a); a = fmaf(wr[kk][1], w0[j + 1], a); a = fmaf(wr[kk][2], w0[j + 2], a);
                a = fmaf(wr[kk][3], w1[j], a); a = fmaf(wr[kk][4], w1[j + 1], a); a = fmaf(wr[kk][5], w1[j + 2], a);
                a = fmaf(wr[kk][6], w2[j], a); a = fmaf(wr[kk][7], w2[j + 1], a); a = fmaf(wr[kk][8], w2[j + 2], a);
                o4[j] = a;
              }
              const uint32_t off = a_tile_offset(ci * KPL + kk, m);
              if (X3) {
                const float4 h = make_float4(tf32_hi(o4[0]), tf32_hi(o4[1]), tf32_hi(o4[2]), tf32_hi(o4[3]));
                *reinterpret_cast<float4*>(my_op + off) = h;
                *reinterpret_cast<float4*>(my_op + L::OFF_ALO + off) = make_float4(o4[0] - h.x, o4[1] - h.y, o4[2] - h.z, o4[3] - h.w);
              } else {
                *reinterpret_cast<float4*>(my_op + off) = make_float4(o4[0], o4[1], o4[2], o4[3]);
              }
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        mbar_arrive(&a_full[sa]);
        mbar_arrive(&in_empty[s]);
        if (rec) {
          const long long tk3 = clock64();
          atomicAdd(&g_ds_timing[0], (unsigned long long)(tk1 - tk0));
          atomicAdd(&g_ds_timing[1], (unsigned long long)(tk2 - tk1));
          atomicAdd(&g_ds_timing[2], (unsigned long long)(tk3 - tk2));
          atomicAdd(&g_ds_timing[3], 1ull);
        }
      }
    }
  }
  __syncthreads();
  if ((p.timing & 1) && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_ds_timing[12], (unsigned long long)(clock64() - t_kernel0));
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

template <int N_TILE, int KPL, int PW, bool X3>
static int launch_ds(const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& mw, const CUtensorMap& mwl, DsParams p,
                     int B, cudaStream_t st) {
  using L = DsCfg<N_TILE, KPL, PW, X3>;
  auto kern = dsconv_fused_kernel<N_TILE, KPL, PW, X3>;
  static std::atomic<uint64_t> attr_mask{0};   // cudaFuncSetAttribute is per device
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "dsconv: smem attribute (%d B): %s", L::TOTAL, cudaGetErrorString(e));
  }
  p.tiles_x = ceil_div(p.W, PW);
  p.tiles_y = ceil_div(p.H, L::PH);
  const int64_t total = (int64_t)B * p.tiles_x * p.tiles_y;
  SMAAT_REQUIRE(total < (1ll << 31), "dsconv: too many tiles");
  p.total_tiles = (int)total;
  p.nchunks = ceil_div(p.C0 + p.C1, L::CC);
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kern<<<grid, L::THREADS, L::TOTAL, st>>>(m0, m1, mw, mwl, p);
  SMAAT_LAUNCH_CHECK("smaat_dsconv_fwd");
  return SMAAT_OK;
}

// patch width: 32 (PH = 4) or 16 (PH = 8), whichever wastes fewer MMA lanes; 0 = not worth fusing
static int pick_pw(int H, int W) {
  double best = 1e9;
  int pw = 0;
  const int cand[2] = {32, 16};
  for (int i = 0; i < 2; ++i) {
    const int c = cand[i], ph = TC_BM / c;
    const double waste = ((double)ceil_div(W, c) * c / W) * ((double)ceil_div(H, ph) * ph / H);
    if (waste < best - 1e-9) {
      best = waste;
      pw = c;
    }
  }
  return best <= 1.35 ? pw : 0;
}

static bool ds_eligible(const float* x0, int C0, int64_t bs0, const float* x1, int C1, int64_t bs1, const float* pw_w,
                        const float* pw_w_lo, int H, int W, int k, int Cout) {
  if (k != 1 && k != 2) return false;
  if (Cout > 128 || Cout < 8) return false;
  if (W % 4 != 0 || !aligned16(x0) || bs0 % 4 != 0) return false;
  if (C1 > 0 && (!aligned16(x1) || bs1 % 4 != 0 || C0 % (TC_BK / k) != 0)) return false;
  const int K = k * (C0 + C1);
  if (K % 4 != 0 || !aligned16(pw_w) || (pw_w_lo && !aligned16(pw_w_lo))) return false;
  return pick_pw(H, W) != 0;
}

// second-generation kernel (dsconv_tmem.cu): A operand in tensor memory; k = 2, Cout <= 128, no batch statistics
bool dsconv_tmem_eligible(const float* x0, int C0, int64_t bs0, const float* x1, int C1, int64_t bs1, const float* pw_w,
                          const float* pw_w_lo, int H, int W, int k, int Cout);
int dsconv_tmem_run(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride, const float* dw_w,
                    const float* dw_b, const float* pw_w, const float* pw_w_lo, const float* scale, const float* shift, float* y,
                    int64_t y_bstride, const float* oc_w, const float* oc_b, float* oc_y, int B, int H, int W, int Cout, int relu, int mode,
                    cudaStream_t st);

// 0 = auto (TMEM-operand kernel where it applies, else the shared-memory-operand kernel), 1 = shared-memory-operand kernel only,
// 2 = TMEM-operand kernel only (shapes it does not take are refused).  SMAAT_DS_IMPL presets it.
static std::atomic<int> g_ds_impl{-1};
static int ds_impl() {
  int v = g_ds_impl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("SMAAT_DS_IMPL");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
    g_ds_impl.store(v, std::memory_order_relaxed);
  }
  return v;
}

}  // namespace smaat

using namespace smaat;

extern "C" int smaat_set_dsconv_impl(int impl) {
  SMAAT_REQUIRE(impl >= 0 && impl <= 2, "set_dsconv_impl: 0 = auto, 1 = shared-memory A operand, 2 = TMEM A operand");
  g_ds_impl.store(impl, std::memory_order_relaxed);
  return SMAAT_OK;
}

extern "C" int smaat_dsconv_eligible2(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                                      const float* pw_w, int H, int W, int k, int Cout, int with_stats) {
  const int impl = ds_impl();
  // batch statistics in the epilogue exist in the shared-memory-operand kernel only
  if (impl != 1 && !with_stats && dsconv_tmem_eligible(x0, C0, x0_bstride, x1, C1, x1_bstride, pw_w, nullptr, H, W, k, Cout)) return 1;
  if (impl == 2) return 0;
  return ds_eligible(x0, C0, x0_bstride, x1, C1, x1_bstride, pw_w, nullptr, H, W, k, Cout) ? 1 : 0;
}

extern "C" int smaat_dsconv_eligible(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                                     const float* pw_w, int H, int W, int k, int Cout) {
  return smaat_dsconv_eligible2(x0, C0, x0_bstride, x1, C1, x1_bstride, pw_w, H, W, k, Cout, 0);
}

static int dsconv_run(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride, const float* dw_w,
                      const float* dw_b, const float* pw_w, const float* pw_w_lo, const float* scale, const float* shift, float* y,
                      int64_t y_bstride, double* stats, const float* oc_w, const float* oc_b, float* oc_y, int B, int H, int W,
                      int k, int Cout, int relu, int mode, void* stream) {
  SMAAT_REQUIRE(x0 && dw_w && pw_w && (y || oc_y), "dsconv: null pointer");
  SMAAT_REQUIRE(B > 0 && C0 > 0 && C1 >= 0 && H > 0 && W > 0 && Cout > 0, "dsconv: bad shape");
  SMAAT_REQUIRE(C1 == 0 || x1, "dsconv: C1=%d but x1 is null", C1);
  SMAAT_REQUIRE(mode == SMAAT_PW_TF32 || mode == SMAAT_PW_TF32X3, "dsconv: mode must be SMAAT_PW_TF32 or SMAAT_PW_TF32X3");
  SMAAT_REQUIRE(mode != SMAAT_PW_TF32X3 || pw_w_lo, "dsconv: TF32X3 needs pw_w_lo (see smaat_split_tf32)");
  SMAAT_REQUIRE(oc_y || y_bstride >= (int64_t)Cout * H * W, "dsconv: y batch stride too small");
  SMAAT_REQUIRE(!oc_y || (oc_w && !stats), "dsconv+outconv: needs the OutConv weight and no batch statistics");
  const int impl = ds_impl();
  if (impl != 1 && !stats && dsconv_tmem_eligible(x0, C0, x0_bstride, x1, C1, x1_bstride, pw_w, pw_w_lo, H, W, k, Cout))
    return dsconv_tmem_run(x0, C0, x0_bstride, x1, C1, x1_bstride, dw_w, dw_b, pw_w, pw_w_lo, scale, shift, y, y_bstride, oc_w, oc_b, oc_y,
                           B, H, W, Cout, relu, mode, (cudaStream_t)stream);
  if (impl == 2 || !ds_eligible(x0, C0, x0_bstride, x1, C1, x1_bstride, pw_w, pw_w_lo, H, W, k, Cout))
    return fail(SMAAT_E_UNSUPPORTED, "dsconv: shape not taken by the fused kernel (k=%d Cout=%d H=%d W=%d); use dw3x3 + pw1x1", k,
                Cout, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  const int pw = pick_pw(H, W);
  const int ph = TC_BM / pw;
  const int n_tile = Cout > 64 ? 128 : 64;
  const int cc = TC_BK / k;
  const bool x3 = mode == SMAAT_PW_TF32X3;
  const int K = k * (C0 + C1);

  CUtensorMap m0, m1, mw, mwl;
  const uint32_t box[4] = {(uint32_t)(pw + 8), (uint32_t)(ph + 2), (uint32_t)cc, 1u};
  {
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C0, (uint64_t)B};
    const uint64_t str[4] = {0, (uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)x0_bstride * 4};
    int r = make_tmap_f32(&m0, x0, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, "dsconv(x0)");
    if (r) return r;
    m1 = m0;
  }
  if (C1 > 0) {
    const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C1, (uint64_t)B};
    const uint64_t str[4] = {0, (uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)x1_bstride * 4};
    int r = make_tmap_f32(&m1, x1, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, "dsconv(x1)");
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)Cout};
    const uint64_t str[2] = {0, (uint64_t)K * 4};
    const uint32_t wbox[2] = {(uint32_t)TC_BK, (uint32_t)n_tile};
    int r = make_tmap_f32(&mw, pw_w, 2, dims, str, wbox, CU_TENSOR_MAP_SWIZZLE_128B, "dsconv(w)");
    if (r) return r;
    mwl = mw;
    if (x3) {
      r = make_tmap_f32(&mwl, pw_w_lo, 2, dims, str, wbox, CU_TENSOR_MAP_SWIZZLE_128B, "dsconv(w_lo)");
      if (r) return r;
    }
  }
  DsParams p;
  p.dw_w = dw_w; p.dw_b = dw_b; p.scale = scale; p.shift = shift; p.y = y; p.y_bstride = y_bstride; p.stats = stats;
  p.oc_w = oc_w; p.oc_b = oc_b; p.oc_y = oc_y;
  static const int timing_on = [] { const char* e = getenv("SMAAT_DSCONV_TIMING"); return e ? atoi(e) : 0; }();
  p.timing = timing_on;
  p.C0 = C0; p.C1 = C1; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu; p.K = K;
  p.tiles_x = p.tiles_y = p.total_tiles = p.nchunks = 0;

#define DS_DISPATCH(NT, KP, PWv)                                                   \
  return x3 ? launch_ds<NT, KP, PWv, true>(m0, m1, mw, mwl, p, B, st)              \
            : launch_ds<NT, KP, PWv, false>(m0, m1, mw, mwl, p, B, st)
  if (n_tile == 64) {
    if (k == 2) { if (pw == 32) { DS_DISPATCH(64, 2, 32); } else { DS_DISPATCH(64, 2, 16); } }
    else        { if (pw == 32) { DS_DISPATCH(64, 1, 32); } else { DS_DISPATCH(64, 1, 16); } }
  } else {
    if (k == 2) { if (pw == 32) { DS_DISPATCH(128, 2, 32); } else { DS_DISPATCH(128, 2, 16); } }
    else        { if (pw == 32) { DS_DISPATCH(128, 1, 32); } else { DS_DISPATCH(128, 1, 16); } }
  }
#undef DS_DISPATCH
}

extern "C" int smaat_dsconv_fwd(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                                const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_w_lo,
                                const float* scale, const float* shift, float* y, int64_t y_bstride, double* stats, int B, int H,
                                int W, int k, int Cout, int relu, int mode, void* stream) {
  SMAAT_REQUIRE(y, "dsconv: null output");
  return dsconv_run(x0, C0, x0_bstride, x1, C1, x1_bstride, dw_w, dw_b, pw_w, pw_w_lo, scale, shift, y, y_bstride, stats, nullptr,
                    nullptr, nullptr, B, H, W, k, Cout, relu, mode, stream);
}

/* The network's last two modules in one kernel: DS conv -> BN/ReLU -> OutConv(Cout -> 1) (reference models/SmaAt_UNet.py:55-56,
 * unet_parts.py:67-73).  The Cout-channel activation never reaches HBM: each epilogue thread owns one pixel's Cout values
 * (TMEM lane = pixel) and reduces them against oc_w.  logits: (B, 1, H, W). */
extern "C" int smaat_dsconv_outconv_fwd(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                                        const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_w_lo,
                                        const float* scale, const float* shift, const float* oc_w, const float* oc_b,
                                        float* logits, int B, int H, int W, int k, int Cout, int relu, int mode, void* stream) {
  SMAAT_REQUIRE(oc_w && logits, "dsconv+outconv: null pointer");
  return dsconv_run(x0, C0, x0_bstride, x1, C1, x1_bstride, dw_w, dw_b, pw_w, pw_w_lo, scale, shift, nullptr, 0, nullptr, oc_w, oc_b,
                    logits, B, H, W, k, Cout, relu, mode, stream);
}

/* Debug hook: copies the fused kernel's stage timers of CTA 0 (16 counters, clock64 cycles; layout in dsconv_fused.cu)
 * to `out` and clears them.  Synchronises the device. */
extern "C" int smaat_debug_dsconv_timing(unsigned long long* out) {
  SMAAT_REQUIRE(out, "debug_dsconv_timing: null pointer");
  cudaError_t e = cudaMemcpyFromSymbol(out, g_ds_timing, sizeof(g_ds_timing));
  if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "debug_dsconv_timing: %s", cudaGetErrorString(e));
  unsigned long long z[16] = {0};
  e = cudaMemcpyToSymbol(g_ds_timing, z, sizeof(z));
  if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "debug_dsconv_timing: %s", cudaGetErrorString(e));
  return SMAAT_OK;
}
