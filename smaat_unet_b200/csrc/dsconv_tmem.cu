// dsconv_tmem.cu -- DepthwiseSeparableConv forward (kernels_per_layer = 2) as one kernel, second generation:
// the depthwise result goes from the CUDA cores STRAIGHT INTO TENSOR MEMORY (tcgen05.st) and is consumed by
// tcgen05.mma as its A operand from TMEM -- it never touches shared memory, let alone HBM.
//
// Replaces DepthwiseSeparableConv.forward (reference models/layers.py:47-50: depthwise 3x3, groups = Cin, k outputs per
// channel, then pointwise 1x1, nothing in between) + eval BatchNorm2d + ReLU (parts_ds.py:25-26,34-35) [+ OutConv,
// unet_parts.py:67-73, for the network's last conv].
//
// Why (round-1 ncu, profiles/r01d_ncu_summary.md): the first-generation kernel (dsconv_fused.cu) staged the A operand in
// shared memory: per 128-pixel x 32-channel chunk it wrote 32 KB (hi + lo tf32 parts), the tensor core read them back
// (32 KB) plus the weights (24-48 KB), the producers read the 15 KB input box -- ~120 KB through a 128 B/clk port, i.e.
// ~1000+ cycles per chunk against an MMA floor of 384-768, with 41 % bank-conflict wavefronts on top, and the weight
// chunk re-fetched from L2 for every tile (L2 -> SM traffic ~7.5 TB/s on the K = 512 layers).  Here:
//   * A operand in TMEM: no shared-memory write, no tensor-core read of A from shared memory;
//   * a work unit is a PAIR of vertically adjacent 128-pixel tiles (PW x 8 or 16 x 16 pixels = two M = 128 MMAs): one
//     input box, one depthwise-weight stage and ONE pointwise-weight chunk serve both, halving the weight traffic from L2;
//   * the stencil runs as packed FFMA2 (fma.rn.f32x2): the two depthwise outputs of a channel share every input value.
//
// Thread <-> data mapping (dictated by the tcgen05.st.16x256b fragment layout, tc_common.cuh): a producer warp owns the
// TMEM lane quarter q = warp % 4 of both half tiles; its lane T owns TMEM lanes c + 8r (c = T / 4, r = 0..3) and the K
// columns 8i + 2(T % 4) + {0, 1}, i = 0..3 -- i.e. input channel 4i + T % 4 of the chunk with both of its depthwise
// outputs (k = 2).  Lane 8r + c of quarter q in half h is pixel x = 4 (c % NX) + r of pair row RQ q + 2 (c / NX) + h
// (NX = PW / 4): a thread owns 4 consecutive pixels of two vertically adjacent rows (one per half tile), so a row of the
// window is one LDS.128 plus two edge values shuffled from the neighbouring lanes, and a quarter is whole 32-pixel (two
// 16-pixel) row segments, so the epilogue warp's store of a channel is 128 (2 x 64) contiguous bytes.  [A first version
// mapped a quarter to an 8 x 4 pixel block: its epilogue stores were 4 x 32-byte segments per instruction and ran 5x
// slower -- 94 cycles per column -- although every sector was fully written.]
//
// Warps (512 threads, 1 CTA per SM, persistent over tile pairs):
//   0        TMA: input halo boxes (PW + 8) x (PHP + 3) x 16 channels, OOB zero fill = padding 1; virtual concat [x0, x1]
//   1, 3     MMA issuers: one thread each (a single-thread elect.sync region), warp 1 the MMAs of half tile 0, warp 3 those of
//            half tile 1: per unit and half 4 k-steps x {A_hi B_hi, A_hi B_lo, A_lo B_hi}; warp 1 also allocates TMEM
//   2        pointwise-weight ring loader (TMA, K-major SW128, hi | lo)
//            (the layer's depthwise weights are staged once, by all threads, before the roles split)
//   4..7     epilogue: tcgen05.ld (lane = pixel, 32 columns per step) -> BN affine + ReLU -> NCHW stores (or OutConv dot)
//   8..15    two depthwise producer groups (group g takes every second unit)
// Barriers are tested by phase parity, which is only unambiguous while a waiter can never be two completions behind: the rings
// that are shared between the two producer groups (input fills, A-stage hand-backs) therefore have one barrier per (stage, group).
#include <stdlib.h>

#include "tc_common.cuh"

// Instrumentation (stage timers, event trace, knock-out flags) is compiled in only with -DSMAAT_DT_INSTRUMENT=1
// (SMAAT_DT_INSTRUMENT=1 bash build.sh): even predicated off, its instructions sat in the single-lane MMA issue loop, whose
// length -- not the tensor pipe -- set the pace of the kernel (profiles/r02_dsconv_tmem_trace.txt).
#ifndef SMAAT_DT_INSTRUMENT
#define SMAAT_DT_INSTRUMENT 0
#endif

namespace smaat {

// Stage timers of CTA 0 (clock64 cycles, accumulated over launches until read; SMAAT_DSCONV_TIMING=1): see
// smaat_debug_dsconv_tmem_timing.  [0] producer group 0: wait input stage  [1] wait free A stage  [2] stencil + TMEM stores
// [3] units | [4] MMA: wait A  [5] wait B  [6] wait drained accumulator  [7] issue  [8] units | [9] epilogue warp 4: wait
// accumulators  [10] drain + store  [11] pairs | [12] kernel cycles | [13] TMA: wait free input stage  [14] units |
// [15] stager: wait free stage  [16] total  [17] units | [18] weight loader: wait free stage  [19] units
__device__ unsigned long long g_dt_timing[24];
// Per-CTA record of the last timed launch (SMAAT_DSCONV_TIMING=1): [3 * cta + 0] globaltimer ns at role dispatch, [+1] at exit,
// [+2] SM id -- the spread shows how evenly the static pair -> CTA assignment finishes (smaat_debug_dsconv_tmem_cta_timing).
// Event trace of CTA 0 (SMAAT_DSCONV_TIMING=2; stage-timer atomics off): clock64 at [16 * unit + k] for the first DT_TRACE_UNITS
// units -- k: 0 TMA box issued  1 producer: box landed  2 stencil done  3 A stage free  4 A stored + arrived |
// 5 issuer: operands seen  6 half-0 batch starts  7 half-0 batch issued  8 half-1 starts  9 half-1 issued + commits |
// epilogue (at the pair's first unit): 10 half-0 accumulator seen  11 half-0 drained  12 half-1 seen  13 half-1 drained
constexpr int DT_TRACE_UNITS = 256;
__device__ long long g_dt_trace[16 * DT_TRACE_UNITS];
constexpr int DT_MAX_CTAS = 256;
__device__ unsigned long long g_dt_cta[3 * DT_MAX_CTAS];
__device__ __forceinline__ unsigned long long dt_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct DtParams {
  const float* dw_w;
  const float* dw_b;
  const float* scale;
  const float* shift;
  float* y;
  int64_t y_bstride;
  const float* oc_w;   // fused OutConv (1 class): logits = sum_c oc_w[c] * act[c] + oc_b, written instead of y
  const float* oc_b;
  float* oc_y;
  int C0, C1, H, W, Cout, relu, K;
  int px_tiles, py_tiles, total_pairs, nchunks;
  int flags;           // tuning experiments (SMAAT_DT_FLAGS): 1 = L2 prefetch of the next pair's boxes (measured 3 % slower: off);
                       // 64 = one MMA-issuing warp instead of two (2.5 % slower).  Instrumented builds only, knock-outs that give WRONG
                       // results and exist to find the binding stage: 2 = epilogue drains but does not store, 4 = producers skip the
                       // stencil, 8 = issuer skips the two tf32x3 correction MMAs, 16 = issuer issues no MMA, 32 = every input box is
                       // the CTA's first one (L2 hits)
  int npass;           // output-channel passes of N_TILE channels each (Cout > 128: the depthwise work is repeated per pass)
  int timing;
};

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// instruction descriptor for A from TMEM (always K-major), B K-major: as make_idesc_tf32 without the a_major bit
__host__ __device__ constexpr uint32_t make_idesc_tf32_ts(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
// Fragment-shaped TMEM store, 16 lanes x 8 columns: thread T writes f[e] -> (lane T/4, column 2(T%4) + e) and
// f[2 + e] -> (lane T/4 + 8, same column), e = 0, 1 (the mirror of tmem_ld_16x256b_x4's layout, one column group).
__device__ __forceinline__ void tmem_st_16x256b_x1(uint32_t taddr, uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(f0), "r"(f1), "r"(f2), "r"(f3)
               : "memory");
}
// packed fp32 pair arithmetic (sm_100 FFMA2): d = a * (b, b) + c on (lo, hi) pairs held in 64-bit registers; the scalar
// operand is broadcast by the instruction itself (SASS: FFMA2 Rd, Ra.F32x2, Rb.F32, Rc.F32x2)
__device__ __forceinline__ uint64_t fma2_bcast(uint64_t a, float b, uint64_t c) {
  uint64_t d, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(bb), "l"(c));
  return d;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int N_TILE, int PW, bool X3, bool BIG>
struct DtCfg {
  static constexpr int KPL = 2;
  static constexpr int CC = TC_BK / KPL;                 // 16 input channels per chunk
  static constexpr int NX = PW / 4;                       // 4-pixel groups per row: 8 (PW = 32) or 4 (PW = 16)
  static constexpr int RQ = 2 * (8 / NX);                 // rows of the pair owned by one TMEM lane quarter: 2 or 4
  static constexpr int PHP = 4 * RQ;                      // rows of a tile PAIR: 8 (PW = 32) or 16 (PW = 16)
  static constexpr int BW = PW + 8;
  static constexpr int BH = PHP + 3;                      // 2 halo rows + 1 unused row: channel stride = 8 or 24 mod 32 words
  static constexpr int CHS = BH * BW;                     // words between channels of the staged box
  static_assert(CHS % 32 == 8 || CHS % 32 == 24, "the 4 channel phases of a warp must fall on disjoint bank octets");
  static constexpr int IN_BYTES = CC * CHS * 4;
  static_assert(IN_BYTES % 128 == 0, "TMA destination alignment");
  static constexpr int WD_FLOATS = 20;                    // per input channel: 9 taps x (output 0, output 1) + 2 biases (16-byte aligned rows)
  static constexpr int WD_MAXC = BIG ? 512 : 256;         // the whole layer's depthwise weights stay resident in shared memory:
  static constexpr int WD_MAX_BYTES = WD_MAXC * WD_FLOATS * 4;   // nchunks * 16 rows of 80 B at the END of the carve-up (run-time size)
  static constexpr int B_BYTES = N_TILE * TC_BK * 4;
  static constexpr int BST_BYTES = (X3 ? 2 : 1) * B_BYTES;
  // producer groups (128 threads each).  THREE groups (two-pass stencil with 32 accumulator registers so that 640 threads fit,
  // a fourth input stage) were built when the event trace showed the producers as the pacing stage: parity-green, but slower
  // (12 layers 3.64 ms against 3.35 ms) -- more warps fight for the same issue slots, shared-memory port and FP32 pipe.
  static constexpr int NG = 2;
  // TMEM: accumulators [pair buffer][half] x N_TILE columns, then the A ring (per stage: half 0 hi | lo, half 1 hi | lo)
  static constexpr int ACC_PAIRS = (N_TILE <= 64) ? 2 : 1;
  static constexpr int ACC_COLS = ACC_PAIRS * 2 * N_TILE;
  static constexpr int AH_COLS = X3 ? 64 : 32;            // A columns of one half tile
  static constexpr int AST_COLS = 2 * AH_COLS;
  static constexpr int AS = ((512 - ACC_COLS) / AST_COLS) > 4 ? 4 : ((512 - ACC_COLS) / AST_COLS);
  static_assert(AS >= 2, "A ring");
  // BIG (more than 256 input channels: 40 KB of depthwise weights, >= 16 chunks per pair) trades ring depth for the table
  static constexpr int BS = BIG ? 2 : 3;                  // pointwise-weight ring: its TMA loads must cover an L2 round trip
  static constexpr int IS = (N_TILE <= 64 && !BIG) ? 5 : 3;   // input ring (the TMA thread runs far ahead of the producers anyway)
  static_assert(IS >= 2, "input ring");
  static constexpr int OFF_BR = ((IS * IN_BYTES + 1023) / 1024) * 1024;
  static constexpr int OFF_BAR = OFF_BR + BS * BST_BYTES;
  static constexpr int BAR_BYTES = 512;
  static_assert((IS * NG + IS + AS + AS * NG + 2 * BS + 8) * 8 + 8 <= BAR_BYTES, "barrier block");
  static constexpr int AFF_N = 512;                        // epilogue affine of ALL output channels (up to 4 passes of 128)
  static constexpr int OFF_WD = OFF_BAR + BAR_BYTES + 3 * AFF_N * 4;
  static constexpr int FIXED = OFF_WD + 1024;             // + alignment slack; + the depthwise-weight table = dynamic shared memory
  static_assert(FIXED + WD_MAX_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(N_TILE <= AFF_N, "epilogue affine staging");
  static constexpr int THREADS = 128 + 128 + 128 * NG;
};

template <int N_TILE, int PW, bool X3, bool BIG>
__global__ void __launch_bounds__(DtCfg<N_TILE, PW, X3, BIG>::THREADS, 1)
    dsconv_tmem_kernel(const __grid_constant__ CUtensorMap map_in0, const __grid_constant__ CUtensorMap map_in1,
                       const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo, const DtParams p) {
  using L = DtCfg<N_TILE, PW, X3, BIG>;
  constexpr int IS = L::IS, AS = L::AS, BS = L::BS, CC = L::CC, BW = L::BW, CHS = L::CHS, NX = L::NX, RQ = L::RQ, PHP = L::PHP;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* b_base = smem + L::OFF_BR;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  // [IS][NG] input box landed (TMA tx) -- one barrier per (stage, producer group that reads the fill): with IS odd and two groups a
  // group meets a stage only at every SECOND fill, and a phase-parity test on a per-stage barrier cannot tell "my fill landed"
  // from "the fill before the other group's landed late" (TMA loads may complete out of order).  The group would read the
  // wrong box and arrive on in_empty a phase early: over-arrival, an intermittent "unspecified launch failure" (seen on
  // 576 x 576 inputs in tf32 mode).  Per (stage, group) every visit needs exactly one more completion.
  uint64_t* in_full = bars;
  uint64_t* in_empty = in_full + IS * L::NG;      // [IS] producer group done with the stage (128 arrivals)
  uint64_t* a_full = in_empty + IS;               // [AS] A operand of both halves in TMEM (128 arrivals)
  // [AS][NG] MMAs reading the A stage retired (commit) -- one barrier per (stage, producer group that writes the stage NEXT): a
  // group visits "its" barrier of a stage once per lcm(AS, NG) units and every visit needs exactly one more completion than the
  // last, so the phase-parity test cannot alias.  With one barrier per stage and NG > AS a group could find the stage two
  // completions behind (parity equal again), walk through and overwrite operands still being read.
  uint64_t* a_empty = a_full + AS;
  uint64_t* b_full = a_empty + AS * L::NG;        // [BS]
  uint64_t* b_empty = b_full + BS;                // [BS]
  uint64_t* tmem_full = b_empty + BS;             // [2][2] per (pair buffer, half): its MMAs retired (the epilogue starts on half 0 while half 1 finishes)
  uint64_t* tmem_empty = tmem_full + 4;           // [2][2] per (pair buffer, half): drained by the epilogue (128 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 4);
  float* aff = reinterpret_cast<float*>(smem + L::OFF_BAR + L::BAR_BYTES);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nch = p.nchunks;
  const int pairs_per_img = p.px_tiles * p.py_tiles;
  // this CTA's pairs: blockIdx.x, + gridDim.x, ...; its units: (pair, chunk) in that order
  const int my_pairs = (p.total_pairs > (int)blockIdx.x) ? (p.total_pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

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
      for (int g = 0; g < L::NG; ++g) mbar_init(&a_empty[s * L::NG + g], (p.flags & 64) ? 1 : 2);
    }
    for (int s = 0; s < BS; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], (p.flags & 64) ? 1 : 2);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&tmem_full[s], 1);
    for (int s = 0; s < 4; ++s) mbar_init(&tmem_empty[s], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  for (int c = threadIdx.x; c < L::AFF_N; c += blockDim.x) {
    aff[c] = (c < p.Cout && p.scale) ? __ldg(p.scale + c) : 1.f;
    aff[L::AFF_N + c] = (c < p.Cout && p.shift) ? __ldg(p.shift + c) : 0.f;
    aff[2 * L::AFF_N + c] = (c < p.Cout && p.oc_w) ? __ldg(p.oc_w + c) : 0.f;
  }
  {
    // the layer's depthwise weights, once: row c = (w[2c][tap], w[2c + 1][tap]) x 9 taps, then the 2 biases (LDS.128-able,
    // already paired for FFMA2); channels past Cin (the last chunk's tail) are zero rows.  [A per-unit staging warp was the
    // first design: its global-load latency, ~2500 cycles per unit, set the pace of the whole kernel.]
    float* wd_all = reinterpret_cast<float*>(smem + L::OFF_WD);
    const int Cin = p.C0 + p.C1;
    const int rows = nch * CC;
    for (int idx = threadIdx.x; idx < rows * L::WD_FLOATS; idx += blockDim.x) {
      const int c = idx / L::WD_FLOATS, f = idx - c * L::WD_FLOATS;
      float x = 0.f;
      if (c < Cin) {
        if (f < 18) x = __ldg(p.dw_w + (int64_t)c * 18 + (f & 1) * 9 + (f >> 1));
        else if (p.dw_b) x = __ldg(p.dw_b + c * 2 + (f - 18));
      }
      wd_all[idx] = x;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t a_ring = tmem_base + (uint32_t)L::ACC_COLS;
  const bool rec0 = SMAAT_DT_INSTRUMENT && p.timing == 1 && blockIdx.x == 0 && lane == 0;     // stage timers: CTA 0, one lane per role
  const bool trc = SMAAT_DT_INSTRUMENT && p.timing == 2 && blockIdx.x == 0 && lane == 0;
#define DT_FLAG(x) (SMAAT_DT_INSTRUMENT && (p.flags & (x)))
#define DT_TR(unit, k) do { if (trc && (unit) < (uint32_t)DT_TRACE_UNITS) g_dt_trace[16 * (unit) + (k)] = clock64(); } while (0)
  const long long t_kernel0 = (rec0 && warp == 0) ? clock64() : 0;
  const bool rec_cta = SMAAT_DT_INSTRUMENT && p.timing && threadIdx.x == 0 && blockIdx.x < DT_MAX_CTAS;
  if (rec_cta) g_dt_cta[3 * blockIdx.x] = dt_globaltimer();
#define DT_T(var) const long long var = rec0 ? clock64() : 0
#define DT_ADD(idx, a, b) do { if (rec0) atomicAdd(&g_dt_timing[idx], (unsigned long long)((b) - (a))); } while (0)
#define DT_INC(idx) do { if (rec0) atomicAdd(&g_dt_timing[idx], 1ull); } while (0)

  // work item = (pair, channel pass), pass fastest: the passes of a pair run on neighbouring SMs at the same time and
  // share the pair's input boxes through L2
  auto pair_origin = [&](int item, int& b, int& x0, int& y0) {
    const int pair = item / p.npass;
    b = pair / pairs_per_img;
    const int t2 = pair - b * pairs_per_img;
    const int ty = t2 / p.px_tiles, tx = t2 - ty * p.px_tiles;
    x0 = tx * PW;
    y0 = ty * PHP;
  };

  if (warp == 0) {
    // ===== TMA: one input halo box per unit =====
    // [plain lane test: these loops issue 1-3 instructions per unit, the single-thread elect.sync region of the MMA issuers
    // has nothing to gain here]
    if (lane == 0) {
      uint32_t u = 0;
      for (int j = 0; j < my_pairs; ++j) {
        const int pair = blockIdx.x + j * gridDim.x;
        int b, x0, y0;
        pair_origin(DT_FLAG(32) ? (int)blockIdx.x : pair, b, x0, y0);
        int nb = 0, nx0 = 0, ny0 = 0;
        const bool has_next = j + 1 < my_pairs;
        if (has_next) pair_origin(pair + gridDim.x, nb, nx0, ny0);
        for (int i = 0; i < nch; ++i, ++u) {
          const int s = u % IS;
          DT_T(tq0);
          mbar_wait(&in_empty[s], ((u / IS) & 1u) ^ 1u);
          DT_T(tq1);
          DT_ADD(13, tq0, tq1);
          DT_INC(14);
          DT_TR(u, 0);
          uint64_t* full = &in_full[s * L::NG + (int)(u % (uint32_t)L::NG)];      // the barrier of the group that reads unit u
          mbar_arrive_expect_tx(full, L::IN_BYTES);
          const int cb = i * CC;
          const CUtensorMap* m = (cb < p.C0) ? &map_in0 : &map_in1;
          const int cc = (cb < p.C0) ? cb : cb - p.C0;
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                  "r"(smem_u32(smem + s * L::IN_BYTES)),
              "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(full)), "r"(x0 - 4), "r"(y0 - 1), "r"(cc), "r"(b)
              : "memory");
          // the same chunk of this CTA's NEXT pair goes to L2 now, so its TMA load later pays L2 latency only
          if (has_next && (p.flags & 1))
            asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                         "r"(nx0 - 4), "r"(ny0 - 1), "r"(cc), "r"(nb)
                         : "memory");
        }
      }
    }
  } else if (warp == 2) {
    // ===== pointwise-weight ring: K-major SW128 chunks (hi rows | lo rows), one per unit, shared by both half tiles =====
    if (lane == 0) {
      uint32_t u = 0;
      for (int j = 0; j < my_pairs; ++j) {
        const int n0 = ((blockIdx.x + j * gridDim.x) % p.npass) * N_TILE;
        for (int i = 0; i < nch; ++i, ++u) {
          const int sb = u % BS;
          DT_T(tq0);
          mbar_wait(&b_empty[sb], ((u / BS) & 1u) ^ 1u);
          DT_T(tq1);
          DT_ADD(18, tq0, tq1);
          DT_INC(19);
          mbar_arrive_expect_tx(&b_full[sb], L::BST_BYTES);
          tma_load_2d(b_base + sb * L::BST_BYTES, &map_w, &b_full[sb], i * TC_BK, n0);
          if (X3) tma_load_2d(b_base + sb * L::BST_BYTES + L::B_BYTES, &map_wlo, &b_full[sb], i * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1 || (warp == 3 && !(p.flags & 64))) {
    // ===== MMA issuers: ONE lane of warp 1 (half tile 0) and of warp 3 (half tile 1) run the loop; flag 64: warp 1 takes both =====
    // A lone warp issues its dependent scalar instructions a few cycles apart, so the loop is kept to the MMAs, their operand
    // addresses, the commits and the barrier tests: the first version (elect.sync per batch, index arithmetic by division,
    // predicated-off timers) spent ~550 instructions ~ 2000 cycles per unit here while the 24 MMAs need 800-1600.
    // The waits for unit u + 1's operands are taken BETWEEN the two half-tile batches of unit u (behind queued MMAs).
    if (elect_one()) {      // elect.sync, not lane == 0: the compiler then knows ONE thread runs the region and emits each
                            // tcgen05.mma once -- under a plain lane test every MMA sits in its own per-active-lane loop
      constexpr uint32_t idesc = make_idesc_tf32_ts(N_TILE);
      const bool dual = (p.flags & 64) == 0;
      const bool rec0_role = rec0;
      const bool rec0 = rec0_role && warp == 1;          // stage timers: the first issuer only
      const uint32_t total_units = (uint32_t)my_pairs * (uint32_t)nch;
      const int nk_last = (min(TC_BK, p.K - (nch - 1) * TC_BK) + 7) >> 3;
      const uint64_t bdesc0 = make_b_desc(smem_u32(b_base));          // stage sb: + sb * BST_BYTES / 16 in the address field
      uint32_t u = 0, sa = 0, sb = 0, pha = 0, phb = 0;               // ring positions and phase bits of the CURRENT unit
      uint32_t gnx = (uint32_t)(AS % L::NG);                          // producer group of unit u + AS: the next writer of stage sa
      // one half-tile batch of unit (sa, sb): 4 k-steps x (hi*hi [+ hi*lo + lo*hi])
      auto batch = [&](uint32_t h, uint32_t pb, int i, int nk) {
        const uint32_t d_tmem = tmem_base + (pb * 2 + h) * N_TILE;
        const uint32_t a_hi = a_ring + sa * (uint32_t)L::AST_COLS + h * (uint32_t)L::AH_COLS;
        const uint64_t bd_hi = bdesc0 + (uint64_t)(sb * (uint32_t)(L::BST_BYTES >> 4));
        const uint64_t bd_lo = bd_hi + (uint64_t)(L::B_BYTES >> 4);
        // k-step: 8 TMEM columns of A, 8 tf32 = 32 B along the K-major weight rows (descriptor address unit = 16 B)
        auto kstep = [&](int kk, uint32_t acc) {
          if (DT_FLAG(16)) return;
          umma_tf32_ts(d_tmem, a_hi + 8u * kk, bd_hi + (uint64_t)(kk * 2), idesc, acc);
          if (X3 && !DT_FLAG(8)) {
            umma_tf32_ts(d_tmem, a_hi + 8u * kk, bd_lo + (uint64_t)(kk * 2), idesc, 1u);
            umma_tf32_ts(d_tmem, a_hi + 32u + 8u * kk, bd_hi + (uint64_t)(kk * 2), idesc, 1u);
          }
        };
        if (nk == TC_BK / 8) {      // the common case: MMAs back to back
          kstep(0, i > 0 ? 1u : 0u);
          kstep(1, 1u);
          kstep(2, 1u);
          kstep(3, 1u);
        } else {
          for (int kk = 0; kk < nk; ++kk) kstep(kk, (i > 0 || kk > 0) ? 1u : 0u);
        }
      };
      auto wait_next = [&]() {      // operands of unit u + 1
        const uint32_t san = (sa + 1 == (uint32_t)AS) ? 0u : sa + 1, sbn = (sb + 1 == (uint32_t)BS) ? 0u : sb + 1;
        DT_T(tm0);
        mbar_wait(&a_full[san], san ? pha : pha ^ 1u);
        DT_T(tm1);
        mbar_wait(&b_full[sbn], sbn ? phb : phb ^ 1u);
        DT_T(tm2);
        tc_fence_after();
        DT_ADD(4, tm0, tm1);
        DT_ADD(5, tm1, tm2);
        DT_INC(8);
        if (warp == 1) DT_TR(u + 1, 5);
      };
      if (total_units > 0) {
        mbar_wait(&a_full[0], 0u);
        mbar_wait(&b_full[0], 0u);
        tc_fence_after();
      }
      const uint32_t hmine = (warp == 3) ? 1u : 0u;
      for (int j = 0; j < my_pairs; ++j) {
        const uint32_t pb = (L::ACC_PAIRS == 2) ? (uint32_t)(j & 1) : 0u;
        const uint32_t use = (L::ACC_PAIRS == 2) ? (uint32_t)(j >> 1) : (uint32_t)j;    // how often this pair buffer was used before
        for (int i = 0; i < nch; ++i, ++u) {
          const int nk = (i == nch - 1) ? nk_last : TC_BK / 8;
          if (dual) {
            DT_T(tm3);
            if (i == 0) {
              mbar_wait(&tmem_empty[pb * 2 + hmine], (use & 1u) ^ 1u);     // the epilogue drained this accumulator
              tc_fence_after();
            }
            DT_T(tm4);
            DT_ADD(6, tm3, tm4);
            DT_TR(u, 6 + 2 * hmine);
            batch(hmine, pb, i, nk);
            if (i == nch - 1) umma_commit(&tmem_full[pb * 2 + hmine]);    // a commit tracks all MMAs issued so far by this thread
            umma_commit(&a_empty[sa * (uint32_t)L::NG + gnx]);            // both issuers arrive: count 2
            umma_commit(&b_empty[sb]);
            DT_T(tm5);
            DT_ADD(7, tm4, tm5);
            DT_TR(u, 7 + 2 * hmine);
            if (u + 1 < total_units) wait_next();
          } else {
            DT_T(tm3);
            if (i == 0) {
              mbar_wait(&tmem_empty[pb * 2 + 0], (use & 1u) ^ 1u);
              tc_fence_after();
            }
            DT_T(tm4);
            DT_TR(u, 6);
            batch(0u, pb, i, nk);
            if (i == nch - 1) umma_commit(&tmem_full[pb * 2 + 0]);
            DT_T(tm5);
            DT_TR(u, 7);
            if (u + 1 < total_units) wait_next();
            DT_T(tm6);
            if (i == 0) {
              mbar_wait(&tmem_empty[pb * 2 + 1], (use & 1u) ^ 1u);
              tc_fence_after();
            }
            DT_T(tm7);
            DT_TR(u, 8);
            batch(1u, pb, i, nk);
            if (i == nch - 1) umma_commit(&tmem_full[pb * 2 + 1]);
            umma_commit(&a_empty[sa * (uint32_t)L::NG + gnx]);
            umma_commit(&b_empty[sb]);
            DT_T(tm8);
            DT_TR(u, 9);
            DT_ADD(6, tm3, tm4);
            DT_ADD(6, tm6, tm7);
            DT_ADD(7, tm4, tm5);
            DT_ADD(7, tm7, tm8);
          }
          if (++sa == (uint32_t)AS) { sa = 0; pha ^= 1u; }
          if (++sb == (uint32_t)BS) { sb = 0; phb ^= 1u; }
          if (++gnx == (uint32_t)L::NG) gnx = 0;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue warps 4..7: TMEM lane quarter q = warp % 4 =====
    const int q = warp & 3;
    const int r = lane >> 3, c = lane & 7;          // TMEM lane 8r + c of the quarter <-> pixel x = 4 cx + r of row RQ q + 2 cy + h
    const int cx = c % NX, cy = c / NX;
    const float act_lo = p.relu ? 0.f : -INFINITY;
    const int64_t P = (int64_t)p.H * p.W;
    for (int j = 0; j < my_pairs; ++j) {
      const int pair = blockIdx.x + j * gridDim.x;
      int b, x0, y0;
      pair_origin(pair, b, x0, y0);
      const int n0 = (pair % p.npass) * N_TILE;       // first output channel of this pass
      const uint32_t pb = (L::ACC_PAIRS == 2) ? (uint32_t)(j & 1) : 0u;
      const uint32_t use = (L::ACC_PAIRS == 2) ? (uint32_t)(j >> 1) : (uint32_t)j;
      if (warp == 4) DT_INC(11);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        DT_T(te0);
        mbar_wait(&tmem_full[pb * 2 + h], use & 1u);
        DT_T(te1);
        if (warp == 4) DT_ADD(9, te0, te1);
        if (warp == 4) DT_TR((uint32_t)j * (uint32_t)nch, 10 + 2 * h);
        tc_fence_after();
        const int gy = y0 + RQ * q + 2 * cy + h, gx = x0 + 4 * cx + r;
        const bool pvalid = (gy < p.H) && (gx < p.W);
        float* ypix = p.y + (int64_t)b * p.y_bstride + (int64_t)gy * p.W + gx;
        float oc_dot = 0.f;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (pb * 2 + h) * N_TILE;
        // Accumulator columns in groups of 32, double-buffered in registers: the TMEM load of group g + 1 is in flight while
        // group g is stored, and the accumulator is handed back to the MMA issuer as soon as its LAST group sits in registers
        // (before that group's stores) -- with one accumulator buffer per half tile (N_TILE = 128) the drain is exposed.
        constexpr int NGRP = N_TILE / 32;
        const int ngrp = min(NGRP, (p.Cout - n0 + 31) >> 5);         // warp-uniform, >= 1
        const bool no_store = DT_FLAG(2);
        uint32_t v2[2][32];
        tmem_ld32(tacc, v2[0]);
#pragma unroll
        for (int gi = 0; gi < NGRP; ++gi) {
          if (gi < ngrp) {
            const int c0 = 32 * gi;
            uint32_t(&v)[32] = v2[gi & 1];
            tmem_ld_wait();
            if (gi + 1 < ngrp) {
              tmem_ld32(tacc + (uint32_t)(c0 + 32), v2[(gi + 1) & 1]);
            } else {
              tc_fence_before();
              mbar_arrive(&tmem_empty[pb * 2 + h]);
            }
            const int nchn = min(32, p.Cout - (n0 + c0));      // warp-uniform
            float* yp = ypix + (int64_t)(n0 + c0) * P;
            const float* sc_p = aff + n0 + c0;                  // the affine of these 32 channels: broadcast LDS.128, 4 channels at a time
            const float* sh_p = aff + L::AFF_N + n0 + c0;
            if (p.oc_y) {
              // channels past Cout: zero accumulators, identity affine, zero OutConv weight -> no mask needed
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 a = *reinterpret_cast<const float4*>(sc_p + 4 * j4), t = *reinterpret_cast<const float4*>(sh_p + 4 * j4);
                const float4 w4 = *reinterpret_cast<const float4*>(aff + 2 * L::AFF_N + n0 + c0 + 4 * j4);
                oc_dot = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), a.x, t.x), act_lo), w4.x, oc_dot);
                oc_dot = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), a.y, t.y), act_lo), w4.y, oc_dot);
                oc_dot = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), a.z, t.z), act_lo), w4.z, oc_dot);
                oc_dot = fmaf(fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), a.w, t.w), act_lo), w4.w, oc_dot);
              }
            } else if (nchn == 32) {
              // hot path: FFMA, FMNMX, pointer bump, STG per channel; each store instruction of the warp = 128 (2 x 64) contiguous bytes
              if (pvalid && !no_store) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 a = *reinterpret_cast<const float4*>(sc_p + 4 * j4), t = *reinterpret_cast<const float4*>(sh_p + 4 * j4);
                  yp[0] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 0]), a.x, t.x), act_lo);
                  yp[P] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 1]), a.y, t.y), act_lo);
                  yp[2 * P] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 2]), a.z, t.z), act_lo);
                  yp[3 * P] = fmaxf(fmaf(__uint_as_float(v[4 * j4 + 3]), a.w, t.w), act_lo);
                  yp += 4 * P;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (pvalid && j < nchn) yp[(int64_t)j * P] = fmaxf(fmaf(__uint_as_float(v[j]), sc_p[j], sh_p[j]), act_lo);
            }
          }
        }
        if (p.oc_y && pvalid) p.oc_y[(int64_t)b * P + (int64_t)gy * p.W + gx] = oc_dot + (p.oc_b ? __ldg(p.oc_b) : 0.f);
        DT_T(te2);
        if (warp == 4) DT_ADD(10, te1, te2);
        if (warp == 4) DT_TR((uint32_t)j * (uint32_t)nch, 11 + 2 * h);
      }
    }
  } else if (warp >= 8) {
    // ===== depthwise producers: 2 groups of 4 warps; group g takes units u with u % 2 == g =====
    // Thread (c = lane / 4, ph = lane % 4) of the warp with TMEM quarter q owns, in BOTH half tiles h = 0, 1, the four
    // consecutive pixels x = 4 cx .. 4 cx + 3 of row RQ q + 2 cy + h (cx = c % NX, cy = c / NX): TMEM lane 8r + c <-> pixel
    // x = 4 cx + r.  A quarter is then whole 32-pixel (or two 16-pixel) row segments: the epilogue warp's store of one
    // channel is 128 (2 x 64) contiguous bytes.  Per input channel: 4 input rows x (one LDS.128 + edge values by shuffle
    // from the neighbouring lanes of the same channel phase) feed 2 output rows x 4 pixels x 2 depthwise outputs = 72
    // packed FFMA2 (fma.rn.f32x2: the two outputs of a channel share the input value, weights come interleaved).
    const int g = (warp - 8) >> 2;
    const int q = warp & 3;
    const int c = lane >> 2, ph = lane & 3;
    const int cx = c % NX, cy = c / NX;
    const bool lb = (cx == 0), rb = (cx == NX - 1);
    const uint32_t lane_base = ((uint32_t)(q * 32) << 16);
    // box row of pair row y is y + 1, box column of pair column x is x + 4: this thread's window starts at input row
    // RQ q + 2 cy (box rows +0 .. +3), its LDS.128 at box column 4 cx + 4, the edge values at 4 cx + 3 / 4 cx + 8
    const int win_off = (RQ * q + 2 * cy) * BW + 4 * cx + 4;
    const int edge_off = lb ? -1 : 4;
    uint32_t u = 0, aph = 0, iph = 0;   // phase bit per A stage (hand-back barriers) / per input stage (fill barriers) of THIS group
    for (int j = 0; j < my_pairs; ++j) {
      for (int i = 0; i < nch; ++i, ++u) {
        if ((int)(u % (uint32_t)L::NG) != g) continue;
        const int s = u % IS, sa = u % AS;
        DT_T(tp0);
        mbar_wait(&in_full[s * L::NG + g], (iph >> s) & 1u);
        iph ^= 1u << s;
        DT_T(tp1);
        if ((warp & 3) == 0) DT_TR(u, 1);
        const float* in_stage = reinterpret_cast<const float*>(smem + s * L::IN_BYTES);
        const float* wd = reinterpret_cast<const float*>(smem + L::OFF_WD) + (size_t)i * CC * L::WD_FLOATS;
        uint64_t acc[4][2][4];       // [channel i4][half / output row h][pixel] = (depthwise output 2 ci, 2 ci + 1)
        if (DT_FLAG(4)) {
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int x = 0; x < 4; ++x) acc[i4][h][x] = (uint64_t)(u + x);
        } else
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int ci = 4 * i4 + ph;
          const ulonglong2* wv = reinterpret_cast<const ulonglong2*>(wd + ci * L::WD_FLOATS);
          uint64_t w[9], bias;       // w[tap] = (weight of output 2 ci, weight of output 2 ci + 1)
          {
            const ulonglong2 t0 = wv[0], t1 = wv[1], t2 = wv[2], t3 = wv[3], t4 = wv[4];
            w[0] = t0.x; w[1] = t0.y; w[2] = t1.x; w[3] = t1.y; w[4] = t2.x; w[5] = t2.y; w[6] = t3.x; w[7] = t3.y; w[8] = t4.x;
            bias = t4.y;
          }
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[i4][h][x] = bias;
          const float* src = in_stage + ci * CHS + win_off;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const float4 a = *reinterpret_cast<const float4*>(src + rr * BW);
            float left = __shfl_up_sync(0xffffffffu, a.w, 4), right = __shfl_down_sync(0xffffffffu, a.x, 4);
            if (lb | rb) {
              const float e = src[rr * BW + edge_off];
              if (lb) left = e; else right = e;
            }
            const float v[6] = {left, a.x, a.y, a.z, a.w, right};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int dy = rr - h;         // output row h reads input rows h .. h + 2
              if (dy < 0 || dy > 2) continue;
#pragma unroll
              for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) acc[i4][h][x] = fma2_bcast(w[3 * dy + dx], v[x + dx], acc[i4][h][x]);
            }
          }
        }
        // the A stage is needed only now: the stencil of this unit overlapped the MMAs still reading the stage
        if ((warp & 3) == 0) DT_TR(u, 2);
        if (u >= (uint32_t)AS) {       // the stage had a reader: wait for this group's own hand-back barrier of the stage
          mbar_wait(&a_empty[sa * L::NG + g], (aph >> sa) & 1u);
          aph ^= 1u << sa;
        }
        DT_T(tp2);
        if ((warp & 3) == 0) DT_TR(u, 3);
        tc_fence_after();
        const uint32_t a_st = a_ring + (uint32_t)(sa * L::AST_COLS) + lane_base;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              // 16-lane group g2 of the quarter: TMEM lanes c + 8 (2 g2) and c + 8 (2 g2 + 1) = pixels x = 2 g2, 2 g2 + 1;
              // K columns 8 i4 + 2 ph + {0, 1}
              const uint32_t t = a_st + (uint32_t)(h * L::AH_COLS + 8 * i4) + ((uint32_t)(16 * g2) << 16);
              float o0, o1, o2, o3;
              unpack2(acc[i4][h][2 * g2], o0, o1);
              unpack2(acc[i4][h][2 * g2 + 1], o2, o3);
              if (X3) {
                const float h0 = tf32_hi(o0), h1 = tf32_hi(o1), h2 = tf32_hi(o2), h3 = tf32_hi(o3);
                tmem_st_16x256b_x1(t, __float_as_uint(h0), __float_as_uint(h1), __float_as_uint(h2), __float_as_uint(h3));
                tmem_st_16x256b_x1(t + 32u, __float_as_uint(o0 - h0), __float_as_uint(o1 - h1), __float_as_uint(o2 - h2),
                                   __float_as_uint(o3 - h3));
              } else {
                tmem_st_16x256b_x1(t, __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
              }
            }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a_full[sa]);
        mbar_arrive(&in_empty[s]);
        DT_T(tp3);
        if ((warp & 3) == 0) DT_TR(u, 4);
        if (warp == 8) { DT_ADD(0, tp0, tp1); DT_ADD(1, tp1, tp2); DT_ADD(2, tp2, tp3); DT_INC(3); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (rec0 && warp == 0) atomicAdd(&g_dt_timing[12], (unsigned long long)(clock64() - t_kernel0));
  if (rec_cta) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_dt_cta[3 * blockIdx.x + 1] = dt_globaltimer();
    g_dt_cta[3 * blockIdx.x + 2] = smid;
  }
#undef DT_T
#undef DT_ADD
#undef DT_INC
#undef DT_TR
#undef DT_FLAG
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int N_TILE, int PW, bool X3, bool BIG>
static int launch_dt(const CUtensorMap& m0, const CUtensorMap& m1, const CUtensorMap& mw, const CUtensorMap& mwl, DtParams p, int B,
                     cudaStream_t st) {
  using L = DtCfg<N_TILE, PW, X3, BIG>;
  auto kern = dsconv_tmem_kernel<N_TILE, PW, X3, BIG>;
  static std::atomic<uint64_t> attr_mask{0};   // cudaFuncSetAttribute is per device
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::FIXED + L::WD_MAX_BYTES);
    if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "dsconv(tmem): smem attribute (%d B): %s", L::FIXED + L::WD_MAX_BYTES, cudaGetErrorString(e));
  }
  p.px_tiles = ceil_div(p.W, PW);
  p.py_tiles = ceil_div(p.H, L::PHP);
  p.npass = ceil_div(p.Cout, N_TILE);
  const int64_t total = (int64_t)B * p.px_tiles * p.py_tiles * p.npass;
  SMAAT_REQUIRE(total < (1ll << 31), "dsconv(tmem): too many tiles");
  p.total_pairs = (int)total;
  p.nchunks = ceil_div(p.C0 + p.C1, L::CC);
  const int grid = p.total_pairs < num_sms() ? p.total_pairs : num_sms();
  kern<<<grid, L::THREADS, L::FIXED + p.nchunks * L::CC * L::WD_FLOATS * 4, st>>>(m0, m1, mw, mwl, p);
  SMAAT_LAUNCH_CHECK("smaat_dsconv_fwd(tmem)");
  return SMAAT_OK;
}

// pair shape: 32 x 8 or 16 x 16 pixels, whichever wastes fewer MMA rows; 0 = not worth it
static int pick_pw_pair(int H, int W) {
  double best = 1e9;
  int pw = 0;
  const int cand[2] = {32, 16};
  for (int i = 0; i < 2; ++i) {
    const int c = cand[i], php = 8 * (4 / (c / 8));
    const double waste = ((double)ceil_div(W, c) * c / W) * ((double)ceil_div(H, php) * php / H);
    if (waste < best - 1e-9) {
      best = waste;
      pw = c;
    }
  }
  return best <= 1.35 ? pw : 0;
}

bool dsconv_tmem_eligible(const float* x0, int C0, int64_t bs0, const float* x1, int C1, int64_t bs1, const float* pw_w,
                          const float* pw_w_lo, int H, int W, int k, int Cout) {
  if (k != 2) return false;
  if (Cout < 8 || Cout > 512 || (Cout > 128 && Cout % 128 != 0)) return false;   // > 128: whole passes of 128 channels
  if (W % 4 != 0 || !aligned16(x0) || bs0 % 4 != 0) return false;
  if (C1 > 0 && (!aligned16(x1) || bs1 % 4 != 0 || C0 % 16 != 0)) return false;
  if (((C0 + C1 + 15) / 16) * 16 > 512) return false;     // the depthwise weights of all channels stay resident in shared memory
  const int K = k * (C0 + C1);
  if (K % 4 != 0 || !aligned16(pw_w) || (pw_w_lo && !aligned16(pw_w_lo))) return false;
  return pick_pw_pair(H, W) != 0;
}

int dsconv_tmem_run(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride, const float* dw_w,
                    const float* dw_b, const float* pw_w, const float* pw_w_lo, const float* scale, const float* shift, float* y,
                    int64_t y_bstride, const float* oc_w, const float* oc_b, float* oc_y, int B, int H, int W, int Cout, int relu, int mode,
                    cudaStream_t st) {
  SMAAT_REQUIRE(!oc_y || Cout <= 128, "dsconv+outconv: the fused OutConv needs all Cout <= 128 channels in one pass");
  const int pw = pick_pw_pair(H, W);
  const int php = 8 * (4 / (pw / 8));
  const int n_tile = Cout > 64 ? 128 : 64;
  const bool x3 = mode == SMAAT_PW_TF32X3;
  const int K = 2 * (C0 + C1);

  CUtensorMap m0, m1, mw, mwl;
  const uint32_t box[4] = {(uint32_t)(pw + 8), (uint32_t)(php + 3), 16u, 1u};
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
  DtParams p;
  p.dw_w = dw_w; p.dw_b = dw_b; p.scale = scale; p.shift = shift; p.y = y; p.y_bstride = y_bstride;
  p.oc_w = oc_w; p.oc_b = oc_b; p.oc_y = oc_y;
  p.C0 = C0; p.C1 = C1; p.H = H; p.W = W; p.Cout = Cout; p.relu = relu; p.K = K;
  p.px_tiles = p.py_tiles = p.total_pairs = p.nchunks = p.npass = 0;
  static const int timing_on = [] { const char* e = getenv("SMAAT_DSCONV_TIMING"); return e ? atoi(e) : 0; }();
  p.timing = timing_on;
  static const int flags_on = [] { const char* e = getenv("SMAAT_DT_FLAGS"); return e ? atoi(e) : 0; }();
  p.flags = flags_on;

  const bool big = ((C0 + C1 + 15) / 16) * 16 > 256;
#define DT_DISPATCH(NT, PWv)                                                                               \
  return x3 ? (big ? launch_dt<NT, PWv, true, true>(m0, m1, mw, mwl, p, B, st)                             \
                   : launch_dt<NT, PWv, true, false>(m0, m1, mw, mwl, p, B, st))                           \
            : (big ? launch_dt<NT, PWv, false, true>(m0, m1, mw, mwl, p, B, st)                            \
                   : launch_dt<NT, PWv, false, false>(m0, m1, mw, mwl, p, B, st))
  if (n_tile == 64) {
    if (pw == 32) { DT_DISPATCH(64, 32); } else { DT_DISPATCH(64, 16); }
  } else {
    if (pw == 32) { DT_DISPATCH(128, 32); } else { DT_DISPATCH(128, 16); }
  }
#undef DT_DISPATCH
}

}  // namespace smaat

/* Debug hook: the event trace of CTA 0 of the last launch made with SMAAT_DSCONV_TIMING=2 (16 clock64 stamps per unit, layout
 * at g_dt_trace) for the first `n_units` <= 256 units, to the HOST array `out`.  Synchronises the device. */
extern "C" int smaat_debug_dsconv_tmem_trace(long long* out, int n_units) {
  using namespace smaat;
  SMAAT_REQUIRE(out && n_units > 0 && n_units <= DT_TRACE_UNITS, "debug_dsconv_tmem_trace: bad arguments");
  cudaError_t e = cudaMemcpyFromSymbol(out, g_dt_trace, sizeof(long long) * 16 * n_units);
  if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "debug_dsconv_tmem_trace: %s", cudaGetErrorString(e));
  return SMAAT_OK;
}

/* Debug hook: per-CTA (start ns, end ns, SM id) triples of the last timed launch of the TMEM-operand kernel, `n_ctas` <= 256 of
 * them, to the HOST array `out` (3 * n_ctas entries).  Synchronises the device. */
extern "C" int smaat_debug_dsconv_tmem_cta_timing(unsigned long long* out, int n_ctas) {
  using namespace smaat;
  SMAAT_REQUIRE(out && n_ctas > 0 && n_ctas <= DT_MAX_CTAS, "debug_dsconv_tmem_cta_timing: bad arguments");
  cudaError_t e = cudaMemcpyFromSymbol(out, g_dt_cta, sizeof(unsigned long long) * 3 * n_ctas);
  if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "debug_dsconv_tmem_cta_timing: %s", cudaGetErrorString(e));
  return SMAAT_OK;
}

/* Debug hook: copies the TMEM-operand kernel's stage timers of CTA 0 (24 counters, clock64 cycles; layout above) to the HOST
 * array `out` and clears them.  Synchronises the device. */
extern "C" int smaat_debug_dsconv_tmem_timing(unsigned long long* out) {
  using namespace smaat;
  SMAAT_REQUIRE(out, "debug_dsconv_tmem_timing: null pointer");
  cudaError_t e = cudaMemcpyFromSymbol(out, g_dt_timing, sizeof(g_dt_timing));
  if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "debug_dsconv_tmem_timing: %s", cudaGetErrorString(e));
  unsigned long long z[24] = {0};
  e = cudaMemcpyToSymbol(g_dt_timing, z, sizeof(z));
  if (e != cudaSuccess) return fail(SMAAT_E_CUDA, "debug_dsconv_tmem_timing: %s", cudaGetErrorString(e));
  return SMAAT_OK;
}
