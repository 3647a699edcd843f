// Tensor-core (tcgen05) variants of the fused evaluator for the 512-wide (xxxlarge) and 256-wide (xxlarge) models.
//
// Same per-tile pipeline as nfb_ffma.cuh -- featurise + Mahalanobis in fp64, every layer on chip, fused
// un-flip / average / decode epilogue -- but each layer is a tcgen05.mma (kind::f16) GEMM:
//
//   D[batch rows][features] (fp32, TMEM)  +=  A[batch rows][K] (fp16, smem)  x  B[features][K] (fp16, smem)
//
//   * A = the activation tile (queries in the first half of the rows, their flipped twins in the second), K-major,
//     SWIZZLE_128B atoms (8 rows x 128 B), written by the epilogue warps themselves (thread = batch row, 16-byte
//     conflict-free stores).
//   * B = weights, fp16 with a per-layer power-of-two scale, pre-packed on the host as K-major no-swizzle core
//     matrices, H/2 output features x 32 k per slice, streamed L2 -> smem by a producer thread through a 4-deep TMA
//     ring; the kernel runs as clusters of two CTAs that each fetch half of every slice and multicast it to both.
//   * D in tensor memory, two N-blocks of H/2 features; the whole MMA warp runs the issue loop, one elected lane issues.
//   * Layers are software-pipelined in halves (see the barrier comment in the kernel); the activation tile stays
//     single-buffered.
//
// Two precisions (template parameter SPLIT):
//   SPLIT = 1  "tensor":  operands rounded to fp16 (TF32's 10-bit mantissa), one MMA per product, 128 rows per tile,
//              SiLU by tanh.approx.  First layer fp32-exact by K-concatenation [x_hi | x_lo | x_hi].[W_hi | W_hi | W_lo].
//              TF32-class accuracy (latent error median 3e-4); tolerance: tests/parity.py, "tensor".
//   SPLIT = 3  "tensor_fp32":  every operand split into two fp16 halves, a = a_hi + a_lo (22 significant bits), and
//              each product evaluated as a_hi.w_hi + a_lo.w_hi + a_hi.w_lo with fp32 accumulation -- three MMAs.
//              A NumPy emulation gives error statistics indistinguishable from an fp32 evaluation, so this
//              variant is held to the fp32 tolerance.  Both activation halves must be on chip: 2 x 2 bytes per element,
//              so a 512-wide tile is 64 rows (M = 64 MMAs: same 128 cycles as M = 128, i.e. half rate -- measured,
//              scripts/umma_probe_m64.cu), a 256-wide tile 128 rows.  Accurate SiLU (expf), as the FFMA kernel.
//
// Why these shapes (measured with scripts/umma_probe*.cu on B200): with A in shared memory an
// M = 128 MMA never takes less than ~115 cycles however small N is, so N must be large, which puts the batch on the
// TMEM lanes; an MMA issued from a lone `if (lane == 0)` thread costs ~220 cycles in address arithmetic and election
// loops, so the issue loop is warp-converged with uniform descriptors (128 cycles per M128 N256 K16 MMA = the peak).
//
// Roles (320 threads): warp 0 lane 0 = TMA producer; warp 1 = TMEM allocator + MMA issue; warps 2..9 = epilogue
// (thread = batch row; the two warps of a TMEM lane quarter split each half's features).
#pragma once
#include <cuda_fp16.h>

#include "nfb_ffma.cuh"  // featurise_query, silu, decode_and_store, PTX helpers

namespace nfb {
namespace tc {

constexpr int SLICE_K = 32;            // k per streamed weight slice
constexpr int STAGES = 4;
constexpr int ATOM_K = 64;             // fp16 per 128-byte swizzle row
constexpr int THREADS = 320;
constexpr int EPI_WARPS = 8, EPI_THREADS = 256;
constexpr int MAX_TC_LAYERS = 7;
constexpr int LAST_ROWS = 256;         // the last layer's 224 packed rows, padded to whole N-blocks

template <int H_, int SPLIT_>
struct TcCfg {
  static constexpr int H = H_;
  static constexpr int SPLIT = SPLIT_;
  static constexpr int ROWS = (SPLIT == 3 && H == 512) ? 64 : 128;   // batch rows per tile == MMA M
  static constexpr int NQT = ROWS / 2;                               // queries per tile
  static constexpr int K0 = (SPLIT == 3) ? 32 : 96;                  // first layer's K ([hi | lo | hi] concatenation when SPLIT = 1)
  static constexpr int SLICE_N = H / 2;                              // output features per slice == MMA N (256 or 128)
  static constexpr uint32_t SLICE_BYTES = SLICE_N * SLICE_K * 2;     // 16 or 8 KB
  static constexpr int SLOTS_PER_STEP = (SPLIT == 3) ? 2 : 1;        // ring slots per (N-block, k-slice): W_hi then W_lo
  static constexpr int LAST_BLOCKS = LAST_ROWS / SLICE_N;            // N-blocks of the last layer (1 or 2)
  static constexpr uint32_t ATOM_BYTES = ROWS * 128;                 // ROWS/8 row groups x 1024 B
  static constexpr uint32_t A_BYTES = (H / ATOM_K) * ATOM_BYTES;     // one fp16 copy of the activation tile
  static constexpr uint32_t ACT_BYTES = A_BYTES * (SPLIT == 3 ? 2 : 1);  // [A_hi][A_lo]
  static constexpr uint32_t STAGE_BYTES = M_LAST * ROWS * 4;         // last layer's latents, fp32 [224][ROWS]
  // SPLIT = 3 keeps the two small product terms in an accumulator of their own (D_lo): the tensor core truncates
  // (rounds toward zero) at every accumulation, so a bias of ~half an ulp of D builds up per MMA; with the ~2^-11
  // smaller a_lo.w_hi / a_hi.w_lo blocks accumulated separately, the big accumulator sees 32 truncations per
  // 512-long dot product instead of 96 (measured: median relative error 1.3e-5 -> see DESIGN.md).  D_lo lives in the
  // lanes an M = 64 tile leaves free (lane + 16, "interleaved" allocation), or in the upper 256 columns (H = 256).
  static constexpr uint32_t D_LO_OFFSET = (SPLIT != 3) ? 0u : (ROWS == 64 ? (16u << 16) : 256u);
  static constexpr uint32_t TMEM_COLS = (SPLIT == 3 && ROWS == 128) ? 512 : H;   // power of two >= 32
  static constexpr int PART = H / 4;                                 // features per epilogue warp per half (128 or 64)
  // shared-memory map (bytes from a 1024-aligned base)
  static constexpr uint32_t OFF_ACT = 0;
  static constexpr uint32_t OFF_RING = OFF_ACT + ACT_BYTES;
  static constexpr uint32_t OFF_BIAS = OFF_RING + STAGES * SLICE_BYTES;   // fp32 [L][H] (+ scales [L] at MAX_TC_LAYERS * H)
  static constexpr uint32_t BIAS_FLOATS = MAX_TC_LAYERS * H + 16;
  static constexpr uint32_t OFF_MAHA = OFF_BIAS + BIAS_FLOATS * 4;        // fp32 [ROWS]
  static constexpr uint32_t OFF_RE = OFF_MAHA + ROWS * 4;                 // fp32 [NQT]
  static constexpr uint32_t OFF_BARS = OFF_RE + NQT * 4;                  // mbarriers
  static constexpr uint32_t OFF_TMEM = OFF_BARS + 16 * 8;
  // the staging block re-uses the idle activation tile when that is big enough, else gets its own room
  static constexpr uint32_t OFF_STAGE = (ACT_BYTES >= STAGE_BYTES) ? OFF_ACT : ((OFF_TMEM + 16 + 127) & ~127u);
  static constexpr uint32_t SMEM_BYTES = ((ACT_BYTES >= STAGE_BYTES) ? OFF_TMEM + 16 : OFF_STAGE + STAGE_BYTES) + 1024;
  static_assert(H == 512 || H == 256, "tensor-core variants: 512- or 256-wide models");
  static_assert(SPLIT == 1 || SPLIT == 3, "one MMA per product, or the three-term fp16 split");
  static_assert(SMEM_BYTES <= 232448, "shared-memory budget");

  __host__ __device__ static constexpr int k_slices(int l) { return (l == 0 ? K0 : H) / SLICE_K; }
  __host__ __device__ static constexpr int n_blocks(int l, int L) { return l == L - 1 ? LAST_BLOCKS : 2; }
  __host__ __device__ static constexpr int slices_in_layer(int l, int L) { return k_slices(l) * n_blocks(l, L) * SLOTS_PER_STEP; }
};

struct TcParams {
  EvalParams ev;            // inputs / outputs / n / n_layers (blob unused)
  const __half* wblob;      // slices in streaming order, Cfg::SLICE_BYTES each
  const float* bias;        // [L][H] fp32, last layer in packed row order (224 used)
  const float* inv_scale;   // [L]
  long long* dbg;           // optional: clock64 stamps of CTA 0's second tile (scripts/gpu_debug_tc.py), else NULL
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// commit that arrives on the same-offset mbarrier of every CTA in `mask` (weight slices are shared by the CTA pair)
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// TMA bulk copy global -> the same smem offset of every CTA in `mask`, completing on each one's mbarrier
__device__ __forceinline__ void bulk_copy_g2s_multicast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// two fp32 -> packed fp16x2 (lo = first argument), round to nearest, saturate to +-65504
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// SiLU for the fp16-operand path: x * sigmoid(x) = h + h * tanh(h), h = x / 2; one MUFU (tanh.approx, rel. error 2^-11,
// the precision the result is about to be rounded to anyway)
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// byte offset of the 16-byte chunk holding k..k+7 of batch row n in an activation copy (SWIZZLE_128B K-major atoms)
template <class Cfg>
__device__ __forceinline__ uint32_t act_chunk_offset(int n, int k) {
  const int atom = k >> 6, chunk = (k & 63) >> 3;
  return uint32_t(atom) * Cfg::ATOM_BYTES + uint32_t(n >> 3) * 1024 + uint32_t(n & 7) * 128 + uint32_t((chunk ^ (n & 7)) << 4);
}

// ---- featurise one batch row (nfb::featurise_query) -> fp16 hi / lo halves of its 32 (25 used) first-layer inputs ----
__device__ __forceinline__ void featurise_pack(const EvalParams& p, long long q, bool valid, bool twin,
                                               uint32_t (&hi)[16], uint32_t (&lo)[16], float& m, float& re) {
  double x[N_IN];
  featurise_query(p, q, valid, twin, x, m, re);
  // hi / lo split in fp16 (22 significant bits together)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float f0 = (2 * i < N_IN) ? static_cast<float>(x[2 * i]) : 0.0f;
    float f1 = (2 * i + 1 < N_IN) ? static_cast<float>(x[2 * i + 1]) : 0.0f;
    const __half h0 = __float2half_rn(f0), h1 = __float2half_rn(f1);
    hi[i] = uint32_t(__half_as_ushort(h0)) | (uint32_t(__half_as_ushort(h1)) << 16);
    const float l0 = (2 * i < N_IN) ? static_cast<float>(x[2 * i] - double(__half2float(h0))) : 0.0f;
    const float l1 = (2 * i + 1 < N_IN) ? static_cast<float>(x[2 * i + 1] - double(__half2float(h1))) : 0.0f;
    lo[i] = pack_f16x2(l0, l1);
  }
}

// ... and write the first layer's fp16 input row n into the activation tile:
//   SPLIT = 1: [x_hi | x_lo | x_hi] in k = 0..95 of the single copy;  SPLIT = 3: x_hi -> A_hi, x_lo -> A_lo, k = 0..31
template <class Cfg>
__device__ __forceinline__ void featurise_row(const EvalParams& p, long long q, bool valid, bool twin, int n,
                                              unsigned char* act, float* maha, float* re_s) {
  uint32_t hi[16], lo[16];
  float m, re;
  featurise_pack(p, q, valid, twin, hi, lo, m, re);
  maha[n] = m;
  if (!twin) re_s[n] = re;
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) {
    const uint4 vh = make_uint4(hi[4 * c8], hi[4 * c8 + 1], hi[4 * c8 + 2], hi[4 * c8 + 3]);
    const uint4 vl = make_uint4(lo[4 * c8], lo[4 * c8 + 1], lo[4 * c8 + 2], lo[4 * c8 + 3]);
    if constexpr (Cfg::SPLIT == 1) {  // k = 0..31 hi, 32..63 lo, 64..95 hi
      *reinterpret_cast<uint4*>(act + act_chunk_offset<Cfg>(n, 8 * c8)) = vh;
      *reinterpret_cast<uint4*>(act + act_chunk_offset<Cfg>(n, 32 + 8 * c8)) = vl;
      *reinterpret_cast<uint4*>(act + act_chunk_offset<Cfg>(n, 64 + 8 * c8)) = vh;
    } else {
      *reinterpret_cast<uint4*>(act + act_chunk_offset<Cfg>(n, 8 * c8)) = vh;
      *reinterpret_cast<uint4*>(act + Cfg::A_BYTES + act_chunk_offset<Cfg>(n, 8 * c8)) = vl;
    }
  }
}

// Launched as clusters of two CTAs (two SMs of one TPC).  The pair walks the same slice sequence; each CTA's
// producer fetches HALF of every weight slice and multicasts it into both CTAs' rings, which halves the L2 -> SM
// weight traffic.  Every CTA runs the same number of tiles (surplus tiles are all-padding), so the two rings never
// diverge.
template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) nfb_tc_kernel(const TcParams tp) {
  constexpr int H = Cfg::H, SLICE_N = Cfg::SLICE_N, PART = Cfg::PART, ROWS = Cfg::ROWS, NQT = Cfg::NQT, SPLIT = Cfg::SPLIT;
  constexpr uint32_t SLICE_BYTES = Cfg::SLICE_BYTES, ATOM_BYTES = Cfg::ATOM_BYTES;
  const EvalParams& p = tp.ev;
  extern __shared__ unsigned char smem_unaligned[];
  // (1024-aligned by pointer arithmetic on the __shared__ array, not by integer round-trip: the latter makes every
  //  access below a generic LD / ST instead of LDS / STS)
  unsigned char* const smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* const act = smem + Cfg::OFF_ACT;   // A_hi (the only copy when SPLIT = 1), then A_lo
  unsigned char* const ring = smem + Cfg::OFF_RING;
  float* const bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* const maha = reinterpret_cast<float*>(smem + Cfg::OFF_MAHA);
  float* const re_s = reinterpret_cast<float*>(smem + Cfg::OFF_RE);
  const uint32_t bar_full = smem_u32(smem + Cfg::OFF_BARS), bar_empty = bar_full + 8 * STAGES;
  // Layer pipeline barriers.  "Half" h = features [H/2 h, H/2 h + H/2) = TMEM columns of N-block h = K region h
  // (the second half of the activation atoms when h = 1) of the next layer.  The MMA warp walks a layer as four quarters
  //   q1 = (N-block 0, K region 0), q2 = (N-block 0, K region 1), q3 = (N-block 1, K region 0), q4 = (N-block 1, K region 1)
  // so that the epilogue of half 0 overlaps q3/q4 and the next layer's q1 overlaps the epilogue of half 1.
  const uint32_t bar_a0 = bar_empty + 8 * STAGES;   // epilogue -> MMA: K region 0 of the layer input is in smem, N-block 0's TMEM drained
  const uint32_t bar_a1 = bar_a0 + 8;               // ... K region 1, N-block 1's TMEM
  const uint32_t bar_d0 = bar_a1 + 8;               // MMA -> epilogue: N-block 0 complete (after q2)
  const uint32_t bar_k0 = bar_d0 + 8;               // MMA -> epilogue: K region 0 no longer read (after q3)
  const uint32_t bar_d1 = bar_k0 + 8;               // MMA -> epilogue: N-block 1 complete, K region 1 no longer read (after q4)
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const long long n_tiles = (p.n + NQT - 1) / NQT;
  const uint32_t my_tiles = uint32_t((n_tiles + gridDim.x - 1) / gridDim.x);  // same for every CTA
  const uint32_t cta_rank = cluster_ctarank();
  int slices_per_tile = 0;
  for (int l = 0; l < L; ++l) slices_per_tile += Cfg::slices_in_layer(l, L);

  for (int i = tid; i < L * H; i += THREADS) bias_s[i] = tp.bias[i];
  if (tid < L) bias_s[MAX_TC_LAYERS * H + tid] = tp.inv_scale[tid];
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 2);  // both CTAs of the pair have consumed the slot
    }
    mbar_init(bar_a0, EPI_WARPS);
    mbar_init(bar_a1, EPI_WARPS);
    mbar_init(bar_d0, 1);
    mbar_init(bar_k0, 1);
    mbar_init(bar_d1, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  cluster_sync_all();  // the peer's barriers exist before anything is multicast at them

  if (warp == 0) {
    // ============================ weight producer ============================
    if (lane == 0) {
      const uint32_t total = my_tiles * uint32_t(slices_per_tile);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(tp.wblob);
      for (uint32_t it = 0, j = 0; it < total; ++it) {
        const uint32_t s = it & (STAGES - 1), ph = (it / STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        mbar_arrive_expect_tx(bar_full + 8 * s, SLICE_BYTES);  // my half + the peer's half
        bulk_copy_g2s_multicast(smem_u32(ring + s * SLICE_BYTES) + cta_rank * (SLICE_BYTES / 2),
                                src + size_t(j) * SLICE_BYTES + cta_rank * (SLICE_BYTES / 2), SLICE_BYTES / 2,
                                bar_full + 8 * s, uint16_t(3));
        if (++j == uint32_t(slices_per_tile)) j = 0;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // The whole warp runs this loop converged and one elected lane issues: descriptors then live in uniform registers.
    const uint32_t idesc = (1u << 4) | (uint32_t(SLICE_N >> 3) << 17) | (uint32_t(ROWS >> 4) << 24);  // f16 x f16 -> f32
    // descriptor high words: A = SWIZZLE_128B K-major (SBO 1024 B), B = no-swizzle core matrices (LBO 128 B, SBO 512 B)
    const uint64_t a_hi = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
    const uint64_t b_hi = (uint64_t(128 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46);
    // (14-bit start-address field; in a cluster the 32-bit shared address carries the CTA rank in its top bits)
    const uint32_t a0 = (smem_u32(act) >> 4) & 0x3FFF, b0 = (smem_u32(ring) >> 4) & 0x3FFF;
    uint32_t it = 0, ph_a0 = 0, ph_a1 = 0;
    auto issue = [&](int nb, int ks_begin, int ks_end) {  // K-slices [ks_begin, ks_end) of N-block nb, in streaming order
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        const uint32_t a_lo = a0 + uint32_t(ks >> 1) * (ATOM_BYTES >> 4) + uint32_t(ks & 1) * 4;  // 64 B per 32 k
#pragma unroll
        for (int half = 0; half < Cfg::SLOTS_PER_STEP; ++half, ++it) {  // SPLIT = 3: the W_hi slot, then the W_lo slot
          const uint32_t s = it & (STAGES - 1), ph = (it / STAGES) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t b_lo = b0 + s * (SLICE_BYTES >> 4);
          const uint32_t d = tmem + nb * SLICE_N, d_small = d + Cfg::D_LO_OFFSET;
          if (elect_one_sync()) {
            if (SPLIT == 1 || half == 0) {
              umma_f16(d, a_hi | a_lo, b_hi | b_lo, idesc, ks != 0);
              umma_f16(d, a_hi | (a_lo + 2), b_hi | (b_lo + 16), idesc, 1u);
              if (SPLIT == 3) {  // a_lo . w_hi -> D_lo
                constexpr uint32_t LO = Cfg::A_BYTES >> 4;
                umma_f16(d_small, a_hi | (a_lo + LO), b_hi | b_lo, idesc, ks != 0);
                umma_f16(d_small, a_hi | (a_lo + LO + 2), b_hi | (b_lo + 16), idesc, 1u);
              }
            } else {             // a_hi . w_lo -> D_lo
              umma_f16(d_small, a_hi | a_lo, b_hi | b_lo, idesc, 1u);
              umma_f16(d_small, a_hi | (a_lo + 2), b_hi | (b_lo + 16), idesc, 1u);
            }
            umma_commit_multicast(bar_empty + 8 * s, uint16_t(3));  // slot is free once BOTH CTAs' MMAs have read it
          }
          __syncwarp();
        }
      }
    };
    auto commit = [&](uint32_t bar) {
      if (elect_one_sync()) umma_commit(bar);
      __syncwarp();
    };
    for (uint32_t t = 0; t < my_tiles; ++t)
      for (int l = 0; l < L; ++l) {
        const int kt = Cfg::k_slices(l);                       // K-slices of the layer
        const int kh = (l == 0) ? kt : kt / 2;                 // of which in K region 0 (the first layer's inputs all are)
        mbar_wait(bar_a0, ph_a0);
        ph_a0 ^= 1;
        tc_fence_after();
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32 + 2 * l] = clock64();
        issue(0, 0, kh);                                       // q1
        mbar_wait(bar_a1, ph_a1);
        ph_a1 ^= 1;
        tc_fence_after();
        issue(0, kh, kt);                                      // q2
        if (l != L - 1) {
          commit(bar_d0);
          issue(1, 0, kh);                                     // q3
          commit(bar_k0);
          issue(1, kh, kt);                                    // q4
          commit(bar_d1);
        } else {                                               // last layer: 256 packed rows, one signal when all are done
          if (Cfg::LAST_BLOCKS == 2) issue(1, 0, kt);
          commit(bar_d0);
        }
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32 + 2 * l + 1] = clock64();
      }
  } else {
    // ============================ epilogue warps ============================
    const int e = warp - 2, q = warp & 3, part = e >> 2;  // part: which PART of a half's features this warp takes
    // batch row of this thread == TMEM lane 32 q + lane.  M = 64 tiles only occupy lanes 0..15 of each quarter.
    const bool has_row = (ROWS == 128) || lane < 16;
    const int n = (ROWS == 128) ? 32 * q + lane : 16 * q + (lane & 15);
    const int et = (e << 5) | lane;          // 0..255
    const uint32_t t_lane = tmem + (uint32_t(32 * q) << 16);
    const bool vec_ok =
        p.out_f64 ? ((p.out_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                  : ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    uint32_t ph_d0 = 0, ph_k0 = 0, ph_d1 = 0;
    // scale, bias, SiLU and fp16 packing of 32 accumulators: w = the fp16 activations (the hi halves when SPLIT = 3),
    // wl = their fp16 remainders (SPLIT = 3 only)
    auto activate = [&](const uint32_t (&v)[32], const float* b, float inv_s, uint32_t (&w)[16], uint32_t (&wl)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float y0 = fmaf(__uint_as_float(v[2 * i]), inv_s, b[2 * i]);
        const float y1 = fmaf(__uint_as_float(v[2 * i + 1]), inv_s, b[2 * i + 1]);
        if constexpr (SPLIT == 1) {
          w[i] = pack_f16x2(silu_fast(y0), silu_fast(y1));
        } else {
          const float x0 = silu(y0), x1 = silu(y1);
          w[i] = pack_f16x2(x0, x1);
          const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
          wl[i] = pack_f16x2(x0 - back.x, x1 - back.y);
        }
      }
    };
    // 32 accumulator columns of this thread's row; SPLIT = 3 adds the separately accumulated small terms (D_lo)
    auto load_acc = [&](uint32_t col, uint32_t (&v)[32]) {
      tmem_ld32(t_lane + col, v);
      if constexpr (SPLIT == 3) {
        if constexpr (ROWS == 64) {  // lanes 16..31 of the quarter hold D_lo of rows lane - 16
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = __float_as_uint(__uint_as_float(v[i]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[i]), 16));
        } else {
          uint32_t u[32];
          tmem_ld32(t_lane + col + Cfg::D_LO_OFFSET, u);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
        }
      }
    };
    auto store32 = [&](int k, const uint32_t (&w)[16], const uint32_t (&wl)[16]) {  // 32 consecutive features k..k+31 of row n
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t off = act_chunk_offset<Cfg>(n, k + 8 * c8);
        *reinterpret_cast<uint4*>(act + off) = make_uint4(w[4 * c8], w[4 * c8 + 1], w[4 * c8 + 2], w[4 * c8 + 3]);
        if constexpr (SPLIT == 3)
          *reinterpret_cast<uint4*>(act + Cfg::A_BYTES + off) = make_uint4(wl[4 * c8], wl[4 * c8 + 1], wl[4 * c8 + 2], wl[4 * c8 + 3]);
      }
    };
    auto publish = [&](uint32_t bar) {  // my smem writes -> async proxy, my TMEM reads are done; tell the MMA warp
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    for (uint32_t t = 0; t < my_tiles; ++t) {
      const long long tile = blockIdx.x + (long long)t * gridDim.x;  // may be past the end: an all-padding tile
      const long long tile_q0 = tile * NQT;
      const bool stamp = tp.dbg && blockIdx.x == 0 && t == 1 && e == 0 && lane == 0;
      if (stamp) tp.dbg[0] = clock64();
      if (part == 0 && has_row) {  // one thread per batch row (row n: query n, row NQT + n: its twin)
        const bool twin = n >= NQT;
        const long long qi = tile_q0 + (twin ? n - NQT : n);
        featurise_row<Cfg>(p, qi, qi < p.n, twin, n, act, maha, re_s);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_a0);
        mbar_arrive(bar_a1);
      }
      if (stamp) tp.dbg[1] = clock64();

      for (int l = 0; l < L - 1; ++l) {
        const float inv_s = bias_s[MAX_TC_LAYERS * H + l];
        const float* b = bias_s + l * H + part * PART;
        // ---- half 0: N-block 0 is complete after q2; its outputs may only land in K region 0 after q3 ----
        mbar_wait(bar_d0, ph_d0);
        ph_d0 ^= 1;
        tc_fence_after();
        if (stamp) tp.dbg[2 + 2 * l] = clock64();
        if constexpr (SPLIT == 1) {
          uint32_t w[PART / 32][16], unused[16];
#pragma unroll
          for (int c = 0; c < PART / 32; ++c) {
            uint32_t v[32];
            load_acc(uint32_t(part * PART + 32 * c), v);
            activate(v, b + 32 * c, inv_s, w[c], unused);
          }
          mbar_wait(bar_k0, ph_k0);
          ph_k0 ^= 1;
#pragma unroll
          for (int c = 0; c < PART / 32; ++c) store32(part * PART + 32 * c, w[c], unused);
        } else {
          // (hi and lo halves of a whole PART do not fit in registers: keep the first chunk, then stream)
          uint32_t w0[16], wl0[16];
          {
            uint32_t v[32];
            load_acc(uint32_t(part * PART), v);
            activate(v, b, inv_s, w0, wl0);
          }
          mbar_wait(bar_k0, ph_k0);
          ph_k0 ^= 1;
          if (has_row) store32(part * PART, w0, wl0);
#pragma unroll 1
          for (int c = 1; c < PART / 32; ++c) {
            uint32_t v[32], w[16], wl[16];
            load_acc(uint32_t(part * PART + 32 * c), v);
            activate(v, b + 32 * c, inv_s, w, wl);
            if (has_row) store32(part * PART + 32 * c, w, wl);
          }
        }
        publish(bar_a0);
        // ---- half 1: N-block 1 complete (and K region 1 free) after q4; overlaps the next layer's q1 ----
        mbar_wait(bar_d1, ph_d1);
        ph_d1 ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < PART / 32; ++c) {
          uint32_t v[32], w[16], wl[16];
          load_acc(uint32_t(H / 2 + part * PART + 32 * c), v);
          activate(v, b + H / 2 + 32 * c, inv_s, w, wl);
          if (has_row) store32(H / 2 + part * PART + 32 * c, w, wl);
        }
        publish(bar_a1);
        if (stamp) tp.dbg[3 + 2 * l] = clock64();
      }

      // ---- last layer: latents -> fp32 staging [224 packed rows][ROWS] (the idle activation tile when it fits) ----
      mbar_wait(bar_d0, ph_d0);
      ph_d0 ^= 1;
      tc_fence_after();
      if (stamp) tp.dbg[20] = clock64();
      {
        float* stage = reinterpret_cast<float*>(smem + Cfg::OFF_STAGE);
        const float inv_s = bias_s[MAX_TC_LAYERS * H + (L - 1)];
        const float* b = bias_s + (L - 1) * H;
        // feature ranges: part 0 -> [0, 128), part 1 -> [128, 224)
        const int f_begin = part * 128, f_end = part ? M_LAST : 128;
#pragma unroll 1
        for (int c0 = f_begin; c0 < f_end; c0 += 32) {
          uint32_t v[32];
          load_acc(uint32_t(c0), v);
          if (has_row) {
#pragma unroll
            for (int i = 0; i < 32; ++i) stage[(c0 + i) * ROWS + n] = fmaf(__uint_as_float(v[i]), inv_s, b[c0 + i]);
          }
        }
        tc_fence_before();
        named_bar_sync(1, EPI_THREADS);
        if (stamp) tp.dbg[21] = clock64();
        for (int item = et; item < (M_LAST / 4) * (NQT / 4); item += EPI_THREADS) {
          const int g = item / (NQT / 4), qb = (item % (NQT / 4)) * 4;
          float acc[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 y = *reinterpret_cast<const float4*>(stage + (4 * g + i) * ROWS + qb);
            const float4 f = *reinterpret_cast<const float4*>(stage + (4 * g + i) * ROWS + NQT + qb);
            acc[i][0] = y.x; acc[i][1] = y.y; acc[i][2] = y.z; acc[i][3] = y.w;
            acc[i][4] = f.x; acc[i][5] = f.y; acc[i][6] = f.z; acc[i][7] = f.w;
          }
          const long long q0 = tile_q0 + qb;
          const long long left = p.n - q0;
          const int nvalid = left >= 4 ? 4 : (left > 0 ? int(left) : 0);
          decode_and_store(p, g, q0, nvalid, vec_ok, acc, maha + qb, maha + NQT + qb, re_s + qb);
        }
        named_bar_sync(1, EPI_THREADS);  // staging (and maha / re) fully consumed before the next tile overwrites them
        if (stamp) tp.dbg[22] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while the peer can still multicast into this CTA
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace nfb
