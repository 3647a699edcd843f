// Tensor-core evaluators on CTA PAIRS: tcgen05.mma.cta_group::2 (the kernels behind the "tensor" and "tensor_fp32"
// variants; the one-SM kernels of nfb_tc.cuh remain as the reference implementation of the same pipeline).
//
// The two CTAs of a cluster form ONE MMA: each CTA contributes its own batch rows of A (the activation tile) and HALF
// of the B slice (half of the output features of an N-block); the leader CTA issues for both tensor cores.
//
//   * Why (measured): (1) the 512-wide three-term kernel must keep both fp16 halves of the
//     activations on chip, which leaves room for 64 batch rows only, and a one-SM M = 64 MMA costs the same 128 cycles
//     as an M = 128 one; as half of a pair-wide M = 128 MMA the same 64 rows run in 64 cycles (scripts/umma_probe_2sm.cu:
//     4093 MAC/clk/SM).  (2) The one-SM kernels are shared-memory-bandwidth bound, not tensor bound: per 128-cycle MMA the
//     tensor core reads 4 KB of A and 8 KB of B, and TMA writes another 8 KB of weights -- 160 B/clk against the SM's
//     128.  A pair splits B: each SM stores and reads only its half.
//   * D layouts (measured, same probe).  Pair M = 256 (128 rows per CTA): row r, column n -> TMEM lane r, column n.
//     Pair M = 128 (64 rows per CTA): row r, column n -> lane r + 64 (n / (N/2)), column n % (N/2); all 128 lanes are
//     used, so every epilogue lane has work (the one-SM M = 64 layout idles half of them).
//   * Weights: the fp16 slice streams of the one-SM kernels (H/2 features x 32 k per slice; SPLIT = 3: a W_hi and a
//     W_lo slice per step).  CTA c fetches rows [c N/2, c N/2 + N/2) of each -- half the bytes, no multicast.  The peer
//     CTA's MMA warp has nothing to issue; it relays "my half of slot s has landed" to the leader's full barrier.
//   * Layer pipeline: the half-layer schedule of nfb_tc.cuh (quarters q1..q4), with the epilogue -> MMA barriers in
//     the leader CTA (both CTAs' epilogue warps arrive there) and the MMA -> epilogue barriers signalled in both CTAs
//     by multicast commits.  K region 1 of a layer is swept chunk-major (k_slice_at) and its readiness is signalled
//     chunk by chunk (barriers a1[0..NCH)), so q2 starts while the rest of the previous layer's N-block 1 is still
//     being activated.
//   * Featurisation runs ahead: two extra warps compute the NEXT tile's 25 inputs and Mahalanobis terms (fp64, ~7,000
//     latency-bound cycles per tile) while the current tile's layers run, and drop them into a small first-layer input
//     buffer of their own as soon as layer 0 of the current tile has consumed it.  The next tile's layer 0 then starts
//     the moment the last layer's accumulators are staged, overlapping the decode.
//   * All barriers use default (CTA-scope) semantics: a cluster-scope release compiles to MEMBAR.ALL.GPU and a
//     cluster-scope acquiring wait to CCTL.IVALL (an L1 invalidate) per wait -- measured 3x slower MMA issue.  Nothing
//     is handed over through the generic proxy across CTAs: the data these barriers order is written to the writer's own
//     shared memory, made visible to the async proxy by fence.proxy.async before the arrive, and read by that same SM's
//     tensor core.
#pragma once
#include "nfb_tc.cuh"

namespace nfb {
namespace tc2 {

using namespace nfb::tc;  // PTX wrappers, SLICE_K, ATOM_K, EPI_WARPS, MAX_TC_LAYERS, LAST_ROWS

constexpr int THREADS2 = 384;   // warp 0 producer, warp 1 MMA issue / relay, warps 2..9 epilogue, warps 10..11 featurise-ahead
constexpr int FW_WARP0 = 10, FW_WARPS = 2;
constexpr int N_BARS = 40;

template <int H_, int SPLIT_>
struct Tc2Cfg {
  static constexpr int H = H_;
  static constexpr int SPLIT = SPLIT_;
  static constexpr int ROWS = (SPLIT == 3 && H == 512) ? 64 : 128;   // batch rows per CTA; the MMA's M = 2 ROWS spans the pair
  static constexpr int NQT = ROWS / 2;                               // queries per CTA tile (rows NQT.. are the twins)
  static constexpr int K0 = (SPLIT == 3) ? 32 : 128;                 // first layer's K (SPLIT = 1: [x_hi | x_lo | x_hi | -] . [W_hi | W_hi | W_lo | 0])
  static constexpr int BLOCK_N = H / 2;                              // MMA N = features per N-block
  static constexpr int HALF_N = BLOCK_N / 2;                         // B rows held by each CTA
  // Weight stream (pack_model_tc2 in nfb_api.cu): per CTA rank, the ring slots of one tile in the order the kernel
  // consumes them.  A slot = SPS consecutive positions of a K sweep, each position my half (HALF_N rows x 32 k, K-major
  // no-swizzle core matrices) of the W_hi slice and, when SPLIT = 3, of the W_lo slice.  One bulk copy per slot: issuing
  // a copy costs the producer ~260 cycles whatever its size (scripts/tma_mma_probe.cu), more than a one-term position's
  // MMAs last, hence two positions per slot there.
  static constexpr int SPS = (SPLIT == 3) ? 1 : 2;
  static constexpr uint32_t PIECE_BYTES = HALF_N * SLICE_K * 2;      // my rows of one slice
  static constexpr uint32_t POS_BYTES = (SPLIT == 3 ? 2 : 1) * PIECE_BYTES;
  static constexpr uint32_t SLOT_BYTES = SPS * POS_BYTES;
  static constexpr uint32_t ATOM_BYTES = ROWS * 128;
  static constexpr uint32_t A_BYTES = (H / ATOM_K) * ATOM_BYTES;
  static constexpr uint32_t ACT_BYTES = A_BYTES * (SPLIT == 3 ? 2 : 1);  // [A_hi][A_lo]
  static constexpr uint32_t REGION_BYTES = ACT_BYTES;
  // first-layer input rows [x_hi | x_lo] at k = 0..63 of one atom, outside the activation tile so that they can be
  // written a tile ahead.  (SPLIT = 1 reads them as the K = 96 concatenation [x_hi | x_lo | x_hi]: its third k-slice
  // points at the first again.)
  static constexpr uint32_t IN_BYTES = ATOM_BYTES;
  // (Measured: reading the 512-wide models' 14 KB of biases through L1 instead of from shared memory, to buy weight-ring
  //  slots, loses 5 %: with 227 KB of shared memory carved out the L1 is too small to keep them.)
  // TMEM: D_hi of N-block nb at column nb * TILE_COLS; SPLIT = 3 keeps the two small product terms in accumulators of
  // their own (the tensor core truncates at every accumulation, see nfb_tc.cuh) at + 256
  static constexpr uint32_t TILE_COLS = (ROWS == 64) ? BLOCK_N / 2 : BLOCK_N;
  static constexpr uint32_t D_LO_OFFSET = (SPLIT == 3) ? 256u : 0u;
  static constexpr uint32_t TMEM_COLS = (SPLIT == 3) ? 512 : H;
  // epilogue geometry: G thread groups along the features of an N-block, each thread NCH chunks of 32 features per half
  static constexpr int G = (ROWS == 64) ? 4 : 2;
  static constexpr int NCH = BLOCK_N / (32 * G);
  static constexpr uint32_t BIAS_FLOATS = MAX_TC_LAYERS * H + 16;   // biases [L][H], then 1 / scale [L]
  static constexpr uint32_t FIXED_BYTES = REGION_BYTES + IN_BYTES + BIAS_FLOATS * 4 + 2 * ROWS * 4 + 2 * NQT * 4 + N_BARS * 8 + 16 + 1024;
  static constexpr int RING_FIT = int((232448 - FIXED_BYTES) / SLOT_BYTES);
  static constexpr int RING = RING_FIT > 10 ? 10 : RING_FIT;          // weight ring depth: what fits, at most 10
  static constexpr uint32_t OFF_ACT = 0;
  static constexpr uint32_t OFF_IN = REGION_BYTES;
  static constexpr uint32_t OFF_RING = OFF_IN + IN_BYTES;
  static constexpr uint32_t OFF_BIAS = OFF_RING + RING * SLOT_BYTES;
  static constexpr uint32_t OFF_MAHA = OFF_BIAS + BIAS_FLOATS * 4;    // fp32 [2][ROWS], by tile parity
  static constexpr uint32_t OFF_RE = OFF_MAHA + 2 * ROWS * 4;         // fp32 [2][NQT]
  static constexpr uint32_t OFF_BARS = OFF_RE + 2 * NQT * 4;
  static constexpr uint32_t OFF_TMEM = OFF_BARS + N_BARS * 8;
  static constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
  static_assert(H == 512 || H == 256, "tensor-core variants: 512- or 256-wide models");
  static_assert(SPLIT == 1 || SPLIT == 3, "one MMA per product, or the three-term fp16 split");
  static_assert(NCH == 2 || NCH == 4, "chunks per thread per half");
  static_assert(RING >= 3 && 2 * RING + 12 <= N_BARS, "weight ring / barrier block");
  static_assert(SMEM_BYTES <= 232448, "shared-memory budget");

  __host__ __device__ static constexpr int k_slices(int l) { return (l == 0 ? K0 : H) / SLICE_K; }
  __host__ __device__ static constexpr int n_blocks(int l, int L) { return l == L - 1 ? LAST_ROWS / BLOCK_N : 2; }
  __host__ __device__ static constexpr int slots_in_layer(int l, int L) { return k_slices(l) * n_blocks(l, L) / SPS; }
  // Position in a layer's K sweep -> k-slice.  K region 1 (the second half of the inputs) is swept chunk-major: first the
  // G slices that hold the chunk every epilogue thread of the previous layer writes first, then those of the second
  // chunk, ... so the sweep can start on a chunk while the later ones are still being activated (barriers a1[chunk]).
  // Thread group g writes slices g NCH .. g NCH + NCH - 1 of the region, one per chunk.
  __host__ __device__ static constexpr int k_slice_at(int pos, int kt) {
    if (kt != H / SLICE_K || pos < kt / 2) return pos;  // (the first layer's few slices all count as region 0)
    const int R = kt / 2, j = pos - R;
    return R + (j % G) * NCH + j / G;
  }
  static_assert(G % SPS == 0, "a chunk's slices are whole ring slots");
};

__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the same-offset mbarrier of both CTAs once every MMA issued so far (on both tensor cores) has finished
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(uint16_t(3)) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive on a barrier of another CTA of the cluster (default semantics on purpose, see the header comment)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// SiLU x / (1 + e^-x) from the two MUFU approximations directly (ex2.approx: 2 ulp, rcp.approx: 1 ulp) -- the three-term
// epilogue is MUFU / issue bound, and its operands are about to be split into 22-bit fp16 pairs anyway.
__device__ __forceinline__ float silu_ex2(float h) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fminf(-h, 80.0f) * 1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return h * r;
}

// Un-flip + average + squash + decode of ONE query's group g (packed rows 4 g .. 4 g + 3 of the last layer, see
// nfb_common.cuh: pack_last_layer_row) from the query's latents y and its twin's f, stored as scalars.  The same
// formulas as nfb::decode_and_store [main.py:274-316]; used by the pair kernels, which decode straight from the
// accumulators (query and twin on neighbouring TMEM lanes) instead of staging the latents in shared memory.
__device__ __forceinline__ void decode_group_store1(const EvalParams& p, int g, long long q, const float (&y)[4], const float (&f)[4],
                                                    float maha_o, float maha_t, float re) {
  float o[4];
  int row[4], n_out = 4;
  if (g < N_BL) {  // theta_u, ue_u, theta_l, ue_l of station g
    const float zt_u = 0.5f * (y[0] + f[2]), zu_u = 0.5f * (y[1] - f[3]);
    const float zt_l = 0.5f * (y[2] + f[0]), zu_l = 0.5f * (y[3] - f[1]);
    o[0] = (exp10f(zt_u) - 0.1f) / (fabsf(zu_u) * re);
    o[1] = zu_u;
    o[2] = (exp10f(zt_l) - 0.1f) / (fabsf(zu_l) * re);
    o[3] = zu_l;
    row[0] = 6 + g; row[1] = 6 + 2 * N_BL + g; row[2] = 6 + 3 * N_BL + g; row[3] = 6 + 5 * N_BL + g;
  } else if (g < 48) {  // H_u, H_l of stations s, s + 1
    const int s = 2 * (g - N_BL);
    o[0] = 2.6f * expf(0.5f * (y[0] + f[2]));
    o[1] = 2.6f * expf(0.5f * (y[1] + f[3]));
    o[2] = 2.6f * expf(0.5f * (y[2] + f[0]));
    o[3] = 2.6f * expf(0.5f * (y[3] + f[1]));
    row[0] = 6 + N_BL + s; row[1] = 6 + N_BL + s + 1; row[2] = 6 + 4 * N_BL + s; row[3] = 6 + 4 * N_BL + s + 1;
  } else if (g == 48) {  // analysis_confidence, CL, CD, CM
    const float z0 = 0.5f * ((y[0] - maha_o) + (f[0] - maha_t));
    const float z1 = 0.5f * (y[1] - f[1]), z2 = 0.5f * (y[2] + f[2]), z3 = 0.5f * (y[3] - f[3]);
    o[0] = 1.0f / (1.0f + expf(-z0));
    o[1] = 0.5f * z1;
    o[2] = expf((z2 - 2.0f) * 2.0f);
    o[3] = z3 / 20.0f;
    row[0] = 0; row[1] = 1; row[2] = 2; row[3] = 3;
  } else if (g == 49) {  // Top_Xtr, Bot_Xtr
    o[0] = clip01_keep_nan(0.5f * (y[0] + f[1]));
    o[1] = clip01_keep_nan(0.5f * (y[1] + f[0]));
    o[2] = o[3] = 0.0f;
    row[0] = 4; row[1] = 5; row[2] = row[3] = 0;
    n_out = 2;
  } else {
    return;  // padding rows
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < n_out) {
      if (p.out_bl != nullptr && row[k] >= 6)
        __stcs(static_cast<unsigned short*>(p.out_bl) + (row[k] - 6) * p.out_ld + q, static_cast<unsigned short>(pack_bf16x2(o[k], 0.0f) & 0xFFFFu));
      else if (p.out_f64) __stcs(static_cast<double*>(p.out) + row[k] * p.out_ld + q, double(o[k]));
      else __stcs(static_cast<float*>(p.out) + row[k] * p.out_ld + q, o[k]);
    }
}

__device__ __forceinline__ void store1(const EvalParams& p, int row, long long q, float v) {
  if (p.out_bl != nullptr && row >= 6) {  // mixed-precision outputs: boundary-layer rows as bfloat16
    __stcs(static_cast<unsigned short*>(p.out_bl) + (row - 6) * p.out_ld + q, static_cast<unsigned short>(pack_bf16x2(v, 0.0f) & 0xFFFFu));
    return;
  }
  if (p.out_f64) __stcs(static_cast<double*>(p.out) + row * p.out_ld + q, double(v));
  else __stcs(static_cast<float*>(p.out) + row * p.out_ld + q, v);
}
// Four decode groups of the same kind at once (g0 .. g0 + 3, all boundary-layer theta / ue groups or all H groups): the
// sixteen outputs are independent chains, which is what hides the MUFU and division latencies with two warps per
// scheduler.  exp / exp10 / division by the fast intrinsics (relative error < 1e-6, two orders of magnitude inside the
// decode tolerance of the tensor variants; the fp32 kernels keep the accurate library versions).
__device__ __forceinline__ void decode_groups4_store(const EvalParams& p, int g0, long long q, const float (&y)[4][4], const float (&f)[4][4],
                                                     float re) {
  float o[4][4];
  if (g0 < N_BL) {  // theta_u, ue_u, theta_l, ue_l of stations g0 .. g0 + 3
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
      const float zt_u = 0.5f * (y[gg][0] + f[gg][2]), zu_u = 0.5f * (y[gg][1] - f[gg][3]);
      const float zt_l = 0.5f * (y[gg][2] + f[gg][0]), zu_l = 0.5f * (y[gg][3] - f[gg][1]);
      o[gg][0] = __fdividef(exp2f(zt_u * 3.3219280948873623f) - 0.1f, fabsf(zu_u) * re);
      o[gg][1] = zu_u;
      o[gg][2] = __fdividef(exp2f(zt_l * 3.3219280948873623f) - 0.1f, fabsf(zu_l) * re);
      o[gg][3] = zu_l;
    }
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
      store1(p, 6 + g0 + gg, q, o[gg][0]);
      store1(p, 6 + 2 * N_BL + g0 + gg, q, o[gg][1]);
      store1(p, 6 + 3 * N_BL + g0 + gg, q, o[gg][2]);
      store1(p, 6 + 5 * N_BL + g0 + gg, q, o[gg][3]);
    }
  } else {  // H_u, H_l of stations s, s + 1 for groups g0 .. g0 + 3 (all in [32, 48))
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
      o[gg][0] = 2.6f * __expf(0.5f * (y[gg][0] + f[gg][2]));
      o[gg][1] = 2.6f * __expf(0.5f * (y[gg][1] + f[gg][3]));
      o[gg][2] = 2.6f * __expf(0.5f * (y[gg][2] + f[gg][0]));
      o[gg][3] = 2.6f * __expf(0.5f * (y[gg][3] + f[gg][1]));
    }
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) {
      const int s = 2 * (g0 + gg - N_BL);
      store1(p, 6 + N_BL + s, q, o[gg][0]);
      store1(p, 6 + N_BL + s + 1, q, o[gg][1]);
      store1(p, 6 + 4 * N_BL + s, q, o[gg][2]);
      store1(p, 6 + 4 * N_BL + s + 1, q, o[gg][3]);
    }
  }
}

// tcgen05.ld without the wait, and the wait as a separate statement that "produces" the registers, so that the compiler
// cannot move their consumers above it: lets the next chunk's TMEM read overlap the current chunk's arithmetic.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1) nfb_tc2_kernel(const TcParams tp) {
  constexpr int H = Cfg::H, ROWS = Cfg::ROWS, NQT = Cfg::NQT, SPLIT = Cfg::SPLIT, RING = Cfg::RING, NCH = Cfg::NCH;
  constexpr uint32_t SLOT_BYTES = Cfg::SLOT_BYTES, PIECE_BYTES = Cfg::PIECE_BYTES, ATOM_BYTES = Cfg::ATOM_BYTES;
  const EvalParams& p = tp.ev;
  extern __shared__ unsigned char smem_unaligned[];
  // (1024-aligned by pointer arithmetic on the __shared__ array, not by integer round-trip: the latter makes every
  //  access below a generic LD / ST instead of LDS / STS)
  unsigned char* const smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* const act = smem + Cfg::OFF_ACT;   // A_hi (the only copy when SPLIT = 1), then A_lo
  unsigned char* const inbuf = smem + Cfg::OFF_IN;  // the first layer's input rows (written a tile ahead)
  unsigned char* const ring = smem + Cfg::OFF_RING;
  float* const bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* const scale_s = bias_s + MAX_TC_LAYERS * H;
  float* const maha2 = reinterpret_cast<float*>(smem + Cfg::OFF_MAHA);
  float* const re2 = reinterpret_cast<float*>(smem + Cfg::OFF_RE);
  const uint32_t bar_full = smem_u32(smem + Cfg::OFF_BARS), bar_empty = bar_full + 8 * RING;
  // Layer pipeline barriers (nfb_tc.cuh).  a0 / a1[c] are waited on in the leader CTA only: K region 0 / chunk c of K region 1 /
  // the second half of K region 1's slices are in shared memory and the TMEM they came from is drained.
  const uint32_t bar_a0 = bar_empty + 8 * RING, bar_a1 = bar_a0 + 8, bar_d0 = bar_a1 + 8 * 4, bar_k0 = bar_d0 + 8,
                 bar_d1 = bar_k0 + 8;
  // Tile hand-over barriers.  in (leader): the featurise warps of both CTAs have written the tile's input rows.
  // drained (leader): the epilogue warps of both CTAs have read the last layer's accumulators.  infree (both, by commit):
  // layer 0 has consumed the input rows.  tiledone (local): this CTA's epilogue warps have finished the tile's decode.
  const uint32_t bar_in = bar_d1 + 8, bar_drained = bar_in + 8, bar_infree = bar_drained + 8, bar_tiledone = bar_infree + 8;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const long long n_tiles = (p.n + NQT - 1) / NQT;
  const uint32_t my_tiles = uint32_t((n_tiles + gridDim.x - 1) / gridDim.x);  // same for every CTA
  const uint32_t cta_rank = cluster_ctarank();
  int slots_per_tile = 0;
  for (int l = 0; l < L; ++l) slots_per_tile += Cfg::slots_in_layer(l, L);

  for (int i = tid; i < L * H; i += THREADS2) bias_s[i] = tp.bias[i];
  if (tid < L) scale_s[tid] = tp.inv_scale[tid];
  if (tid == 0) {
    for (int s = 0; s < RING; ++s) {
      mbar_init(bar_full + 8 * s, cta_rank == 0 ? 2 : 1);  // leader: its own producer + the peer's relay
      mbar_init(bar_empty + 8 * s, 1);                     // one multicast commit
    }
    mbar_init(bar_a0, 2 * EPI_WARPS);                      // the epilogue warps of both CTAs
    for (int c = 0; c < NCH; ++c) mbar_init(bar_a1 + 8 * c, 2 * EPI_WARPS);
    mbar_init(bar_d0, 1);
    mbar_init(bar_k0, 1);
    mbar_init(bar_d1, 1);
    mbar_init(bar_in, 2 * FW_WARPS);
    mbar_init(bar_drained, 2 * EPI_WARPS);
    mbar_init(bar_infree, 1);
    mbar_init(bar_tiledone, EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ============================ weight producer: my rank's slot stream, one bulk copy per slot ============================
    // Runs warp-converged with one elected lane issuing, so that the loop state lives in uniform registers.
    const unsigned char* src = reinterpret_cast<const unsigned char*>(tp.wblob) + size_t(cta_rank) * size_t(slots_per_tile) * SLOT_BYTES;
    const uint32_t total = my_tiles * uint32_t(slots_per_tile);
    uint32_t s = 0, ph = 0;
    long long wait_empty = 0;  // diagnostics (tp.dbg only)
    for (uint32_t it = 0, j = 0; it < total; ++it) {
      const bool timing = tp.dbg && it >= uint32_t(slots_per_tile) && it < 2u * uint32_t(slots_per_tile);
      const long long c0 = timing ? clock64() : 0;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (timing) wait_empty += clock64() - c0;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar_full + 8 * s, SLOT_BYTES);
        bulk_copy_g2s(smem_u32(ring + s * SLOT_BYTES), src + size_t(j) * SLOT_BYTES, SLOT_BYTES, bar_full + 8 * s);
      }
      __syncwarp();
      if (++j == uint32_t(slots_per_tile)) j = 0;
      if (++s == RING) { s = 0; ph ^= 1; }
    }
    if (tp.dbg && lane == 0 && blockIdx.x < 2) tp.dbg[51 + blockIdx.x] = wait_empty;
  } else if (warp == 1 && cta_rank != 0) {
    // ============================ peer CTA: relay "slot landed" to the leader ============================
    // One lane per ring slot, so that the round trip of a remote arrive does not pace the stream.
    if (lane < RING) {
      const uint32_t total = my_tiles * uint32_t(slots_per_tile);
      const uint32_t mine = bar_full + 8 * lane, leaders = map_to_rank(mine, 0);
      uint32_t ph = 0;
      for (uint32_t it = lane; it < total; it += RING, ph ^= 1) {
        mbar_wait(mine, ph);
        mbar_arrive_cluster(leaders);
      }
    }
  } else if (warp == 1) {
    // ============================ leader CTA: MMA issue for the pair ============================
    // The whole warp runs this loop converged and one elected lane issues: descriptors then live in uniform registers.
    const uint32_t idesc = (1u << 4) | (uint32_t(Cfg::BLOCK_N >> 3) << 17) | (uint32_t((2 * ROWS) >> 4) << 24);  // f16 x f16 -> f32, M over two CTAs
    // descriptor high words: A = SWIZZLE_128B K-major (SBO 1024 B), B = no-swizzle core matrices (LBO 128 B, SBO 512 B)
    const uint64_t a_hi = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
    const uint64_t b_hi = (uint64_t(128 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46);
    // (14-bit start-address field; in a cluster the 32-bit shared address carries the CTA rank in its top bits)
    const uint32_t a_act = (smem_u32(act) >> 4) & 0x3FFF, a_in = (smem_u32(inbuf) >> 4) & 0x3FFF, b0 = (smem_u32(ring) >> 4) & 0x3FFF;
    constexpr uint32_t WLO = PIECE_BYTES >> 4;
    uint32_t a0 = a_in, LO = 4;  // A base and distance of the lo half: layer 0 reads [x_hi | x_lo] of the input buffer, the others A_hi / A_lo
    uint32_t s = 0, ph = 0, ph_a0 = 0, ph_a1 = 0, ph_in = 0, ph_dr = 0;  // (the a1 barriers flip together)
    long long wait_full = 0, wait_a0 = 0, wait_a1 = 0;  // diagnostics (tp.dbg only): cycles this warp spent waiting
    const bool timing = tp.dbg && blockIdx.x == 0;
    // MMAs of one position of a K sweep (k-slice ks of the A tile; its weights at byte offset b_ofs >> 4 of the ring)
    // into N-block nb; called by the elected lane
    auto mma_pos = [&](int nb, int ks, uint32_t b_lo, bool first) {
      if (SPLIT == 1 && a0 == a_in) ks = (ks == 1);  // layer 0 of the one-term kernel: [x_hi | x_lo | x_hi | x_hi] . [W_hi | W_hi | W_lo | 0]
      const uint32_t a_lo = a0 + uint32_t(ks >> 1) * (ATOM_BYTES >> 4) + uint32_t(ks & 1) * 4;  // 64 B per 32 k
      const uint32_t d = tmem + nb * Cfg::TILE_COLS, d_small = d + Cfg::D_LO_OFFSET;
      umma_f16_pair(d, a_hi | a_lo, b_hi | b_lo, idesc, !first);                            // a_hi . w_hi -> D_hi
      umma_f16_pair(d, a_hi | (a_lo + 2), b_hi | (b_lo + 16), idesc, 1u);
      // (Measured and not taken: a_hi . w_hi and a_hi . w_lo back to back with the A slice kept in the collector
      //  (.collector::a::fill / ::lastuse, SASS A_KEEP / A_REUSE) -- correct, but no faster: A is not what binds.)
      if constexpr (SPLIT == 3) {
        umma_f16_pair(d_small, a_hi | (a_lo + LO), b_hi | b_lo, idesc, !first);             // a_lo . w_hi -> D_lo
        umma_f16_pair(d_small, a_hi | (a_lo + LO + 2), b_hi | (b_lo + 16), idesc, 1u);
        umma_f16_pair(d_small, a_hi | a_lo, b_hi | (b_lo + WLO), idesc, 1u);                // a_hi . w_lo -> D_lo
        umma_f16_pair(d_small, a_hi | (a_lo + 2), b_hi | (b_lo + WLO + 16), idesc, 1u);
      }
    };
    // Positions [pos_begin, pos_end) of N-block nb's K sweep (whole ring slots: SPS positions each)
    auto issue = [&](int nb, int kt, int pos_begin, int pos_end) {
      for (int pos = pos_begin; pos < pos_end; pos += Cfg::SPS) {
        const long long c0 = timing ? clock64() : 0;
        mbar_wait(bar_full + 8 * s, ph);
        if (timing) wait_full += clock64() - c0;
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t b_lo = b0 + s * (SLOT_BYTES >> 4);
#pragma unroll
          for (int u = 0; u < Cfg::SPS; ++u) mma_pos(nb, Cfg::k_slice_at(pos + u, kt), b_lo + u * (Cfg::POS_BYTES >> 4), pos + u == 0);
          umma_commit_pair(bar_empty + 8 * s);  // the slot is free in both CTAs once both tensor cores have read it
        }
        __syncwarp();
        if (++s == RING) { s = 0; ph ^= 1; }
      }
    };
    auto commit = [&](uint32_t bar) {
      if (elect_one_sync()) umma_commit_pair(bar);
      __syncwarp();
    };
    for (uint32_t t = 0; t < my_tiles; ++t) {
      {  // ---- layer 0: from the input buffer, as soon as it is written and the previous tile's accumulators are read ----
        const int kt = Cfg::k_slices(0);
        if (timing && t == 1) wait_full = wait_a0 = wait_a1 = 0;
        const long long c0 = timing ? clock64() : 0;
        mbar_wait(bar_in, ph_in);
        ph_in ^= 1;
        if (t != 0) {
          mbar_wait(bar_drained, ph_dr);
          ph_dr ^= 1;
        }
        if (timing) wait_a0 += clock64() - c0;
        tc_fence_after();
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32] = clock64();
        a0 = a_in;
        LO = 4;
        issue(0, kt, 0, kt);
        commit(bar_d0);
        issue(1, kt, 0, kt);
        commit(bar_k0);
        commit(bar_d1);
        commit(bar_infree);
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[33] = clock64();
        a0 = a_act;
        LO = Cfg::A_BYTES >> 4;
      }
      for (int l = 1; l < L; ++l) {
        const int kt = Cfg::k_slices(l);
        const int kh = kt / 2;                                 // slices in K region 0
        long long c0 = timing ? clock64() : 0;
        mbar_wait(bar_a0, ph_a0);
        if (timing) wait_a0 += clock64() - c0;
        ph_a0 ^= 1;
        tc_fence_after();
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32 + 2 * l] = clock64();
        issue(0, kt, 0, kh);                                   // q1
        for (int c = 0; c < NCH; ++c) {                        // q2, chunk by chunk as the previous layer's epilogue delivers
          c0 = timing ? clock64() : 0;
          mbar_wait(bar_a1 + 8 * c, ph_a1);
          if (timing) wait_a1 += clock64() - c0;
          tc_fence_after();
          issue(0, kt, kh + c * Cfg::G, kh + (c + 1) * Cfg::G);
        }
        ph_a1 ^= 1;
        if (l != L - 1) {
          commit(bar_d0);
          issue(1, kt, 0, kh);                                 // q3
          commit(bar_k0);
          issue(1, kt, kh, kt);                                // q4
          commit(bar_d1);
        } else {                                               // last layer: 256 packed rows, one signal when all are done
          if (Cfg::n_blocks(l, L) == 2) issue(1, kt, 0, kt);
          commit(bar_d0);
        }
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) {
          tp.dbg[32 + 2 * l + 1] = clock64();
          if (l == L - 1) { tp.dbg[48] = wait_full; tp.dbg[49] = wait_a0; tp.dbg[50] = wait_a1; }
        }
      }
    }
  } else if (warp >= FW_WARP0) {
    // ============================ featurise-ahead warps (both CTAs) ============================
    // Thread ft owns batch rows ft (, ft + 64): row 2 i is query i of the tile, row 2 i + 1 its flipped twin -- neighbouring
    // TMEM lanes, so that the final decode can pair them with a warp shuffle.
    constexpr int RPT = ROWS / (32 * FW_WARPS);
    const int ft = tid - 32 * FW_WARP0;
    const uint32_t leader_in = map_to_rank(bar_in, 0);
    uint32_t ph_free = 0, ph_done = 0;
    for (uint32_t t = 0; t < my_tiles; ++t) {
      const long long tile_q0 = (blockIdx.x + (long long)t * gridDim.x) * NQT;  // may be past the end: an all-padding tile
      float* const maha = maha2 + (t & 1) * ROWS;
      float* const re_s = re2 + (t & 1) * NQT;
#pragma unroll 1  // (one copy of the featurisation code: the kernel is instruction-cache heavy as it is)
      for (int j = 0; j < RPT; ++j) {
        const int n = ft + 64 * j;
        const bool twin = n & 1;
        const long long qi = tile_q0 + (n >> 1);
        uint32_t hi[16], lo[16];
        float m, re;
        featurise_pack(p, qi, qi < p.n, twin, hi, lo, m, re);
        if (j == 0) {
          if (t >= 1) {  // layer 0 of the previous tile has consumed the input rows
            mbar_wait(bar_infree, ph_free);
            ph_free ^= 1;
          }
          if (t >= 2) {  // the decode of the tile before that has finished with this parity's maha / re
            mbar_wait(bar_tiledone, ph_done);
            ph_done ^= 1;
          }
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const uint4 vh = make_uint4(hi[4 * c8], hi[4 * c8 + 1], hi[4 * c8 + 2], hi[4 * c8 + 3]);
          const uint4 vl = make_uint4(lo[4 * c8], lo[4 * c8 + 1], lo[4 * c8 + 2], lo[4 * c8 + 3]);
          *reinterpret_cast<uint4*>(inbuf + act_chunk_offset<Cfg>(n, 8 * c8)) = vh;
          *reinterpret_cast<uint4*>(inbuf + act_chunk_offset<Cfg>(n, 32 + 8 * c8)) = vl;
        }
        maha[n] = m;
        if (!twin) re_s[n >> 1] = re;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_in);
    }
  } else {
    // ============================ epilogue warps (both CTAs) ============================
    // ROWS = 128: TMEM lane 32 q + lane = batch row; the two warps of a lane quarter split an N-block's features.
    // ROWS = 64:  TMEM lane 32 q + lane = batch row (32 q + lane) % 64, holding features [BLOCK_N/2 (q / 2), ...) of the
    //             N-block in TILE_COLS columns; the two warps of a lane quarter split those.
    const int e = warp - 2, q = warp & 3, part = e >> 2;
    const int n = (ROWS == 64) ? 32 * (q & 1) + lane : 32 * q + lane;    // batch row in this CTA
    const int g = (ROWS == 64) ? 2 * (q >> 1) + part : part;             // feature group of this thread
    const int fofs = 32 * NCH * g;                                       // first of my features within an N-block
    const int fcol = (ROWS == 64) ? 32 * NCH * part : fofs;              // ... and their first column within the accumulator tile
    const uint32_t t_lane = tmem + (uint32_t(32 * q) << 16);
    const uint32_t leader_a0 = map_to_rank(bar_a0, 0), leader_a1 = map_to_rank(bar_a1, 0),
                   leader_drained = map_to_rank(bar_drained, 0);
    uint32_t ph_d0 = 0, ph_k0 = 0, ph_d1 = 0;
    // scale, bias, SiLU and fp16 packing of 32 accumulators: w = the fp16 activations (the hi halves when SPLIT = 3),
    // wl = their fp16 remainders (SPLIT = 3 only)
    auto activate = [&](const uint32_t (&v)[32], const float* b, float inv_s, uint32_t (&w)[16], uint32_t (&wl)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float y0 = fmaf(__uint_as_float(v[2 * i]), inv_s, b[2 * i]);
        const float y1 = fmaf(__uint_as_float(v[2 * i + 1]), inv_s, b[2 * i + 1]);
        if constexpr (SPLIT == 1) {
          // (Measured alternative: h + h tanh(h) in packed fp16 with tanh.approx.f16x2 -- one MUFU per two activations,
          //  2.5 instructions each -- makes the epilogue 12 % faster but doubles the variant's error: not taken.)
          w[i] = pack_f16x2(silu_fast(y0), silu_fast(y1));
        } else {
          const float x0 = silu_ex2(y0), x1 = silu_ex2(y1);
          w[i] = pack_f16x2(x0, x1);
          const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
          wl[i] = pack_f16x2(x0 - back.x, x1 - back.y);
        }
      }
    };
    auto load_acc = [&](uint32_t col, uint32_t (&v)[32]) {  // 32 accumulator columns of my lane; SPLIT = 3: D_hi + D_lo
      tmem_ld32(t_lane + col, v);
      if constexpr (SPLIT == 3) {
        uint32_t u[32];
        tmem_ld32(t_lane + col + Cfg::D_LO_OFFSET, u);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
      }
    };
    auto store32 = [&](int k, const uint32_t (&w)[16], const uint32_t (&wl)[16]) {  // features k..k+31 of row n
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t off = act_chunk_offset<Cfg>(n, k + 8 * c8);
        *reinterpret_cast<uint4*>(act + off) = make_uint4(w[4 * c8], w[4 * c8 + 1], w[4 * c8 + 2], w[4 * c8 + 3]);
        if constexpr (SPLIT == 3)
          *reinterpret_cast<uint4*>(act + Cfg::A_BYTES + off) = make_uint4(wl[4 * c8], wl[4 * c8 + 1], wl[4 * c8 + 2], wl[4 * c8 + 3]);
      }
    };
    auto publish = [&](uint32_t leader_bar) {  // my smem writes -> async proxy, my TMEM reads are done; tell the leader's MMA warp
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_bar);
    };
    for (uint32_t t = 0; t < my_tiles; ++t) {
      const long long tile = blockIdx.x + (long long)t * gridDim.x;  // may be past the end: an all-padding tile
      const long long tile_q0 = tile * NQT;
      const bool stamp = tp.dbg && blockIdx.x == 0 && t == 1 && warp == 4 && lane == 0;
      const float* const maha = maha2 + (t & 1) * ROWS;   // written by the featurise warps before this tile's layer 0 ran
      const float* const re_s = re2 + (t & 1) * NQT;
      if (stamp) tp.dbg[0] = tp.dbg[1] = clock64();

      for (int l = 0; l < L - 1; ++l) {
        const float inv_s = scale_s[l];
        const float* b = bias_s + l * H + fofs;
        // ---- half 0: N-block 0 complete after q2; its outputs may only land in K region 0 after q3 ----
        mbar_wait(bar_d0, ph_d0);
        ph_d0 ^= 1;
        tc_fence_after();
        if (stamp) tp.dbg[2 + 2 * l] = clock64();
        {
          uint32_t w[NCH][16], wl[SPLIT == 3 ? NCH : 1][16];
          if constexpr (SPLIT == 1) {  // the next chunk's TMEM read is in flight while this one is activated
            uint32_t v[2][32];
            tmem_ld32_issue(t_lane + fcol, v[0]);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              tmem_ld_wait(v[c & 1]);
              if (c + 1 < NCH) tmem_ld32_issue(t_lane + fcol + 32 * (c + 1), v[(c + 1) & 1]);
              activate(v[c & 1], b + 32 * c, inv_s, w[c], w[c]);
            }
          } else {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              uint32_t v[32];
              load_acc(uint32_t(fcol + 32 * c), v);
              activate(v, b + 32 * c, inv_s, w[c], wl[c]);
            }
          }
          if (stamp && l == 2) tp.dbg[53] = clock64();
          mbar_wait(bar_k0, ph_k0);
          ph_k0 ^= 1;
          if (stamp && l == 2) tp.dbg[54] = clock64();
#pragma unroll
          for (int c = 0; c < NCH; ++c) store32(fofs + 32 * c, w[c], wl[SPLIT == 3 ? c : 0]);
        }
        publish(leader_a0);
        if (stamp && l == 2) tp.dbg[55] = clock64();
        // ---- half 1: N-block 1 complete (and K region 1 free) after q4; overlaps the next layer's q1 and, for its
        //      second NCH / 2 chunks, the first part of q2 ----
        mbar_wait(bar_d1, ph_d1);
        ph_d1 ^= 1;
        tc_fence_after();
        if (stamp && l == 2) tp.dbg[56] = clock64();
        if constexpr (SPLIT == 1) {
          uint32_t v[2][32];
          tmem_ld32_issue(t_lane + Cfg::TILE_COLS + fcol, v[0]);
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            uint32_t w[16];
            tmem_ld_wait(v[c & 1]);
            if (c + 1 < NCH) tmem_ld32_issue(t_lane + Cfg::TILE_COLS + fcol + 32 * (c + 1), v[(c + 1) & 1]);
            activate(v[c & 1], b + Cfg::BLOCK_N + 32 * c, inv_s, w, w);
            store32(Cfg::BLOCK_N + fofs + 32 * c, w, w);
            publish(leader_a1 + 8 * c);
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            uint32_t v[32], w[16], wl[16];
            load_acc(uint32_t(Cfg::TILE_COLS + fcol + 32 * c), v);
            activate(v, b + Cfg::BLOCK_N + 32 * c, inv_s, w, wl);
            store32(Cfg::BLOCK_N + fofs + 32 * c, w, wl);
            publish(leader_a1 + 8 * c);
          }
        }
        if (stamp) tp.dbg[3 + 2 * l] = clock64();
      }

      // ---- last layer: decode straight from the accumulators.  My lane holds 32-row chunks of the packed latents of one
      //      batch row; lane ^ 1 holds the same chunks of its twin (or query).  Of a chunk's 8 decode groups the query's
      //      lane takes the first four and the twin's lane the last four, after swapping the halves they lack. ----
      mbar_wait(bar_d0, ph_d0);
      ph_d0 ^= 1;
      tc_fence_after();
      if (stamp) tp.dbg[20] = clock64();
      {
        const float inv_s = scale_s[L - 1];
        const float* b = bias_s + (L - 1) * H;
        // my packed rows: ROWS = 64: the N-block's [fofs, fofs + 64); ROWS = 128: [128 part, 128 part + 128) (columns = rows)
        const int f_begin = (ROWS == 64) ? fofs : 128 * part, f_count = (ROWS == 64) ? 32 * NCH : 128;
        const int c_begin = (ROWS == 64) ? fcol : f_begin;
        const bool odd = n & 1;                       // my row is a twin (n and lane have the same parity)
        const long long qg = tile_q0 + (n >> 1);      // the query my lane pair decodes
        const bool valid = qg < p.n;
        const float maha_o = maha[n & ~1], maha_t = maha[n | 1], re = re_s[n >> 1];
#pragma unroll 1
        for (int c = 0; c < f_count; c += 32) {
          const int f0 = f_begin + c;    // packed rows f0 .. f0 + 31
          if (f0 >= M_LAST_USED) break;  // (warp-uniform; the rest is padding)
          const bool skip = p.scalars_only && f0 < 4 * 48;  // (warp-uniform) boundary-layer chunks not wanted
          uint32_t v[32];
          if (!skip) load_acc(uint32_t(c_begin + c), v);
          if (c + 32 >= f_count || f0 + 32 >= M_LAST_USED) {  // that was my last read of the accumulators
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_drained);  // the next tile's layer 0 may overwrite them
            if (stamp) tp.dbg[21] = clock64();
          }
          if (skip) continue;
          float z[32], other[16];
#pragma unroll
          for (int i = 0; i < 32; ++i) z[i] = fmaf(__uint_as_float(v[i]), inv_s, b[f0 + i]);
#pragma unroll
          for (int k = 0; k < 16; ++k) other[k] = __shfl_xor_sync(0xffffffffu, odd ? z[k] : z[16 + k], 1);
          float y[4][4], f[4][4];
#pragma unroll
          for (int gg = 0; gg < 4; ++gg)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              y[gg][i] = odd ? other[4 * gg + i] : z[4 * gg + i];
              f[gg][i] = odd ? z[16 + 4 * gg + i] : other[4 * gg + i];
            }
          const int g0 = (f0 >> 2) + (odd ? 4 : 0);
          if (valid) {
            if (f0 < 4 * 48) {  // (warp-uniform) boundary-layer and H groups: four of a kind
              decode_groups4_store(p, g0, qg, y, f, re);
            } else {            // the scalars' chunk: groups 48 (confidence, CL, CD, CM), 49 (Xtr), padding
#pragma unroll
              for (int gg = 0; gg < 2; ++gg) decode_group_store1(p, g0 + gg, qg, y[gg], f[gg], maha_o, maha_t, re);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tiledone);
        if (stamp) tp.dbg[22] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair still works
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

}  // namespace tc2
}  // namespace nfb
