// Tensor-core (tcgen05) evaluators for the NARROW models: padded width 128 (large, xlarge -- the reference's default model,
// main.py:71) and 64 (xxsmall .. medium).  The "tensor" / "tensor_fp32" variants of nfb_tc2.cuh, re-cut for layers that are
// a single MMA wide.
//
// Why another kernel.  In nfb_tc2.cuh a layer is two N-blocks of H/2 features on a CTA pair, because at H >= 256 an N-block
// is >= 128 columns and an M = 128 MMA with its A operand in shared memory costs ~115 cycles however small N is
// (scripts/umma_probe*.cu).  At H = 128 that cut would issue N = 64 MMAs at the 115-cycle floor -- a quarter of the tensor
// core's rate -- so here a whole layer is ONE N-block: D[128 batch rows][H features] += A[128][K] . B[H][K], N = H (128 or 64),
// and the last layer's 208 packed rows are one N = 208 MMA.  One CTA per SM, cta_group::1, no cluster: every barrier is
// CTA-local.
//
//   * A = the activation tile, fp16, K-major SWIZZLE_128B atoms, rows 2 i / 2 i + 1 = query i of the tile / its flipped twin
//     (neighbouring TMEM lanes: the decode pairs them with one shuffle).  Written by the epilogue warps.
//   * B = weights, fp16 with a per-layer power-of-two scale, K-major no-swizzle core matrices, pre-packed on the host in the
//     order the kernel consumes them (nfb_api.cu: pack_model_tcn).  A ring slot is two "pieces" of (N rows x 32 k):
//     SPLIT = 1: two consecutive k-slices; SPLIT = 3: the W_hi and the W_lo piece of one k-slice.  One TMA bulk copy per slot
//     (issuing a copy costs ~260 cycles whatever its size: scripts/tma_mma_probe.cu), 5 slots deep.  The weights of these
//     models are 0.1 - 0.4 MB: they stay in L2 and every CTA streams them once per tile.
//   * D in tensor memory: columns 0..207; SPLIT = 3 keeps the two small product terms in an accumulator of their own at
//     column 256 (the tensor core truncates at every accumulation, nfb_tc.cuh).
//   * Precisions as in nfb_tc.cuh: SPLIT = 1 "tensor" (fp16 operands, first layer fp32-exact by K-concatenation
//     [x_hi | x_lo | x_hi | -] . [W_hi | W_hi | W_lo | 0]); SPLIT = 3 "tensor_fp32" (a_hi w_hi + a_lo w_hi + a_hi w_lo).
//   * Roles (448 threads): warp 0 weight producer; warp 1 MMA issue; warps 2..9 epilogue (tcgen05.ld -> scale, bias, SiLU
//     -> fp16 -> smem; the two warps of a TMEM lane quarter split the features) and, after the last layer, the decode
//     straight from the accumulators (nfb_tc2.cuh's); warps 10..13 featurise the NEXT tile (fp64, one row per thread) into a first-layer input
//     buffer of its own while the current tile's layers run.
//   * Schedule: a tile's layers run one after the other -- MMA sweep, then epilogue, then the next sweep -- with the next
//     tile's layer 0 starting as soon as the last accumulators are read, i.e. under the decode.  (Not yet: two tiles in
//     flight so that one tile's epilogue runs under the other's MMAs.)
// Exact invariants: query and twin go through the same MMAs as neighbouring rows, and the un-flip / average expressions are
// those of decode_and_store, so symmetric-airfoil, mirror and batch-position identities hold bit for bit, as on the other paths.
#pragma once
#include "nfb_tc2.cuh"

namespace nfb {
namespace tcn {

using namespace nfb::tc;   // PTX wrappers, SLICE_K, ATOM_K, MAX_TC_LAYERS, TcParams
using nfb::tc2::decode_group_store1;
using nfb::tc2::decode_groups4_store;
using nfb::tc2::silu_ex2;

constexpr int TCN_THREADS = 448;
constexpr int TCN_EPI_WARPS = 8, TCN_FW_WARP0 = 10, TCN_FW_WARPS = 4;
constexpr int TCN_LAST_N = 208;        // packed rows of the last layer (200 used), a multiple of 16
constexpr int TCN_BIAS_LD = 256;       // floats per layer in the bias block

template <int H_, int SPLIT_>
struct TcnCfg {
  static constexpr int H = H_;
  static constexpr int SPLIT = SPLIT_;
  static constexpr int ROWS = 128, NQT = 64;
  static constexpr int K0 = (SPLIT == 3) ? 32 : 128;                  // SPLIT = 1: [x_hi | x_lo | x_hi | -] . [W_hi | W_hi | W_lo | 0]
  static constexpr int SPS = (SPLIT == 3) ? 1 : 2;                    // k-slices per ring slot
  static constexpr uint32_t PIECE_MAX = TCN_LAST_N * SLICE_K * 2;     // 13,312 B: one (N x 32 k) piece of the last layer
  static constexpr uint32_t SLOT_BYTES = 2 * PIECE_MAX;               // two pieces per slot (slot stride; copies are sized per layer)
  static constexpr int RING = 5;
  static constexpr uint32_t ATOM_BYTES = ROWS * 128;
  static constexpr uint32_t A_BYTES = (H / ATOM_K) * ATOM_BYTES;
  static constexpr uint32_t ACT_BYTES = A_BYTES * (SPLIT == 3 ? 2 : 1);
  static constexpr uint32_t IN_BYTES = ATOM_BYTES;                    // [x_hi | x_lo] at k = 0..63 of one atom
  static constexpr uint32_t D_LO_OFFSET = (SPLIT == 3) ? 256u : 0u;
  static constexpr uint32_t TMEM_COLS = (SPLIT == 3) ? 512 : 256;
  static constexpr int NCH = H / 64;                                  // 32-feature chunks per epilogue thread per hidden layer
  static constexpr uint32_t BIAS_FLOATS = MAX_TC_LAYERS * TCN_BIAS_LD + 16;
  static constexpr uint32_t OFF_ACT = 0;
  static constexpr uint32_t OFF_IN = ACT_BYTES;
  static constexpr uint32_t OFF_RING = OFF_IN + IN_BYTES;
  static constexpr uint32_t OFF_BIAS = OFF_RING + RING * SLOT_BYTES;
  static constexpr uint32_t OFF_MAHA = OFF_BIAS + BIAS_FLOATS * 4;    // fp32 [2][ROWS], by tile parity
  static constexpr uint32_t OFF_RE = OFF_MAHA + 2 * ROWS * 4;         // fp32 [2][NQT]
  static constexpr uint32_t OFF_BARS = OFF_RE + 2 * NQT * 4;
  static constexpr uint32_t OFF_TMEM = OFF_BARS + (2 * RING + 8) * 8;
  static constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
  static_assert(H == 128 || H == 64, "narrow tensor-core kernels: 128- or 64-wide models");
  static_assert(SPLIT == 1 || SPLIT == 3, "one MMA per product, or the three-term fp16 split");
  static_assert(SMEM_BYTES <= 232448, "shared-memory budget");

  __host__ __device__ static constexpr int n_of(int l, int L) { return l == L - 1 ? TCN_LAST_N : H; }          // MMA N of layer l
  __host__ __device__ static constexpr int k_slices(int l) { return (l == 0 ? K0 : H) / SLICE_K; }
  __host__ __device__ static constexpr int slots_in_layer(int l) { return k_slices(l) / SPS; }
  __host__ __device__ static constexpr uint32_t piece_bytes(int l, int L) { return uint32_t(n_of(l, L)) * SLICE_K * 2; }
};

template <class Cfg>
__global__ void __launch_bounds__(TCN_THREADS, 1) nfb_tcn_kernel(const TcParams tp) {
  constexpr int H = Cfg::H, ROWS = Cfg::ROWS, NQT = Cfg::NQT, SPLIT = Cfg::SPLIT, RING = Cfg::RING, NCH = Cfg::NCH;
  constexpr uint32_t SLOT_BYTES = Cfg::SLOT_BYTES, ATOM_BYTES = Cfg::ATOM_BYTES;
  const EvalParams& p = tp.ev;
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* const smem = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* const act = smem + Cfg::OFF_ACT;   // A_hi (the only copy when SPLIT = 1), then A_lo
  unsigned char* const inbuf = smem + Cfg::OFF_IN;  // the first layer's input rows (written a tile ahead)
  unsigned char* const ring = smem + Cfg::OFF_RING;
  float* const bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* const scale_s = bias_s + MAX_TC_LAYERS * TCN_BIAS_LD;
  float* const maha2 = reinterpret_cast<float*>(smem + Cfg::OFF_MAHA);
  float* const re2 = reinterpret_cast<float*>(smem + Cfg::OFF_RE);
  const uint32_t bar_full = smem_u32(smem + Cfg::OFF_BARS), bar_empty = bar_full + 8 * RING;
  const uint32_t bar_a = bar_empty + 8 * RING;     // epilogue -> MMA: the next layer's input is in shared memory, the accumulators are read
  const uint32_t bar_d = bar_a + 8;                // MMA -> epilogue: the layer's accumulators are complete (and its input no longer read)
  const uint32_t bar_in = bar_d + 8;               // featurise -> MMA: the tile's first-layer input rows are written
  const uint32_t bar_drained = bar_in + 8;         // epilogue -> MMA: the last layer's accumulators are read
  const uint32_t bar_infree = bar_drained + 8;     // MMA -> featurise: layer 0 has consumed the input rows
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_TMEM);
  // epilogue -> featurise: epilogue warps that have finished a tile's decode (8 per tile).  A COUNTER, not an mbarrier: the
  // featurise warps wait for tile t - 2 while they work on tile t, and a narrow model's tile (3 - 6 small layers) can be
  // finished faster than one featurisation (fp64, ~7,000 cycles) -- two phases of a parity barrier could then complete
  // between two waits, and a parity wait that has been lapped never returns (found the hard way: xxsmall, 400 tiles per CTA).
  volatile uint32_t* const decodes_done = tmem_slot + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const long long n_tiles = (p.n + NQT - 1) / NQT;
  const uint32_t my_tiles = blockIdx.x < n_tiles ? uint32_t((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  int slots_per_tile = 0;
  for (int l = 0; l < L; ++l) slots_per_tile += Cfg::slots_in_layer(l);

  for (int i = tid; i < L * TCN_BIAS_LD; i += TCN_THREADS) bias_s[i] = tp.bias[i];
  if (tid < L) scale_s[tid] = tp.inv_scale[tid];
  if (tid == 0) {
    for (int s = 0; s < RING; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_a, TCN_EPI_WARPS);
    mbar_init(bar_d, 1);
    mbar_init(bar_in, TCN_FW_WARPS);
    mbar_init(bar_drained, TCN_EPI_WARPS);
    mbar_init(bar_infree, 1);
    *decodes_done = 0;
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

  if (warp == 0) {
    // ============================ weight producer: one bulk copy per slot, sized by the slot's layer ============================
    const unsigned char* const src = reinterpret_cast<const unsigned char*>(tp.wblob);
    uint32_t s = 0, ph = 0;
    for (uint32_t t = 0; t < my_tiles; ++t) {
      uint32_t j = 0;
      for (int l = 0; l < L; ++l) {
        const uint32_t bytes = 2 * Cfg::piece_bytes(l, L);
        for (int k = 0; k < Cfg::slots_in_layer(l); ++k, ++j) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
            bulk_copy_g2s(smem_u32(ring + s * SLOT_BYTES), src + size_t(j) * SLOT_BYTES, bytes, bar_full + 8 * s);
          }
          __syncwarp();
          if (++s == RING) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issue (warp-converged, one elected lane issues) ============================
    // descriptor high words: A = SWIZZLE_128B K-major (SBO 1024 B), B = no-swizzle core matrices (LBO 128 B, SBO 512 B)
    const uint64_t a_hi = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
    const uint64_t b_hi = (uint64_t(128 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46);
    const uint32_t a_act = (smem_u32(act) >> 4) & 0x3FFF, a_in = (smem_u32(inbuf) >> 4) & 0x3FFF, b0 = (smem_u32(ring) >> 4) & 0x3FFF;
    uint32_t s = 0, ph = 0, ph_a = 0, ph_in = 0, ph_dr = 0;
    // One ring slot of layer l: k position `pos` (in slots) of the layer's K sweep
    auto issue_layer = [&](int l) {
      const int n = Cfg::n_of(l, L);
      const uint32_t idesc = (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(ROWS >> 4) << 24);  // f16 x f16 -> f32, M = 128
      const uint32_t piece = Cfg::piece_bytes(l, L) >> 4;
      const uint32_t d = tmem, d_small = tmem + Cfg::D_LO_OFFSET;
      for (int pos = 0; pos < Cfg::slots_in_layer(l); ++pos) {
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t b_lo = b0 + s * (SLOT_BYTES >> 4);
          if constexpr (SPLIT == 1) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int ks = 2 * pos + u;
              // layer 0 reads [x_hi | x_lo] of the input buffer as the K = 128 concatenation [x_hi | x_lo | x_hi | x_hi]
              const uint32_t a_lo = (l == 0) ? a_in + uint32_t(ks == 1) * 4
                                             : a_act + uint32_t(ks >> 1) * (ATOM_BYTES >> 4) + uint32_t(ks & 1) * 4;
              const uint32_t b = b_lo + u * piece;
              umma_f16(d, a_hi | a_lo, b_hi | b, idesc, ks != 0);
              umma_f16(d, a_hi | (a_lo + 2), b_hi | (b + 16), idesc, 1u);
            }
          } else {
            const int ks = pos;
            const uint32_t a_lo = (l == 0) ? a_in : a_act + uint32_t(ks >> 1) * (ATOM_BYTES >> 4) + uint32_t(ks & 1) * 4;
            const uint32_t LO = (l == 0) ? 4u : (Cfg::A_BYTES >> 4);  // distance of the lo half of the activations
            umma_f16(d, a_hi | a_lo, b_hi | b_lo, idesc, ks != 0);                              // a_hi . w_hi -> D_hi
            umma_f16(d, a_hi | (a_lo + 2), b_hi | (b_lo + 16), idesc, 1u);
            umma_f16(d_small, a_hi | (a_lo + LO), b_hi | b_lo, idesc, ks != 0);                 // a_lo . w_hi -> D_lo
            umma_f16(d_small, a_hi | (a_lo + LO + 2), b_hi | (b_lo + 16), idesc, 1u);
            umma_f16(d_small, a_hi | a_lo, b_hi | (b_lo + piece), idesc, 1u);                   // a_hi . w_lo -> D_lo
            umma_f16(d_small, a_hi | (a_lo + 2), b_hi | (b_lo + piece + 16), idesc, 1u);
          }
          umma_commit(bar_empty + 8 * s);  // the slot is free once the tensor core has read it
        }
        __syncwarp();
        if (++s == RING) { s = 0; ph ^= 1; }
      }
    };
    auto commit = [&](uint32_t bar) {
      if (elect_one_sync()) umma_commit(bar);
      __syncwarp();
    };
    for (uint32_t t = 0; t < my_tiles; ++t) {
      mbar_wait(bar_in, ph_in);
      ph_in ^= 1;
      if (t != 0) {
        mbar_wait(bar_drained, ph_dr);
        ph_dr ^= 1;
      }
      tc_fence_after();
      if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32] = clock64();
      issue_layer(0);
      commit(bar_d);
      commit(bar_infree);
      if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[33] = clock64();
      for (int l = 1; l < L; ++l) {
        mbar_wait(bar_a, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32 + 2 * l] = clock64();
        issue_layer(l);
        commit(bar_d);
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32 + 2 * l + 1] = clock64();
      }
    }
  } else if (warp >= TCN_FW_WARP0) {
    // ============================ featurise-ahead warps ============================
    // Thread ft owns batch row ft: row 2 i is query i of the tile, row 2 i + 1 its flipped twin.  One row per thread: a row's
    // featurisation is ~10,000 latency-bound cycles (fp64 sin / cos / log and the 325-term quadratic form), and with two rows
    // per thread it was what paced every narrow model (measured: profiles/r2_tcn_phase_stamps_first_version.txt).
    const int ft = tid - 32 * TCN_FW_WARP0;
    uint32_t ph_free = 0;
    for (uint32_t t = 0; t < my_tiles; ++t) {
      const long long tile_q0 = (blockIdx.x + (long long)t * gridDim.x) * NQT;
      float* const maha = maha2 + (t & 1) * ROWS;
      float* const re_s = re2 + (t & 1) * NQT;
      if (tp.dbg && blockIdx.x == 0 && t == 2 && ft == 0) tp.dbg[46] = clock64();
#pragma unroll 1
      for (int j = 0; j < ROWS / (32 * TCN_FW_WARPS); ++j) {
        const int n = ft + 32 * TCN_FW_WARPS * j;
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
            while (*decodes_done < uint32_t(TCN_EPI_WARPS) * (t - 1)) {}
            __threadfence_block();
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
      if (lane == 0) mbar_arrive(bar_in);
      if (tp.dbg && blockIdx.x == 0 && t == 2 && ft == 0) tp.dbg[47] = clock64();
    }
  } else {
    // ============================ epilogue warps ============================
    // TMEM lane 32 q + lane = batch row; the two warps of a lane quarter split a layer's features.
    const int e = warp - 2, q = warp & 3, part = e >> 2;
    const int n = 32 * q + lane;                       // batch row
    const int fofs = (H / 2) * part;                   // first of my features
    const uint32_t t_lane = tmem + (uint32_t(32 * q) << 16);
    uint32_t ph_d = 0;
    auto activate = [&](const uint32_t (&v)[32], const float* b, float inv_s, uint32_t (&w)[16], uint32_t (&wl)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float y0 = fmaf(__uint_as_float(v[2 * i]), inv_s, b[2 * i]);
        const float y1 = fmaf(__uint_as_float(v[2 * i + 1]), inv_s, b[2 * i + 1]);
        if constexpr (SPLIT == 1) {
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
    for (uint32_t t = 0; t < my_tiles; ++t) {
      const long long tile = blockIdx.x + (long long)t * gridDim.x;
      const long long tile_q0 = tile * NQT;
      const bool stamp = tp.dbg && blockIdx.x == 0 && t == 1 && warp == 4 && lane == 0;
      const float* const maha = maha2 + (t & 1) * ROWS;
      const float* const re_s = re2 + (t & 1) * NQT;
      if (stamp) tp.dbg[0] = clock64();

      for (int l = 0; l < L - 1; ++l) {
        const float inv_s = scale_s[l];
        const float* b = bias_s + l * TCN_BIAS_LD + fofs;
        mbar_wait(bar_d, ph_d);  // the sweep is complete: accumulators final, the layer's input no longer read
        ph_d ^= 1;
        tc_fence_after();
        if (stamp) tp.dbg[2 + 2 * l] = clock64();
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t v[32], w[16], wl[16];
          load_acc(uint32_t(fofs + 32 * c), v);
          activate(v, b + 32 * c, inv_s, w, wl);
          store32(fofs + 32 * c, w, wl);
        }
        // my smem writes -> async proxy, my TMEM reads are done; tell the MMA warp
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a);
        if (stamp) tp.dbg[3 + 2 * l] = clock64();
      }

      // ---- last layer: decode straight from the accumulators (nfb_tc2.cuh).  My lane holds 32-row chunks of the packed
      //      latents of one batch row; lane ^ 1 holds the same chunks of its twin (or query). ----
      mbar_wait(bar_d, ph_d);
      ph_d ^= 1;
      tc_fence_after();
      if (stamp) tp.dbg[20] = clock64();
      {
        const float inv_s = scale_s[L - 1];
        const float* b = bias_s + (L - 1) * TCN_BIAS_LD;
        const int f_begin = 128 * part, f_count = 128;   // packed rows [128 part, 128 part + 128) (columns = rows)
        const bool odd = n & 1;                          // my row is a twin
        const long long qg = tile_q0 + (n >> 1);         // the query my lane pair decodes
        const bool valid = qg < p.n;
        const float maha_o = maha[n & ~1], maha_t = maha[n | 1], re = re_s[n >> 1];
#pragma unroll 1
        for (int c = 0; c < f_count; c += 32) {
          const int f0 = f_begin + c;    // packed rows f0 .. f0 + 31
          if (f0 >= M_LAST_USED) break;  // (warp-uniform; the rest is padding)
          const bool skip = p.scalars_only && f0 < 4 * 48;  // (warp-uniform) boundary-layer chunks not wanted
          uint32_t v[32];
          if (!skip) load_acc(uint32_t(f0), v);
          if (c + 32 >= f_count || f0 + 32 >= M_LAST_USED) {  // that was my last read of the accumulators
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_drained);  // the next tile's layer 0 may overwrite them
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
        if (lane == 0) {
          __threadfence_block();  // my warp's reads of maha / re come before the count
          atomicAdd(const_cast<uint32_t*>(decodes_done), 1u);
        }
        if (stamp) tp.dbg[22] = clock64();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

}  // namespace tcn
}  // namespace nfb
