// Three-term ("tensor_fp32") tensor-core evaluator for the 512-wide model on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// The one-SM three-term kernel (nfb_tc.cuh, SPLIT = 3) must keep both fp16 halves of the activation tile on chip, which
// at 512 features leaves room for 64 batch rows only -- and an M = 64 MMA costs the same 128 cycles as an M = 128 one.
// Here the two CTAs of a cluster form ONE M = 128 MMA: each CTA contributes its own 64 batch rows of A and HALF of the
// B slice (128 of the 256 output features of an N-block), the leader CTA issues for both tensor cores, and each SM runs
// its 64 x 256 x 16 share in 64 cycles -- the full rate (measured: scripts/umma_probe_2sm.cu, 4093 MAC/clk/SM).
//
//   * D layout (measured, same probe): batch row r of a CTA, column n of the 256-wide N-block -> that CTA's TMEM lane
//     r + 64 (n / 128), column n % 128.  All 128 lanes are used (the M = 64 one-SM layout idles half of them), so every
//     epilogue lane has work: lanes 0..63 = rows with features [0, 128) of the N-block, lanes 64..127 = the same rows
//     with features [128, 256).
//   * TMEM (512 columns): D_hi of N-block 0 / 1 at columns 0 / 128, D_lo (the two small product terms, accumulated
//     apart because the tensor core truncates at every accumulation) at 256 / 384.
//   * Weights: the same fp16 W_hi / W_lo slice stream as the one-SM kernel (256 features x 32 k per slice); CTA c
//     fetches rows [128 c, 128 c + 128) of both -- half the bytes, no multicast.  The peer CTA's MMA warp has nothing
//     to issue; it relays "my half of slot s has landed" to the leader's full barrier instead.
//   * Layer pipeline: the half-layer schedule of nfb_tc.cuh (q1..q4, barriers a0 / a1 / d0 / k0 / d1), with the
//     epilogue -> MMA barriers living in the leader CTA (both CTAs' epilogue warps arrive there) and the MMA -> epilogue
//     barriers signalled in both CTAs by multicast commits.
#pragma once
#include "nfb_tc.cuh"

namespace nfb {
namespace tc2 {

using namespace nfb::tc;  // PTX wrappers, SLICE_K, ATOM_K, THREADS, EPI_WARPS, MAX_TC_LAYERS, LAST_ROWS

constexpr int RING = 5;  // weight ring depth (16 KB slots)

template <int H_>
struct Tc2Cfg {
  static constexpr int H = H_;
  static constexpr int SPLIT = 3;
  static constexpr int ROWS = 64;                                    // batch rows per CTA; the MMA's M = 128 spans the pair
  static constexpr int NQT = ROWS / 2;                               // queries per CTA tile (rows NQT.. are the twins)
  static constexpr int K0 = 32;
  static constexpr int BLOCK_N = 256;                                // MMA N = features per N-block
  static constexpr int HALF_N = BLOCK_N / 2;                         // B rows held by each CTA
  static constexpr uint32_t PIECE_BYTES = HALF_N * SLICE_K * 2;      // one CTA's share of a slice: 8 KB
  static constexpr uint32_t SLICE_BYTES = 2 * PIECE_BYTES;           // a packed slice in the blob (256 features x 32 k)
  static constexpr uint32_t SLOT_BYTES = 2 * PIECE_BYTES;            // ring slot: my W_hi piece, my W_lo piece
  static constexpr uint32_t ATOM_BYTES = ROWS * 128;
  static constexpr uint32_t A_BYTES = (H / ATOM_K) * ATOM_BYTES;
  static constexpr uint32_t ACT_BYTES = 2 * A_BYTES;                 // [A_hi][A_lo]
  static constexpr uint32_t STAGE_BYTES = M_LAST * ROWS * 4;         // last layer's latents, fp32 [224][ROWS], in the idle activation tile
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t D_LO_OFFSET = 256;
  static constexpr uint32_t OFF_ACT = 0;
  static constexpr uint32_t OFF_RING = OFF_ACT + ACT_BYTES;
  static constexpr uint32_t OFF_BIAS = OFF_RING + RING * SLOT_BYTES;
  static constexpr uint32_t BIAS_FLOATS = MAX_TC_LAYERS * H + 16;
  static constexpr uint32_t OFF_MAHA = OFF_BIAS + BIAS_FLOATS * 4;
  static constexpr uint32_t OFF_RE = OFF_MAHA + ROWS * 4;
  static constexpr uint32_t OFF_BARS = OFF_RE + NQT * 4;
  static constexpr uint32_t OFF_TMEM = OFF_BARS + 24 * 8;
  static constexpr uint32_t OFF_STAGE = OFF_ACT;
  static constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
  static_assert(H == 512, "the pair kernel exists for the 512-wide model (two 256-feature N-blocks)");
  static_assert(ACT_BYTES >= STAGE_BYTES, "staging re-uses the activation tile");
  static_assert(SMEM_BYTES <= 232448, "shared-memory budget");

  __host__ __device__ static constexpr int k_slices(int l) { return (l == 0 ? K0 : H) / SLICE_K; }
  __host__ __device__ static constexpr int n_blocks(int l, int L) { return l == L - 1 ? LAST_ROWS / BLOCK_N : H / BLOCK_N; }
  __host__ __device__ static constexpr int steps_in_layer(int l, int L) { return k_slices(l) * n_blocks(l, L); }
  // Position in a layer's K sweep -> k-slice.  K region 1 (the second half of the inputs) is swept chunk-major: first the
  // four 32-feature chunks the epilogue threads of the previous layer write first, then the four they write second, so
  // the sweep can start on the first four while the second four are still being activated (barriers a1a / a1b).
  __host__ __device__ static constexpr int k_slice_at(int pos, int kt) {
    return (pos < kt / 2 || kt < 4) ? pos : kt / 2 + 2 * ((pos - kt / 2) & 3) + ((pos - kt / 2) >> 2);
  }
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
// Arrive on a barrier of another CTA of the cluster.  Default (CTA-scope) semantics on purpose, as for every other
// barrier here: a cluster-scope release compiles to MEMBAR.ALL.GPU and a cluster-scope acquiring wait to CCTL.IVALL
// (an L1 invalidate) per wait -- measured 3x slower MMA issue.  Nothing is handed over through the generic proxy
// across CTAs: the data these barriers order is written to the writer's own shared memory, made visible to the async
// proxy by fence.proxy.async before the arrive, and read by that same SM's tensor core.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// SiLU x / (1 + e^-x) from the two MUFU approximations directly (ex2.approx: 2 ulp, rcp.approx: 1 ulp) -- the epilogue
// here is MUFU / issue bound, and the operands are about to be split into 22-bit fp16 pairs anyway.
__device__ __forceinline__ float silu_ex2(float h) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fminf(-h, 80.0f) * 1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return h * r;
}

template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) nfb_tc2_kernel(const TcParams tp) {
  constexpr int H = Cfg::H, ROWS = Cfg::ROWS, NQT = Cfg::NQT;
  constexpr uint32_t SLOT_BYTES = Cfg::SLOT_BYTES, PIECE_BYTES = Cfg::PIECE_BYTES, ATOM_BYTES = Cfg::ATOM_BYTES;
  const EvalParams& p = tp.ev;
  extern __shared__ unsigned char smem_unaligned[];
  unsigned char* const smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_unaligned) + 1023) & ~uintptr_t(1023));
  unsigned char* const act = smem + Cfg::OFF_ACT;   // A_hi, then A_lo
  unsigned char* const ring = smem + Cfg::OFF_RING;
  float* const bias_s = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* const maha = reinterpret_cast<float*>(smem + Cfg::OFF_MAHA);
  float* const re_s = reinterpret_cast<float*>(smem + Cfg::OFF_RE);
  const uint32_t bar_full = smem_u32(smem + Cfg::OFF_BARS), bar_empty = bar_full + 8 * RING;
  // layer pipeline barriers, see nfb_tc.cuh; a0 / a1 are waited on in the leader CTA only
  const uint32_t bar_a0 = bar_empty + 8 * RING, bar_a1a = bar_a0 + 8, bar_a1b = bar_a1a + 8, bar_d0 = bar_a1b + 8, bar_k0 = bar_d0 + 8,
                 bar_d1 = bar_k0 + 8;  // a1a / a1b: the first / second four k-slices of K region 1 are in smem
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_TMEM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const long long n_tiles = (p.n + NQT - 1) / NQT;
  const uint32_t my_tiles = uint32_t((n_tiles + gridDim.x - 1) / gridDim.x);  // same for every CTA
  const uint32_t cta_rank = cluster_ctarank();
  int steps_per_tile = 0;
  for (int l = 0; l < L; ++l) steps_per_tile += Cfg::steps_in_layer(l, L);

  for (int i = tid; i < L * H; i += THREADS) bias_s[i] = tp.bias[i];
  if (tid < L) bias_s[MAX_TC_LAYERS * H + tid] = tp.inv_scale[tid];
  if (tid == 0) {
    for (int s = 0; s < RING; ++s) {
      mbar_init(bar_full + 8 * s, cta_rank == 0 ? 2 : 1);  // leader: its own producer + the peer's relay
      mbar_init(bar_empty + 8 * s, 1);                     // one multicast commit
    }
    mbar_init(bar_a0, 2 * EPI_WARPS);                      // the epilogue warps of both CTAs
    mbar_init(bar_a1a, 2 * EPI_WARPS);
    mbar_init(bar_a1b, 2 * EPI_WARPS);
    mbar_init(bar_d0, 1);
    mbar_init(bar_k0, 1);
    mbar_init(bar_d1, 1);
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
    // ============================ weight producer: my half of every slice ============================
    if (lane == 0) {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(tp.wblob) + cta_rank * PIECE_BYTES;
      uint32_t s = 0, ph = 0;
      for (uint32_t t = 0; t < my_tiles; ++t)
        for (int l = 0, base = 0; l < L; base += Cfg::steps_in_layer(l, L), ++l) {
          const int kt = Cfg::k_slices(l);
          for (int nb = 0; nb < Cfg::n_blocks(l, L); ++nb)
            for (int pos = 0; pos < kt; ++pos) {
              const size_t j = size_t(base + nb * kt + Cfg::k_slice_at(pos, kt));  // blob order: layer, N-block, k-slice; [W_hi slice][W_lo slice]
              mbar_wait(bar_empty + 8 * s, ph ^ 1);
              mbar_arrive_expect_tx(bar_full + 8 * s, SLOT_BYTES);
              const uint32_t dst = smem_u32(ring + s * SLOT_BYTES);
              bulk_copy_g2s(dst, src + j * (2 * Cfg::SLICE_BYTES), PIECE_BYTES, bar_full + 8 * s);                                   // W_hi rows
              bulk_copy_g2s(dst + PIECE_BYTES, src + j * (2 * Cfg::SLICE_BYTES) + Cfg::SLICE_BYTES, PIECE_BYTES, bar_full + 8 * s);  // W_lo rows
              if (++s == RING) { s = 0; ph ^= 1; }
            }
        }
    }
  } else if (warp == 1 && cta_rank != 0) {
    // ============================ peer CTA: relay "slot landed" to the leader ============================
    const uint32_t total = my_tiles * uint32_t(steps_per_tile);
    const uint32_t leader_full = map_to_rank(bar_full, 0);
    uint32_t s = 0, ph = 0;
    for (uint32_t it = 0; it < total; ++it) {
      mbar_wait(bar_full + 8 * s, ph);
      if (lane == 0) mbar_arrive_cluster(leader_full + 8 * s);
      __syncwarp();
      if (++s == RING) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ============================ leader CTA: MMA issue for the pair ============================
    const uint32_t idesc = (1u << 4) | (uint32_t(Cfg::BLOCK_N >> 3) << 17) | (uint32_t(128 >> 4) << 24);  // f16 x f16 -> f32, M = 128 over two CTAs
    const uint64_t a_hi = (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
    const uint64_t b_hi = (uint64_t(128 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46);
    const uint32_t a0 = (smem_u32(act) >> 4) & 0x3FFF, b0 = (smem_u32(ring) >> 4) & 0x3FFF;
    constexpr uint32_t LO = Cfg::A_BYTES >> 4, WLO = PIECE_BYTES >> 4;
    uint32_t s = 0, ph = 0, ph_a0 = 0, ph_a1 = 0;  // (a1a and a1b flip together)
    long long wait_full = 0, wait_a0 = 0, wait_a1 = 0;  // diagnostics (tp.dbg only): cycles this warp spent waiting
    const bool timing = tp.dbg && blockIdx.x == 0;
    auto issue = [&](int nb, int kt, int pos_begin, int pos_end) {  // positions [pos_begin, pos_end) of N-block nb's K sweep
      for (int pos = pos_begin; pos < pos_end; ++pos) {
        const int ks = Cfg::k_slice_at(pos, kt);
        const uint32_t a_lo = a0 + uint32_t(ks >> 1) * (ATOM_BYTES >> 4) + uint32_t(ks & 1) * 4;  // 64 B per 32 k
        const long long c0 = timing ? clock64() : 0;
        mbar_wait(bar_full + 8 * s, ph);
        if (timing) wait_full += clock64() - c0;
        tc_fence_after();
        const uint32_t b_lo = b0 + s * (SLOT_BYTES >> 4);
        const uint32_t d = tmem + nb * Cfg::HALF_N, d_small = d + Cfg::D_LO_OFFSET;
        if (elect_one_sync()) {
          umma_f16_pair(d, a_hi | a_lo, b_hi | b_lo, idesc, pos != 0);                          // a_hi . w_hi -> D_hi
          umma_f16_pair(d, a_hi | (a_lo + 2), b_hi | (b_lo + 16), idesc, 1u);
          umma_f16_pair(d_small, a_hi | (a_lo + LO), b_hi | b_lo, idesc, pos != 0);             // a_lo . w_hi -> D_lo
          umma_f16_pair(d_small, a_hi | (a_lo + LO + 2), b_hi | (b_lo + 16), idesc, 1u);
          umma_f16_pair(d_small, a_hi | a_lo, b_hi | (b_lo + WLO), idesc, 1u);                  // a_hi . w_lo -> D_lo
          umma_f16_pair(d_small, a_hi | (a_lo + 2), b_hi | (b_lo + WLO + 16), idesc, 1u);
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
    for (uint32_t t = 0; t < my_tiles; ++t)
      for (int l = 0; l < L; ++l) {
        const int kt = Cfg::k_slices(l);
        const int kh = (l == 0) ? kt : kt / 2;
        if (timing && t == 1 && l == 0) wait_full = wait_a0 = wait_a1 = 0;
        long long c0 = timing ? clock64() : 0;
        mbar_wait(bar_a0, ph_a0);
        if (timing) wait_a0 += clock64() - c0;
        ph_a0 ^= 1;
        tc_fence_after();
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) tp.dbg[32 + 2 * l] = clock64();
        issue(0, kt, 0, kh);                                   // q1
        const int km = kh + (kt - kh) / 2;
        c0 = timing ? clock64() : 0;
        mbar_wait(bar_a1a, ph_a1);
        if (timing) wait_a1 += clock64() - c0;
        tc_fence_after();
        issue(0, kt, kh, km);                                  // q2, first four slices
        c0 = timing ? clock64() : 0;
        mbar_wait(bar_a1b, ph_a1);
        if (timing) wait_a1 += clock64() - c0;
        ph_a1 ^= 1;
        tc_fence_after();
        issue(0, kt, km, kt);                                  // q2, the rest
        commit(bar_d0);
        if (l != L - 1) {
          issue(1, kt, 0, kh);                                 // q3
          commit(bar_k0);
          issue(1, kt, kh, kt);                                // q4
          commit(bar_d1);
        }
        if (tp.dbg && blockIdx.x == 0 && t == 1 && lane == 0) {
          tp.dbg[32 + 2 * l + 1] = clock64();
          if (l == L - 1) { tp.dbg[48] = wait_full; tp.dbg[49] = wait_a0; tp.dbg[50] = wait_a1; }
        }
      }
  } else {
    // ============================ epilogue warps (both CTAs) ============================
    // TMEM lane 32 q + lane = batch row (32 q + lane) % 64 with features [128 (q / 2), 128 (q / 2) + 128) of the N-block;
    // the two warps of a lane quarter split those 128 columns.
    const int e = warp - 2, q = warp & 3, part = e >> 2;
    const int n = 32 * (q & 1) + lane;          // batch row in this CTA
    const int fcol = 64 * part;                 // first of my 64 TMEM columns within an accumulator tile
    const int fofs = 128 * (q >> 1) + fcol;     // ... = feature offset within the N-block
    const int et = (e << 5) | lane;             // 0..255
    const uint32_t t_lane = tmem + (uint32_t(32 * q) << 16);
    const uint32_t leader_a0 = map_to_rank(bar_a0, 0), leader_a1a = map_to_rank(bar_a1a, 0), leader_a1b = map_to_rank(bar_a1b, 0);
    const bool vec_ok =
        p.out_f64 ? ((p.out_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                  : ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    uint32_t ph_d0 = 0, ph_k0 = 0, ph_d1 = 0;
    auto activate = [&](const uint32_t (&v)[32], const float* b, float inv_s, uint32_t (&w)[16], uint32_t (&wl)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x0 = silu_ex2(fmaf(__uint_as_float(v[2 * i]), inv_s, b[2 * i]));
        const float x1 = silu_ex2(fmaf(__uint_as_float(v[2 * i + 1]), inv_s, b[2 * i + 1]));
        w[i] = pack_f16x2(x0, x1);
        const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        wl[i] = pack_f16x2(x0 - back.x, x1 - back.y);
      }
    };
    auto load_acc = [&](uint32_t col, uint32_t (&v)[32]) {  // D_hi + D_lo
      uint32_t u[32];
      tmem_ld32(t_lane + col, v);
      tmem_ld32(t_lane + col + Cfg::D_LO_OFFSET, u);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
    };
    auto store32 = [&](int k, const uint32_t (&w)[16], const uint32_t (&wl)[16]) {  // features k..k+31 of row n, hi and lo
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t off = act_chunk_offset<Cfg>(n, k + 8 * c8);
        *reinterpret_cast<uint4*>(act + off) = make_uint4(w[4 * c8], w[4 * c8 + 1], w[4 * c8 + 2], w[4 * c8 + 3]);
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
      const bool stamp = tp.dbg && blockIdx.x == 0 && t == 1 && warp == 4 && lane == 0;  // a featurising warp
      if (stamp) tp.dbg[0] = clock64();
      if (q < 2 && part == 0) {  // one thread per batch row (row n: query n, row NQT + n: its twin)
        const bool twin = n >= NQT;
        const long long qi = tile_q0 + (twin ? n - NQT : n);
        featurise_row<Cfg>(p, qi, qi < p.n, twin, n, act, maha, re_s);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(leader_a0);
        mbar_arrive_cluster(leader_a1a);
        mbar_arrive_cluster(leader_a1b);
      }
      if (stamp) tp.dbg[1] = clock64();

      for (int l = 0; l < L - 1; ++l) {
        const float inv_s = bias_s[MAX_TC_LAYERS * H + l];
        const float* b = bias_s + l * H + fofs;
        // ---- half 0: N-block 0 complete after q2; its outputs may only land in K region 0 after q3 ----
        mbar_wait(bar_d0, ph_d0);
        ph_d0 ^= 1;
        tc_fence_after();
        if (stamp) tp.dbg[2 + 2 * l] = clock64();
        {
          uint32_t w[2][16], wl[2][16];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            load_acc(uint32_t(fcol + 32 * c), v);
            activate(v, b + 32 * c, inv_s, w[c], wl[c]);
          }
          mbar_wait(bar_k0, ph_k0);
          ph_k0 ^= 1;
#pragma unroll
          for (int c = 0; c < 2; ++c) store32(fofs + 32 * c, w[c], wl[c]);
        }
        publish(leader_a0);
        // ---- half 1: N-block 1 complete (and K region 1 free) after q4; overlaps the next layer's q1 ----
        mbar_wait(bar_d1, ph_d1);
        ph_d1 ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {  // chunk c of every thread = the c-th four k-slices of the next layer's K region 1
          uint32_t v[32], w[16], wl[16];
          load_acc(uint32_t(Cfg::HALF_N + fcol + 32 * c), v);
          activate(v, b + Cfg::BLOCK_N + 32 * c, inv_s, w, wl);
          store32(Cfg::BLOCK_N + fofs + 32 * c, w, wl);
          publish(c == 0 ? leader_a1a : leader_a1b);
        }
        if (stamp) tp.dbg[3 + 2 * l] = clock64();
      }

      // ---- last layer: latents -> fp32 staging [224 packed rows][ROWS] in the idle activation tile ----
      mbar_wait(bar_d0, ph_d0);
      ph_d0 ^= 1;
      tc_fence_after();
      if (stamp) tp.dbg[20] = clock64();
      {
        float* stage = reinterpret_cast<float*>(smem + Cfg::OFF_STAGE);
        const float inv_s = bias_s[MAX_TC_LAYERS * H + (L - 1)];
        const float* b = bias_s + (L - 1) * H;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int f0 = fofs + 32 * c;  // packed rows f0 .. f0 + 31
          if (f0 >= M_LAST) break;       // (warp-uniform)
          uint32_t v[32];
          load_acc(uint32_t(fcol + 32 * c), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) stage[(f0 + i) * ROWS + n] = fmaf(__uint_as_float(v[i]), inv_s, b[f0 + i]);
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
        named_bar_sync(1, EPI_THREADS);  // staging (and maha / re) consumed before the next tile's featurise overwrites them
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
