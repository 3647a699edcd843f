// fp32 (FFMA2) evaluator for the models up to 256 wide: the same arithmetic as nfb_ffma.cuh, scheduled so that the
// phases that do not use the FMA pipe (fp64 featurisation, SiLU, the MUFU-heavy decode, waiting for weights)
// overlap the GEMM of other warps.
//
// A CTA still owns a tile of NC = 32768 / H columns in shared memory, but the tile is split into WN = NC / 64
// independent COLUMN GROUPS (32 queries and their 32 twins, WM warps each).  A group does everything for its own
// queries -- featurise, every hidden layer, the last layer, the decode, the stores -- and synchronises only with
// itself (one named barrier per group).  Nothing in the tile loop is CTA-wide, so the groups drift apart: while one
// group evaluates SiLUs or decodes, the others keep the FMA pipe busy (measured on B200, xlarge: 108.8 -> see
// DESIGN.md section 6).
//
// The only thing the groups share is the weight stream: one ring of TMA bulk copies per CTA, which every warp
// consumes in the same order.  The ring is deep, because its depth is what bounds the drift (see wait_slice).
//
// Last layer: 256 / H SWEEPS that look exactly like hidden layers (H packed output rows each, same 8 x TN thread
// tile, same slice size) whose epilogue is un-flip + average + decode + store instead of SiLU.  Packed row
// p = sweep * H + r holds latent row pack_last_layer_row(p); a thread owns two row groups of four per sweep.
// Every accumulator sees the bias, then one IEEE fma per k in k order: results are bit-identical to
// nfb_ffma_kernel, nfb_small_kernel and the value columns of nfb_jac_kernel (tests/test_gpu_parity.py).
#pragma once
#include "nfb_ffma.cuh"

namespace nfb {

constexpr int M_LAST_SWEPT = 256;  // last layer's packed rows in the sweep layout (64 groups of four, 50 used)

template <int H_, int KC_, int STAGES_, int TN_>
struct GroupCfg {
  static constexpr int H = H_;
  static constexpr int KC = KC_;
  static constexpr int STAGES = STAGES_;
  static constexpr int TN = TN_;
  static constexpr int WARPS = 128 / TN;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int LN = 64 / TN;            // lanes along columns: a warp always spans 64 columns
  static constexpr int LM = 32 / LN;            // lanes along rows
  static constexpr int NC = 32768 / H;
  static constexpr int NQ = NC / 2;
  static constexpr int WM = H / (8 * LM);       // warps per column group
  static constexpr int WN = WARPS / WM;         // column groups per CTA
  static constexpr int GROUP_THREADS = WM * 32;
  static constexpr int NS = M_LAST_SWEPT / H;   // last-layer sweeps
  static constexpr int SLOT_FLOATS = KC * H;
  static constexpr int MAX_SLICES = K0_PAD / KC + (NFB_MAX_LAYERS_K - 2) * (H / KC) + NS * (H / KC);
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t(H) * NC + size_t(STAGES) * SLOT_FLOATS + NC + NQ) +
                                       16 * STAGES + 8 * size_t(MAX_SLICES);
  static_assert(TN == 8 || TN == 16, "thread tile is 8 or 16 columns wide");
  static_assert(H <= 256 && M_LAST_SWEPT % H == 0 && H % (8 * LM) == 0 && WARPS % WM == 0 && WN * 64 == NC, "bad width");
  static_assert(GROUP_THREADS >= 64, "a group featurises its 64 columns with one thread each");
  static_assert((STAGES & (STAGES - 1)) == 0 && STAGES >= 4, "STAGES must be a power of two >= 4");
  static_assert(K0_PAD % KC == 0 && H % KC == 0, "KC must divide every K");
  static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

// The sweep-layout copy of the last layer sits behind the blob of nfb_ffma.cuh: Wt[NS][H (k)][H (r)], then bias[256].
template <int H>
__host__ __device__ constexpr size_t blob_sweeps_offset(int n_layers) { return blob_total_floats<H>(n_layers); }
template <int H>
__host__ __device__ constexpr size_t blob_total_floats_swept(int n_layers) {
  return blob_total_floats<H>(n_layers) + size_t(M_LAST_SWEPT) * H + M_LAST_SWEPT;
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) nfb_ffma_groups_kernel(const EvalParams p) {
  constexpr int H = Cfg::H, NC = Cfg::NC, NQ = Cfg::NQ, KC = Cfg::KC, STAGES = Cfg::STAGES, TN = Cfg::TN;
  constexpr int WM = Cfg::WM, SLOT = Cfg::SLOT_FLOATS, LM = Cfg::LM, LN = Cfg::LN, NS = Cfg::NS;
  constexpr int WARPS = Cfg::WARPS, THREADS = Cfg::THREADS, QT = TN / 2;  // QT queries per thread

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const act = reinterpret_cast<float*>(smem_raw);  // [H][NC]
  float* const ring = act + H * NC;                        // [STAGES][SLOT]
  float* const maha = ring + STAGES * SLOT;                // [NC]
  float* const re_s = maha + NC;                           // [NQ]
  const uint32_t bar_full = smem_u32(re_s + NQ);           // STAGES x 8 B
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  uint2* const table = reinterpret_cast<uint2*>(re_s + NQ + 4 * STAGES);  // {float offset, bytes} per slice

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const long long n_tiles = (p.n + NQ - 1) / NQ;
  const uint32_t slices_per_tile = K0_PAD / KC + (L - 2 + NS) * (H / KC);
  const uint32_t my_tiles = uint32_t((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t total_slices = my_tiles * slices_per_tile;

  for (uint32_t j = tid; j < slices_per_tile; j += THREADS) {
    uint32_t off;
    if (j < K0_PAD / KC) {
      off = j * KC * H;
    } else {
      const uint32_t r = j - K0_PAD / KC, l = 1 + r / (H / KC), c = r % (H / KC);
      off = (l < uint32_t(L - 1)) ? uint32_t(blob_layer_offset<H>(int(l))) + c * KC * H
                                  : uint32_t(blob_sweeps_offset<H>(L)) + (l - (L - 1)) * H * H + c * KC * H;
    }
    table[j] = make_uint2(off, KC * H * uint32_t(sizeof(float)));
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int lm = lane / LN, ln = lane % LN;
  const int wm = warp % WM, wn = warp / WM;   // wn: this warp's column group
  const int h_row = (wm * LM + lm) * 4;       // rows h_row..+3 and H/2 + h_row..+3 of a layer (or of a sweep)
  const int h_col = wn * 32 + ln * QT;        // columns h_col..+QT-1 and NC/2 + h_col..
  const int gtid = tid - wn * Cfg::GROUP_THREADS;
  auto group_bar = [&]() { named_bar_sync(1 + wn, Cfg::GROUP_THREADS); };

  const bool vec_ok =
      p.out_f64 ? ((p.out_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                : ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);

  // Weight stream: slice j of the CTA's run lands in ring slot j % STAGES and is issued by lane 0 of warp j % WARPS when
  // that warp starts slice j - LEAD.  The issuer waits until every warp has released slice j - STAGES, so a warp runs at
  // most STAGES - LEAD slices ahead of the slowest one without blocking, and finds its weights as long as it is less
  // than LEAD slices ahead of the issuers: LEAD = STAGES / 2 bounds the drift of the groups by half a ring either way.
  // (Measured and rejected: feeding the ring opportunistically -- whichever warp finds the next slot free claims the
  //  slice with a shared-memory atomic -- puts a shared-memory round trip on every warp's path at every slice: -3 %.)
  constexpr uint32_t LEAD = STAGES / 2;
  uint32_t it = 0;  // slices this warp has consumed
  auto issue = [&](uint32_t j) {
    const uint2 e = table[j % slices_per_tile];
    const uint32_t s = j & (STAGES - 1), ph = (j / STAGES) & 1;
    mbar_wait(bar_empty + 8 * s, ph ^ 1);
    mbar_arrive_expect_tx(bar_full + 8 * s, e.y);
    bulk_copy_g2s(smem_u32(ring + s * SLOT), p.blob + e.x, e.y, bar_full + 8 * s);
  };
  if (tid == 0)
    for (uint32_t j = 0; j < LEAD && j < total_slices; ++j) issue(j);
  auto wait_slice = [&](uint32_t& s_out) {
    const uint32_t j = it + LEAD;
    if (lane == 0 && (j & (WARPS - 1)) == uint32_t(warp) && j < total_slices) issue(j);
    s_out = it & (STAGES - 1);
    mbar_wait(bar_full + 8 * s_out, (it / STAGES) & 1);
  };
  auto release_slice = [&](uint32_t s) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    ++it;
  };

  const float* const sweep_w = p.blob + blob_sweeps_offset<H>(L);
  const float* const sweep_bias = sweep_w + size_t(M_LAST_SWEPT) * H;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long tile_q0 = tile * NQ;

    group_bar();  // this group's readers of its `act` columns, `maha`, `re_s` (previous tile) are done
    if (gtid < 64) {
      const bool twin = gtid >= 32;
      const int c = wn * 32 + (gtid & 31);
      const long long q = tile_q0 + c;
      featurise_column<NC>(p, q, q < p.n, twin, twin ? NQ + c : c, act, maha, re_s);
    }
    group_bar();

    // ---------------- hidden layers ----------------
    for (int l = 0; l < L - 1; ++l) {
      const int K = (l == 0) ? K0_PAD : H;
      const float* bias = p.blob + blob_layer_offset<H>(l) + size_t(K) * H;
      float2 acc[8][TN / 2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float b_lo = __ldg(bias + h_row + i), b_hi = __ldg(bias + H / 2 + h_row + i);
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) {
          acc[i][jj] = make_float2(b_lo, b_lo);
          acc[4 + i][jj] = make_float2(b_hi, b_hi);
        }
      }
      for (int c = 0; c < K / KC; ++c) {
        uint32_t s;
        wait_slice(s);
        ffma_slice<8, TN, H, NC, KC>(ring + s * SLOT + h_row, act + (c * KC) * NC + h_col, acc);
        release_slice(s);
      }
      group_bar();  // the group has finished reading this layer's input
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (i < 4) ? h_row + i : H / 2 + h_row + (i - 4);
#pragma unroll
        for (int v = 0; v < TN / 8; ++v) {
          float4 lo, hi;
          lo.x = silu(acc[i][2 * v].x); lo.y = silu(acc[i][2 * v].y);
          lo.z = silu(acc[i][2 * v + 1].x); lo.w = silu(acc[i][2 * v + 1].y);
          hi.x = silu(acc[i][TN / 4 + 2 * v].x); hi.y = silu(acc[i][TN / 4 + 2 * v].y);
          hi.z = silu(acc[i][TN / 4 + 2 * v + 1].x); hi.w = silu(acc[i][TN / 4 + 2 * v + 1].y);
          *reinterpret_cast<float4*>(act + row * NC + h_col + 4 * v) = lo;
          *reinterpret_cast<float4*>(act + row * NC + NC / 2 + h_col + 4 * v) = hi;
        }
      }
      group_bar();
    }

    // ---------------- last layer: NS sweeps of H packed rows, decode fused ----------------
    for (int sw = 0; sw < NS; ++sw) {
      // scalars-only output needs row groups 48 and 49 only: the other sweeps are consumed without arithmetic
      const bool wanted = !p.scalars_only || (sw + 1) * (H / 4) > 48;
      float2 acc[8][TN / 2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float b_lo = __ldg(sweep_bias + sw * H + h_row + i), b_hi = __ldg(sweep_bias + sw * H + H / 2 + h_row + i);
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) {
          acc[i][jj] = make_float2(b_lo, b_lo);
          acc[4 + i][jj] = make_float2(b_hi, b_hi);
        }
      }
      for (int c = 0; c < H / KC; ++c) {
        uint32_t s;
        wait_slice(s);
        if (wanted) ffma_slice<8, TN, H, NC, KC>(ring + s * SLOT + h_row, act + (c * KC) * NC + h_col, acc);
        release_slice(s);
      }
      if (!wanted) continue;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int g = sw * (H / 4) + half * (H / 8) + wm * LM + lm;  // packed row group: rows 4g..4g+3
#pragma unroll
        for (int v = 0; v < TN / 8; ++v) {  // groups of 4 queries (+ their twins)
          float yf[4][8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2(&a)[TN / 2] = acc[4 * half + i];
            yf[i][0] = a[2 * v].x; yf[i][1] = a[2 * v].y;
            yf[i][2] = a[2 * v + 1].x; yf[i][3] = a[2 * v + 1].y;
            yf[i][4] = a[TN / 4 + 2 * v].x; yf[i][5] = a[TN / 4 + 2 * v].y;
            yf[i][6] = a[TN / 4 + 2 * v + 1].x; yf[i][7] = a[TN / 4 + 2 * v + 1].y;
          }
          const int c4 = h_col + 4 * v;
          const long long q0 = tile_q0 + c4;
          const long long left = p.n - q0;
          const int nvalid = left >= 4 ? 4 : (left > 0 ? int(left) : 0);
          decode_and_store(p, g, q0, nvalid, vec_ok, yf, maha + c4, maha + NQ + c4, re_s + c4);
        }
      }
    }
  }
}

}  // namespace nfb
