// fp32 (FFMA2) evaluator, column-group schedule: the same arithmetic as nfb_ffma.cuh, scheduled so that the
// phases that do not use the FMA pipe (fp64 featurisation, SiLU, the MUFU-heavy decode, waiting for weights)
// overlap the GEMM of other warps.
//
// A CTA still owns a tile of NC = 32768 / H columns in shared memory, but the tile is split into WN independent
// COLUMN GROUPS (QG queries and their QG twins, WM warps each; QG = 32, or 16 for H = 512).  A group does everything
// for its own queries -- featurise, every hidden layer, the last layer, the decode, the stores -- and synchronises
// only with itself (one named barrier per group).  Nothing in the tile loop is CTA-wide, so the groups drift apart:
// while one group evaluates SiLUs or decodes, the others keep the FMA pipe busy.  Measured on B200 against the
// CTA-synchronous kernel of nfb_ffma.cuh: xlarge 108.8 -> 143 M queries/s, xxsmall 478 -> 740, xxxlarge 10.09 -> 10.38
// (profiles/r1_size_sweep_column_group_kernels.txt).
//
// The only thing the groups share is the weight stream: one ring of TMA bulk copies per CTA, which every warp
// consumes in the same order.  The ring is deep, because its depth is what bounds the drift (see wait_slice).
//
// Last layer: SWEEPS that look exactly like hidden layers (same 8 x TN thread tile, same slices) whose epilogue is
// un-flip + average + decode + store instead of SiLU.  Packed row p = 4 g + i holds latent row pack_last_layer_row(p);
// a thread owns two row groups g per full sweep.  The 198 rows are 48 groups of boundary-layer rows and 6 scalars:
//   H =  64: three full sweeps (groups 0..47) + the scalar pass
//   H = 128: one full sweep (groups 0..31) + one half sweep (4-row thread tile, groups 32..47) + the scalar pass
//   H = 256: one full sweep of 256 packed rows (50 of 64 groups used)
//   H = 512: two column groups of 16 queries (four warps each); the 256 packed rows are one half sweep
// The scalar pass is one more slice of the stream (8 packed rows x K): the group's first warp evaluates the six
// scalar latents of its 32 queries and their twins, one query per lane, and stores them as 128-byte rows.  Without
// it the scalars would cost a whole sweep of their own (H = 64: a quarter of the last layer, which is 73 % of
// xxsmall).  With NFB_OUTPUT_SCALARS the boundary-layer sweeps are not even streamed.
// K is trimmed to the model's real hidden width rounded up to a slice (48-wide models: 6 of 8 slices per layer).
// Every accumulator sees the bias, then one IEEE fma per k in k order: results are bit-identical to
// nfb_ffma_kernel, nfb_small_kernel and the value columns of nfb_jac_kernel (tests/test_gpu_parity.py).
#pragma once
#include "nfb_ffma.cuh"
#include "nfb_ring.cuh"

#ifndef NFB_GRP_DECODE_ONCE
#define NFB_GRP_DECODE_ONCE 0
#endif
// Timing experiments only (scripts/groups_probe.cu; results are wrong with any bit set): 1 no featurisation, 2 no SiLU
// arithmetic, 4 no decode, 8 no group barriers, 64 no last layer.
#ifndef NFB_EXP
#define NFB_EXP 0
#endif

namespace nfb {

constexpr int M_LAST_SWEPT = 256;  // last layer's packed rows in the sweep layout (64 groups of four, 50 used)

// The sweep-layout copy of the last layer sits behind the blob of nfb_ffma.cuh:
//   full sweeps Wt[NFULL][k][H], half sweep Wt[k][H / 2], scalar rows Wt[k][8], bias[256] by packed row.
template <int H>
struct SweepBlob {
  static constexpr int NFULL = H == 64 ? 3 : H == 512 ? 0 : 1;     // H = 512: the 256 packed rows are one half sweep
  static constexpr int NHALF = (H == 128 || H == 512) ? 1 : 0;
  static constexpr int NSCAL = H >= 256 ? 0 : 1;                   // H >= 256: the scalars ride in the sweep
  static constexpr int HALF_ROW0 = NFULL * H;  // first packed row of the half sweep (H = 128: 128)
  static constexpr int SCAL_ROW0 = 192;        // packed rows 192..197: the six scalar latents in latent order
  __host__ __device__ static constexpr size_t full_off(int n_layers) { return blob_total_floats<H>(n_layers); }
  __host__ __device__ static constexpr size_t half_off(int n_layers) { return full_off(n_layers) + size_t(NFULL) * H * H; }
  __host__ __device__ static constexpr size_t scal_off(int n_layers) { return half_off(n_layers) + size_t(NHALF) * H * (H / 2); }
  __host__ __device__ static constexpr size_t bias_off(int n_layers) { return scal_off(n_layers) + size_t(NSCAL) * H * 8; }
  __host__ __device__ static constexpr size_t total(int n_layers) { return bias_off(n_layers) + M_LAST_SWEPT; }
};

template <int H_, int KC_, int STAGES_, int TN_, int QG_ = 32, int LEAD_ = 0>
struct GroupCfg {
  static constexpr int H = H_;
  static constexpr int KC = KC_;
  static constexpr int STAGES = STAGES_;
  static constexpr int TN = TN_;
  static constexpr int LEAD = LEAD_ > 0 ? LEAD_ : STAGES_ / 2;  // slices the copies run ahead of their issuer
  static constexpr int QG = QG_;                // queries per column group: 32, or 16 for the 512-wide model
  static constexpr int WARPS = 128 / TN;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int LN = 2 * QG / TN;        // lanes along columns: a warp spans its group's 2 QG columns
  static constexpr int LM = 32 / LN;            // lanes along rows
  static constexpr int NC = 32768 / H;
  static constexpr int NQ = NC / 2;
  static constexpr int WM = H / (8 * LM);       // warps per column group
  static constexpr int WN = WARPS / WM;         // column groups per CTA
  static constexpr int GROUP_THREADS = WM * 32;
  static constexpr int NFULL = SweepBlob<H>::NFULL;  // full sweeps: H packed rows each
  static constexpr int NHALF = SweepBlob<H>::NHALF;  // half sweep: H / 2 packed rows (4-row thread tile)
  static constexpr int NSCAL = SweepBlob<H>::NSCAL;  // scalar pass: packed rows 192..199
  static constexpr int SLOT_FLOATS = KC * H;
  static constexpr int MAX_SLICES = K0_PAD / KC + (NFB_MAX_LAYERS_K - 2 + NFULL + NHALF) * (H / KC) + NSCAL;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t(H) * NC + size_t(STAGES) * SLOT_FLOATS + NC + NQ) +
                                       16 * STAGES + 8 * size_t(MAX_SLICES);
  static_assert(TN == 8 || TN == 16, "thread tile is 8 or 16 columns wide");
  static_assert(KC >= 8 && H % (8 * LM) == 0 && WARPS % WM == 0 && WN * 2 * QG == NC, "bad width");
  static_assert(NSCAL == 0 || QG == 32, "the scalar pass is one query per lane");
  static_assert(STAGES >= 4 && LEAD >= 1 && LEAD < STAGES, "ring too shallow");
  static_assert(K0_PAD % KC == 0 && H % KC == 0, "KC must divide every K");
  static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) nfb_ffma_groups_kernel(const EvalParams p) {
  constexpr int H = Cfg::H, NC = Cfg::NC, NQ = Cfg::NQ, KC = Cfg::KC, STAGES = Cfg::STAGES, TN = Cfg::TN;
  constexpr int WM = Cfg::WM, SLOT = Cfg::SLOT_FLOATS, LM = Cfg::LM, LN = Cfg::LN;
  constexpr int NFULL = Cfg::NFULL, NHALF = Cfg::NHALF, NSCAL = Cfg::NSCAL, QG = Cfg::QG;
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
  using SB = SweepBlob<H>;
  const int nkh = (p.k_hidden + KC - 1) / KC;                   // K-slices of a hidden / last layer (trimmed K)
  const bool bl_wanted = !p.scalars_only;                        // boundary-layer rows wanted?
  const int n_full = (bl_wanted || NSCAL == 0) ? NFULL : 0;      // H >= 256: the sweep also holds the scalars
  const int n_half = (bl_wanted || NSCAL == 0) ? NHALF : 0;
  const uint32_t slices_per_tile = K0_PAD / KC + (L - 2 + n_full + n_half) * nkh + NSCAL;
  const uint32_t my_tiles = uint32_t((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t total_slices = my_tiles * slices_per_tile;

  if (tid == 0) {  // the tile's slice sequence: {float offset in the blob, bytes}
    uint32_t j = 0;
    for (int c = 0; c < K0_PAD / KC; ++c) table[j++] = make_uint2(c * KC * H, KC * H * 4);
    for (int l = 1; l < L - 1; ++l)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(blob_layer_offset<H>(l)) + c * KC * H, KC * H * 4);
    for (int sw = 0; sw < n_full; ++sw)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(SB::full_off(L)) + (sw * H + c * KC) * H, KC * H * 4);
    for (int sw = 0; sw < n_half; ++sw)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(SB::half_off(L)) + c * KC * (H / 2), KC * (H / 2) * 4);
    if (NSCAL) table[j++] = make_uint2(uint32_t(SB::scal_off(L)), uint32_t(p.k_hidden) * 8 * 4);
  }
  const int lm = lane / LN, ln = lane % LN;
  const int wm = warp % WM, wn = warp / WM;   // wn: this warp's column group
  const int h_row = (wm * LM + lm) * 4;       // rows h_row..+3 and H/2 + h_row..+3 of a layer (or of a sweep)
  const int h_col = wn * QG + ln * QT;        // columns h_col..+QT-1 and NC/2 + h_col..
  const int gtid = tid - wn * Cfg::GROUP_THREADS;
#if NFB_EXP & 8
  auto group_bar = [&]() {};
#else
  auto group_bar = [&]() { named_bar_sync(1 + wn, Cfg::GROUP_THREADS); };
#endif

  const bool vec_ok =
      p.out_f64 ? ((p.out_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                : ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);

  // Weight stream: see nfb_ring.cuh.  LEAD = STAGES / 2 bounds the drift of the groups by half a ring either way.
  // (Measured and rejected: feeding the ring opportunistically -- whichever warp finds the next slot free claims the
  //  slice with a shared-memory atomic -- puts a shared-memory round trip on every warp's path at every slice: -3 %.)
  WeightRing<STAGES, Cfg::LEAD, WARPS, SLOT> wr;
  wr.init(ring, table, p.blob, bar_full, bar_empty, slices_per_tile, total_slices);

  const float* const sweep_bias = p.blob + SB::bias_off(L);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long tile_q0 = tile * NQ;

    group_bar();  // this group's readers of its `act` columns, `maha`, `re_s` (previous tile) are done
    for (int t = gtid; t < 2 * QG; t += Cfg::GROUP_THREADS) {
      const bool twin = t >= QG;
      const int c = wn * QG + (t % QG);
      const long long q = tile_q0 + c;
#if NFB_EXP & 1
      const int col = twin ? NQ + c : c;
      maha[col] = 0.f;
      if (!twin) re_s[c] = 1e6f;
      for (int k = 0; k < K0_PAD; ++k) act[k * NC + col] = 1e-3f * float(k + (q & 7));
#else
      featurise_column<NC>(p, q, q < p.n, twin, twin ? NQ + c : c, act, maha, re_s);
#endif
    }
    group_bar();

    // ---------------- hidden layers ----------------
    for (int l = 0; l < L - 1; ++l) {
      const float* bias = p.blob + blob_layer_offset<H>(l) + size_t(l == 0 ? K0_PAD : H) * H;
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
      const int nk = (l == 0) ? K0_PAD / KC : nkh;
      for (int c = 0; c < nk; ++c) {
        const float* ws = wr.wait();
        ffma_slice<8, TN, H, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
        wr.release();
      }
      group_bar();  // the group has finished reading this layer's input
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (i < 4) ? h_row + i : H / 2 + h_row + (i - 4);
#pragma unroll
        for (int v = 0; v < TN / 8; ++v) {
          float4 lo, hi;
#if NFB_EXP & 2
#define silu(x) (x)
#endif
          lo.x = silu(acc[i][2 * v].x); lo.y = silu(acc[i][2 * v].y);
          lo.z = silu(acc[i][2 * v + 1].x); lo.w = silu(acc[i][2 * v + 1].y);
          hi.x = silu(acc[i][TN / 4 + 2 * v].x); hi.y = silu(acc[i][TN / 4 + 2 * v].y);
          hi.z = silu(acc[i][TN / 4 + 2 * v + 1].x); hi.w = silu(acc[i][TN / 4 + 2 * v + 1].y);
#if NFB_EXP & 2
#undef silu
#endif
          *reinterpret_cast<float4*>(act + row * NC + h_col + 4 * v) = lo;
          *reinterpret_cast<float4*>(act + row * NC + NC / 2 + h_col + 4 * v) = hi;
        }
      }
      group_bar();
    }

    // ---------------- last layer: sweeps with the decode fused (the activations are only read) ----------------
    auto decode_rows = [&](const float2(&a4)[4][TN / 2], int g) {  // four packed rows 4g..4g+3 of this thread's columns
#pragma unroll
      for (int v = 0; v < TN / 8; ++v) {  // groups of 4 queries (+ their twins)
        float yf[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          yf[i][0] = a4[i][2 * v].x; yf[i][1] = a4[i][2 * v].y;
          yf[i][2] = a4[i][2 * v + 1].x; yf[i][3] = a4[i][2 * v + 1].y;
          yf[i][4] = a4[i][TN / 4 + 2 * v].x; yf[i][5] = a4[i][TN / 4 + 2 * v].y;
          yf[i][6] = a4[i][TN / 4 + 2 * v + 1].x; yf[i][7] = a4[i][TN / 4 + 2 * v + 1].y;
        }
        const int c4 = h_col + 4 * v;
        const long long q0 = tile_q0 + c4;
        const long long left = p.n - q0;
        const int nvalid = left >= 4 ? 4 : (left > 0 ? int(left) : 0);
#if NFB_EXP & 4
        float sum = 0.f;
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 8; ++j) sum += yf[i][j];
        if (sum == 123.456f) static_cast<float*>(p.out)[q0] = sum + float(g + nvalid);
#else
        decode_and_store(p, g, q0, nvalid, vec_ok, yf, maha + c4, maha + NQ + c4, re_s + c4);
#endif
      }
    };
#if NFB_EXP & 64
    continue;
#endif
    for (int sw = 0; sw < n_full; ++sw) {
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
      for (int c = 0; c < nkh; ++c) {
        const float* ws = wr.wait();
        ffma_slice<8, TN, H, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
        wr.release();
      }
      const int g = sw * (H / 4) + wm * LM + lm;  // packed row groups g and g + H / 8
#if NFB_GRP_DECODE_ONCE
      // A/B only: one copy of the decode code for both row groups (register selects instead of a second inlined
      // instance), 10 % less SASS for the desynchronised groups to walk.  Measured: no gain (64-wide -1 %, others within
      // the +-2 % run-to-run spread).
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        float2 a4[4][TN / 2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < TN / 2; ++jj) a4[i][jj] = half ? acc[4 + i][jj] : acc[i][jj];
        decode_rows(a4, g + half * (H / 8));
      }
#else
      decode_rows(reinterpret_cast<const float2(&)[4][TN / 2]>(acc[0]), g);
      decode_rows(reinterpret_cast<const float2(&)[4][TN / 2]>(acc[4]), g + H / 8);
#endif
    }
    if constexpr (NHALF > 0) {
      if (n_half) {  // packed rows HALF_ROW0 + h_row..+3: one row group per thread
        float2 acc[4][TN / 2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float b = __ldg(sweep_bias + SB::HALF_ROW0 + h_row + i);
#pragma unroll
          for (int jj = 0; jj < TN / 2; ++jj) acc[i][jj] = make_float2(b, b);
        }
        for (int c = 0; c < nkh; ++c) {
          const float* ws = wr.wait();
          ffma_slice<4, TN, H / 2, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
          wr.release();
        }
        decode_rows(acc, SB::HALF_ROW0 / 4 + wm * LM + lm);
      }
    }
    if constexpr (NSCAL > 0) {
      // scalar pass: lane = query of the group; accumulators for the query (yq) and its twin (yt), latent rows 0..5
      const float* ws = wr.wait();
      if (wm == 0) {
        const int c = wn * QG + lane;
        float yq[6], yt[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) yq[i] = yt[i] = __ldg(sweep_bias + SB::SCAL_ROW0 + i);
        const float* w = ws;
        const float* aq = act + c;
        const int kh = p.k_hidden;
#pragma unroll 4
        for (int k = 0; k < kh; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(w + 8 * k);
          const float2 w1 = *reinterpret_cast<const float2*>(w + 8 * k + 4);
          const float xq = aq[k * NC], xt = aq[k * NC + NQ];
          yq[0] = fmaf(w0.x, xq, yq[0]); yt[0] = fmaf(w0.x, xt, yt[0]);
          yq[1] = fmaf(w0.y, xq, yq[1]); yt[1] = fmaf(w0.y, xt, yt[1]);
          yq[2] = fmaf(w0.z, xq, yq[2]); yt[2] = fmaf(w0.z, xt, yt[2]);
          yq[3] = fmaf(w0.w, xq, yq[3]); yt[3] = fmaf(w0.w, xt, yt[3]);
          yq[4] = fmaf(w1.x, xq, yq[4]); yt[4] = fmaf(w1.x, xt, yt[4]);
          yq[5] = fmaf(w1.y, xq, yq[5]); yt[5] = fmaf(w1.y, xt, yt[5]);
        }
        // the expressions of decode_and_store's groups 48 and 49 [main.py:244-246, 292-303]
        const float z0 = 0.5f * ((yq[0] - maha[c]) + (yt[0] - maha[NQ + c]));
        const float z1 = 0.5f * (yq[1] - yt[1]);
        const float z2 = 0.5f * (yq[2] + yt[2]);
        const float z3 = 0.5f * (yq[3] - yt[3]);
        float o[6];
        o[0] = 1.0f / (1.0f + expf(-z0));
        o[1] = 0.5f * z1;
        o[2] = expf((z2 - 2.0f) * 2.0f);
        o[3] = z3 / 20.0f;
        o[4] = clip01_keep_nan(0.5f * (yq[4] + yt[5]));
        o[5] = clip01_keep_nan(0.5f * (yq[5] + yt[4]));
        const long long q = tile_q0 + c;
        if (q < p.n) {
          if (p.out_f64) {
#pragma unroll
            for (int i = 0; i < 6; ++i) __stcs(static_cast<double*>(p.out) + i * p.out_ld + q, double(o[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) __stcs(static_cast<float*>(p.out) + i * p.out_ld + q, o[i]);
          }
        }
      }
      __syncwarp();  // the scalar pass diverges on q < p.n
      wr.release();
    }
  }
}

}  // namespace nfb
