// Reverse-mode gradient of the SCALAR outputs for large batches (the "fused reverse pass").
//
//   grad[t][q] = sum_{i < 6} cot[i][q] * d out_i[q] / d raw_t[q],      out = (analysis_confidence, CL, CD, CM, Top_Xtr, Bot_Xtr)
//
// i.e. the gradient, with respect to the 23 raw inputs, of any scalar objective of the six aerodynamic scalars -- what a
// design optimiser asks for (lift-to-drag, drag at fixed lift, ...) -- for every query of a batch, in about 1.6 forward
// passes.  The forward-mode kernel (nfb_jac.cuh) pays 24 passes per query for the same numbers.
//
// Same column-group schedule, weight ring and thread tile as nfb_ffma_groups.cuh.  Per group and tile:
//   forward   featurise; hidden layers as in the value kernel, each epilogue also storing SiLU'(z) of its element in a
//             per-CTA scratch block in global memory ([L-1][H][NC] fp32, L2-resident: 5 KB per query and layer stack);
//   turn      the scalar pass: the six latents of the query and of its twin (one query per lane), the six outputs
//             (stored if wanted), their cotangents pulled back through decode, squash and the un-flip average
//             [main.py:274-303], then d a_{L-2} = W_scal^T d latent, already multiplied by the stored slope, written
//             over the (no longer needed) activations;
//   backward  for l = L-2 .. 0: d a_{l-1} = W_l^T d z_l as the SAME register-blocked FFMA2 GEMM on a transposed copy of
//             the weights (slices of KC output units x H inputs), epilogue "times the stored slope" instead of SiLU;
//   finish    one thread per query: d x of the query + the twin's d x pulled back through the flip (a signed permutation),
//             the Mahalanobis term of the confidence logit (fp64, as in the forward pass), the chain rule of the
//             featurisation [main.py:171-184], 23 coalesced stores.
// FULL = true (nfb_eval_vjp_reverse_device) takes cotangents of all 198 outputs: the last layer's sweeps run as in the value
// kernel, their epilogue pulls the cotangents of theta, H and ue/Vinf back to the latents of the query and of its twin
// [main.py:274-316] and parks them in the scratch block; after the scalar pass (which parks its six rows there too) the
// parked rows come back into the tile H at a time and W_last^T (packed rows, transposed) takes them to d a_{L-2}.  theta's
// direct dependence on Re [main.py:309,314] is summed per query in a fixed order in the finish step.
#pragma once
#include "nfb_ffma_groups.cuh"
#include "nfb_jac.cuh"  // maha_gradient_regs

// Timing experiments only (results are wrong with any bit set): 1 no slope traffic, 2 trivial boundary-layer epilogue (no
// cotangent loads, no transcendental work, no output stores), 4 no finish step, 8 no featurisation, 16 no parked rows,
// 32 no turn.
#ifndef NFB_GRAD_EXP
#define NFB_GRAD_EXP 0
#endif

namespace nfb {

struct GradParams {
  EvalParams ev;      // inputs, n, n_layers, k_hidden; ev.blob = the GRADIENT blob below; ev.out = 6 x out_ld values or NULL
  const void* cot;    // 6 x cot_ld cotangents, type of the outputs (ev.out_f64)
  void* grad;         // 23 x grad_ld
  long long cot_ld, grad_ld;
  float* scratch;     // gridDim.x blocks of (L-1) * H * NC floats
};

// Gradient blob (floats): the whole value blob (nfb_ffma_groups.cuh: layers and last-layer sweeps), then the scalar rows
// of the last layer Wt[k][8] and their 8 biases, then the transposed layers WT_l[m][k] = W_l[m][k]: WT_0 as H x 32, WT_1 .. WT_{L-2}
// as H x H, then the last layer transposed by packed row, [208][H].
constexpr int GRAD_LAST_ROWS = 208;  // packed rows of the last layer the reverse sweep contracts over (200 used; a multiple of KC)

template <int H>
struct GradBlob {
  __host__ __device__ static constexpr size_t scal_off(int L) { return SweepBlob<H>::total(L); }          // behind the whole value blob
  __host__ __device__ static constexpr size_t scal_bias_off(int L) { return scal_off(L) + size_t(H) * 8; }
  __host__ __device__ static constexpr size_t wt0_off(int L) { return scal_bias_off(L) + 8; }             // WT_0[m][32]
  __host__ __device__ static constexpr size_t wt_off(int L, int l) {                                      // WT_l[m][H], l >= 1
    return wt0_off(L) + size_t(H) * K0_PAD + size_t(l - 1) * H * H;
  }
  __host__ __device__ static constexpr size_t wst_off(int L) { return wt_off(L, L - 1); }                 // W_last^T: [packed row][k]
  __host__ __device__ static constexpr size_t total(int L) { return wst_off(L) + size_t(GRAD_LAST_ROWS) * H; }
};

// per-CTA scratch (floats): SiLU slopes [L-1][H][NC]; FULL: parked latent cotangents [256][NC] and theta's d / d Re partials [32][NC]
template <int H, int NC>
__host__ __device__ constexpr size_t grad_scratch_floats(int L, bool full) {
  return size_t(L - 1) * H * NC + (full ? size_t(256 + N_BL) * NC : 0);
}

// Boundary-layer row group g of decode_and_store, reverse mode: from the four packed rows of 4 queries (yf[i][0..3]) and
// their twins (yf[i][4..7]) compute the outputs (stored if p.out), read their cotangents and return the cotangents of
// the 32 latents in dl (same layout) plus, for theta, its direct d / d Re share per query.  Groups 48.. belong to the scalar pass.
__device__ __forceinline__ void bl_group_vjp(const EvalParams& p, const GradParams& gp, int g, long long q0, int nvalid, bool vec_ok,
                                             const float (&yf)[4][8], const float* __restrict__ re4, float (&dl)[4][8],
                                             float (&re_part)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dl[i][j] = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) re_part[j] = 0.0f;
  if (g >= 48) return;
#if NFB_GRAD_EXP & 2
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 8; ++j) dl[i][j] = 0.5f * yf[i][j];
  return;
#endif
  int r[4];
  if (g < N_BL) {
    r[0] = 6 + g; r[1] = 6 + 2 * N_BL + g; r[2] = 6 + 3 * N_BL + g; r[3] = 6 + 5 * N_BL + g;
  } else {
    const int s = 2 * (g - N_BL);
    r[0] = 6 + N_BL + s; r[1] = r[0] + 1; r[2] = 6 + 4 * N_BL + s; r[3] = r[2] + 1;
  }
  float ct[4][4], o[4][4];
  // cotangents: one 16-byte load per row where the four queries exist and the block is aligned (the usual case)
  const bool cot_vec = nvalid == 4 && (p.out_f64 ? ((gp.cot_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(gp.cot) & 15) == 0 && (q0 & 1) == 0)
                                                 : ((gp.cot_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(gp.cot) & 15) == 0 && (q0 & 3) == 0));
  if (cot_vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (p.out_f64) {
        const double2* src = reinterpret_cast<const double2*>(static_cast<const double*>(gp.cot) + r[i] * gp.cot_ld + q0);
        const double2 a = __ldg(src), b = __ldg(src + 1);
        ct[i][0] = float(a.x); ct[i][1] = float(a.y); ct[i][2] = float(b.x); ct[i][3] = float(b.y);
      } else {
        const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(gp.cot) + r[i] * gp.cot_ld + q0));
        ct[i][0] = a.x; ct[i][1] = a.y; ct[i][2] = a.z; ct[i][3] = a.w;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ct[i][j] = 0.0f;
        if (j < nvalid)
          ct[i][j] = p.out_f64 ? float(__ldg(static_cast<const double*>(gp.cot) + r[i] * gp.cot_ld + q0 + j))
                               : __ldg(static_cast<const float*>(gp.cot) + r[i] * gp.cot_ld + q0 + j);
      }
  }
  if (g < N_BL) {  // theta_u, ue_u, theta_l, ue_l of station g
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float zt_u = 0.5f * (yf[0][j] + yf[2][4 + j]);
      const float zu_u = 0.5f * (yf[1][j] - yf[3][4 + j]);
      const float zt_l = 0.5f * (yf[2][j] + yf[0][4 + j]);
      const float zu_l = 0.5f * (yf[3][j] - yf[1][4 + j]);
      const float re = re4[j];
      const float pu = exp10f(zt_u), pl = exp10f(zt_l);
      const float th_u = (pu - 0.1f) / (fabsf(zu_u) * re), th_l = (pl - 0.1f) / (fabsf(zu_l) * re);
      o[0][j] = th_u; o[1][j] = zu_u; o[2][j] = th_l; o[3][j] = zu_l;
      // a zero cotangent contributes zero even where theta or its derivative is infinite (|ue| -> 0, 10^z overflow)
      const bool cu = ct[0][j] != 0.0f, cl = ct[2][j] != 0.0f;
      // the outputs above are the value kernel's expressions, bit for bit; the derivative terms below only feed the gradient and
      // use the two-instruction division (2 ulp) -- seven IEEE divisions per query and station were a sixth of this kernel's time
      const float g_ztu = cu ? __fdividef(ct[0][j] * 2.302585092994046f * pu, fabsf(zu_u) * re) : 0.0f;
      const float g_zuu = ct[1][j] - (cu ? __fdividef(ct[0][j] * th_u, zu_u) : 0.0f);
      const float g_ztl = cl ? __fdividef(ct[2][j] * 2.302585092994046f * pl, fabsf(zu_l) * re) : 0.0f;
      const float g_zul = ct[3][j] - (cl ? __fdividef(ct[2][j] * th_l, zu_l) : 0.0f);
      re_part[j] = -__fdividef((cu ? ct[0][j] * th_u : 0.0f) + (cl ? ct[2][j] * th_l : 0.0f), re);
      dl[0][j] = 0.5f * g_ztu; dl[2][4 + j] = 0.5f * g_ztu;
      dl[1][j] = 0.5f * g_zuu; dl[3][4 + j] = -0.5f * g_zuu;
      dl[2][j] = 0.5f * g_ztl; dl[0][4 + j] = 0.5f * g_ztl;
      dl[3][j] = 0.5f * g_zul; dl[1][4 + j] = -0.5f * g_zul;
    }
  } else {  // H_u, H_l of stations s, s + 1
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[0][j] = 2.6f * expf(0.5f * (yf[0][j] + yf[2][4 + j]));
      o[1][j] = 2.6f * expf(0.5f * (yf[1][j] + yf[3][4 + j]));
      o[2][j] = 2.6f * expf(0.5f * (yf[2][j] + yf[0][4 + j]));
      o[3][j] = 2.6f * expf(0.5f * (yf[3][j] + yf[1][4 + j]));
      const float g0 = ct[0][j] != 0.0f ? 0.5f * ct[0][j] * o[0][j] : 0.0f, g1 = ct[1][j] != 0.0f ? 0.5f * ct[1][j] * o[1][j] : 0.0f;
      const float g2 = ct[2][j] != 0.0f ? 0.5f * ct[2][j] * o[2][j] : 0.0f, g3 = ct[3][j] != 0.0f ? 0.5f * ct[3][j] * o[3][j] : 0.0f;
      dl[0][j] = g0; dl[2][4 + j] = g0;
      dl[1][j] = g1; dl[3][4 + j] = g1;
      dl[2][j] = g2; dl[0][4 + j] = g2;
      dl[3][j] = g3; dl[1][4 + j] = g3;
    }
  }
  if (p.out) {
#pragma unroll
    for (int i = 0; i < 4; ++i) store_row4(p, r[i], q0, nvalid, vec_ok, o[i]);
  }
}

template <class Cfg>
struct GradCfg {
  static constexpr int MAX_SLICES = K0_PAD / Cfg::KC + (2 * (NFB_MAX_LAYERS_K - 1) + Cfg::NFULL + Cfg::NHALF) * (Cfg::H / Cfg::KC) + 1 +
                                    GRAD_LAST_ROWS / Cfg::KC;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t(Cfg::H) * Cfg::NC + size_t(Cfg::STAGES) * Cfg::SLOT_FLOATS +
                                                        Cfg::NC + Cfg::NQ) + 16 * Cfg::STAGES + 8 * size_t(MAX_SLICES);
  static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

template <class Cfg, bool FULL>
__global__ void __launch_bounds__(Cfg::THREADS, 1) nfb_grad_kernel(const GradParams gp) {
  constexpr int H = Cfg::H, NC = Cfg::NC, NQ = Cfg::NQ, KC = Cfg::KC, STAGES = Cfg::STAGES, TN = Cfg::TN;
  constexpr int WM = Cfg::WM, SLOT = Cfg::SLOT_FLOATS, LM = Cfg::LM, LN = Cfg::LN, QG = Cfg::QG;
  constexpr int WARPS = Cfg::WARPS, QT = TN / 2;
  constexpr int NFULL = FULL ? Cfg::NFULL : 0, NHALF = FULL ? Cfg::NHALF : 0, NB = GRAD_LAST_ROWS;
  const EvalParams& p = gp.ev;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const act = reinterpret_cast<float*>(smem_raw);  // [H][NC]: activations, then their cotangents
  float* const ring = act + H * NC;                        // [STAGES][SLOT]
  float* const maha = ring + STAGES * SLOT;                // [NC]: d2 / 50 of each column, later the cotangent of it
  float* const re_s = maha + NC;                           // [NQ] (unused by the scalars; featurise_column writes it)
  const uint32_t bar_full = smem_u32(re_s + NQ);
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  uint2* const table = reinterpret_cast<uint2*>(re_s + NQ + 4 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const long long n_tiles = (p.n + NQ - 1) / NQ;
  using GB = GradBlob<H>;
  const int nkh = (p.k_hidden + KC - 1) / KC;
  const uint32_t slices_per_tile = K0_PAD / KC + (L - 2 + NFULL + NHALF) * nkh + 1 + (FULL ? NB / KC : 0) + (L - 1) * nkh;
  const uint32_t my_tiles = uint32_t((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t total_slices = my_tiles * slices_per_tile;
  float* const slopes = gp.scratch + size_t(blockIdx.x) * grad_scratch_floats<H, NC>(L, FULL);  // [L-1][H][NC]
  float* const dlat = slopes + size_t(L - 1) * H * NC;  // FULL: [256][NC] parked latent cotangents, [32][NC] d theta / d Re partials
  using SB = SweepBlob<H>;

  if (tid == 0) {
    uint32_t j = 0;
    for (int c = 0; c < K0_PAD / KC; ++c) table[j++] = make_uint2(c * KC * H, KC * H * 4);
    for (int l = 1; l < L - 1; ++l)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(blob_layer_offset<H>(l)) + c * KC * H, KC * H * 4);
    for (int sw = 0; sw < NFULL; ++sw)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(SB::full_off(L)) + (sw * H + c * KC) * H, KC * H * 4);
    for (int sw = 0; sw < NHALF; ++sw)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(SB::half_off(L)) + c * KC * (H / 2), KC * (H / 2) * 4);
    table[j++] = make_uint2(uint32_t(GB::scal_off(L)), uint32_t(p.k_hidden) * 8 * 4);
    if (FULL)
      for (int c = 0; c < NB / KC; ++c) table[j++] = make_uint2(uint32_t(GB::wst_off(L)) + c * KC * H, KC * H * 4);
    for (int l = L - 2; l >= 1; --l)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(GB::wt_off(L, l)) + c * KC * H, KC * H * 4);
    for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(GB::wt0_off(L)) + c * KC * K0_PAD, KC * K0_PAD * 4);
  }

  const int lm = lane / LN, ln = lane % LN;
  const int wm = warp % WM, wn = warp / WM;
  const int h_row = (wm * LM + lm) * 4;
  const int h_col = wn * QG + ln * QT;
  const int gtid = tid - wn * Cfg::GROUP_THREADS;
  auto group_bar = [&]() { named_bar_sync(1 + wn, Cfg::GROUP_THREADS); };

  // the weight stream (nfb_ring.cuh); init ends with the CTA barrier that also publishes `table`
  WeightRing<STAGES, Cfg::LEAD, WARPS, SLOT> wr;
  wr.init(ring, table, p.blob, bar_full, bar_empty, slices_per_tile, total_slices);
  // this thread's tile element (row i of 8, float4 v of its query columns / twin columns) in a [.][H][NC] block
  auto tile_offset = [&](int i, int v, bool twin) {
    const int row = (i < 4) ? h_row + i : H / 2 + h_row + (i - 4);
    return row * NC + (twin ? NC / 2 : 0) + h_col + 4 * v;
  };

  // Slopes come back from the scratch block (mostly from DRAM: the block is larger than the L2) in the epilogue of each reverse
  // GEMM; asking the L2 for them when that GEMM starts hides the trip (xlarge: the slope traffic cost 13 % of the kernel).
  auto prefetch_slopes = [&](const float* sl) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int v = 0; v < TN / 8; ++v)
#pragma unroll
        for (int tw = 0; tw < 2; ++tw)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(sl + tile_offset(i, v, tw == 1)));
  };

  // Featurisation of a tile's columns of this group: threads gtid < 2 QG, one column each.
  auto featurise_tile = [&](long long tq0) {
    for (int t = gtid; t < 2 * QG; t += Cfg::GROUP_THREADS) {
      const bool twin = t >= QG;
      const int c = wn * QG + (t % QG);
      const long long q = tq0 + c;
#if NFB_GRAD_EXP & 8
      const int col = twin ? NQ + c : c;
      maha[col] = 0.f;
      if (!twin) re_s[c] = 1e6f;
      for (int k = 0; k < K0_PAD; ++k) act[k * NC + col] = 1e-3f * float(k + (q & 7));
#else
      featurise_column<NC>(p, q, q < p.n, twin, twin ? NQ + c : c, act, maha, re_s);
#endif
    }
  };
  // The finish step of tile t and the featurisation of tile t + 1 are both fp64 work of a few threads (two lanes per query /
  // one thread per column) during which the group's GEMM stands still.  Where the group has warps to spare (4 QG <= its threads)
  // they run AT THE SAME TIME on different warps: the finish warps take their 26 inputs out of the tile into registers, one
  // barrier later the featurise warps may overwrite it.  (H = 64: two warps per group, both needed by either step: in turn.)
  constexpr bool SPLIT_STEPS = 4 * QG <= Cfg::GROUP_THREADS;
  constexpr int FIN0 = SPLIT_STEPS ? Cfg::GROUP_THREADS - 2 * QG : 0;  // first finish thread of the group
  static_assert(QG % 16 == 0 && 2 * QG <= Cfg::GROUP_THREADS && FIN0 % 32 == 0, "finish step: 16 queries and their twins per warp");

  if (blockIdx.x < n_tiles) {
    featurise_tile((long long)blockIdx.x * NQ);
    group_bar();
  }
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long tile_q0 = tile * NQ;

    // ---------------- forward: hidden layers, SiLU and its slope ----------------
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
      group_bar();
      float* const sl = slopes + size_t(l) * H * NC;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int v = 0; v < TN / 8; ++v) {
#pragma unroll
          for (int tw = 0; tw < 2; ++tw) {
            const float2 a0 = acc[i][tw * (TN / 4) + 2 * v], a1 = acc[i][tw * (TN / 4) + 2 * v + 1];
            const float h[4] = {a0.x, a0.y, a1.x, a1.y};
            float y[4], d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // silu(h) = h r, silu'(h) = r (1 + h (1 - r)), r = 1 / (1 + exp(-h))
              float ex, r;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(h[e] * -1.4426950408889634f));
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex));
              y[e] = h[e] * r;
              d[e] = r * fmaf(h[e], 1.0f - r, 1.0f);
            }
            const int off = tile_offset(i, v, tw == 1);
            *reinterpret_cast<float4*>(act + off) = make_float4(y[0], y[1], y[2], y[3]);
#if !(NFB_GRAD_EXP & 1)
            __stcg(reinterpret_cast<float4*>(sl + off), make_float4(d[0], d[1], d[2], d[3]));
#endif
          }
        }
      }
      group_bar();  // also orders the slope stores before the reads of other threads of the group
    }

    // ---------------- FULL: the last layer's sweeps; epilogue = outputs (if wanted) and the latents' cotangents, parked ----------------
    if constexpr (FULL) {
      const bool vec_ok =
          p.out_f64 ? ((p.out_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                    : ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
      auto vjp_rows = [&](const float2(&a4)[4][TN / 2], int g) {  // four packed rows 4g..4g+3 of this thread's columns
#pragma unroll
        for (int v = 0; v < TN / 8; ++v) {
          float yf[4][8], dl[4][8], re_part[4];
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
          bl_group_vjp(p, gp, g, q0, nvalid, vec_ok, yf, re_s + c4, dl, re_part);
#if NFB_GRAD_EXP & 16
          if (dl[0][0] == 123.456f) dlat[c4] = dl[1][1] + dl[2][2] + dl[3][3] + re_part[0];
#else
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float* const row = dlat + size_t(4 * g + i) * NC;
            __stcg(reinterpret_cast<float4*>(row + c4), make_float4(dl[i][0], dl[i][1], dl[i][2], dl[i][3]));
            __stcg(reinterpret_cast<float4*>(row + NQ + c4), make_float4(dl[i][4], dl[i][5], dl[i][6], dl[i][7]));
          }
          if (g < N_BL)
            __stcg(reinterpret_cast<float4*>(dlat + size_t(256 + g) * NC + c4), make_float4(re_part[0], re_part[1], re_part[2], re_part[3]));
#endif
        }
      };
      for (int sw = 0; sw < NFULL; ++sw) {
        float2 acc[8][TN / 2];
        const float* const sweep_bias = p.blob + SB::bias_off(L);
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
        const int g = sw * (H / 4) + wm * LM + lm;
        vjp_rows(reinterpret_cast<const float2(&)[4][TN / 2]>(acc[0]), g);
        vjp_rows(reinterpret_cast<const float2(&)[4][TN / 2]>(acc[4]), g + H / 8);
      }
      if constexpr (NHALF > 0) {
        float2 acc[4][TN / 2];
        const float* const sweep_bias = p.blob + SB::bias_off(L);
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
        vjp_rows(acc, SB::HALF_ROW0 / 4 + wm * LM + lm);
      }
      group_bar();  // the parked rows of groups 48, 49 (zeros) are written before the scalar pass overwrites them
    }

    // ---------------- turn: scalar latents, outputs, their cotangents, d a_{L-2} ----------------
    {
      const float* ws = wr.wait();
      if (wm == 0 && lane < QG && !(NFB_GRAD_EXP & 32)) {
        const int c = wn * QG + lane;
        const long long q = tile_q0 + c;
        const float* const bs = p.blob + GB::scal_bias_off(L);
        float yq[6], yt[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) yq[i] = yt[i] = __ldg(bs + i);
        const float* w = ws;
        float* const aq = act + c;
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
        // outputs: the expressions of decode_and_store's groups 48 and 49 [main.py:244-246, 292-303]
        const float z0 = 0.5f * ((yq[0] - maha[c]) + (yt[0] - maha[NQ + c]));
        const float z1 = 0.5f * (yq[1] - yt[1]);
        const float z2 = 0.5f * (yq[2] + yt[2]);
        const float z3 = 0.5f * (yq[3] - yt[3]);
        const float z4 = 0.5f * (yq[4] + yt[5]), z5 = 0.5f * (yq[5] + yt[4]);
        const float conf = 1.0f / (1.0f + expf(-z0)), cd = expf((z2 - 2.0f) * 2.0f);
        const bool valid = q < p.n;
        float ct[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (valid) {
          if (p.out_f64) {
#pragma unroll
            for (int i = 0; i < 6; ++i) ct[i] = float(__ldg(static_cast<const double*>(gp.cot) + i * gp.cot_ld + q));
          } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) ct[i] = __ldg(static_cast<const float*>(gp.cot) + i * gp.cot_ld + q);
          }
          if (p.out) {
            const float o[6] = {conf, 0.5f * z1, cd, z3 / 20.0f, clip01_keep_nan(z4), clip01_keep_nan(z5)};
            if (p.out_f64) {
#pragma unroll
              for (int i = 0; i < 6; ++i) static_cast<double*>(p.out)[i * p.out_ld + q] = double(o[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 6; ++i) static_cast<float*>(p.out)[i * p.out_ld + q] = o[i];
            }
          }
        }
        // cotangents of z0..z5, then of the latents of the query (dq) and of its twin (dt)
        const float g0 = ct[0] * conf * (1.0f - conf);
        const float g1 = ct[1] * 0.5f;
        const float g2 = ct[2] != 0.0f ? ct[2] * 2.0f * cd : 0.0f;  // CD = exp(.) may overflow far from the training data
        const float g3 = ct[3] / 20.0f;
        const float g4 = (z4 > 0.0f && z4 < 1.0f) ? ct[4] : 0.0f;
        const float g5 = (z5 > 0.0f && z5 < 1.0f) ? ct[5] : 0.0f;
        const float dq[6] = {0.5f * g0, 0.5f * g1, 0.5f * g2, 0.5f * g3, 0.5f * g4, 0.5f * g5};
        const float dt[6] = {0.5f * g0, -0.5f * g1, 0.5f * g2, -0.5f * g3, 0.5f * g5, 0.5f * g4};
        maha[c] = -0.5f * g0;       // cotangent of the query's d2 / 50
        maha[NQ + c] = -0.5f * g0;  // ... and of the twin's
        if constexpr (FULL) {  // park the six rows (packed rows 192..197) with the others
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            __stcg(dlat + size_t(SB::SCAL_ROW0 + i) * NC + c, dq[i]);
            __stcg(dlat + size_t(SB::SCAL_ROW0 + i) * NC + NQ + c, dt[i]);
          }
          // rows 198..207 are padding of the reverse sweep (zero weights): they must read as zero, never as whatever an
          // earlier launch left in the scratch block (a NaN there would poison this column: NaN * 0)
#pragma unroll
          for (int i = 6; i < NB - SB::SCAL_ROW0; ++i) {
            __stcg(dlat + size_t(SB::SCAL_ROW0 + i) * NC + c, 0.0f);
            __stcg(dlat + size_t(SB::SCAL_ROW0 + i) * NC + NQ + c, 0.0f);
          }
        } else {
          const float* const sl = slopes + size_t(L - 2) * H * NC + c;
#pragma unroll 4
          for (int k = 0; k < kh; ++k) {
            const float4 w0 = *reinterpret_cast<const float4*>(w + 8 * k);
            const float2 w1 = *reinterpret_cast<const float2*>(w + 8 * k + 4);
            const float sq = w0.x * dq[0] + w0.y * dq[1] + w0.z * dq[2] + w0.w * dq[3] + w1.x * dq[4] + w1.y * dq[5];
            const float st = w0.x * dt[0] + w0.y * dt[1] + w0.z * dt[2] + w0.w * dt[3] + w1.x * dt[4] + w1.y * dt[5];
            aq[k * NC] = sq * __ldcg(sl + k * NC);
            aq[k * NC + NQ] = st * __ldcg(sl + k * NC + NQ);
          }
        }
      }
      __syncwarp();  // the turn diverges (one query per lane, q < n); every lane has finished with the slot
      wr.release();
    }
    group_bar();

    // ---------------- FULL: d a_{L-2} = W_last^T d latent, the parked rows brought back H at a time ----------------
    if constexpr (FULL) {
      float2 acc[8][TN / 2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) acc[i][jj] = make_float2(0.0f, 0.0f);
      if (!(NFB_GRAD_EXP & 1)) prefetch_slopes(slopes + size_t(L - 2) * H * NC);
      for (int r0 = 0; r0 < NB; r0 += H) {
        const int rows_here = (NB - r0 < H) ? NB - r0 : H;
        if (r0 > 0) group_bar();  // the previous chunk has been read
        constexpr int V4 = 2 * QG / 4;  // float4s per row of the group's columns: queries, then twins
        for (int idx = gtid; idx < rows_here * V4; idx += Cfg::GROUP_THREADS) {
          const int r = idx / V4, cc = idx % V4;
          const int col = (cc < QG / 4) ? wn * QG + 4 * cc : NQ + wn * QG + 4 * (cc - QG / 4);
#if !(NFB_GRAD_EXP & 16)
          *reinterpret_cast<float4*>(act + r * NC + col) = __ldcg(reinterpret_cast<const float4*>(dlat + size_t(r0 + r) * NC + col));
#endif
        }
        group_bar();
        for (int c = 0; c < rows_here / KC; ++c) {
          const float* ws = wr.wait();
          ffma_slice<8, TN, H, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
          wr.release();
        }
      }
      group_bar();
      const float* const sl = slopes + size_t(L - 2) * H * NC;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int v = 0; v < TN / 8; ++v) {
#pragma unroll
          for (int tw = 0; tw < 2; ++tw) {
            const float2 a0 = acc[i][tw * (TN / 4) + 2 * v], a1 = acc[i][tw * (TN / 4) + 2 * v + 1];
            const int off = tile_offset(i, v, tw == 1);
            const float4 m = (NFB_GRAD_EXP & 1) ? make_float4(1.f, 1.f, 1.f, 1.f) : __ldcg(reinterpret_cast<const float4*>(sl + off));
            *reinterpret_cast<float4*>(act + off) = make_float4(a0.x * m.x, a0.y * m.y, a1.x * m.z, a1.y * m.w);
          }
        }
      }
      group_bar();
    }

    // ---------------- backward: d a_{l-1} = W_l^T d z_l ----------------
    for (int l = L - 2; l >= 1; --l) {
      float2 acc[8][TN / 2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) acc[i][jj] = make_float2(0.0f, 0.0f);
      if (!(NFB_GRAD_EXP & 1)) prefetch_slopes(slopes + size_t(l - 1) * H * NC);
      for (int c = 0; c < nkh; ++c) {
        const float* ws = wr.wait();
        ffma_slice<8, TN, H, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
        wr.release();
      }
      group_bar();  // the group has finished reading d z_l
      const float* const sl = slopes + size_t(l - 1) * H * NC;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int v = 0; v < TN / 8; ++v) {
#pragma unroll
          for (int tw = 0; tw < 2; ++tw) {
            const float2 a0 = acc[i][tw * (TN / 4) + 2 * v], a1 = acc[i][tw * (TN / 4) + 2 * v + 1];
            const int off = tile_offset(i, v, tw == 1);
            // d z_{l-1} = d a_{l-1} * SiLU'(z_{l-1}); this thread stored those slopes itself
            const float4 m = (NFB_GRAD_EXP & 1) ? make_float4(1.f, 1.f, 1.f, 1.f) : __ldcg(reinterpret_cast<const float4*>(sl + off));
            *reinterpret_cast<float4*>(act + off) = make_float4(a0.x * m.x, a0.y * m.y, a1.x * m.z, a1.y * m.w);
          }
        }
      }
      group_bar();
    }
    {  // layer 0: d x = W_0^T d z_0 has 32 rows only -- a 4-row thread tile on the threads whose rows exist
      const bool mine = h_row < K0_PAD;
      float2 acc[4][TN / 2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) acc[i][jj] = make_float2(0.0f, 0.0f);
      for (int c = 0; c < nkh; ++c) {
        const float* ws = wr.wait();
        if (mine) ffma_slice<4, TN, K0_PAD, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
        wr.release();
      }
      group_bar();
      if (mine) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int v = 0; v < TN / 8; ++v) {
#pragma unroll
            for (int tw = 0; tw < 2; ++tw) {
              const float2 a0 = acc[i][tw * (TN / 4) + 2 * v], a1 = acc[i][tw * (TN / 4) + 2 * v + 1];
              *reinterpret_cast<float4*>(act + tile_offset(i, v, tw == 1)) = make_float4(a0.x, a0.y, a1.x, a1.y);
            }
          }
        }
      }
      group_bar();
    }

    // ---------------- finish: rows 0..24 of the tile now hold d f / d x of the query and of its twin ----------------
    // Two lanes per query: lane l < 16 of a warp takes the query, lane l + 16 its twin (the 625 fp64 multiply-adds of the
    // Mahalanobis gradient, once per twin, were 12 % of the kernel when one thread did both in turn); the twin's vector comes
    // over by shuffle.  QG is 16 or 32: one or two warps of the group, all 32 lanes in step.
    const bool fin = gtid >= FIN0 && gtid < FIN0 + 2 * QG;
    const int fin_tw = lane >> 4;
    const int fin_c = wn * QG + ((gtid - FIN0) >> 5) * 16 + (lane & 15);
    float fin_v[N_IN], fin_m = 0.0f;
    if (fin) {
#pragma unroll
      for (int k = 0; k < N_IN; ++k) fin_v[k] = act[k * NC + fin_tw * NQ + fin_c];
      fin_m = maha[fin_tw * NQ + fin_c];
    }
    group_bar();  // the tile may be overwritten
    const long long next_tile = tile + gridDim.x;
    if (SPLIT_STEPS && next_tile < n_tiles) featurise_tile(next_tile * NQ);  // threads gtid < 2 QG: not the finish warps
    if (fin) {
      const int tw = fin_tw, c = fin_c;
      const long long q = tile_q0 + c;
      const bool valid = q < p.n;
#if NFB_GRAD_EXP & 4
      if (valid && tw == 0)
        for (int i = 0; i < N_RAW; ++i) static_cast<float*>(gp.grad)[i * gp.grad_ld + q] = fin_v[i];
#else
      double x[N_IN];
      float m_unused, re_unused;
      featurise_query<false>(p, q, valid, false, x, m_unused, re_unused);
      double G[N_IN], T[N_IN];
      {
        double xx[N_IN], gg[N_IN];
#pragma unroll
        for (int i = 0; i < N_IN; ++i) xx[i] = x[i];
        if (tw) {  // main.py:253-265
#pragma unroll
          for (int i = 0; i < 8; ++i) { xx[i] = -x[8 + i]; xx[8 + i] = -x[i]; }
          xx[16] = -x[16]; xx[18] = -x[18]; xx[23] = x[24]; xx[24] = x[23];
        }
        maha_gradient_regs(p, xx, gg);
        const double mc = double(fin_m) / (2.0 * N_IN);
#pragma unroll
        for (int k = 0; k < N_IN; ++k) {
          G[k] = double(fin_v[k]) + mc * gg[k];
          T[k] = __shfl_down_sync(0xffffffffu, G[k], 16);  // the query's lane receives its twin's
        }
      }
      if (valid && tw == 0) {
      // pull the twin's gradient back through the flip (its transpose: the same signed permutation)
#pragma unroll
      for (int i = 0; i < 8; ++i) { G[8 + i] -= T[i]; G[i] -= T[8 + i]; }
      G[16] -= T[16]; G[17] += T[17]; G[18] -= T[18]; G[19] += T[19]; G[20] += T[20]; G[21] += T[21]; G[22] += T[22];
      G[23] += T[24]; G[24] += T[23];
      // featurisation [main.py:171-184]
      constexpr double kdeg = 3.141592653589793238462643383279502884 / 180.0;
      const double alpha = p.in_f64 ? __ldg(static_cast<const double*>(p.in[18]) + q * p.in_stride[18])
                                    : double(__ldg(static_cast<const float*>(p.in[18]) + q * p.in_stride[18]));
      const double rey = p.in_f64 ? __ldg(static_cast<const double*>(p.in[19]) + q * p.in_stride[19])
                                  : double(__ldg(static_cast<const float*>(p.in[19]) + q * p.in_stride[19]));
      const double a = alpha * kdeg;
      double out[N_RAW];
#pragma unroll
      for (int i = 0; i < 17; ++i) out[i] = G[i];
      out[17] = 50.0 * G[17];
      out[18] = kdeg * (2.0 * cos(2.0 * a) * G[18] - sin(a) * G[19] + 2.0 * cos(a) * sin(a) * G[20]);
      out[19] = G[21] / (3.5 * rey);
      if constexpr (FULL) {  // theta = (10^z - 0.1) / (|ue| Re): the direct dependence, station by station in a fixed order
        float sre = 0.0f;
        for (int g = 0; g < N_BL; ++g) sre += __ldcg(dlat + size_t(256 + g) * NC + c);
        out[19] += double(sre);
      }
      out[20] = G[22] / 4.5;
      out[21] = G[23];
      out[22] = G[24];
      if (p.out_f64) {
#pragma unroll
        for (int i = 0; i < N_RAW; ++i) static_cast<double*>(gp.grad)[i * gp.grad_ld + q] = out[i];
      } else {
#pragma unroll
        for (int i = 0; i < N_RAW; ++i) static_cast<float*>(gp.grad)[i * gp.grad_ld + q] = float(out[i]);
      }
      }  // valid && tw == 0
#endif
    }
    if (!SPLIT_STEPS && next_tile < n_tiles) featurise_tile(next_tile * NQ);  // the same threads: after their finish step
    group_bar();  // the next tile's inputs are in place
  }
}

}  // namespace nfb
