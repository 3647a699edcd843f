// Fused fp32 (FFMA) evaluator for NeuralFoil's learned core -- one persistent kernel per model width.
//
// What one CTA does per tile of NQ queries (reference neuralfoil/main.py line numbers in brackets):
//   1. featurise: 23 raw scalars -> 25 latent inputs in fp64, for the query and for its flipped twin
//      [171-184, 253-265]; squared Mahalanobis distance of both in fp64 [47-61]; the 25 inputs are
//      rounded to fp32 into the activation tile in shared memory.
//   2. every hidden layer: act <- SiLU(W act + b) [218-241] as a register-blocked FFMA GEMM, the
//      tile (H x NC fp32) never leaving shared memory; weights arrive in K-slices through a ring of
//      TMA bulk copies (L2 -> smem) issued by lane 0 of the warps in turn (see issue_slice).
//   3. last layer with its rows permuted so that each thread owns, for four queries AND their twins,
//      exactly the rows that un-flip, average, squash and decode need [244-246, 268-316]; results go
//      straight from registers to global memory as 16-byte streaming stores, full 128-byte lines.
//
// Column c < NQ of a tile is query c; column NQ + c is its twin.  Both run through the same weights
// in the same order, so the reference's exact symmetries (tests/test_invariants.py) hold bit for bit.
#pragma once
#include "nfb_common.cuh"

#ifndef NFB_GROUP_BARS
#define NFB_GROUP_BARS 0
#endif

#ifndef NFB_SLICE_ORDER
#define NFB_SLICE_ORDER 0
#endif

namespace nfb {

template <int H_, int KC_, int STAGES_, int TN_>
struct FfmaCfg {
  static constexpr int H = H_;                  // padded hidden width: 64, 128, 256 or 512
  static constexpr int KC = KC_;                // K rows per streamed weight slice
  static constexpr int STAGES = STAGES_;        // ring depth
  static constexpr int TN = TN_;                // columns per thread: 8 (4 queries + twins) or 16 (8 + twins)
  static constexpr int WARPS = 128 / TN;        // 64 accumulators x 512 threads == 128 x 256 threads
  static constexpr int THREADS = WARPS * 32;
  static constexpr int LN = 64 / TN;            // lanes along columns: a warp always spans 64 columns
  static constexpr int LM = 32 / LN;            // lanes along rows
  static constexpr int NC = 32768 / H;          // columns per tile (queries + twins)
  static constexpr int NQ = NC / 2;             // queries per tile
  static constexpr int WM = H / (8 * LM);       // warps along rows (hidden layers, 8 rows per thread)
  static constexpr int WN = WARPS / WM;         // warps along columns == last-layer passes == NC / 64
  static constexpr int SLOT_FLOATS = KC * (H > M_LAST ? H : M_LAST);
  static constexpr int MAX_SLICES =  // slices per tile at the deepest supported network
      K0_PAD / KC + (NFB_MAX_LAYERS_K - 2) * (H / KC) + WN * (H / KC);
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t(H) * NC + size_t(STAGES) * SLOT_FLOATS + NC + NQ) +
                                       16 * STAGES + 8 * size_t(MAX_SLICES);
  static_assert(TN == 8 || TN == 16, "thread tile is 8 or 16 columns wide");
  static_assert(H % (8 * LM) == 0 && WARPS % WM == 0 && WN * 64 == NC, "bad width");
  static_assert(WARPS * LM * 4 >= M_LAST_USED, "last layer does not fit the row groups");
  static_assert((STAGES & (STAGES - 1)) == 0 && STAGES >= 4, "STAGES must be a power of two >= 4");
  static_assert(K0_PAD % KC == 0 && H % KC == 0, "KC must divide every K");
};

// Packed-blob offsets (floats).  Layer 0: Wt[32][H], b[H]; hidden l: Wt[H][H], b[H]; last: Wt[H][208], b[208].
template <int H>
__host__ __device__ constexpr size_t blob_layer_offset(int l) {
  return l == 0 ? 0 : size_t(K0_PAD + 1) * H + size_t(l - 1) * (size_t(H) * H + H);
}
template <int H>
__host__ __device__ constexpr size_t blob_total_floats(int n_layers) {
  return blob_layer_offset<H>(n_layers - 1) + size_t(H) * M_LAST + M_LAST;
}

// SiLU as the reference's swish: x / (1 + exp(-x)) [main.py:239].  exp overflow must give -0, not NaN.
// Five instructions: FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL.  exp(-h) = ex2(-h log2(e)) with the product rounded to fp32
// carries a relative error of |h| * 8.6e-8, but what reaches the activation is at most h^2 e^-|h| * 8.6e-8 < 5e-8 in
// absolute terms -- below the rounding of the activation itself -- so libdevice's expf (a dozen FMA-pipe instructions
// per activation, 14 % of the pipe's time on the 128-wide models once the GEMM overlaps everything else) buys nothing.
// Large negative h: ex2 -> inf, rcp(inf) = 0, h * 0 = -0 as in the reference; NaN propagates.
__device__ __forceinline__ float silu(float h) {
#ifdef NFB_SILU_LIBDEVICE  // A/B only: the previous formulation
  const float e = expf(fminf(-h, 80.0f));
  return __fdividef(h, 1.0f + e);
#else
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(h * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return h * r;
#endif
}

// One K-slice of the register-blocked GEMM: acc[TM][8] += W[k][rows] * act[k][cols], k = 0..KC-1.
// `w` points at this thread's first row inside the slice (row stride MP floats per k);
// `a` points at this thread's first column in the activation tile (row stride NC).
//
// The multiply-adds are issued as packed FFMA2 (fma.rn.f32x2, sm_100+): one weight scalar times a
// pair of adjacent columns.  Measured on B200 (scripts/ffma_peak.cu): the scalar 8x8 outer product
// tops out at 85 of 128 FFMA lanes/clk/SM because every FFMA reads two new 32-bit registers that ptxas
// cannot keep in opposite banks; FFMA2 with the column pair held in the operand-reuse cache while the
// eight rows stream past it (j outer, i inner) reaches 126.  Each accumulator still sees one IEEE
// fma per k, in k order, so results are bit-identical to the scalar loop.
// acc[i][jj] holds columns (2 jj, 2 jj + 1) of row i: jj < TN/4 -> queries, the rest -> their twins.
template <int TM, int TN, int MP, int NC>
__device__ __forceinline__ void load_fragments(const float* __restrict__ w, const float* __restrict__ a,
                                               float (&wv)[TM], float2 (&av)[TN / 2]) {
  {
    const float4 t = *reinterpret_cast<const float4*>(w);
    wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
  }
  if constexpr (TM == 8) {
    const float4 t = *reinterpret_cast<const float4*>(w + MP / 2);
    wv[4] = t.x; wv[5] = t.y; wv[6] = t.z; wv[7] = t.w;
  }
#pragma unroll
  for (int v = 0; v < TN / 8; ++v) {  // queries, then twins (NC / 2 columns further)
    const float4 t = *reinterpret_cast<const float4*>(a + 4 * v);
    av[2 * v] = make_float2(t.x, t.y); av[2 * v + 1] = make_float2(t.z, t.w);
    const float4 u = *reinterpret_cast<const float4*>(a + NC / 2 + 4 * v);
    av[TN / 4 + 2 * v] = make_float2(u.x, u.y); av[TN / 4 + 2 * v + 1] = make_float2(u.z, u.w);
  }
}

// Two schedules of the same arithmetic.  For the 8 x 8 tile the k loop is software-pipelined by hand
// (fragments of k + 1 are requested before the FFMA2 block of k); for the 8 x 16 tile, which already
// uses every register, the plain loop lets ptxas place the loads (measured: DESIGN.md section 6).
template <int TM, int TN, int MP, int NC, int KC>
__device__ __forceinline__ void ffma_slice(const float* __restrict__ w, const float* __restrict__ a,
                                           float2 (&acc)[TM][TN / 2]) {
  if constexpr (TN == 8) {
    float wv[2][TM];
    float2 av[2][TN / 2];
    load_fragments<TM, TN, MP, NC>(w, a, wv[0], av[0]);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const int cur = kk & 1;
      if (kk + 1 < KC) load_fragments<TM, TN, MP, NC>(w + (kk + 1) * MP, a + (kk + 1) * NC, wv[cur ^ 1], av[cur ^ 1]);
#pragma unroll
      for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
        for (int i = 0; i < TM; ++i)
          acc[i][jj] = __ffma2_rn(make_float2(wv[cur][i], wv[cur][i]), av[cur][jj], acc[i][jj]);
    }
  } else {
#if NFB_SLICE_ORDER == 3
    // Transposed pairing: the packed operand is a pair of ROWS (two adjacent weights, exactly as LDS.128 delivers them)
    // of one column; it stays in the operand-reuse cache while the TN column scalars stream past.  The accumulators are
    // the same scalars, only paired differently in the register file: acc[2 ip][jj].x sits beside acc[2 ip + 1][jj].x.
    static_assert(TM % 2 == 0, "rows are paired");
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float wv[TM];
      float2 av[TN / 2];
      load_fragments<TM, TN, MP, NC>(w + kk * MP, a + kk * NC, wv, av);
#pragma unroll
      for (int ip = 0; ip < TM / 2; ++ip) {
        const float2 wp = make_float2(wv[2 * ip], wv[2 * ip + 1]);
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) {
          float2 t = __ffma2_rn(wp, make_float2(av[jj].x, av[jj].x), make_float2(acc[2 * ip][jj].x, acc[2 * ip + 1][jj].x));
          acc[2 * ip][jj].x = t.x; acc[2 * ip + 1][jj].x = t.y;
          t = __ffma2_rn(wp, make_float2(av[jj].y, av[jj].y), make_float2(acc[2 * ip][jj].y, acc[2 * ip + 1][jj].y));
          acc[2 * ip][jj].y = t.x; acc[2 * ip + 1][jj].y = t.y;
        }
      }
    }
#elif NFB_SLICE_ORDER == 2  // fragments double-buffered by hand, as for the 8 x 8 tile
    float wv[2][TM];
    float2 av[2][TN / 2];
    load_fragments<TM, TN, MP, NC>(w, a, wv[0], av[0]);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const int cur = kk & 1;
      if (kk + 1 < KC) load_fragments<TM, TN, MP, NC>(w + (kk + 1) * MP, a + (kk + 1) * NC, wv[cur ^ 1], av[cur ^ 1]);
#pragma unroll
      for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
        for (int ii = 0; ii < TM; ++ii) {
          const int i = (jj & 1) ? TM - 1 - ii : ii;
          acc[i][jj] = __ffma2_rn(make_float2(wv[cur][i], wv[cur][i]), av[cur][jj], acc[i][jj]);
        }
    }
#else
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float wv[TM];
      float2 av[TN / 2];
      load_fragments<TM, TN, MP, NC>(w + kk * MP, a + kk * NC, wv, av);
#pragma unroll
      for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
        for (int ii = 0; ii < TM; ++ii) {
#if NFB_SLICE_ORDER == 1  // snake: the row index turns round at every column pair, so the weight operand carries over
          const int i = (jj & 1) ? TM - 1 - ii : ii;
#else
          const int i = ii;
#endif
          acc[i][jj] = __ffma2_rn(make_float2(wv[i], wv[i]), av[jj], acc[i][jj]);
        }
    }
#endif
  }
}

// ---- step 1: featurisation + Mahalanobis (shared by every kernel) -----------------------------------
// One query (or its flipped twin): 23 raw scalars -> the 25 latent inputs x (fp64) [main.py:171-184, 253-265],
// d2 / 50 of the Mahalanobis correction [main.py:47-61, 244-246] and Re.  fp64 throughout: cond(S^-1) ~ 7e7.
// WITH_MAHA = false (the reverse pass's finish step, which only needs x again) leaves the quadratic form out.
template <bool WITH_MAHA = true>
__device__ __forceinline__ void featurise_query(const EvalParams& p, long long q, bool valid, bool twin,
                                                double (&x)[N_IN], float& maha, float& re) {
  double r[N_RAW];
  if (valid) {
    if (p.in_f64) {
#pragma unroll
      for (int i = 0; i < N_RAW; ++i)
        r[i] = __ldg(static_cast<const double*>(p.in[i]) + q * p.in_stride[i]);
    } else {
#pragma unroll
      for (int i = 0; i < N_RAW; ++i)
        r[i] = static_cast<double>(__ldg(static_cast<const float*>(p.in[i]) + q * p.in_stride[i]));
    }
  } else {  // padding column of a partial tile: any finite, harmless query
#pragma unroll
    for (int i = 0; i < N_RAW; ++i) r[i] = 0.0;
    r[19] = 1.0e6; r[20] = 9.0; r[21] = 1.0; r[22] = 1.0;
  }

#pragma unroll
  for (int i = 0; i < 17; ++i) x[i] = r[i];
  x[17] = r[17] * 50.0;
  constexpr double kPi = 3.141592653589793238462643383279502884;
  const double alpha = r[18];
  // aerosandbox.numpy.sind/cosd: argument * pi / 180, evaluated left to right.
  x[18] = sin((2.0 * alpha) * kPi / 180.0);
  const double c = cos(alpha * kPi / 180.0);
  x[19] = c;
  x[20] = __dsub_rn(1.0, __dmul_rn(c, c));  // no fma contraction: the reference rounds c**2 first
  x[21] = (log(r[19]) - 12.5) / 3.5;
  x[22] = (r[20] - 9.0) / 4.5;
  x[23] = r[21];
  x[24] = r[22];

  if (twin) {  // main.py:253-265
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double u = x[i], l = x[8 + i];
      x[i] = -l;
      x[8 + i] = -u;
    }
    x[16] = -x[16];
    x[18] = -x[18];
    const double t = x[23];
    x[23] = x[24];
    x[24] = t;
  }

  re = static_cast<float>(r[19]);
  if constexpr (!WITH_MAHA) {
    maha = 0.0f;
    return;
  }
  // (x - mu)^T S^-1 (x - mu) with the packed symmetric form
  double d[N_IN];
#pragma unroll
  for (int i = 0; i < N_IN; ++i) d[i] = x[i] - p.mean[i];
  double d2 = 0.0;
  int idx = 0;
#pragma unroll
  for (int i = 0; i < N_IN; ++i) {
    double t = 0.0;
#pragma unroll
    for (int j = i; j < N_IN; ++j) t = fma(p.quad[idx++], d[j], t);
    d2 = fma(d[i], t, d2);
  }
  maha = static_cast<float>(d2 / (2.0 * N_IN));  // main.py:244-246
}

// fp32 kernels: one thread per tile column; the 25 inputs go, rounded to fp32, into rows 0..31 of the activation tile
template <int NC>
__device__ __forceinline__ void featurise_column(const EvalParams& p, long long q, bool valid,
                                                 bool twin, int col, float* __restrict__ act,
                                                 float* __restrict__ maha, float* __restrict__ re_s) {
  double x[N_IN];
  float m, re;
  featurise_query(p, q, valid, twin, x, m, re);
  maha[col] = m;
  if (!twin) re_s[col] = re;
#pragma unroll
  for (int k = 0; k < N_IN; ++k) act[k * NC + col] = static_cast<float>(x[k]);
#pragma unroll
  for (int k = N_IN; k < K0_PAD; ++k) act[k * NC + col] = 0.0f;
}

// ---- step 3 helpers: output rows ------------------------------------------------------------------
__device__ __forceinline__ void store_row4(const EvalParams& p, int row, long long q0, int nvalid,
                                           bool vec_ok, const float (&v)[4]) {
  if (nvalid <= 0) return;
  if (p.out_bl != nullptr && row >= 6) {  // mixed-precision outputs: boundary-layer rows as bfloat16
    unsigned short* o = static_cast<unsigned short*>(p.out_bl) + (row - 6) * p.out_ld + q0;
    if (nvalid == 4 && (p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out_bl) & 7) == 0) {
      __stcs(reinterpret_cast<uint2*>(o), make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3])));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) o[j] = static_cast<unsigned short>(pack_bf16x2(v[j], 0.0f) & 0xFFFFu);
    }
    return;
  }
  if (p.out_f64) {
    double* o = static_cast<double*>(p.out) + row * p.out_ld + q0;
    if (vec_ok && nvalid == 4) {
      __stcs(reinterpret_cast<double2*>(o), make_double2(v[0], v[1]));
      __stcs(reinterpret_cast<double2*>(o) + 1, make_double2(v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) o[j] = v[j];
    }
  } else {
    float* o = static_cast<float*>(p.out) + row * p.out_ld + q0;
    if (vec_ok && nvalid == 4) {
      __stcs(reinterpret_cast<float4*>(o), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nvalid) o[j] = v[j];
    }
  }
}

__device__ __forceinline__ float clip01_keep_nan(float z) {  // np.clip propagates NaN
  return (z != z) ? z : fminf(fmaxf(z, 0.0f), 1.0f);
}

// Un-flip + average + squash + decode for one thread's 4 rows x (4 queries + 4 twins) and store.
// acc[i][j]: packed row 4g+i, query j (j < 4) or twin of query j-4.   [main.py:274-316]
// The four branches only compute; the stores are shared code behind them (the function is inlined several times per
// kernel, and the column groups of nfb_ffma_groups_kernel run different parts of it at the same time: code size is
// instruction-cache pressure).
__device__ __forceinline__ void decode_and_store(const EvalParams& p, int g, long long q0, int nvalid,
                                                 bool vec_ok, const float (&acc)[4][8],
                                                 const float* __restrict__ maha_o,
                                                 const float* __restrict__ maha_t,
                                                 const float* __restrict__ re4) {
  float o0[4], o1[4], o2[4], o3[4];
  int r0, r1, r2 = -1, r3 = -1;
  if (p.scalars_only && g < 48) return;  // boundary-layer rows not wanted
  if (g < N_BL) {  // theta_u, ue_u, theta_l, ue_l of station g
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float zt_u = 0.5f * (acc[0][j] + acc[2][4 + j]);
      const float zu_u = 0.5f * (acc[1][j] - acc[3][4 + j]);
      const float zt_l = 0.5f * (acc[2][j] + acc[0][4 + j]);
      const float zu_l = 0.5f * (acc[3][j] - acc[1][4 + j]);
      const float re = re4[j];
      o0[j] = (exp10f(zt_u) - 0.1f) / (fabsf(zu_u) * re);
      o1[j] = zu_u;
      o2[j] = (exp10f(zt_l) - 0.1f) / (fabsf(zu_l) * re);
      o3[j] = zu_l;
    }
    r0 = 6 + g; r1 = 6 + 2 * N_BL + g; r2 = 6 + 3 * N_BL + g; r3 = 6 + 5 * N_BL + g;
  } else if (g < 48) {  // H_u, H_l of stations s, s + 1
    const int s = 2 * (g - N_BL);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o0[j] = 2.6f * expf(0.5f * (acc[0][j] + acc[2][4 + j]));
      o1[j] = 2.6f * expf(0.5f * (acc[1][j] + acc[3][4 + j]));
      o2[j] = 2.6f * expf(0.5f * (acc[2][j] + acc[0][4 + j]));
      o3[j] = 2.6f * expf(0.5f * (acc[3][j] + acc[1][4 + j]));
    }
    r0 = 6 + N_BL + s; r1 = r0 + 1; r2 = 6 + 4 * N_BL + s; r3 = r2 + 1;
  } else if (g == 48) {  // analysis_confidence, CL, CD, CM
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z0 = 0.5f * ((acc[0][j] - maha_o[j]) + (acc[0][4 + j] - maha_t[j]));
      const float z1 = 0.5f * (acc[1][j] - acc[1][4 + j]);
      const float z2 = 0.5f * (acc[2][j] + acc[2][4 + j]);
      const float z3 = 0.5f * (acc[3][j] - acc[3][4 + j]);
      o0[j] = 1.0f / (1.0f + expf(-z0));  // fp32 saturates where the reference clips at +-707
      o1[j] = 0.5f * z1;
      o2[j] = expf((z2 - 2.0f) * 2.0f);
      o3[j] = z3 / 20.0f;
    }
    r0 = 0; r1 = 1; r2 = 2; r3 = 3;
  } else if (g == 49) {  // Top_Xtr, Bot_Xtr
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o0[j] = clip01_keep_nan(0.5f * (acc[0][j] + acc[1][4 + j]));
      o1[j] = clip01_keep_nan(0.5f * (acc[1][j] + acc[0][4 + j]));
      o2[j] = 0.0f; o3[j] = 0.0f;
    }
    r0 = 4; r1 = 5;
  } else {
    return;
  }
  store_row4(p, r0, q0, nvalid, vec_ok, o0);
  store_row4(p, r1, q0, nvalid, vec_ok, o1);
  if (r2 >= 0) {
    store_row4(p, r2, q0, nvalid, vec_ok, o2);
    store_row4(p, r3, q0, nvalid, vec_ok, o3);
  }
}

// ---- the kernel ----------------------------------------------------------------------------------
// Weight streaming: the slices of one tile form a fixed sequence (layer 0's K-slices, each hidden
// layer's, then the last layer's once per pass) described by a table in shared memory.  Slice j of
// the CTA's whole run lands in ring slot j % STAGES.  There is no producer warp (an extra warp would
// cap the compute warps' registers): when a warp starts slice `it`, its lane 0 -- for the one warp
// with (it + STAGES - 2) % WARPS == warp -- issues the TMA bulk copy of slice it + STAGES - 2, whose
// slot was last read by slice it - 2, which every warp has normally long finished.
// (Checked again for the 8-warp configuration: a ninth, producer warp makes the register allocator budget for 12 warps
//  -- 168 registers per thread instead of 255 -- and the 8 x 16 thread tile spills.)
template <class Cfg>
__device__ __forceinline__ void issue_slice(uint32_t j, uint32_t slices_per_tile, const uint2* __restrict__ table,
                                            const float* __restrict__ blob, float* ring, uint32_t bar_full,
                                            uint32_t bar_empty) {
  constexpr int STAGES = Cfg::STAGES;
  const uint2 e = table[j % slices_per_tile];
  const uint32_t s = j & (STAGES - 1), ph = (j / STAGES) & 1;
  mbar_wait(bar_empty + 8 * s, ph ^ 1);  // previous occupant (slice j - STAGES) released by every warp
  mbar_arrive_expect_tx(bar_full + 8 * s, e.y);
  bulk_copy_g2s(smem_u32(ring + s * Cfg::SLOT_FLOATS), blob + e.x, e.y, bar_full + 8 * s);
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) nfb_ffma_kernel(const EvalParams p) {
  constexpr int H = Cfg::H, NC = Cfg::NC, NQ = Cfg::NQ, KC = Cfg::KC, STAGES = Cfg::STAGES, TN = Cfg::TN;
  constexpr int WM = Cfg::WM, WN = Cfg::WN, SLOT = Cfg::SLOT_FLOATS, LM = Cfg::LM, LN = Cfg::LN;
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
  const uint32_t slices_per_tile = K0_PAD / KC + (L - 2) * (H / KC) + WN * (H / KC);
  const uint32_t my_tiles = uint32_t((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t total_slices = my_tiles * slices_per_tile;

  for (uint32_t j = tid; j < slices_per_tile; j += THREADS) {
    uint32_t l, c;  // layer and K-slice index of slice j
    if (j < K0_PAD / KC) {
      l = 0; c = j;
    } else {
      const uint32_t r = j - K0_PAD / KC;
      l = 1 + r / (H / KC); c = r % (H / KC);
      if (l > uint32_t(L - 1)) l = L - 1;  // last layer's repeated passes
    }
    const uint32_t M = (l == uint32_t(L - 1)) ? M_LAST : H;
    table[j] = make_uint2(uint32_t(blob_layer_offset<H>(int(l))) + c * KC * M, KC * M * uint32_t(sizeof(float)));
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0)  // prologue: slices 0 .. STAGES-3 (the rest are issued from the slice loops below)
    for (uint32_t j = 0; j < STAGES - 2 && j < total_slices; ++j)
      issue_slice<Cfg>(j, slices_per_tile, table, p.blob, ring, bar_full, bar_empty);

  const int lm = lane / LN, ln = lane % LN;
  // hidden layers: warp grid WM x WN, thread tile 8 rows (two groups of 4) x TN columns (QT + QT twins)
  const int wm = warp % WM, wn = warp / WM;
  const int h_row = (wm * LM + lm) * 4;   // rows h_row..+3 and H/2 + h_row..+3
  const int h_col = wn * 32 + ln * QT;    // columns h_col..+QT-1 and NC/2 + h_col..
  // last layer: every warp along rows, thread tile 4 packed rows x TN columns, WN passes of 32 queries
  const int g_last = warp * LM + lm;      // packed row group; rows 4g..4g+3
  const bool last_active = (warp * LM * 4 < M_LAST_USED);

  const bool vec_ok =
      p.out_f64 ? ((p.out_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                : ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);

  uint32_t it = 0;
  auto wait_slice = [&](uint32_t& s_out) {  // make slice `it` visible; keep the ring fed
    const uint32_t j = it + STAGES - 2;
    if (lane == 0 && (j & (WARPS - 1)) == uint32_t(warp) && j < total_slices)
      issue_slice<Cfg>(j, slices_per_tile, table, p.blob, ring, bar_full, bar_empty);
    s_out = it & (STAGES - 1);
    mbar_wait(bar_full + 8 * s_out, (it / STAGES) & 1);
  };
  auto release_slice = [&](uint32_t s) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
    ++it;
  };

  // Hidden layers touch only the 64 columns of this warp's column group (WM warps): with NFB_GROUP_BARS the groups
  // synchronise among themselves and drift apart as far as the weight ring allows.
  auto layer_bar = [&]() {
#if NFB_GROUP_BARS
    named_bar_sync(2 + wn, WM * 32);
#else
    named_bar_sync(1, THREADS);
#endif
  };

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long tile_q0 = tile * NQ;

    named_bar_sync(1, THREADS);  // previous tile's readers of `act`, `maha`, `re_s` are done
    for (int col = tid; col < NC; col += THREADS) {
      const bool twin = col >= NQ;
      const long long q = tile_q0 + (twin ? col - NQ : col);
      featurise_column<NC>(p, q, q < p.n, twin, col, act, maha, re_s);
    }
    named_bar_sync(1, THREADS);

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
      // (Measured and rejected: waiting for two K-slices at a time -- half the barrier waits and warp syncs -- is 17 %
      //  slower: a slot is then released a slice later, and the ring is latency-critical.)
      for (int c = 0; c < K / KC; ++c) {
        uint32_t s;
        wait_slice(s);
        ffma_slice<8, TN, H, NC, KC>(ring + s * SLOT + h_row, act + (c * KC) * NC + h_col, acc);
        release_slice(s);
      }
      layer_bar();  // everyone has finished reading this layer's input
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
      layer_bar();
    }
#if NFB_GROUP_BARS
    named_bar_sync(1, THREADS);  // the last layer reads every column group's rows
#endif

    // ---------------- last layer + fused epilogue, WN passes of 32 queries ----------------
    {
      const float* bias = p.blob + blob_layer_offset<H>(L - 1) + size_t(H) * M_LAST;
      for (int pass = 0; pass < WN; ++pass) {
        const int col = pass * 32 + ln * QT;
        float2 acc[4][TN / 2];
        if (last_active) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float b = __ldg(bias + g_last * 4 + i);
#pragma unroll
            for (int jj = 0; jj < TN / 2; ++jj) acc[i][jj] = make_float2(b, b);
          }
        }
        for (int c = 0; c < H / KC; ++c) {
          uint32_t s;
          wait_slice(s);
          if (last_active)
            ffma_slice<4, TN, M_LAST, NC, KC>(ring + s * SLOT + g_last * 4, act + (c * KC) * NC + col, acc);
          release_slice(s);
        }
        if (last_active) {
#pragma unroll
          for (int v = 0; v < TN / 8; ++v) {  // groups of 4 queries (+ their twins)
            float yf[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              yf[i][0] = acc[i][2 * v].x; yf[i][1] = acc[i][2 * v].y;
              yf[i][2] = acc[i][2 * v + 1].x; yf[i][3] = acc[i][2 * v + 1].y;
              yf[i][4] = acc[i][TN / 4 + 2 * v].x; yf[i][5] = acc[i][TN / 4 + 2 * v].y;
              yf[i][6] = acc[i][TN / 4 + 2 * v + 1].x; yf[i][7] = acc[i][TN / 4 + 2 * v + 1].y;
            }
            const int c4 = col + 4 * v;
            const long long q0 = tile_q0 + c4;
            const long long left = p.n - q0;
            const int nvalid = left >= 4 ? 4 : (left > 0 ? int(left) : 0);
            decode_and_store(p, g_last, q0, nvalid, vec_ok, yf, maha + c4, maha + NQ + c4, re_s + c4);
          }
        }
      }
    }
  }
}

// ---- fp32 peak microbenchmark -----------------------------------------------------------------------
// 16 independent x <- x * a + b chains per thread, 32 warps per SM: the FFMA issue-rate ceiling of the
// chip (measured 126.8 of 128 lanes/clk/SM on B200; scripts/ffma_peak.cu has the other patterns).
__global__ void __launch_bounds__(1024, 1) nfb_ffma_peak_kernel(float* sink, int iters) {
  float x[16];
  const float a = 1.0f + 1e-6f * threadIdx.x, b = 1e-3f;
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = float(i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 32; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) sink[0] = s;  // never true in practice; defeats dead-code elimination
}

}  // namespace nfb
