// Jacobian kernel: the 198 outputs of the path AND their derivatives with respect to the 23 raw inputs,
// for design optimisers that today differentiate the reference by tracing main.py through CasADi
// (studies/sequential_multifidelity_optimization.py:21-56).
//
// Forward mode, fp32, one CTA per query.  Every quantity of the path carries 23 tangents along: the CTA's activation
// tile has 48 columns -- [query | its 23 tangents | twin | the twin's 23 tangents] -- and an affine layer treats them all
// alike (a tangent goes through W like a value; the bias only enters the two primal columns).  One thread per output
// row, weights read straight from the L2-resident packed fp32 blob of the fp32 kernels (K-major: coalesced), the
// activation tile updated in place between two barriers.  What differs from the value path:
//   featurise   d x / d raw is sparse (identity on the Kulfan rows, trig / log terms for alpha and Re); the flip of the
//               twin is a signed permutation, so it applies to tangents unchanged; the Mahalanobis term contributes
//               -(2 / 50) (x - mu)^T S^-1 dx to the confidence logit's tangents (fp64, like the value).
//   SiLU        value: z s(z); tangent: s (1 + z (1 - s)) dz with s = sigmoid(z) of the SAME thread's primal column.
//   decode      chain rule through the un-flip / average (linear), sigmoid, exp, 10^z / (|ue| Re) and the Xtr clip.
// The primal columns see exactly the arithmetic of the fp32 kernels (acc = bias, one fma per k in k order, the same SiLU
// and decode), so the values this kernel returns are bit-identical to theirs.
// Output: values 198 x ld (optional) and jac[q][row][t], 198 x 23 contiguous floats (or doubles) per query.
#pragma once
#include "nfb_ffma.cuh"

namespace nfb {

constexpr int JAC_T = N_RAW;               // tangents per twin
constexpr int JAC_HALF = 1 + JAC_T;        // columns per twin: primal, tangents
constexpr int JAC_COLS = 2 * JAC_HALF;     // 48

struct JacParams {
  EvalParams ev;    // inputs, n, blob, n_layers; ev.out / ev.out_ld = values (may be NULL)
  void* jac;        // n x 198 x 23, ev.out_f64 decides the type (NULL in VJP mode)
  // VJP mode (cot != NULL): instead of the Jacobian itself, grad[t][q] = sum_row cot[row][q] * d out_row / d raw_t
  const void* cot;  // 198 x cot_ld cotangents, same type as the outputs
  void* grad;       // 23 x grad_ld
  long long cot_ld, grad_ld;
};

template <int H>
struct JacCfg {
  // One output row per thread; TWO for the 512-wide model (256 threads, 96 accumulators each, 255 registers): an
  // activation row read from shared memory then feeds 96 FFMAs instead of 48 -- with one row per thread the broadcast
  // LDS.128s (12 per 48 FFMA) tie with the FFMA pipe, and 512 threads cap the registers at 128.
  static constexpr int RPT = H > 256 ? 2 : 1;
  static constexpr int THREADS = 256;                // >= 224 rows of the last layer
  static constexpr int KMAX = H > K0_PAD ? H : K0_PAD;
  static constexpr size_t SMEM_BYTES =
      sizeof(float) * (size_t(KMAX) * JAC_COLS + size_t(M_LAST) * JAC_COLS + JAC_COLS + 8) + sizeof(double) * 2 * N_IN;
};

// d2 = sum_i S_ii d_i^2 + sum_{i<j} Q_ij d_i d_j (EvalParams::quad: packed upper triangle, off-diagonals pre-summed):
// g_i = d d2 / d d_i = 2 S_ii d_i + sum_{j != i} Q_ij d_j.  Everything in registers: written as loops that index g[] at run time
// it goes through local memory -- a 650-step dependent chain per query that nothing hides (it cost the reverse-mode kernel
// of nfb_grad.cuh a factor 1.6 .. 2.9).
__device__ __forceinline__ void maha_gradient_regs(const EvalParams& p, const double (&x)[N_IN], double (&g)[N_IN]) {
  double d[N_IN];
#pragma unroll
  for (int i = 0; i < N_IN; ++i) {
    d[i] = x[i] - p.mean[i];
    g[i] = 0.0;
  }
  int idx = 0;
#pragma unroll
  for (int i = 0; i < N_IN; ++i) {
#pragma unroll
    for (int j = i; j < N_IN; ++j) {
      const double q = p.quad[idx++];
      if (i == j) {
        g[i] = fma(2.0 * q, d[i], g[i]);
      } else {
        g[i] = fma(q, d[j], g[i]);
        g[j] = fma(q, d[i], g[j]);
      }
    }
  }
}

__device__ __forceinline__ float silu_slope(float z) {  // d/dz [z sigmoid(z)]
  const float s = __fdividef(1.0f, 1.0f + expf(fminf(-z, 80.0f)));
  return s * fmaf(z, 1.0f - s, 1.0f);
}

// (two CTAs per SM for the narrow models, whose single CTA of 8 warps is latency bound: ncu shows 22 % issue activity)
template <int H>
__global__ void __launch_bounds__(JacCfg<H>::THREADS, (H <= 256) ? 2 : 1) nfb_jac_kernel(const JacParams jp) {
  using Cfg = JacCfg<H>;
  const EvalParams& p = jp.ev;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* const act = reinterpret_cast<float*>(smem_raw);         // [KMAX][48]
  float* const stage = act + Cfg::KMAX * JAC_COLS;               // [224][48] last layer, packed row order
  float* const dmaha = stage + M_LAST * JAC_COLS;                // [48] d2 / 50 and its tangents, per column
  float* const re_s = dmaha + JAC_COLS;                          // [1] (+ padding)
  double* const grad = reinterpret_cast<double*>(re_s + 8);      // [2][25] Mahalanobis gradients of the twins

  const int tid = threadIdx.x, L = p.n_layers;
  const long long q = blockIdx.x;

  // ---- featurise: values (fp64 -> fp32) and the sparse d x / d raw ----
  for (int i = tid; i < K0_PAD * JAC_COLS; i += Cfg::THREADS) act[i] = 0.0f;
  __syncthreads();
  if (tid < 2) {
    const bool twin = tid == 1;
    double x[N_IN];
    float m, re;
    featurise_query(p, q, true, twin, x, m, re);
    float* const col = act + tid * JAC_HALF;  // column of my primal; tangent t at col + 1 + t
#pragma unroll
    for (int k = 0; k < N_IN; ++k) col[k * JAC_COLS] = static_cast<float>(x[k]);
    dmaha[tid * JAC_HALF] = m;
    if (!twin) re_s[0] = re;
    {
      double g[N_IN];
      maha_gradient_regs(p, x, g);
#pragma unroll
      for (int k = 0; k < N_IN; ++k) grad[tid * N_IN + k] = g[k];
    }
    // tangents of the query's features; for the twin row k of the flipped vector is sign * (row perm(k) of the query's)
    const double kdeg = 3.141592653589793238462643383279502884 / 180.0;
    const double alpha = p.in_f64 ? __ldg(static_cast<const double*>(p.in[18]) + q * p.in_stride[18])
                                  : double(__ldg(static_cast<const float*>(p.in[18]) + q * p.in_stride[18]));
    const double rey = p.in_f64 ? __ldg(static_cast<const double*>(p.in[19]) + q * p.in_stride[19])
                                : double(__ldg(static_cast<const float*>(p.in[19]) + q * p.in_stride[19]));
    const double a = alpha * kdeg;
    auto put = [&](int k, int t, double v) {  // d x_k / d raw_t of the QUERY -> my twin's row and sign [main.py:253-265]
      int kk = k;
      double s = 1.0;
      if (twin) {
        if (k < 8) { kk = k + 8; s = -1.0; }
        else if (k < 16) { kk = k - 8; s = -1.0; }
        else if (k == 16 || k == 18) { s = -1.0; }
        else if (k == 23) { kk = 24; }
        else if (k == 24) { kk = 23; }
      }
      col[kk * JAC_COLS + 1 + t] = static_cast<float>(s * v);
    };
    for (int i = 0; i < 17; ++i) put(i, i, 1.0);
    put(17, 17, 50.0);
    put(18, 18, 2.0 * kdeg * cos(2.0 * a));
    put(19, 18, -kdeg * sin(a));
    put(20, 18, 2.0 * kdeg * cos(a) * sin(a));
    put(21, 19, 1.0 / (3.5 * rey));
    put(22, 20, 1.0 / 4.5);
    put(23, 21, 1.0);
    put(24, 22, 1.0);
  }
  __syncthreads();
  if (tid < JAC_COLS && tid % JAC_HALF != 0) {  // tangents of d2 / 50 [main.py:244-246]: g . dx / 50
    const double* g = grad + (tid / JAC_HALF) * N_IN;
    double s = 0.0;
    for (int k = 0; k < N_IN; ++k) s = fma(g[k], double(act[k * JAC_COLS + tid]), s);
    dmaha[tid] = static_cast<float>(s / (2.0 * N_IN));
  }
  __syncthreads();

  // ---- layers: RPT output rows per thread (rows tid, tid + 256), 48 columns each in registers ----
  constexpr int RPT = Cfg::RPT;
  auto affine = [&](const float* __restrict__ w, int ldw, int K, int rows_here, const float* __restrict__ bias, float (&acc)[RPT][JAC_COLS]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float b = (r < rows_here) ? __ldg(bias + tid + 256 * r) : 0.0f;
#pragma unroll
      for (int c = 0; c < JAC_COLS; ++c) acc[r][c] = (c % JAC_HALF == 0) ? b : 0.0f;
    }
    constexpr int U = (RPT == 1) ? 8 : 4;  // weight loads in flight per row
    for (int k0 = 0; k0 < K; k0 += U) {
      float wv[RPT][U];
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int u = 0; u < U; ++u) wv[r][u] = (r < rows_here) ? __ldg(w + size_t(k0 + u) * ldw + tid + 256 * r) : 0.0f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4* a4 = reinterpret_cast<const float4*>(act + (k0 + u) * JAC_COLS);
#pragma unroll
        for (int c4 = 0; c4 < JAC_COLS / 4; ++c4) {
          const float4 a = a4[c4];
#pragma unroll
          for (int r = 0; r < RPT; ++r) {  // packed FFMA2 (fma.rn.f32x2): one IEEE fma per lane, half the issue slots
            const float2 w2 = make_float2(wv[r][u], wv[r][u]);
            const float2 lo = __ffma2_rn(w2, make_float2(a.x, a.y), make_float2(acc[r][4 * c4 + 0], acc[r][4 * c4 + 1]));
            const float2 hi = __ffma2_rn(w2, make_float2(a.z, a.w), make_float2(acc[r][4 * c4 + 2], acc[r][4 * c4 + 3]));
            acc[r][4 * c4 + 0] = lo.x; acc[r][4 * c4 + 1] = lo.y;
            acc[r][4 * c4 + 2] = hi.x; acc[r][4 * c4 + 3] = hi.y;
          }
        }
      }
    }
  };
  auto put_rows = [&](float* dst, int rows_here, const float (&acc)[RPT][JAC_COLS]) {
#pragma unroll
    for (int r = 0; r < RPT; ++r)
      if (r < rows_here) {
#pragma unroll
        for (int c4 = 0; c4 < JAC_COLS / 4; ++c4)
          reinterpret_cast<float4*>(dst + (tid + 256 * r) * JAC_COLS)[c4] =
              make_float4(acc[r][4 * c4], acc[r][4 * c4 + 1], acc[r][4 * c4 + 2], acc[r][4 * c4 + 3]);
      }
  };
  for (int l = 0; l < L - 1; ++l) {
    const int K = (l == 0) ? K0_PAD : H;
    const float* w = p.blob + blob_layer_offset<H>(l);
    const int rows_here = (tid + 256 * (RPT - 1) < H) ? RPT : (tid < H ? 1 : 0);
    float acc[RPT][JAC_COLS];
    if (rows_here) {
      affine(w, H, K, rows_here, w + size_t(K) * H, acc);
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float z = acc[r][half * JAC_HALF], slope = silu_slope(z);
          acc[r][half * JAC_HALF] = silu(z);
#pragma unroll
          for (int t = 1; t < JAC_HALF; ++t) acc[r][half * JAC_HALF + t] *= slope;
        }
    }
    __syncthreads();  // everybody has read the layer's input
    if (rows_here) put_rows(act, rows_here, acc);
    __syncthreads();
  }
  {
    const float* w = p.blob + blob_layer_offset<H>(L - 1);
    if (tid < M_LAST) {  // 224 rows: one per thread
      float acc[RPT][JAC_COLS];
      affine(w, M_LAST, H, 1, w + size_t(H) * M_LAST, acc);
      put_rows(stage, 1, acc);
    }
  }
  __syncthreads();

  // ---- decode: thread (group g, column c): c = 0 the values, c = 1 + t the derivative with respect to input t ----
  // y[i] / f[i]: packed rows 4 g + i of the query / the twin (value or tangent c); Y / F: always the values
  float* const part = act;  // [50][23] partial vector-Jacobian products (the activation tile is no longer needed)
  auto cotangent = [&](int row) -> float {
    return p.out_f64 ? float(__ldg(static_cast<const double*>(jp.cot) + row * jp.cot_ld + q))
                     : __ldg(static_cast<const float*>(jp.cot) + row * jp.cot_ld + q);
  };
  for (int item = tid; item < (M_LAST_USED / 4) * JAC_HALF; item += Cfg::THREADS) {
    const int g = item / JAC_HALF, c = item % JAC_HALF;
    float y[4], f[4], Y[4], F[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float* row = stage + (4 * g + i) * JAC_COLS;
      y[i] = row[c]; f[i] = row[JAC_HALF + c];
      Y[i] = row[0]; F[i] = row[JAC_HALF];
    }
    const float re = re_s[0];
    const bool val = (c == 0);
    const float d_re = (c == 1 + 19) ? 1.0f : 0.0f;  // d Re / d raw_t
    float o[4];
    int rows[4], n_out = 4;
    if (g < N_BL) {  // theta_u, ue_u, theta_l, ue_l of station g [main.py:298-316]
      const float zt_u = 0.5f * (Y[0] + F[2]), zu_u = 0.5f * (Y[1] - F[3]);
      const float zt_l = 0.5f * (Y[2] + F[0]), zu_l = 0.5f * (Y[3] - F[1]);
      const float p_u = exp10f(zt_u), p_l = exp10f(zt_l);
      const float th_u = (p_u - 0.1f) / (fabsf(zu_u) * re), th_l = (p_l - 0.1f) / (fabsf(zu_l) * re);
      if (val) {
        o[0] = th_u; o[1] = zu_u; o[2] = th_l; o[3] = zu_l;
      } else {
        const float dzt_u = 0.5f * (y[0] + f[2]), dzu_u = 0.5f * (y[1] - f[3]);
        const float dzt_l = 0.5f * (y[2] + f[0]), dzu_l = 0.5f * (y[3] - f[1]);
        constexpr float LN10 = 2.302585092994046f;
        o[0] = LN10 * p_u * dzt_u / (fabsf(zu_u) * re) - th_u * (dzu_u / zu_u) - th_u * (d_re / re);
        o[1] = dzu_u;
        o[2] = LN10 * p_l * dzt_l / (fabsf(zu_l) * re) - th_l * (dzu_l / zu_l) - th_l * (d_re / re);
        o[3] = dzu_l;
      }
      rows[0] = 6 + g; rows[1] = 6 + 2 * N_BL + g; rows[2] = 6 + 3 * N_BL + g; rows[3] = 6 + 5 * N_BL + g;
    } else if (g < 48) {  // H_u, H_l of stations s, s + 1
      const int s = 2 * (g - N_BL);
      const float h0 = 2.6f * expf(0.5f * (Y[0] + F[2])), h1 = 2.6f * expf(0.5f * (Y[1] + F[3]));
      const float h2 = 2.6f * expf(0.5f * (Y[2] + F[0])), h3 = 2.6f * expf(0.5f * (Y[3] + F[1]));
      if (val) {
        o[0] = h0; o[1] = h1; o[2] = h2; o[3] = h3;
      } else {
        o[0] = h0 * 0.5f * (y[0] + f[2]); o[1] = h1 * 0.5f * (y[1] + f[3]);
        o[2] = h2 * 0.5f * (y[2] + f[0]); o[3] = h3 * 0.5f * (y[3] + f[1]);
      }
      rows[0] = 6 + N_BL + s; rows[1] = 6 + N_BL + s + 1; rows[2] = 6 + 4 * N_BL + s; rows[3] = 6 + 4 * N_BL + s + 1;
    } else if (g == 48) {  // analysis_confidence, CL, CD, CM
      const float z0 = 0.5f * ((Y[0] - dmaha[0]) + (F[0] - dmaha[JAC_HALF]));
      const float z2 = 0.5f * (Y[2] + F[2]);
      const float conf = 1.0f / (1.0f + expf(-z0)), cd = expf((z2 - 2.0f) * 2.0f);
      if (val) {
        o[0] = conf; o[1] = 0.5f * (0.5f * (Y[1] - F[1])); o[2] = cd; o[3] = (0.5f * (Y[3] - F[3])) / 20.0f;
      } else {
        const float dz0 = 0.5f * ((y[0] - dmaha[c]) + (f[0] - dmaha[JAC_HALF + c]));
        o[0] = conf * (1.0f - conf) * dz0;
        o[1] = 0.5f * (0.5f * (y[1] - f[1]));
        o[2] = 2.0f * cd * (0.5f * (y[2] + f[2]));
        o[3] = (0.5f * (y[3] - f[3])) / 20.0f;
      }
      rows[0] = 0; rows[1] = 1; rows[2] = 2; rows[3] = 3;
    } else {  // g == 49: Top_Xtr, Bot_Xtr
      const float z4 = 0.5f * (Y[0] + F[1]), z5 = 0.5f * (Y[1] + F[0]);
      if (val) {
        o[0] = clip01_keep_nan(z4); o[1] = clip01_keep_nan(z5);
      } else {
        o[0] = (z4 > 0.0f && z4 < 1.0f) ? 0.5f * (y[0] + f[1]) : 0.0f;
        o[1] = (z5 > 0.0f && z5 < 1.0f) ? 0.5f * (y[1] + f[0]) : 0.0f;
      }
      o[2] = o[3] = 0.0f;
      rows[0] = 4; rows[1] = 5; rows[2] = rows[3] = 0;
      n_out = 2;
    }
    if (!val && jp.cot) {  // VJP mode: my four rows' share of grad[c - 1]
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < n_out) {  // a zero cotangent contributes exactly zero even where the derivative is infinite (theta at ue -> 0):
          const float ct = cotangent(rows[k]);  // the convention of the reverse-mode kernels (nfb_grad.cuh: bl_group_vjp)
          if (ct != 0.0f) s = fmaf(ct, o[k], s);
        }
      part[g * JAC_T + (c - 1)] = s;
      continue;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= n_out) continue;
      if (val) {
        if (p.out) {
          if (p.out_f64) static_cast<double*>(p.out)[rows[k] * p.out_ld + q] = double(o[k]);
          else static_cast<float*>(p.out)[rows[k] * p.out_ld + q] = o[k];
        }
      } else {
        const size_t at = (size_t(q) * N_OUT + rows[k]) * JAC_T + (c - 1);
        if (p.out_f64) static_cast<double*>(jp.jac)[at] = double(o[k]);
        else static_cast<float*>(jp.jac)[at] = o[k];
      }
    }
  }
  if (jp.cot) {  // fixed-order reduction over the 50 row groups: deterministic
    __syncthreads();
    if (tid < JAC_T) {
      float s = 0.0f;
      for (int g = 0; g < M_LAST_USED / 4; ++g) s += part[g * JAC_T + tid];
      if (p.out_f64) static_cast<double*>(jp.grad)[tid * jp.grad_ld + q] = double(s);
      else static_cast<float*>(jp.grad)[tid * jp.grad_ld + q] = s;
    }
  }
}

}  // namespace nfb
