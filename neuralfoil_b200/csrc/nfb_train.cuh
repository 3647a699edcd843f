// Training step of the learned core: forward, loss and backward to the weight and bias gradients of every
// layer, for a batch of (scaled input, target) pairs -- what one iteration of the reference's training loop computes
// between `net(x)` and `loss.backward()` (reference training/train.py:56-114 forward, :194-236 loss, :251-255 step).
//
//   forward   y = net(x), y[:,0] -= d2(x) / 50; the same for the flipped input; un-flip; y_fused = (y + y_unflipped) / 2;
//             Top/Bot_Xtr clipped to [0, 1]  (no sigmoid, no decode: the loss is taken in latent space)   [train.py:56-114]
//   loss      column 0: binary cross-entropy with logits; columns 1..197: Huber, delta = 0.05; mean over the batch, times
//             loss_weights; a NaN target is replaced by the prediction                                      [train.py:194-236]
//   backward  d loss / d W_l, d loss / d b_l for every layer (the inputs need no gradient)
//
// Two kernels, fp32 FFMA2 throughout (the reference trains in fp32: torch's TF32 matmul switch is off by default).
//
// 1. nfb_train_fwdbwd_kernel -- the column-group schedule, weight ring and thread tile of nfb_grad.cuh (FULL): a tile of NQ
//    samples and their flipped twins stays in shared memory through the forward pass, the loss and the backward pass down to
//    d z_0.  What the weight gradients need leaves the chip once, sample-major ("[column][feature]"), as it is produced:
//      X    [cols][32]   the inputs of both twins,            A_l  [cols][H]   the activations a_l = SiLU(z_l),
//      DZ_l [cols][H]    d loss / d z_l of the hidden layers, DZ_L [cols][256] d loss / d latent by packed row.
//    SiLU'(z) and the parked latent cotangents stay in the per-CTA scratch block (L2-resident), as in nfb_grad.cuh.
// 2. nfb_train_wgrad_kernel -- d W_l[m][k] = sum_cols DZ_l[col][m] * A_{l-1}[col][k] and d b_l[m] = sum_cols DZ_l[col][m] for
//    every layer in one launch: a register-blocked outer-product GEMM (both operands have the OUTPUT index contiguous, so
//    fragments are plain 16-byte shared-memory loads), cp.async 3-stage pipeline, split over the columns to fill the SMs;
//    the partial sums of the splits are added in a fixed order by nfb_train_reduce_kernel, which also undoes the row
//    permutation of the last layer and drops the padding: the gradients are bit-reproducible from run to run.
// Arithmetic: 3 x the forward pass (forward, d a = W^T d z, d W = d z a^T), both twins.
#pragma once
#include "nfb_grad.cuh"

namespace nfb {

constexpr int TRAIN_LAST_ROWS = 256;  // packed rows of DZ_L per column (200 used; the reverse sweep contracts over 208)

struct TrainParams {
  EvalParams ev;        // blob = the gradient blob (GradBlob<H>); n = batch size; n_layers, k_hidden, mean, quad
  const float* x;       // (batch, 25) scaled inputs, row stride x_ld                         [train.py:251 `x`]
  const float* y;       // (batch, 198) targets, row stride y_ld; NaN = no data                [train.py:252 `y_data`]
  long long x_ld, y_ld;
  const float* loss_w;  // [198] loss weights                                                  [train.py:180-190]
  float inv_batch;      // 1 / batch: the means over the batch                                 [train.py:199-216]
  float delta;          // Huber delta (0.05)                                                  [train.py:213]
  double* loss_out;     // [198] weighted loss components, accumulated with atomics (zeroed before the launch)
  float* scratch;       // gridDim.x blocks: slopes [L-1][H][NC], parked latent cotangents [256][NC]
  float* saved_a;       // X [cols][32], then A_l [cols][H] for l = 1 .. L-1
  float* saved_dz;      // DZ_l [cols][H] for l = 0 .. L-2, then DZ_L [cols][256]
  long long cols;       // n_tiles * NC
};

template <int H>
__host__ __device__ constexpr size_t train_saved_a_offset(int l, long long cols) {  // l = 0: X; l >= 1: A_l
  return l == 0 ? 0 : size_t(K0_PAD) * cols + size_t(l - 1) * H * cols;
}
template <int H>
__host__ __device__ constexpr size_t train_saved_a_floats(int L, long long cols) { return train_saved_a_offset<H>(L, cols); }
template <int H>
__host__ __device__ constexpr size_t train_saved_dz_floats(int L, long long cols) {
  return size_t(L - 1) * H * cols + size_t(TRAIN_LAST_ROWS) * cols;
}

// One element's loss and its derivative with respect to the prediction [train.py:194-216].
__device__ __forceinline__ void huber_term(float pred, float target, float delta, float& loss, float& dloss) {
  if (target != target) {  // NaN target -> replaced by the prediction: zero loss, zero gradient [train.py:197]
    loss = 0.0f; dloss = 0.0f;
    return;
  }
  const float d = pred - target, ad = fabsf(d);
  loss = ad < delta ? 0.5f * d * d : delta * (ad - 0.5f * delta);
  dloss = fminf(fmaxf(d, -delta), delta);
}
__device__ __forceinline__ void bce_logit_term(float pred, float target, float& loss, float& dloss) {
  const float sp = fmaxf(pred, 0.0f) + log1pf(expf(-fabsf(pred)));  // softplus(pred)
  const float sg = 1.0f / (1.0f + expf(-pred));
  if (target != target) {
    // the reference substitutes the (attached) prediction for a NaN target: BCE(p, p) = softplus(p) - p^2, whose derivative
    // has a second term through the target: sigmoid(p) - p - p   [train.py:197-205; checked against torch autograd]
    loss = sp - pred * pred;
    dloss = sg - 2.0f * pred;
    return;
  }
  loss = sp - pred * target;
  dloss = sg - target;
}

template <class Cfg>
struct TrainCfg {
  static constexpr int MAX_SLICES = K0_PAD / Cfg::KC + (2 * (NFB_MAX_LAYERS_K - 1) + Cfg::NFULL + Cfg::NHALF) * (Cfg::H / Cfg::KC) + 1 +
                                    GRAD_LAST_ROWS / Cfg::KC;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (size_t(Cfg::H) * Cfg::NC + size_t(Cfg::STAGES) * Cfg::SLOT_FLOATS +
                                                        Cfg::NC + Cfg::NQ) + 16 * Cfg::STAGES + 8 * size_t(MAX_SLICES);
  static_assert(SMEM_BYTES <= 227 * 1024, "does not fit shared memory");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) nfb_train_fwdbwd_kernel(const TrainParams tp) {
  constexpr int H = Cfg::H, NC = Cfg::NC, NQ = Cfg::NQ, KC = Cfg::KC, STAGES = Cfg::STAGES, TN = Cfg::TN;
  constexpr int WM = Cfg::WM, SLOT = Cfg::SLOT_FLOATS, LM = Cfg::LM, LN = Cfg::LN, QG = Cfg::QG;
  constexpr int WARPS = Cfg::WARPS, QT = TN / 2;
  constexpr int NFULL = Cfg::NFULL, NHALF = Cfg::NHALF, NB = GRAD_LAST_ROWS;
  static_assert(TN == 8, "the training kernel uses the 8 x 8 thread tile");
  const EvalParams& p = tp.ev;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const act = reinterpret_cast<float*>(smem_raw);  // [H][NC]: activations, then cotangents
  float* const ring = act + H * NC;                        // [STAGES][SLOT]
  float* const maha = ring + STAGES * SLOT;                // [NC]: d2 / 50 of each column
  float* const spare = maha + NC;                          // [NQ] (unused)
  const uint32_t bar_full = smem_u32(spare + NQ);
  const uint32_t bar_empty = bar_full + 8 * STAGES;
  uint2* const table = reinterpret_cast<uint2*>(spare + NQ + 4 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.n_layers;
  const long long n_tiles = (p.n + NQ - 1) / NQ;
  using GB = GradBlob<H>;
  using SB = SweepBlob<H>;
  const int nkh = (p.k_hidden + KC - 1) / KC;
  const uint32_t slices_per_tile = K0_PAD / KC + (L - 2 + NFULL + NHALF) * nkh + 1 + NB / KC + (L - 2) * nkh;
  const uint32_t my_tiles = uint32_t((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t total_slices = my_tiles * slices_per_tile;
  float* const slopes = tp.scratch + size_t(blockIdx.x) * grad_scratch_floats<H, NC>(L, true);  // [L-1][H][NC]
  float* const dlat = slopes + size_t(L - 1) * H * NC;                                          // [256][NC] parked latent cotangents
  float* const dz_last = tp.saved_dz + size_t(L - 1) * H * tp.cols;                             // DZ_L [cols][256]

  if (tid == 0) {  // the tile's slice sequence: forward layers, last-layer sweeps, scalar rows, W_last^T, W_l^T for l = L-2 .. 1
    uint32_t j = 0;
    for (int c = 0; c < K0_PAD / KC; ++c) table[j++] = make_uint2(c * KC * H, KC * H * 4);
    for (int l = 1; l < L - 1; ++l)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(blob_layer_offset<H>(l)) + c * KC * H, KC * H * 4);
    for (int sw = 0; sw < NFULL; ++sw)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(SB::full_off(L)) + (sw * H + c * KC) * H, KC * H * 4);
    for (int sw = 0; sw < NHALF; ++sw)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(SB::half_off(L)) + c * KC * (H / 2), KC * (H / 2) * 4);
    table[j++] = make_uint2(uint32_t(GB::scal_off(L)), uint32_t(p.k_hidden) * 8 * 4);
    for (int c = 0; c < NB / KC; ++c) table[j++] = make_uint2(uint32_t(GB::wst_off(L)) + c * KC * H, KC * H * 4);
    for (int l = L - 2; l >= 1; --l)
      for (int c = 0; c < nkh; ++c) table[j++] = make_uint2(uint32_t(GB::wt_off(L, l)) + c * KC * H, KC * H * 4);
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
  // this thread's tile element (row i of 8, its 4 query columns / twin columns) in a [.][H][NC] block
  auto tile_offset = [&](int i, bool twin) {
    const int row = (i < 4) ? h_row + i : H / 2 + h_row + (i - 4);
    return row * NC + (twin ? NC / 2 : 0) + h_col;
  };
  // sum over the lanes that hold the same rows (ln = 0 .. LN-1), then one atomic per row
  auto add_loss = [&](int row, float v) {
#pragma unroll
    for (int o = LN / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (ln == 0 && v != 0.0f) atomicAdd(tp.loss_out + row, double(v));
  };

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long tile_q0 = tile * NQ;
    const long long col0 = tile * NC;  // this tile's first column in the saved tensors

    // ---------------- inputs: x and its flipped twin [train.py:66-79], d2 / 50 of both (fp64) ----------------
    group_bar();
    for (int t = gtid; t < 2 * QG; t += Cfg::GROUP_THREADS) {
      const bool twin = t >= QG;
      const int c = wn * QG + (t % QG);
      const long long q = tile_q0 + c;
      const int col = twin ? NQ + c : c;
      double x[N_IN];
      if (q < p.n) {
#pragma unroll
        for (int i = 0; i < N_IN; ++i) x[i] = double(__ldg(tp.x + q * tp.x_ld + i));
      } else {  // padding column of a partial tile: harmless zeros; its loss cotangents are zero
#pragma unroll
        for (int i = 0; i < N_IN; ++i) x[i] = 0.0;
      }
      if (twin) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double u = x[i], l = x[8 + i];
          x[i] = -l;
          x[8 + i] = -u;
        }
        x[16] = -x[16];
        x[18] = -x[18];
        const double tmp = x[23];
        x[23] = x[24];
        x[24] = tmp;
      }
      double d[N_IN];
#pragma unroll
      for (int i = 0; i < N_IN; ++i) d[i] = x[i] - p.mean[i];
      double d2 = 0.0;
      int idx = 0;
#pragma unroll
      for (int i = 0; i < N_IN; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = i; j < N_IN; ++j) s = fma(p.quad[idx++], d[j], s);
        d2 = fma(d[i], s, d2);
      }
      maha[col] = static_cast<float>(d2 / (2.0 * N_IN));  // train.py:59
      float* const xs = tp.saved_a + size_t(col0 + col) * K0_PAD;
#pragma unroll
      for (int k4 = 0; k4 < K0_PAD / 4; ++k4) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = 4 * k4 + e;
          v[e] = k < N_IN ? static_cast<float>(x[k < N_IN ? k : 0]) : 0.0f;
          act[k * NC + col] = v[e];
        }
        __stcg(reinterpret_cast<float4*>(xs) + k4, make_float4(v[0], v[1], v[2], v[3]));
      }
    }
    group_bar();

    // ---------------- forward: hidden layers; a_l leaves the chip sample-major, SiLU' goes to the scratch block ----------------
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
      float* const a_out = tp.saved_a + train_saved_a_offset<H>(l + 1, tp.cols);
#pragma unroll
      for (int tw = 0; tw < 2; ++tw) {
        float y[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a0 = acc[i][tw * 2], a1 = acc[i][tw * 2 + 1];
          const float h[4] = {a0.x, a0.y, a1.x, a1.y};
          float d[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {  // silu(h) = h r, silu'(h) = r (1 + h (1 - r)), r = 1 / (1 + exp(-h))
            float ex, r;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(h[e] * -1.4426950408889634f));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex));
            y[i][e] = h[e] * r;
            d[e] = r * fmaf(h[e], 1.0f - r, 1.0f);
          }
          const int off = tile_offset(i, tw == 1);
          *reinterpret_cast<float4*>(act + off) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
          __stcg(reinterpret_cast<float4*>(sl + off), make_float4(d[0], d[1], d[2], d[3]));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {  // sample-major copy for the weight-gradient kernel: [column][feature]
          float* const dst = a_out + size_t(col0 + (tw ? NQ : 0) + h_col + e) * H;
          __stcg(reinterpret_cast<float4*>(dst + h_row), make_float4(y[0][e], y[1][e], y[2][e], y[3][e]));
          __stcg(reinterpret_cast<float4*>(dst + H / 2 + h_row), make_float4(y[4][e], y[5][e], y[6][e], y[7][e]));
        }
      }
      group_bar();  // also orders the slope stores before the reads of other threads of the group
    }

    // ---------------- last layer: sweeps; epilogue = un-flip, average, loss and the latents' cotangents ----------------
    {
      // four packed rows 4g .. 4g+3 of this thread's 4 queries and their twins  [train.py:83-112, 194-216]
      auto loss_rows = [&](const float2(&a4)[4][TN / 2], int g) {
        float yf[4][8], dl[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          yf[i][0] = a4[i][0].x; yf[i][1] = a4[i][0].y; yf[i][2] = a4[i][1].x; yf[i][3] = a4[i][1].y;
          yf[i][4] = a4[i][2].x; yf[i][5] = a4[i][2].y; yf[i][6] = a4[i][3].x; yf[i][7] = a4[i][3].y;
#pragma unroll
          for (int j = 0; j < 8; ++j) dl[i][j] = 0.0f;
        }
        const long long q0 = tile_q0 + h_col;
        if (g < 48) {
          int r[4];
          float sign;  // the twin's contribution enters ue/Vinf rows negated [train.py:97-104]
          if (g < N_BL) {
            r[0] = 6 + g; r[1] = 6 + 2 * N_BL + g; r[2] = 6 + 3 * N_BL + g; r[3] = 6 + 5 * N_BL + g;
            sign = -1.0f;
          } else {
            const int s = 2 * (g - N_BL);
            r[0] = 6 + N_BL + s; r[1] = r[0] + 1; r[2] = 6 + 4 * N_BL + s; r[3] = r[2] + 1;
            sign = 1.0f;
          }
          float lsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          float wgt[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) wgt[i] = __ldg(tp.loss_w + r[i]) * tp.inv_batch;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (q0 + j >= p.n) continue;
            // fused rows: (own row of the query) +- (partner row of the twin); partner of i is i ^ 2
            float pred[4];
            pred[0] = 0.5f * (yf[0][j] + yf[2][4 + j]);
            pred[1] = 0.5f * (yf[1][j] + sign * yf[3][4 + j]);
            pred[2] = 0.5f * (yf[2][j] + yf[0][4 + j]);
            pred[3] = 0.5f * (yf[3][j] + sign * yf[1][4 + j]);
            const float* const yrow = tp.y + (q0 + j) * tp.y_ld;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float ls, dls;
              huber_term(pred[i], __ldg(yrow + r[i]), tp.delta, ls, dls);
              lsum[i] = fmaf(wgt[i], ls, lsum[i]);
              const float gi = 0.5f * wgt[i] * dls;
              const float s_i = (i & 1) ? sign : 1.0f;  // rows 1 and 3 are the (possibly negated) second kind
              dl[i][j] = gi;
              dl[i ^ 2][4 + j] = s_i * gi;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) add_loss(r[i], lsum[i]);
        }
        // park the latent cotangents (zeros for padding groups): per-CTA block for the reverse sweep, sample-major for d W
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float* const row = dlat + size_t(4 * g + i) * NC;
          __stcg(reinterpret_cast<float4*>(row + h_col), make_float4(dl[i][0], dl[i][1], dl[i][2], dl[i][3]));
          __stcg(reinterpret_cast<float4*>(row + NQ + h_col), make_float4(dl[i][4], dl[i][5], dl[i][6], dl[i][7]));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float* const dst = dz_last + size_t(col0 + (j < 4 ? 0 : NQ) + h_col + (j & 3)) * TRAIN_LAST_ROWS + 4 * g;
          __stcg(reinterpret_cast<float4*>(dst), make_float4(dl[0][j], dl[1][j], dl[2][j], dl[3][j]));
        }
      };
      const float* const sweep_bias = p.blob + SB::bias_off(L);
      for (int sw = 0; sw < NFULL; ++sw) {
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
        const int g = sw * (H / 4) + wm * LM + lm;
        loss_rows(reinterpret_cast<const float2(&)[4][TN / 2]>(acc[0]), g);
        loss_rows(reinterpret_cast<const float2(&)[4][TN / 2]>(acc[4]), g + H / 8);
      }
      if constexpr (NHALF > 0) {
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
        loss_rows(acc, SB::HALF_ROW0 / 4 + wm * LM + lm);
      }
      group_bar();  // the parked rows of groups 48, 49 (zeros) are written before the scalar pass overwrites them
    }

    // ---------------- turn: the six scalar latents, their loss terms and cotangents [train.py:58-60, 89-92, 107-112] ----------------
    {
      const float* ws = wr.wait();
      if (wm == 0 && lane < QG) {
        const int c = wn * QG + lane;
        const long long q = tile_q0 + c;
        const float* const bs = p.blob + GB::scal_bias_off(L);
        float yq[6], yt[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) yq[i] = yt[i] = __ldg(bs + i);
        const float* w = ws;
        const float* const aq = act + c;
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
        const float z4 = 0.5f * (yq[4] + yt[5]), z5 = 0.5f * (yq[5] + yt[4]);
        const float pred[6] = {0.5f * ((yq[0] - maha[c]) + (yt[0] - maha[NQ + c])), 0.5f * (yq[1] - yt[1]), 0.5f * (yq[2] + yt[2]),
                               0.5f * (yq[3] - yt[3]), clip01_keep_nan(z4), clip01_keep_nan(z5)};
        float gz[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, ls[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (q < p.n) {
          const float* const yrow = tp.y + q * tp.y_ld;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float wi = __ldg(tp.loss_w + i) * tp.inv_batch;
            float l_i, d_i;
            if (i == 0) bce_logit_term(pred[0], __ldg(yrow), l_i, d_i);
            else huber_term(pred[i], __ldg(yrow + i), tp.delta, l_i, d_i);
            ls[i] = wi * l_i;
            gz[i] = wi * d_i;
          }
          // torch.clip passes the gradient on the closed interval [train.py:107-112]
          if (!(z4 >= 0.0f && z4 <= 1.0f)) gz[4] = 0.0f;
          if (!(z5 >= 0.0f && z5 <= 1.0f)) gz[5] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          float v = ls[i];
#pragma unroll
          for (int o = QG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(QG == 32 ? 0xffffffffu : 0x0000ffffu, v, o);
          if (lane == 0 && v != 0.0f) atomicAdd(tp.loss_out + i, double(v));
        }
        const float dq[6] = {0.5f * gz[0], 0.5f * gz[1], 0.5f * gz[2], 0.5f * gz[3], 0.5f * gz[4], 0.5f * gz[5]};
        const float dt[6] = {0.5f * gz[0], -0.5f * gz[1], 0.5f * gz[2], -0.5f * gz[3], 0.5f * gz[5], 0.5f * gz[4]};
        float* const gq = dz_last + size_t(col0 + c) * TRAIN_LAST_ROWS + SB::SCAL_ROW0;
        float* const gt = dz_last + size_t(col0 + NQ + c) * TRAIN_LAST_ROWS + SB::SCAL_ROW0;
#pragma unroll
        for (int i = 0; i < NB - SB::SCAL_ROW0; ++i) {  // rows 198..207: zero padding of the reverse sweep
          const float vq = i < 6 ? dq[i < 6 ? i : 0] : 0.0f, vt = i < 6 ? dt[i < 6 ? i : 0] : 0.0f;
          __stcg(dlat + size_t(SB::SCAL_ROW0 + i) * NC + c, vq);
          __stcg(dlat + size_t(SB::SCAL_ROW0 + i) * NC + NQ + c, vt);
          __stcg(gq + i, vq);
          __stcg(gt + i, vt);
        }
      }
      __syncwarp();  // the turn diverges (one query per lane, q < n); every lane has finished with the slot
      wr.release();
    }
    group_bar();

    // ---------------- backward: d a_{L-2} = W_last^T d latent, the parked rows brought back H at a time ----------------
    auto store_dz = [&](const float2(&acc)[8][TN / 2], int l) {  // d z_l = d a_l * SiLU'(z_l) -> tile, and sample-major for d W
      const float* const sl = slopes + size_t(l) * H * NC;
      float* const dz_out = tp.saved_dz + size_t(l) * H * tp.cols;
#pragma unroll
      for (int tw = 0; tw < 2; ++tw) {
        float y[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a0 = acc[i][tw * 2], a1 = acc[i][tw * 2 + 1];
          const int off = tile_offset(i, tw == 1);
          const float4 m = __ldcg(reinterpret_cast<const float4*>(sl + off));  // this thread stored those slopes itself
          y[i][0] = a0.x * m.x; y[i][1] = a0.y * m.y; y[i][2] = a1.x * m.z; y[i][3] = a1.y * m.w;
          *reinterpret_cast<float4*>(act + off) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float* const dst = dz_out + size_t(col0 + (tw ? NQ : 0) + h_col + e) * H;
          __stcg(reinterpret_cast<float4*>(dst + h_row), make_float4(y[0][e], y[1][e], y[2][e], y[3][e]));
          __stcg(reinterpret_cast<float4*>(dst + H / 2 + h_row), make_float4(y[4][e], y[5][e], y[6][e], y[7][e]));
        }
      }
    };
    {
      float2 acc[8][TN / 2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) acc[i][jj] = make_float2(0.0f, 0.0f);
      for (int r0 = 0; r0 < NB; r0 += H) {
        const int rows_here = (NB - r0 < H) ? NB - r0 : H;
        if (r0 > 0) group_bar();  // the previous chunk has been read
        constexpr int V4 = 2 * QG / 4;  // float4s per row of the group's columns: queries, then twins
        for (int idx = gtid; idx < rows_here * V4; idx += Cfg::GROUP_THREADS) {
          const int r = idx / V4, cc = idx % V4;
          const int col = (cc < QG / 4) ? wn * QG + 4 * cc : NQ + wn * QG + 4 * (cc - QG / 4);
          *reinterpret_cast<float4*>(act + r * NC + col) = __ldcg(reinterpret_cast<const float4*>(dlat + size_t(r0 + r) * NC + col));
        }
        group_bar();
        for (int c = 0; c < rows_here / KC; ++c) {
          const float* ws = wr.wait();
          ffma_slice<8, TN, H, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
          wr.release();
        }
      }
      group_bar();
      store_dz(acc, L - 2);
      group_bar();
    }
    // ---------------- backward: d a_{l-1} = W_l^T d z_l ----------------
    for (int l = L - 2; l >= 1; --l) {
      float2 acc[8][TN / 2];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jj = 0; jj < TN / 2; ++jj) acc[i][jj] = make_float2(0.0f, 0.0f);
      for (int c = 0; c < nkh; ++c) {
        const float* ws = wr.wait();
        ffma_slice<8, TN, H, NC, KC>(ws + h_row, act + (c * KC) * NC + h_col, acc);
        wr.release();
      }
      group_bar();  // the group has finished reading d z_l
      store_dz(acc, l - 1);
      group_bar();
    }
  }
}

// ---- weight gradients -----------------------------------------------------------------------------------------------
// One job per layer: d W[m][k] = sum_c DZ[c][m] * A[c][k], d b[m] = sum_c DZ[c][m], c over all columns (samples and twins).
struct WgradJob {
  const float* dz;   // [cols][ldz]
  const float* a;    // [cols][lda]
  int ldz, lda;
  int m_blocks, n_blocks;   // CTA tiles of BM x BN covering the padded M x N
  int n_valid;              // columns of A that exist (the input layer has 32, of a 64- or 128-wide tile)
  int first_cta;            // prefix sum of m_blocks * n_blocks * k_splits over the jobs
  long long part_off;       // float offset of this job's [M_pad][N_pad] block (then [M_pad] bias sums) in a partial buffer
};
struct WgradParams {
  WgradJob job[NFB_MAX_LAYERS_K];
  int n_jobs;
  int k_splits;
  long long cols;           // multiple of 16
  float* partial;           // [k_splits][part_stride]
  long long part_stride;
};

template <int BM_, int BN_>
struct WgradCfg {
  static constexpr int BM = BM_, BN = BN_, BK = 16, STAGES = 3;
  static constexpr int TX = BN / 8, TY = BM / 8;  // threads along n and m: each an 8 x 8 tile as 2 x 2 quads of 4 x 4
  static constexpr int THREADS = TX * TY;
  static constexpr size_t SMEM_BYTES = sizeof(float) * STAGES * BK * (BM + BN);
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS) nfb_train_wgrad_kernel(const WgradParams wp) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, TX = Cfg::TX, THREADS = Cfg::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* const sm = reinterpret_cast<float*>(smem_raw);  // [STAGES][BK][BM + BN]: DZ fragment rows, then A fragment rows

  int j = 0;
  while (j + 1 < wp.n_jobs && int(blockIdx.x) >= wp.job[j + 1].first_cta) ++j;
  const WgradJob& job = wp.job[j];
  int local = blockIdx.x - job.first_cta;
  const int ks = local % wp.k_splits;
  local /= wp.k_splits;
  const int nb = local % job.n_blocks, mb = local / job.n_blocks;
  const int m0 = mb * BM, n0 = nb * BN;

  const long long steps_total = wp.cols / BK;
  const long long s_begin = steps_total * ks / wp.k_splits, s_end = steps_total * (ks + 1) / wp.k_splits;
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;

  auto load_stage = [&](int stage, long long step) {
    float* const dst = sm + size_t(stage) * BK * (BM + BN);
    const long long c0 = step * BK;
    constexpr int ZV = BM / 4, AV = BN / 4;  // 16-byte chunks per column
    for (int i = tid; i < BK * (ZV + AV); i += THREADS) {
      const int r = i / (ZV + AV), v = i % (ZV + AV);
      if (v < ZV) {
        cp_async16(smem_u32(dst + r * (BM + BN) + 4 * v), job.dz + (c0 + r) * job.ldz + m0 + 4 * v, true);
      } else {
        const int k = n0 + 4 * (v - ZV);
        const bool ok = k < job.n_valid;
        cp_async16(smem_u32(dst + r * (BM + BN) + BM + 4 * (v - ZV)), job.a + (c0 + r) * job.lda + (ok ? k : 0), ok);
      }
    }
  };

  float2 acc[8][4];
  float bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bsum[i] = 0.0f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = make_float2(0.0f, 0.0f);
  }
  const long long n_steps = s_end - s_begin;
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < n_steps) load_stage(s, s_begin + s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (long long st = 0; st < n_steps; ++st) {
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
    __syncthreads();  // stage st has landed for everybody; everybody has finished stage st - 1, whose slot is refilled now
    if (st + STAGES - 1 < n_steps) load_stage(int((st + STAGES - 1) % STAGES), s_begin + st + STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* const cur = sm + size_t(st % STAGES) * BK * (BM + BN);
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      const float* const row = cur + r * (BM + BN);
      const float4 z0 = *reinterpret_cast<const float4*>(row + 4 * ty);
      const float4 z1 = *reinterpret_cast<const float4*>(row + BM / 2 + 4 * ty);
      const float4 a0 = *reinterpret_cast<const float4*>(row + BM + 4 * tx);
      const float4 a1 = *reinterpret_cast<const float4*>(row + BM + BN / 2 + 4 * tx);
      const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
      const float2 av[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][jj] = __ffma2_rn(make_float2(zv[i], zv[i]), av[jj], acc[i][jj]);
      if (nb == 0 && tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) bsum[i] += zv[i];
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  // this split's partial sums: [M_pad][N_pad] row-major, then [M_pad] bias sums
  const int n_pad = job.n_blocks * BN, m_pad = job.m_blocks * BM;
  float* const part = wp.partial + size_t(ks) * wp.part_stride + job.part_off;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? 4 * ty + i : BM / 2 + 4 * ty + (i - 4));
    float* const row = part + size_t(m) * n_pad + n0;
    *reinterpret_cast<float4*>(row + 4 * tx) = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
    *reinterpret_cast<float4*>(row + BN / 2 + 4 * tx) = make_float4(acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y);
    if (nb == 0 && tx == 0) part[size_t(m_pad) * n_pad + m] = bsum[i];
  }
}

// Sum the splits in a fixed order and write the gradients in the reference's layout: d W (out, in) row-major, d b (out).
struct ReduceJob {
  float* grad_w;     // (out_dim, in_dim)
  float* grad_b;     // (out_dim)
  int out_dim, in_dim, n_pad, m_pad;
  int last;          // rows are in packed order (nfb_common.cuh: pack_last_layer_row)
  long long part_off;
  long long first;   // prefix sum of out_dim * (in_dim + 1) over the jobs
};
struct ReduceParams {
  ReduceJob job[NFB_MAX_LAYERS_K];
  int n_jobs, k_splits;
  const float* partial;
  long long part_stride, total;
  const int* packed_row_of;  // [198] inverse of pack_last_layer_row
};

__global__ void nfb_train_reduce_kernel(const ReduceParams rp) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= rp.total) return;
  int j = 0;
  while (j + 1 < rp.n_jobs && idx >= rp.job[j + 1].first) ++j;
  const ReduceJob& job = rp.job[j];
  const long long e = idx - job.first;
  const int o = int(e / (job.in_dim + 1)), i = int(e % (job.in_dim + 1));
  const int m = job.last ? __ldg(rp.packed_row_of + o) : o;
  const long long at = job.part_off + (i < job.in_dim ? (long long)m * job.n_pad + i : (long long)job.m_pad * job.n_pad + m);
  float s = 0.0f;
  for (int k = 0; k < rp.k_splits; ++k) s += rp.partial[size_t(k) * rp.part_stride + at];
  if (i < job.in_dim) job.grad_w[(long long)o * job.in_dim + i] = s;
  else job.grad_b[o] = s;
}

}  // namespace nfb
