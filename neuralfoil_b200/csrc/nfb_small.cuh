// Small-batch (latency) kernel: BASELINE.json configs[3], a design optimiser calling with 1..1000 queries.
//
// The throughput kernel (nfb_ffma.cuh) gives one SM a tile of 32..256 queries; a call with a handful of queries
// would run that whole padded tile on a single SM (measured: 87 us for xxsmall, 174 us for xlarge at N = 1).
// Here a CTA takes FOUR queries (8 columns with their twins), one thread per output row of the layer in flight,
// weights read straight from the L2-resident packed blob (K-major, so a warp's loads are coalesced), activations
// ping-ponged through 2 x 16 KB of shared memory.  N queries spread over ceil(N / 4) SMs.
//
// The arithmetic is the throughput kernel's, instruction for instruction per accumulator: acc = bias, then one
// IEEE fma per k in k order, the same SiLU, the same featurise_column and decode_and_store.  A query therefore
// gets bit-identical results from either kernel (tests/test_gpu_parity.py::test_small_and_large_kernels_agree).
#pragma once
#include "nfb_ffma.cuh"

// weight loads in flight per thread and round (every K is a multiple of 32): a layer is K / U dependent round trips to the L2.
// Measured 16 against 32: one query on xlarge 42.4 -> 41.0 us through nfb_eval_host, but 1,024 queries 199 -> 213 us (registers).
#ifndef NFB_SMALL_U
#define NFB_SMALL_U 16
#endif

namespace nfb {

constexpr int SMALL_QUERIES = 4;                 // per CTA
constexpr int SMALL_COLS = 2 * SMALL_QUERIES;    // + twins

template <int H>
struct SmallCfg {
  static constexpr int THREADS = H > 256 ? H : 256;  // >= 224 rows of the last layer
  static constexpr int KMAX = H > K0_PAD ? H : K0_PAD;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (2 * KMAX * SMALL_COLS + M_LAST * SMALL_COLS + SMALL_COLS + SMALL_QUERIES);
};

// acc[j] += sum_k W[k][m] * act[k][j]; W row stride `ldw` floats, U loads in flight per thread
template <int U>
__device__ __forceinline__ void small_layer(const float* __restrict__ w, int ldw, int K, const float* __restrict__ act,
                                            float (&acc)[SMALL_COLS]) {
  for (int k0 = 0; k0 < K; k0 += U) {
    float wv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) wv[u] = __ldg(w + size_t(k0 + u) * ldw);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 a0 = *reinterpret_cast<const float4*>(act + (k0 + u) * SMALL_COLS);
      const float4 a1 = *reinterpret_cast<const float4*>(act + (k0 + u) * SMALL_COLS + 4);
      acc[0] = fmaf(wv[u], a0.x, acc[0]); acc[1] = fmaf(wv[u], a0.y, acc[1]);
      acc[2] = fmaf(wv[u], a0.z, acc[2]); acc[3] = fmaf(wv[u], a0.w, acc[3]);
      acc[4] = fmaf(wv[u], a1.x, acc[4]); acc[5] = fmaf(wv[u], a1.y, acc[5]);
      acc[6] = fmaf(wv[u], a1.z, acc[6]); acc[7] = fmaf(wv[u], a1.w, acc[7]);
    }
  }
}

template <int H>
__global__ void __launch_bounds__(SmallCfg<H>::THREADS, 1) nfb_small_kernel(const EvalParams p) {
  using Cfg = SmallCfg<H>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* const act0 = reinterpret_cast<float*>(smem_raw);       // [KMAX][8]
  float* const act1 = act0 + Cfg::KMAX * SMALL_COLS;            // [KMAX][8]
  float* const stage = act1 + Cfg::KMAX * SMALL_COLS;           // [224][8] last layer's latents, packed row order
  float* const maha = stage + M_LAST * SMALL_COLS;              // [8]
  float* const re_s = maha + SMALL_COLS;                        // [4]

  const int tid = threadIdx.x, L = p.n_layers;
  const long long q0 = (long long)blockIdx.x * SMALL_QUERIES;

  if (tid < SMALL_COLS) {
    const bool twin = tid >= SMALL_QUERIES;
    const long long q = q0 + (twin ? tid - SMALL_QUERIES : tid);
    featurise_column<SMALL_COLS>(p, q, q < p.n, twin, tid, act0, maha, re_s);
  }
  __syncthreads();

  float* cur = act0;
  float* nxt = act1;
  for (int l = 0; l < L - 1; ++l) {
    const int K = (l == 0) ? K0_PAD : H;
    const float* w = p.blob + blob_layer_offset<H>(l);
    if (tid < H) {
      float acc[SMALL_COLS];
      const float b = __ldg(w + size_t(K) * H + tid);
#pragma unroll
      for (int j = 0; j < SMALL_COLS; ++j) acc[j] = b;
      small_layer<NFB_SMALL_U>(w + tid, H, K, cur, acc);
      float4 lo, hi;
      lo.x = silu(acc[0]); lo.y = silu(acc[1]); lo.z = silu(acc[2]); lo.w = silu(acc[3]);
      hi.x = silu(acc[4]); hi.y = silu(acc[5]); hi.z = silu(acc[6]); hi.w = silu(acc[7]);
      *reinterpret_cast<float4*>(nxt + tid * SMALL_COLS) = lo;
      *reinterpret_cast<float4*>(nxt + tid * SMALL_COLS + 4) = hi;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  {
    const float* w = p.blob + blob_layer_offset<H>(L - 1);
    if (tid < M_LAST) {
      float acc[SMALL_COLS];
      const float b = __ldg(w + size_t(H) * M_LAST + tid);
#pragma unroll
      for (int j = 0; j < SMALL_COLS; ++j) acc[j] = b;
      small_layer<NFB_SMALL_U>(w + tid, M_LAST, H, cur, acc);
#pragma unroll
      for (int j = 0; j < SMALL_COLS; ++j) stage[tid * SMALL_COLS + j] = acc[j];
    }
  }
  __syncthreads();
  if (tid < M_LAST / 4) {  // one thread per packed row group: un-flip, average, decode, store (as the big kernel)
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < SMALL_COLS; ++j) acc[i][j] = stage[(4 * tid + i) * SMALL_COLS + j];
    const long long left = p.n - q0;
    const int nvalid = left >= 4 ? 4 : (left > 0 ? int(left) : 0);
    const bool vec_ok =
        p.out_f64 ? ((p.out_ld & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                  : ((p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    decode_and_store(p, tid, q0, nvalid, vec_ok, acc, maha, maha + SMALL_QUERIES, re_s);
  }
}

}  // namespace nfb
