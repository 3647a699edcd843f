// Shared definitions for the neuralfoil_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfb {

constexpr int N_RAW = 23;        // raw scalars per query (reference main.py:171-184)
constexpr int N_IN = 25;         // latent inputs
constexpr int K0_PAD = 32;       // first layer's K, zero-padded from 25
constexpr int N_OUT = 198;       // latent / physical outputs per query
constexpr int N_BL = 32;         // boundary-layer stations per side (_basic_data_type.py:51)
constexpr int M_LAST_USED = 200;  // last layer's rows after the permutation of pack_last_layer_row
constexpr int M_LAST = 224;      // ... padded to a whole number of warps' row groups (16 or 32 rows)
constexpr int N_MAHA = 325;      // packed upper triangle of the 25x25 quadratic form

constexpr int NFB_MAX_LAYERS_K = 16;  // == NFB_MAX_LAYERS of the C header



// Per-launch arguments.  Passed by value (kernel parameter space, constant bank).
struct EvalParams {
  const void* in[N_RAW];       // device row pointers
  long long in_stride[N_RAW];  // element strides (0 = broadcast)
  void* out;                   // 198 x out_ld
  long long out_ld;
  long long n;                 // queries
  const float* blob;           // packed weights (see nfb_api.cu: pack_model)
  int n_layers;                // affine layers L
  int in_f64;
  int out_f64;
  int scalars_only;            // NFB_OUTPUT_SCALARS: store rows 0..5 only (out is 6 x out_ld)
  int k_hidden;                // widest hidden layer rounded up to 16: rows of K beyond it are zero padding
  // Mean (25) and packed symmetric quadratic form (325, off-diagonals pre-summed S_ij + S_ji) of the scaled-input
  // distribution (reference main.py:27-32, 47-61).  They travel with every launch in the kernel parameter space --
  // the constant bank, so the fp64 featurisation still reads them as constant operands -- and therefore belong to the
  // CONTEXT that launches: two contexts on one device may hold different distributions (a __constant__ symbol, one
  // copy per device, could not), and nothing is shared between threads or streams.
  double mean[N_IN];
  double quad[N_MAHA];
  // Mixed-precision outputs (nfb_eval_*_mixed): when not NULL, the 192 boundary-layer rows (output rows 6..197) are written
  // as bfloat16 (round to nearest) into this 192 x out_ld block instead of `out`, which then holds the six scalar rows only.
  void* out_bl;
};
static_assert(sizeof(EvalParams) + 64 <= 4096, "EvalParams and its wrappers must fit the 4 KB kernel parameter space");

// ---- Row permutation of the last layer -------------------------------------------------------
// Packed row p = 4 * g + i of the last layer holds latent output row pack_last_layer_row(p)
// (or -1 for padding).  Groups of four rows are what one thread owns in the last layer, chosen so
// that un-flip + average + decode (main.py:274-316) need no data from any other thread:
//   g in [0,32)  : station g   -> theta_upper, ue_upper, theta_lower, ue_lower
//   g in [32,48) : stations 2(g-32), +1 -> H_upper x2, H_lower x2
//   g == 48      : analysis_confidence logit, CL, ln CD, CM latents
//   g == 49      : Top_Xtr, Bot_Xtr, pad, pad
//   g in [50,56) : pad
__host__ __device__ constexpr int pack_last_layer_row(int p) {
  const int g = p >> 2, i = p & 3;
  if (g < 32) return (i == 0) ? 6 + g : (i == 1) ? 70 + g : (i == 2) ? 102 + g : 166 + g;
  if (g < 48) {
    const int s = 2 * (g - 32);
    return (i == 0) ? 38 + s : (i == 1) ? 39 + s : (i == 2) ? 134 + s : 135 + s;
  }
  if (g == 48) return i;
  if (g == 49) return (i < 2) ? 4 + i : -1;
  return -1;
}

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                              uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// two fp32 -> packed bf16x2 (lo = first argument), round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

}  // namespace nfb
