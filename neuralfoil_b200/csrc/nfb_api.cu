// C ABI of neuralfoil_b200 (see include/neuralfoil_b200.h).  Single translation unit: the kernels are
// included below.  No torch, no CPU fallback: every entry point that computes needs an sm_100 device.
#include "../../include/neuralfoil_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <cctype>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include <sched.h>

#include "nfb_ffma.cuh"
#include "nfb_ffma_groups.cuh"
#include "nfb_small.cuh"
#include "nfb_tc.cuh"
#include "nfb_tc2.cuh"
#include "nfb_tcn.cuh"
#include "nfb_jac.cuh"
#include "nfb_grad.cuh"
#include "nfb_train.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define NFB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (expr);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(NFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                   \
  } while (0)

struct Model {
  bool loaded = false;
  int n_layers = 0;
  int h_pad = 0;  // 64 / 128 / 256 / 512
  int k_hidden = 0;  // widest hidden layer rounded up to 16
  int dims[NFB_MAX_LAYERS + 1] = {0};
  float* blob = nullptr;  // device
  float* grad_blob = nullptr;  // nfb_grad.cuh: hidden layers, scalar rows, transposed hidden layers
  // tensor-core variants (256- / 512-wide models): fp16 weight slices of the one-term and the three-term kernels,
  // fp32 biases, per-layer 1 / scale
  __half* tc_w = nullptr;
  __half* tc3_w = nullptr;
  __half* tc2_w = nullptr;   // the CTA-pair kernels' streams (nfb_tc2.cuh): one-term, three-term
  __half* tc2_w3 = nullptr;
  float* tc_bias = nullptr;
  float* tc_inv_scale = nullptr;
  // narrow models (64- / 128-wide): weight streams of nfb_tcn.cuh (one-term, three-term), biases [L][256]
  __half* tcn_w = nullptr;
  __half* tcn_w3 = nullptr;
  float* tcn_bias = nullptr;
};

// Kernel tuning per padded width: <H, KC, STAGES>.
// Measured on B200: the 8 x 16 thread tile (8 warps) wins where the GEMM dominates
// (wide models), the 8 x 8 tile (16 warps) where the per-query fp64/MUFU work does (narrow models).
#ifndef NFB_TN_WIDE
#define NFB_TN_WIDE 16
#endif
#ifndef NFB_TN_NARROW
#define NFB_TN_NARROW 8
#endif
#ifndef NFB_KC
#define NFB_KC 8
#endif
#ifndef NFB_KC_NARROW
#define NFB_KC_NARROW 16  // 64-wide models: a K-slice is only 256 FFMA2 cycles per warp; 16-row slices hide the TMA latency (+5 %)
#endif
#ifndef NFB_STAGES_NARROW
#define NFB_STAGES_NARROW 4
#endif
#ifndef NFB_STAGES
#define NFB_STAGES 4
#endif
#ifndef NFB_KC_128
#define NFB_KC_128 NFB_KC
#endif
#ifndef NFB_STAGES_128
#define NFB_STAGES_128 NFB_STAGES_NARROW
#endif
#ifndef NFB_KC_256
#define NFB_KC_256 NFB_KC
#endif
#ifndef NFB_STAGES_256
#define NFB_STAGES_256 NFB_STAGES
#endif
using Cfg64 = nfb::FfmaCfg<64, NFB_KC_NARROW, NFB_STAGES_NARROW, NFB_TN_NARROW>;
using Cfg128 = nfb::FfmaCfg<128, NFB_KC_128, NFB_STAGES_128, NFB_TN_WIDE>;
using Cfg256 = nfb::FfmaCfg<256, NFB_KC_256, NFB_STAGES_256, NFB_TN_WIDE>;
using Cfg512 = nfb::FfmaCfg<512, NFB_KC, NFB_STAGES, NFB_TN_WIDE>;
// Column-group kernels (nfb_ffma_groups.cuh) for the widths up to 256: <H, KC, ring depth, TN>
#ifndef NFB_GRP_KC_64
#define NFB_GRP_KC_64 8
#endif
#ifndef NFB_GRP_STAGES_64
#define NFB_GRP_STAGES_64 16
#endif
#ifndef NFB_GRP_KC_128
#define NFB_GRP_KC_128 8
#endif
#ifndef NFB_GRP_STAGES_128
#define NFB_GRP_STAGES_128 16
#endif
#ifndef NFB_GRP_KC_256
#define NFB_GRP_KC_256 8
#endif
#ifndef NFB_GRP_STAGES_256
#define NFB_GRP_STAGES_256 8
#endif
using Grp64 = nfb::GroupCfg<64, NFB_GRP_KC_64, NFB_GRP_STAGES_64, NFB_TN_NARROW>;
#ifndef NFB_GRP_TN_128
#define NFB_GRP_TN_128 8  // 16 warps x (8 x 8) against 8 warps x (8 x 16): +1..2.5 % at this width, a wash at 256 and 512
#endif
using Grp128 = nfb::GroupCfg<128, NFB_GRP_KC_128, NFB_GRP_STAGES_128, NFB_GRP_TN_128>;
using Grp256 = nfb::GroupCfg<256, NFB_GRP_KC_256, NFB_GRP_STAGES_256, NFB_TN_WIDE>;
#ifndef NFB_GRAD_TN_128
#define NFB_GRAD_TN_128 NFB_GRP_TN_128
#endif
using Grad128 = nfb::GroupCfg<128, NFB_GRP_KC_128, NFB_GRP_STAGES_128, NFB_GRAD_TN_128>;  // the reverse-mode kernels' 128-wide configuration
#ifndef NFB_GRAD_TN_WIDE
#define NFB_GRAD_TN_WIDE 8  // 16 warps hide the single-warp steps of the reverse pass better: xxxlarge +7 % (at 128: 8 x 16 tile -19 %)
#endif
using Grad256 = nfb::GroupCfg<256, NFB_GRP_KC_256, NFB_GRP_STAGES_256, NFB_GRAD_TN_WIDE>;
#ifndef NFB_GRP_STAGES_512
#define NFB_GRP_STAGES_512 5
#endif
#ifndef NFB_GRP_LEAD_512
#define NFB_GRP_LEAD_512 0
#endif
using Grp512 = nfb::GroupCfg<512, 8, NFB_GRP_STAGES_512, NFB_TN_WIDE, 16, NFB_GRP_LEAD_512>;
using Grad512 = nfb::GroupCfg<512, 8, NFB_GRP_STAGES_512, NFB_GRAD_TN_WIDE, 16, NFB_GRP_LEAD_512>;
// the training kernels are written for the 8 x 8 thread tile, whatever the reverse-mode kernels use
using Train256 = nfb::GroupCfg<256, NFB_GRP_KC_256, NFB_GRP_STAGES_256, 8>;
using Train512 = nfb::GroupCfg<512, 8, NFB_GRP_STAGES_512, 8, 16, NFB_GRP_LEAD_512>;

struct HostSlot {  // one in-flight chunk of nfb_eval_host
  cudaStream_t stream = nullptr;
  void* d_in = nullptr;
  void* d_out = nullptr;
  size_t in_bytes = 0, out_bytes = 0;
  void* h_stage = nullptr;  // pinned staging for the small-batch path
  size_t stage_bytes = 0;
  void* h_in = nullptr;     // pinned staging of the large-batch path when the caller's arrays are pageable
  void* h_out = nullptr;
  size_t h_in_bytes = 0, h_out_bytes = 0;
};

}  // namespace

struct nfb_context {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  // Serialises every entry point that touches this context's mutable state (the host entry points' staging slots, the
  // reverse-mode scratch block, the set of configured kernels).  Device entry points hold it only while they enqueue.
  // Recursive: the host entry points call the device ones.
  std::recursive_mutex mu;
  bool have_distribution = false;
  double mean[nfb::N_IN] = {0};   // main.py:27-32; copied into every launch's parameters (nfb_common.cuh: EvalParams)
  double quad[nfb::N_MAHA] = {0};
  std::unordered_set<const void*> configured;  // kernels whose dynamic shared memory limit has been raised on this device
  int64_t tune_small_batch_max = -1;  // nfb_set_tuning; -1 = the measured defaults
  int64_t tune_mapped_max = -1;
  int tune_tc2_max_grid = 0;
  int host_thread_cap = 0;            // 0 = all; nfb_multi_* divides the host cores among its devices
  Model models[NFB_MAX_MODELS];
  HostSlot slots[2];
  void* grad_scratch = nullptr;  // nfb_grad.cuh: SiLU slopes and parked cotangents, one block per CTA
  size_t grad_scratch_bytes = 0;
  void* train_ws = nullptr;      // nfb_train_step_device: saved activations / cotangents and the split partial sums
  size_t train_ws_bytes = 0;
  int* packed_row_of = nullptr;  // [198] inverse of nfb::pack_last_layer_row (device)
  cudaEvent_t grad_event = nullptr;      // recorded behind the last launch that uses grad_scratch ...
  cudaStream_t grad_stream = nullptr;    // ... on this stream: a launch on another stream waits for it first
  bool grad_event_recorded = false;
  void* jac_dev = nullptr;   // nfb_eval_jacobian_host: [inputs | values | jac] device block and its pinned host mirror
  void* jac_host = nullptr;
  size_t jac_dev_bytes = 0, jac_host_bytes = 0;
  std::atomic<int64_t> launches{0};
  long long* tc_debug = nullptr;  // device buffer for nfb_debug_tc_stamps (64 x clock64), normally NULL
};

namespace {

// Host-side repack of the reference's (out, in) row-major weights into the kernel's K-major,
// zero-padded layout; the last layer's rows are permuted by nfb::pack_last_layer_row.
template <int H>
std::vector<float> pack_model(int L, const int* dims, const float* const* W, const float* const* B) {
  std::vector<float> blob(nfb::blob_total_floats<H>(L), 0.0f);
  {  // sweep-layout copy of the last layer for nfb_ffma_groups_kernel (nfb_ffma_groups.cuh)
    using SB = nfb::SweepBlob<H>;
    blob.resize(SB::total(L), 0.0f);
    const int in = dims[L - 1], out = dims[L];
    float* bias = blob.data() + SB::bias_off(L);
    for (int pr = 0; pr < nfb::M_LAST_SWEPT; ++pr) {
      const int src_row = nfb::pack_last_layer_row(pr);
      if (src_row < 0 || src_row >= out) continue;
      bias[pr] = B[L - 1][src_row];
      float* dst;      // element (k = 0) of this packed row
      size_t k_stride;
      if (pr < SB::NFULL * H) {
        dst = blob.data() + SB::full_off(L) + size_t(pr / H) * H * H + pr % H;
        k_stride = H;
      } else if (SB::NHALF && pr < SB::HALF_ROW0 + H / 2) {
        dst = blob.data() + SB::half_off(L) + (pr - SB::HALF_ROW0);
        k_stride = H / 2;
      } else if (SB::NSCAL && pr >= SB::SCAL_ROW0 && pr < SB::SCAL_ROW0 + 8) {
        dst = blob.data() + SB::scal_off(L) + (pr - SB::SCAL_ROW0);
        k_stride = 8;
      } else {
        continue;
      }
      for (int k = 0; k < in; ++k) dst[k * k_stride] = W[L - 1][size_t(src_row) * in + k];
    }
  }
  for (int l = 0; l < L; ++l) {
    const int in = dims[l], out = dims[l + 1];
    const bool last = (l == L - 1);
    const int K = (l == 0) ? nfb::K0_PAD : H;
    const int M = last ? nfb::M_LAST : H;
    float* wt = blob.data() + nfb::blob_layer_offset<H>(l);
    float* bias = wt + size_t(K) * M;
    for (int m = 0; m < M; ++m) {
      const int src_row = last ? nfb::pack_last_layer_row(m) : (m < out ? m : -1);
      if (src_row < 0 || src_row >= out) continue;
      for (int k = 0; k < in; ++k) wt[size_t(k) * M + m] = W[l][size_t(src_row) * in + k];
      bias[m] = B[l][src_row];
    }
  }
  return blob;
}

// Blob of the reverse-mode kernels (nfb_grad.cuh: GradBlob): the value blob, the six scalar rows of the last
// layer as Wt[k][8] with their biases, and the hidden layers transposed, WT_l[m][k] = W_l[m][k] zero-padded to H x 32 (l = 0)
// or H x H.
template <int H>
std::vector<float> pack_grad_blob(int L, const int* dims, const float* const* W, const float* const* B,
                                  const std::vector<float>& value_blob) {
  using GB = nfb::GradBlob<H>;
  std::vector<float> g(GB::total(L), 0.0f);
  std::copy(value_blob.begin(), value_blob.end(), g.begin());  // == SweepBlob<H>::total(L) floats
  const int in_last = dims[L - 1];
  for (int pr = 0; pr < nfb::GRAD_LAST_ROWS; ++pr) {  // the last layer transposed by packed row
    const int src_row = nfb::pack_last_layer_row(pr);
    if (src_row < 0 || src_row >= dims[L]) continue;
    for (int k = 0; k < in_last; ++k) g[GB::wst_off(L) + size_t(pr) * H + k] = W[L - 1][size_t(src_row) * in_last + k];
  }
  for (int i = 0; i < 6; ++i) {
    for (int k = 0; k < in_last; ++k) g[GB::scal_off(L) + size_t(k) * 8 + i] = W[L - 1][size_t(i) * in_last + k];
    g[GB::scal_bias_off(L) + i] = B[L - 1][i];
  }
  for (int l = 0; l < L - 1; ++l) {
    const int in = dims[l], out = dims[l + 1];
    float* wt = g.data() + (l == 0 ? GB::wt0_off(L) : GB::wt_off(L, l));
    const size_t ld = (l == 0) ? nfb::K0_PAD : H;
    for (int m = 0; m < out; ++m)
      for (int k = 0; k < in; ++k) wt[size_t(m) * ld + k] = W[l][size_t(m) * in + k];
  }
  return g;
}

// Tensor-core variant: fp16 weight slices (256 features x 32 k, K-major no-swizzle core matrices: element (r, kk)
// at (r/8)*512 + (kk/8)*128 + (r%8)*16 + (kk%8)*2 bytes) in streaming order, see nfb_tc.cuh.
struct TcPack {
  std::vector<__half> w;
  std::vector<float> bias, inv_scale;
};
template <class Cfg>
TcPack pack_model_tc(int L, const int* dims, const float* const* W, const float* const* B) {
  using namespace nfb::tc;
  constexpr int H = Cfg::H, SLICE_N = Cfg::SLICE_N;
  constexpr uint32_t SLICE_BYTES = Cfg::SLICE_BYTES;
  TcPack out;
  size_t n_slices = 0;
  for (int l = 0; l < L; ++l) n_slices += Cfg::slices_in_layer(l, L);
  out.w.assign(n_slices * (SLICE_BYTES / 2), __float2half_rn(0.0f));
  out.bias.assign(size_t(L) * H, 0.0f);
  out.inv_scale.assign(L, 1.0f);
  size_t slice = 0;
  for (int l = 0; l < L; ++l) {
    const int in = dims[l], outd = dims[l + 1];
    const bool last = (l == L - 1);
    float wmax = 0.0f;
    for (size_t i = 0; i < size_t(in) * outd; ++i) wmax = std::max(wmax, std::fabs(W[l][i]));
    int e = 0;  // scale = 2^e, as large as keeps |w| * scale < 16384, at most 2^12
    if (wmax > 0.0f && std::isfinite(wmax)) e = std::min(12, std::max(-12, int(std::floor(std::log2(16384.0f / wmax)))));
    const float scale = std::ldexp(1.0f, e);
    out.inv_scale[l] = std::ldexp(1.0f, -e);
    const int rows = last ? nfb::M_LAST : H;
    for (int m = 0; m < rows; ++m) {
      const int src = last ? nfb::pack_last_layer_row(m) : (m < outd ? m : -1);
      if (src >= 0 && src < outd) out.bias[size_t(l) * H + m] = B[l][src];
    }
    const int nb_count = last ? Cfg::LAST_BLOCKS : 2;
    const int n_ks = Cfg::k_slices(l);
    for (int nb = 0; nb < nb_count; ++nb)  // streaming order of nfb_tc_kernel: N-block major, K-slices inside
      for (int ks = 0; ks < n_ks; ++ks)
        for (int part = 0; part < Cfg::SLOTS_PER_STEP; ++part, ++slice) {  // SPLIT = 3: the W_hi slice, then the W_lo slice
          __half* dst = out.w.data() + slice * (SLICE_BYTES / 2);
          for (int r = 0; r < SLICE_N; ++r) {
            const int m = nb * SLICE_N + r;
            const int src = last ? (m < nfb::M_LAST ? nfb::pack_last_layer_row(m) : -1) : (m < outd ? m : -1);
            if (src < 0 || src >= outd) continue;
            for (int kk = 0; kk < SLICE_K; ++kk) {
              // SPLIT = 1, layer 0: the k-slices are [W_hi | W_hi | W_lo] over the same 32 (25 used) inputs
              const int k = (Cfg::SPLIT == 1 && l == 0) ? kk : ks * SLICE_K + kk;
              if (k >= in) continue;
              const float ws = W[l][size_t(src) * in + k] * scale;
              const __half hi = __float2half_rn(ws);
              const bool want_lo = (Cfg::SPLIT == 3) ? (part == 1) : (l == 0 && ks == 2);
              const __half v = want_lo ? __float2half_rn(ws - __half2float(hi)) : hi;
              dst[(r / 8) * 256 + (kk / 8) * 64 + (r % 8) * 8 + (kk % 8)] = v;  // 8-row group: 4 k-chunks x 128 B
            }
          }
        }
  }
  return out;
}

// Weight streams of the CTA-pair kernels (nfb_tc2.cuh): for each CTA rank, the ring slots of one tile in the order the
// kernel consumes them -- layer, N-block, K sweep position (Cfg::k_slice_at) -- each position this rank's HALF_N rows of
// the W_hi slice (and of the W_lo slice when SPLIT = 3), K-major no-swizzle core matrices as above.  Same per-layer
// scale as pack_model_tc (whose biases and 1 / scale the kernels share).
template <class Cfg>
std::vector<__half> pack_model_tc2(int L, const int* dims, const float* const* W) {
  using nfb::tc::SLICE_K;
  constexpr int HALF_N = Cfg::HALF_N, BLOCK_N = Cfg::BLOCK_N, PARTS = (Cfg::SPLIT == 3) ? 2 : 1;
  size_t slots = 0;
  for (int l = 0; l < L; ++l) slots += Cfg::slots_in_layer(l, L);
  std::vector<__half> out(2 * slots * (Cfg::SLOT_BYTES / 2), __float2half_rn(0.0f));
  for (int rank = 0; rank < 2; ++rank) {
    size_t slot = 0;
    for (int l = 0; l < L; ++l) {
      const int in = dims[l], outd = dims[l + 1];
      const bool last = (l == L - 1);
      float wmax = 0.0f;
      for (size_t i = 0; i < size_t(in) * outd; ++i) wmax = std::max(wmax, std::fabs(W[l][i]));
      int e = 0;
      if (wmax > 0.0f && std::isfinite(wmax)) e = std::min(12, std::max(-12, int(std::floor(std::log2(16384.0f / wmax)))));
      const float scale = std::ldexp(1.0f, e);
      const int kt = Cfg::k_slices(l);
      for (int nb = 0; nb < Cfg::n_blocks(l, L); ++nb)
        for (int pos0 = 0; pos0 < kt; pos0 += Cfg::SPS, ++slot) {
          __half* const dst_slot = out.data() + (size_t(rank) * slots + slot) * (Cfg::SLOT_BYTES / 2);
          for (int u = 0; u < Cfg::SPS; ++u) {
            const int ks = Cfg::k_slice_at(pos0 + u, kt);
            for (int part = 0; part < PARTS; ++part) {
              __half* const dst = dst_slot + u * (Cfg::POS_BYTES / 2) + part * (Cfg::PIECE_BYTES / 2);
              for (int r = 0; r < HALF_N; ++r) {
                const int m = nb * BLOCK_N + rank * HALF_N + r;
                const int src = last ? (m < nfb::M_LAST ? nfb::pack_last_layer_row(m) : -1) : (m < outd ? m : -1);
                if (src < 0 || src >= outd) continue;
                for (int kk = 0; kk < SLICE_K; ++kk) {
                  // SPLIT = 1, layer 0: the four k-slices are [W_hi | W_hi | W_lo | 0] over the same 32 (25 used) inputs
                  const bool first_x1 = (Cfg::SPLIT == 1 && l == 0);
                  if (first_x1 && ks == 3) continue;
                  const int k = first_x1 ? kk : ks * SLICE_K + kk;
                  if (k >= in) continue;
                  const float ws = W[l][size_t(src) * in + k] * scale;
                  const __half hi = __float2half_rn(ws);
                  const bool want_lo = first_x1 ? (ks == 2) : (part == 1);
                  dst[(r / 8) * 256 + (kk / 8) * 64 + (r % 8) * 8 + (kk % 8)] = want_lo ? __float2half_rn(ws - __half2float(hi)) : hi;
                }
              }
            }
          }
        }
    }
  }
  return out;
}

// Weight streams of the narrow tensor-core kernels (nfb_tcn.cuh): the ring slots of one tile in the order the kernel consumes
// them -- layer, position of the K sweep -- each slot two pieces of (N_l rows x 32 k), K-major no-swizzle core matrices:
// SPLIT = 1: k-slices 2 pos and 2 pos + 1 (layer 0: [W_hi | W_hi | W_lo | 0] over the same 32 inputs); SPLIT = 3: the W_hi and
// the W_lo piece of k-slice pos.  Slot stride Cfg::SLOT_BYTES; the second piece follows the first at piece_bytes(l).
template <class Cfg>
std::vector<__half> pack_model_tcn(int L, const int* dims, const float* const* W, std::vector<float>* inv_scale_out) {
  using nfb::tc::SLICE_K;
  size_t slots = 0;
  for (int l = 0; l < L; ++l) slots += Cfg::slots_in_layer(l);
  std::vector<__half> out(slots * (Cfg::SLOT_BYTES / 2), __float2half_rn(0.0f));
  if (inv_scale_out) inv_scale_out->assign(L, 1.0f);
  size_t slot = 0;
  for (int l = 0; l < L; ++l) {
    const int in = dims[l], outd = dims[l + 1];
    const bool last = (l == L - 1);
    float wmax = 0.0f;
    for (size_t i = 0; i < size_t(in) * outd; ++i) wmax = std::max(wmax, std::fabs(W[l][i]));
    int e = 0;  // scale = 2^e, as large as keeps |w| * scale < 16384, at most 2^12 (as pack_model_tc)
    if (wmax > 0.0f && std::isfinite(wmax)) e = std::min(12, std::max(-12, int(std::floor(std::log2(16384.0f / wmax)))));
    const float scale = std::ldexp(1.0f, e);
    if (inv_scale_out) (*inv_scale_out)[l] = std::ldexp(1.0f, -e);
    const int n_rows = Cfg::n_of(l, L);
    const size_t piece_halfs = Cfg::piece_bytes(l, L) / 2;
    for (int pos = 0; pos < Cfg::slots_in_layer(l); ++pos, ++slot) {
      __half* const dst_slot = out.data() + slot * (Cfg::SLOT_BYTES / 2);
      for (int piece = 0; piece < 2; ++piece) {
        const int ks = (Cfg::SPLIT == 3) ? pos : 2 * pos + piece;
        bool want_lo = (Cfg::SPLIT == 3) ? (piece == 1) : false;
        bool zero = false;
        if (Cfg::SPLIT == 1 && l == 0) {  // [W_hi | W_hi | W_lo | 0]
          want_lo = (ks == 2);
          zero = (ks == 3);
        }
        if (zero) continue;
        __half* const dst = dst_slot + piece * piece_halfs;
        for (int r = 0; r < n_rows; ++r) {
          const int src = last ? nfb::pack_last_layer_row(r) : (r < outd ? r : -1);
          if (src < 0 || src >= outd) continue;
          for (int kk = 0; kk < SLICE_K; ++kk) {
            const int k = (Cfg::SPLIT == 1 && l == 0) ? kk : ks * SLICE_K + kk;
            if (k >= in) continue;
            const float ws = W[l][size_t(src) * in + k] * scale;
            const __half hi = __float2half_rn(ws);
            dst[(r / 8) * 256 + (kk / 8) * 64 + (r % 8) * 8 + (kk % 8)] = want_lo ? __float2half_rn(ws - __half2float(hi)) : hi;
          }
        }
      }
    }
  }
  return out;
}

// Raise a kernel's dynamic shared memory limit once per context (the attribute is per device and per function).
template <class K>
int configure(nfb_context* ctx, K kernel, size_t smem_bytes) {
  const void* key = reinterpret_cast<const void*>(kernel);
  if (ctx->configured.count(key)) return NFB_OK;
  NFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
  ctx->configured.insert(key);
  return NFB_OK;
}
#define NFB_CONFIGURE(kernel, bytes)                          \
  do {                                                        \
    int rc_ = configure(ctx, kernel, bytes);                  \
    if (rc_ != NFB_OK) return rc_;                            \
  } while (0)

template <class Cfg>
int launch_ffma(nfb_context* ctx, const nfb::EvalParams& p, cudaStream_t stream) {
  NFB_CONFIGURE(nfb::nfb_ffma_kernel<Cfg>, Cfg::SMEM_BYTES);
  const long long tiles = (p.n + Cfg::NQ - 1) / Cfg::NQ;
  const int grid = int(std::min<long long>(tiles, ctx->sm_count));
  nfb::nfb_ffma_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return NFB_OK;
}

template <class Cfg>
int launch_ffma_groups(nfb_context* ctx, const nfb::EvalParams& p, cudaStream_t stream) {
  NFB_CONFIGURE(nfb::nfb_ffma_groups_kernel<Cfg>, Cfg::SMEM_BYTES);
  const long long tiles = (p.n + Cfg::NQ - 1) / Cfg::NQ;
  const int grid = int(std::min<long long>(tiles, ctx->sm_count));
  nfb::nfb_ffma_groups_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return NFB_OK;
}

int ensure(void** ptr, size_t* have, size_t need, bool pinned_host);

template <class Cfg, bool FULL>
int launch_grad(nfb_context* ctx, nfb::GradParams& gp, cudaStream_t stream) {
  using GC = nfb::GradCfg<Cfg>;
  NFB_CONFIGURE((nfb::nfb_grad_kernel<Cfg, FULL>), GC::SMEM_BYTES);
  const long long tiles = (gp.ev.n + Cfg::NQ - 1) / Cfg::NQ;
  const int grid = int(std::min<long long>(tiles, ctx->sm_count));
  const size_t need = size_t(grid) * nfb::grad_scratch_floats<Cfg::H, Cfg::NC>(gp.ev.n_layers, FULL) * sizeof(float);
  // The scratch block belongs to the context, not to the stream: launches that use it are chained by an event, so two
  // streams of one context cannot overwrite each other's slopes; concurrency needs one context per stream.
  if (need > ctx->grad_scratch_bytes) {
    // Growing happens when a model or batch larger than any before comes along, never in steady state; cudaMalloc/cudaFree
    // cannot be captured in a graph, so warm the context up with one eager call of the same shape before capturing.
    if (ctx->grad_event_recorded) NFB_CUDA(cudaEventSynchronize(ctx->grad_event));  // nothing may still use the old block
    int rc = ensure(&ctx->grad_scratch, &ctx->grad_scratch_bytes, need, false);
    if (rc != NFB_OK) return rc;
    // parked rows that no kernel writes (packed rows 198..207 of the narrow models) must read as zero, never as NaN;
    // on the caller's stream, i.e. ordered before the kernel below
    NFB_CUDA(cudaMemsetAsync(ctx->grad_scratch, 0, ctx->grad_scratch_bytes, stream));
    ctx->grad_event_recorded = false;
  }
  if (!ctx->grad_event) NFB_CUDA(cudaEventCreateWithFlags(&ctx->grad_event, cudaEventDisableTiming));
  if (ctx->grad_event_recorded && ctx->grad_stream != stream) NFB_CUDA(cudaStreamWaitEvent(stream, ctx->grad_event, 0));
  gp.scratch = static_cast<float*>(ctx->grad_scratch);
  nfb::nfb_grad_kernel<Cfg, FULL><<<grid, Cfg::THREADS, GC::SMEM_BYTES, stream>>>(gp);
  NFB_CUDA(cudaGetLastError());
  NFB_CUDA(cudaEventRecord(ctx->grad_event, stream));
  ctx->grad_event_recorded = true;
  ctx->grad_stream = stream;
  ctx->launches += 1;
  return NFB_OK;
}

// Batches up to this size go to the small-batch kernel (4 queries per CTA); measured crossover, DESIGN.md section 6.
int64_t small_batch_limit(const nfb_context* ctx, int h_pad) {
  if (ctx->tune_small_batch_max >= 0) return ctx->tune_small_batch_max;  // nfb_set_tuning(NFB_TUNE_SMALL_BATCH_MAX)
  return h_pad >= 512 ? 512 : h_pad >= 256 ? 1024 : 2048;
}

template <class Cfg>
int launch_tc(nfb_context* ctx, const nfb::tc::TcParams& tp, cudaStream_t stream) {
  NFB_CONFIGURE(nfb::tc::nfb_tc_kernel<Cfg>, Cfg::SMEM_BYTES);
  const long long tiles = (tp.ev.n + Cfg::NQT - 1) / Cfg::NQT;
  const int grid = int(std::min<long long>((tiles + 1) & ~1LL, ctx->sm_count & ~1));  // whole CTA pairs
  nfb::tc::nfb_tc_kernel<Cfg><<<grid, nfb::tc::THREADS, Cfg::SMEM_BYTES, stream>>>(tp);
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return NFB_OK;
}

// the tensor-core kernels on CTA pairs (nfb_tc2.cuh)
template <class Cfg>
int launch_tc2(nfb_context* ctx, const nfb::tc::TcParams& tp, cudaStream_t stream) {
  NFB_CONFIGURE(nfb::tc2::nfb_tc2_kernel<Cfg>, Cfg::SMEM_BYTES);
  const long long tiles = (tp.ev.n + Cfg::NQT - 1) / Cfg::NQT;
  const int max_grid = ctx->tune_tc2_max_grid > 1 ? std::min(ctx->tune_tc2_max_grid, ctx->sm_count) : ctx->sm_count;  // NFB_TUNE_TC2_MAX_GRID
  const int grid = int(std::min<long long>((tiles + 1) & ~1LL, max_grid & ~1));  // whole CTA pairs
  nfb::tc2::nfb_tc2_kernel<Cfg><<<grid, nfb::tc2::THREADS2, Cfg::SMEM_BYTES, stream>>>(tp);
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return NFB_OK;
}

// the narrow tensor-core kernels (nfb_tcn.cuh): one CTA per SM
template <class Cfg>
int launch_tcn(nfb_context* ctx, const nfb::tc::TcParams& tp, cudaStream_t stream) {
  NFB_CONFIGURE(nfb::tcn::nfb_tcn_kernel<Cfg>, Cfg::SMEM_BYTES);
  const long long tiles = (tp.ev.n + Cfg::NQT - 1) / Cfg::NQT;
  const int grid = int(std::min<long long>(tiles, ctx->sm_count));
  nfb::tcn::nfb_tcn_kernel<Cfg><<<grid, nfb::tcn::TCN_THREADS, Cfg::SMEM_BYTES, stream>>>(tp);
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return NFB_OK;
}

template <int H>
int launch_small(nfb_context* ctx, const nfb::EvalParams& p, cudaStream_t stream) {
  using Cfg = nfb::SmallCfg<H>;
  const int grid = int((p.n + nfb::SMALL_QUERIES - 1) / nfb::SMALL_QUERIES);
  nfb::nfb_small_kernel<H><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return NFB_OK;
}

template <int H>
int launch_jac(nfb_context* ctx, const nfb::JacParams& jp, cudaStream_t stream) {
  using Cfg = nfb::JacCfg<H>;
  NFB_CONFIGURE(nfb::nfb_jac_kernel<H>, Cfg::SMEM_BYTES);
  nfb::nfb_jac_kernel<H><<<unsigned(jp.ev.n), Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(jp);  // one CTA per query
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return NFB_OK;
}

int check_eval_args(nfb_context* ctx, int model_id, int variant, int64_t n, const void* const* in_rows,
                    const int64_t* in_strides, int in_dtype, void* out, int64_t out_ld, int out_dtype) {
  if (!ctx) return fail(NFB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (model_id < 0 || model_id >= NFB_MAX_MODELS)
    return fail(NFB_ERR_INVALID_ARGUMENT, "model_id %d out of range [0, %d)", model_id, NFB_MAX_MODELS);
  if (!ctx->models[model_id].loaded) return fail(NFB_ERR_NOT_LOADED, "model slot %d is empty", model_id);
  if (!ctx->have_distribution)
    return fail(NFB_ERR_NOT_LOADED, "nfb_set_distribution has not been called");
  variant &= ~NFB_VARIANT_FLAGS;
  if (variant != NFB_VARIANT_FP32_FFMA && variant != NFB_VARIANT_TENSOR_FP16 && variant != NFB_VARIANT_TENSOR_FP16X3)
    return fail(NFB_ERR_INVALID_ARGUMENT, "unknown variant %d", variant);
  if (variant != NFB_VARIANT_FP32_FFMA && !ctx->models[model_id].tc_w && !ctx->models[model_id].tcn_w)
    return fail(NFB_ERR_UNSUPPORTED_MODEL,
                "the tensor-core variants exist for models with at most %d layers only; model slot %d has more",
                nfb::tc::MAX_TC_LAYERS, model_id);
  if (n < 0) return fail(NFB_ERR_INVALID_ARGUMENT, "n = %lld is negative", (long long)n);
  if ((in_dtype != NFB_F32 && in_dtype != NFB_F64) || (out_dtype != NFB_F32 && out_dtype != NFB_F64))
    return fail(NFB_ERR_INVALID_ARGUMENT, "dtype must be NFB_F32 or NFB_F64");
  if (n == 0) return NFB_OK;
  if (!in_rows || !in_strides || !out) return fail(NFB_ERR_INVALID_ARGUMENT, "NULL data pointer");
  if (out_ld < n) return fail(NFB_ERR_INVALID_ARGUMENT, "out_ld (%lld) < n (%lld)", (long long)out_ld, (long long)n);
  for (int i = 0; i < NFB_N_RAW_ROWS; ++i) {
    if (!in_rows[i]) return fail(NFB_ERR_INVALID_ARGUMENT, "in_rows[%d] is NULL", i);
    if (in_strides[i] < 0) return fail(NFB_ERR_INVALID_ARGUMENT, "in_strides[%d] is negative", i);
  }
  return NFB_OK;
}

// What every kernel's parameters share: the query rows, the model and the context's distribution statistics.
void fill_eval_params(const nfb_context* ctx, const Model& m, const float* blob, int64_t n, const void* const* in_rows,
                      const int64_t* in_strides, int in_dtype, void* out, int64_t out_ld, int out_dtype, nfb::EvalParams& p) {
  for (int i = 0; i < nfb::N_RAW; ++i) {
    p.in[i] = in_rows[i];
    p.in_stride[i] = in_strides[i];
  }
  p.out = out;
  p.out_ld = out_ld;
  p.n = n;
  p.blob = blob;
  p.n_layers = m.n_layers;
  p.k_hidden = m.k_hidden;
  p.in_f64 = (in_dtype == NFB_F64);
  p.out_f64 = (out_dtype == NFB_F64);
  p.scalars_only = 0;
  p.out_bl = nullptr;
  memcpy(p.mean, ctx->mean, sizeof(p.mean));
  memcpy(p.quad, ctx->quad, sizeof(p.quad));
}

int eval_device_impl(nfb_context* ctx, int model_id, int variant, int64_t n, const void* const* in_rows,
                     const int64_t* in_strides, int in_dtype, void* out, int64_t out_ld, int out_dtype,
                     cudaStream_t stream, void* out_bl = nullptr) {
  const Model& m = ctx->models[model_id];
  nfb::EvalParams p;
  fill_eval_params(ctx, m, m.blob, n, in_rows, in_strides, in_dtype, out, out_ld, out_dtype, p);
  p.out_bl = out_bl;
  p.scalars_only = (variant & NFB_OUTPUT_SCALARS) ? 1 : 0;
  const bool one_sm = (variant & NFB_DEBUG_TC_ONE_SM) != 0;      // cross-check: the one-SM kernels of nfb_tc.cuh
  const bool tiled_only = (variant & NFB_DEBUG_FFMA_TILED) != 0;  // cross-check: the CTA-synchronous kernel, every batch size
  variant &= ~NFB_VARIANT_FLAGS;
  if (variant == NFB_VARIANT_TENSOR_FP16 || variant == NFB_VARIANT_TENSOR_FP16X3) {
    const bool x3 = (variant == NFB_VARIANT_TENSOR_FP16X3);
    nfb::tc::TcParams tp;
    tp.ev = p;
    if (m.tcn_w) {  // 64- / 128-wide models: nfb_tcn.cuh (a layer is one MMA wide; there is no second implementation to cross-check)
      tp.wblob = x3 ? m.tcn_w3 : m.tcn_w;
      tp.bias = m.tcn_bias;
      tp.inv_scale = m.tc_inv_scale;
      tp.dbg = ctx->tc_debug;
      if (m.h_pad == 128)
        return x3 ? launch_tcn<nfb::tcn::TcnCfg<128, 3>>(ctx, tp, stream) : launch_tcn<nfb::tcn::TcnCfg<128, 1>>(ctx, tp, stream);
      return x3 ? launch_tcn<nfb::tcn::TcnCfg<64, 3>>(ctx, tp, stream) : launch_tcn<nfb::tcn::TcnCfg<64, 1>>(ctx, tp, stream);
    }
    tp.wblob = one_sm ? (x3 ? m.tc3_w : m.tc_w) : (x3 ? m.tc2_w3 : m.tc2_w);
    tp.bias = m.tc_bias;
    tp.inv_scale = m.tc_inv_scale;
    tp.dbg = ctx->tc_debug;
    if (!one_sm) {
      if (m.h_pad == 512)
        return x3 ? launch_tc2<nfb::tc2::Tc2Cfg<512, 3>>(ctx, tp, stream) : launch_tc2<nfb::tc2::Tc2Cfg<512, 1>>(ctx, tp, stream);
      return x3 ? launch_tc2<nfb::tc2::Tc2Cfg<256, 3>>(ctx, tp, stream) : launch_tc2<nfb::tc2::Tc2Cfg<256, 1>>(ctx, tp, stream);
    }
    if (m.h_pad == 512)
      return x3 ? launch_tc<nfb::tc::TcCfg<512, 3>>(ctx, tp, stream) : launch_tc<nfb::tc::TcCfg<512, 1>>(ctx, tp, stream);
    return x3 ? launch_tc<nfb::tc::TcCfg<256, 3>>(ctx, tp, stream) : launch_tc<nfb::tc::TcCfg<256, 1>>(ctx, tp, stream);
  }
  if (!tiled_only && n <= small_batch_limit(ctx, m.h_pad)) {
    switch (m.h_pad) {
      case 64: return launch_small<64>(ctx, p, stream);
      case 128: return launch_small<128>(ctx, p, stream);
      case 256: return launch_small<256>(ctx, p, stream);
      case 512: return launch_small<512>(ctx, p, stream);
    }
  }
  if (!tiled_only) {
    switch (m.h_pad) {
      case 64: return launch_ffma_groups<Grp64>(ctx, p, stream);
      case 128: return launch_ffma_groups<Grp128>(ctx, p, stream);
      case 256: return launch_ffma_groups<Grp256>(ctx, p, stream);
      case 512: return launch_ffma_groups<Grp512>(ctx, p, stream);
    }
  }
  switch (m.h_pad) {
    case 64: return launch_ffma<Cfg64>(ctx, p, stream);
    case 128: return launch_ffma<Cfg128>(ctx, p, stream);
    case 256: return launch_ffma<Cfg256>(ctx, p, stream);
    case 512: return launch_ffma<Cfg512>(ctx, p, stream);
  }
  return fail(NFB_ERR_UNSUPPORTED_MODEL, "no kernel for padded width %d", m.h_pad);
}

// ---- host-side staging for pageable caller memory (nfb_eval_host, large batches) ---------------------------------
// cudaMemcpy from / to pageable memory goes through the driver's own bounce buffers on one thread: 4 GB/s measured
// for the 792 MB of a 1 M-query float32 block, 315 ms for the float64 dict of the drop-in call, whatever the model
// size.  Instead the library copies between the caller's arrays and its own pinned buffers with every host core
// (first-touch page faults of a freshly allocated result included) and lets the GPU write float32 even when the
// caller wants float64: the widening happens in that copy, and half the bytes cross PCIe.
int host_threads() {
  static const int n = std::max(1, std::min(int(std::thread::hardware_concurrency()), 32));
  return n;
}

template <class F>
void parallel_units(int64_t units, F&& f, int max_threads = 1 << 30) {  // f(unit) for unit in [0, units), dynamically scheduled
  const int nt = int(std::min<int64_t>(std::min(host_threads(), std::max(1, max_threads)), units));
  if (nt <= 1) {
    for (int64_t u = 0; u < units; ++u) f(u);
    return;
  }
  std::atomic<int64_t> next{0};
  auto work = [&] {
    for (int64_t u = next.fetch_add(1); u < units; u = next.fetch_add(1)) f(u);
  };
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
}

bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

constexpr int64_t STAGE_BLOCK = 1 << 16;  // elements per copy unit

// rows x cn block of float32 in pinned memory (row stride src_ld) -> the caller's block (row stride dst_ld, float32 or float64)
void unstage_output(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, bool dst_f64, int rows, int64_t cn, int thread_cap) {
  const int64_t blocks = (cn + STAGE_BLOCK - 1) / STAGE_BLOCK;
  const int max_threads = std::min(thread_cap, int(rows * cn / (1 << 19)) + 1);  // a thread costs ~30 us to start: one per 512 Ki elements
  parallel_units(rows * blocks, [&](int64_t u) {
    const int64_t r = u / blocks, i0 = (u % blocks) * STAGE_BLOCK, i1 = std::min(cn, i0 + STAGE_BLOCK);
    const float* sp = src + r * src_ld;
    if (dst_f64) {
      double* dp = static_cast<double*>(dst) + r * dst_ld;
      for (int64_t i = i0; i < i1; ++i) dp[i] = double(sp[i]);
    } else {
      memcpy(static_cast<float*>(dst) + r * dst_ld + i0, sp + i0, size_t(i1 - i0) * sizeof(float));
    }
  }, max_threads);
}

int ensure(void** ptr, size_t* have, size_t need, bool pinned_host) {
  if (*have >= need) return NFB_OK;
  if (*ptr) {
    if (pinned_host) NFB_CUDA(cudaFreeHost(*ptr));
    else NFB_CUDA(cudaFree(*ptr));
    *ptr = nullptr;
    *have = 0;
  }
  const size_t want = need + need / 4;
  if (pinned_host) NFB_CUDA(cudaMallocHost(ptr, want));
  else NFB_CUDA(cudaMalloc(ptr, want));
  *have = want;
  return NFB_OK;
}

}  // namespace

extern "C" {

const char* nfb_version(void) { return "neuralfoil_b200 " NFB_VERSION " (sm_100a; fused evaluator: fp32 FFMA2 column groups, tcgen05 cta_group::2 fp16-operand variants, forward-mode Jacobian, fused reverse pass)"; }

const char* nfb_last_error(void) { return g_err; }

int nfb_create(int device, nfb_context** out_ctx) {
  if (!out_ctx) return fail(NFB_ERR_INVALID_ARGUMENT, "out_ctx is NULL");
  *out_ctx = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(NFB_ERR_NO_DEVICE, "no CUDA device visible (%s); neuralfoil_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return fail(NFB_ERR_INVALID_ARGUMENT, "device %d out of range [0, %d)", device, count);
  cudaDeviceProp prop;
  NFB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(NFB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  NFB_CUDA(cudaSetDevice(device));
  nfb_context* ctx = new nfb_context();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  for (auto& s : ctx->slots) {
    cudaError_t se = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) {
      delete ctx;
      return fail(NFB_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(se));
    }
  }
  *out_ctx = ctx;
  return NFB_OK;
}

void nfb_destroy(nfb_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto& m : ctx->models) {
    if (m.blob) cudaFree(m.blob);
    if (m.grad_blob) cudaFree(m.grad_blob);
    if (m.tc_w) cudaFree(m.tc_w);
    if (m.tc3_w) cudaFree(m.tc3_w);
    if (m.tc2_w) cudaFree(m.tc2_w);
    if (m.tc2_w3) cudaFree(m.tc2_w3);
    if (m.tc_bias) cudaFree(m.tc_bias);
    if (m.tc_inv_scale) cudaFree(m.tc_inv_scale);
    if (m.tcn_w) cudaFree(m.tcn_w);
    if (m.tcn_w3) cudaFree(m.tcn_w3);
    if (m.tcn_bias) cudaFree(m.tcn_bias);
  }
  for (auto& s : ctx->slots) {
    if (s.stream) {
      cudaStreamSynchronize(s.stream);
      cudaStreamDestroy(s.stream);
    }
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.h_stage) cudaFreeHost(s.h_stage);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_out) cudaFreeHost(s.h_out);
  }
  if (ctx->tc_debug) cudaFree(ctx->tc_debug);
  if (ctx->jac_dev) cudaFree(ctx->jac_dev);
  if (ctx->grad_scratch) cudaFree(ctx->grad_scratch);
  if (ctx->grad_event) cudaEventDestroy(ctx->grad_event);
  if (ctx->train_ws) cudaFree(ctx->train_ws);
  if (ctx->packed_row_of) cudaFree(ctx->packed_row_of);
  if (ctx->jac_host) cudaFreeHost(ctx->jac_host);
  delete ctx;
}

int nfb_device_info(const nfb_context* ctx, int* device, int* sm_count, int* cc_major, int* cc_minor) {
  if (!ctx) return fail(NFB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (device) *device = ctx->device;
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  return NFB_OK;
}

int nfb_set_distribution(nfb_context* ctx, const double* mean25, const double* inv_cov) {
  if (!ctx || !mean25 || !inv_cov) return fail(NFB_ERR_INVALID_ARGUMENT, "NULL argument");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int idx = 0;
  for (int i = 0; i < nfb::N_IN; ++i)
    for (int j = i; j < nfb::N_IN; ++j)
      ctx->quad[idx++] = (i == j) ? inv_cov[i * nfb::N_IN + i] : inv_cov[i * nfb::N_IN + j] + inv_cov[j * nfb::N_IN + i];
  memcpy(ctx->mean, mean25, sizeof(ctx->mean));
  ctx->have_distribution = true;
  return NFB_OK;
}

int nfb_load_model(nfb_context* ctx, int model_id, int n_layers, const int* dims, const float* const* weights,
                   const float* const* biases) {
  if (!ctx || !dims || !weights || !biases) return fail(NFB_ERR_INVALID_ARGUMENT, "NULL argument");
  if (model_id < 0 || model_id >= NFB_MAX_MODELS)
    return fail(NFB_ERR_INVALID_ARGUMENT, "model_id %d out of range [0, %d)", model_id, NFB_MAX_MODELS);
  if (n_layers < 2 || n_layers > NFB_MAX_LAYERS)
    return fail(NFB_ERR_UNSUPPORTED_MODEL, "n_layers = %d; supported: 2..%d", n_layers, NFB_MAX_LAYERS);
  if (dims[0] != NFB_N_LATENT_INPUTS || dims[n_layers] != NFB_N_OUTPUT_ROWS)
    return fail(NFB_ERR_UNSUPPORTED_MODEL, "network must map %d inputs to %d outputs, got %d -> %d",
                NFB_N_LATENT_INPUTS, NFB_N_OUTPUT_ROWS, dims[0], dims[n_layers]);
  int hmax = 0;
  for (int l = 1; l < n_layers; ++l) {
    if (dims[l] < 1) return fail(NFB_ERR_UNSUPPORTED_MODEL, "hidden width %d", dims[l]);
    hmax = std::max(hmax, dims[l]);
  }
  for (int l = 0; l < n_layers; ++l)
    if (!weights[l] || !biases[l]) return fail(NFB_ERR_INVALID_ARGUMENT, "layer %d has a NULL array", l);
  const int h_pad = hmax <= 64 ? 64 : hmax <= 128 ? 128 : hmax <= 256 ? 256 : hmax <= 512 ? 512 : 0;
  if (!h_pad) return fail(NFB_ERR_UNSUPPORTED_MODEL, "hidden width %d > 512 has no kernel", hmax);

  std::vector<float> blob;
  switch (h_pad) {
    case 64: blob = pack_model<64>(n_layers, dims, weights, biases); break;
    case 128: blob = pack_model<128>(n_layers, dims, weights, biases); break;
    case 256: blob = pack_model<256>(n_layers, dims, weights, biases); break;
    default: blob = pack_model<512>(n_layers, dims, weights, biases); break;
  }
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  Model& m = ctx->models[model_id];
  if (m.blob) {
    NFB_CUDA(cudaDeviceSynchronize());
    NFB_CUDA(cudaFree(m.blob));
    if (m.grad_blob) NFB_CUDA(cudaFree(m.grad_blob));
    if (m.tc_w) NFB_CUDA(cudaFree(m.tc_w));
    if (m.tc3_w) NFB_CUDA(cudaFree(m.tc3_w));
    if (m.tc2_w) NFB_CUDA(cudaFree(m.tc2_w));
    if (m.tc2_w3) NFB_CUDA(cudaFree(m.tc2_w3));
    if (m.tc_bias) NFB_CUDA(cudaFree(m.tc_bias));
    if (m.tc_inv_scale) NFB_CUDA(cudaFree(m.tc_inv_scale));
    if (m.tcn_w) NFB_CUDA(cudaFree(m.tcn_w));
    if (m.tcn_w3) NFB_CUDA(cudaFree(m.tcn_w3));
    if (m.tcn_bias) NFB_CUDA(cudaFree(m.tcn_bias));
    m = Model();
  }
  NFB_CUDA(cudaMalloc(&m.blob, blob.size() * sizeof(float)));
  NFB_CUDA(cudaMemcpy(m.blob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  {
    std::vector<float> gblob;
    switch (h_pad) {
      case 64: gblob = pack_grad_blob<64>(n_layers, dims, weights, biases, blob); break;
      case 128: gblob = pack_grad_blob<128>(n_layers, dims, weights, biases, blob); break;
      case 256: gblob = pack_grad_blob<256>(n_layers, dims, weights, biases, blob); break;
      default: gblob = pack_grad_blob<512>(n_layers, dims, weights, biases, blob); break;
    }
    NFB_CUDA(cudaMalloc(&m.grad_blob, gblob.size() * sizeof(float)));
    NFB_CUDA(cudaMemcpy(m.grad_blob, gblob.data(), gblob.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if ((h_pad == 512 || h_pad == 256) && n_layers <= nfb::tc::MAX_TC_LAYERS) {
    const TcPack tcp = h_pad == 512 ? pack_model_tc<nfb::tc::TcCfg<512, 1>>(n_layers, dims, weights, biases)
                                    : pack_model_tc<nfb::tc::TcCfg<256, 1>>(n_layers, dims, weights, biases);
    const TcPack tcp3 = h_pad == 512 ? pack_model_tc<nfb::tc::TcCfg<512, 3>>(n_layers, dims, weights, biases)
                                     : pack_model_tc<nfb::tc::TcCfg<256, 3>>(n_layers, dims, weights, biases);
    const std::vector<__half> p2 = h_pad == 512 ? pack_model_tc2<nfb::tc2::Tc2Cfg<512, 1>>(n_layers, dims, weights)
                                                : pack_model_tc2<nfb::tc2::Tc2Cfg<256, 1>>(n_layers, dims, weights);
    const std::vector<__half> p23 = h_pad == 512 ? pack_model_tc2<nfb::tc2::Tc2Cfg<512, 3>>(n_layers, dims, weights)
                                                 : pack_model_tc2<nfb::tc2::Tc2Cfg<256, 3>>(n_layers, dims, weights);
    NFB_CUDA(cudaMalloc(&m.tc2_w, p2.size() * sizeof(__half)));
    NFB_CUDA(cudaMemcpy(m.tc2_w, p2.data(), p2.size() * sizeof(__half), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMalloc(&m.tc2_w3, p23.size() * sizeof(__half)));
    NFB_CUDA(cudaMemcpy(m.tc2_w3, p23.data(), p23.size() * sizeof(__half), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMalloc(&m.tc3_w, tcp3.w.size() * sizeof(__half)));
    NFB_CUDA(cudaMemcpy(m.tc3_w, tcp3.w.data(), tcp3.w.size() * sizeof(__half), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMalloc(&m.tc_w, tcp.w.size() * sizeof(__half)));
    NFB_CUDA(cudaMalloc(&m.tc_bias, tcp.bias.size() * sizeof(float)));
    NFB_CUDA(cudaMalloc(&m.tc_inv_scale, tcp.inv_scale.size() * sizeof(float)));
    NFB_CUDA(cudaMemcpy(m.tc_w, tcp.w.data(), tcp.w.size() * sizeof(__half), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMemcpy(m.tc_bias, tcp.bias.data(), tcp.bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMemcpy(m.tc_inv_scale, tcp.inv_scale.data(), tcp.inv_scale.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if ((h_pad == 128 || h_pad == 64) && n_layers <= nfb::tc::MAX_TC_LAYERS) {
    std::vector<float> inv_scale;
    const std::vector<__half> w1 = h_pad == 128 ? pack_model_tcn<nfb::tcn::TcnCfg<128, 1>>(n_layers, dims, weights, &inv_scale)
                                                : pack_model_tcn<nfb::tcn::TcnCfg<64, 1>>(n_layers, dims, weights, &inv_scale);
    const std::vector<__half> w3 = h_pad == 128 ? pack_model_tcn<nfb::tcn::TcnCfg<128, 3>>(n_layers, dims, weights, nullptr)
                                                : pack_model_tcn<nfb::tcn::TcnCfg<64, 3>>(n_layers, dims, weights, nullptr);
    std::vector<float> bias(size_t(n_layers) * nfb::tcn::TCN_BIAS_LD, 0.0f);
    for (int l = 0; l < n_layers; ++l) {
      const bool last = (l == n_layers - 1);
      const int rows = last ? nfb::tcn::TCN_LAST_N : h_pad;
      for (int r = 0; r < rows; ++r) {
        const int src = last ? nfb::pack_last_layer_row(r) : (r < dims[l + 1] ? r : -1);
        if (src >= 0 && src < dims[l + 1]) bias[size_t(l) * nfb::tcn::TCN_BIAS_LD + r] = biases[l][src];
      }
    }
    NFB_CUDA(cudaMalloc(&m.tcn_w, w1.size() * sizeof(__half)));
    NFB_CUDA(cudaMemcpy(m.tcn_w, w1.data(), w1.size() * sizeof(__half), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMalloc(&m.tcn_w3, w3.size() * sizeof(__half)));
    NFB_CUDA(cudaMemcpy(m.tcn_w3, w3.data(), w3.size() * sizeof(__half), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMalloc(&m.tcn_bias, bias.size() * sizeof(float)));
    NFB_CUDA(cudaMemcpy(m.tcn_bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    NFB_CUDA(cudaMalloc(&m.tc_inv_scale, inv_scale.size() * sizeof(float)));
    NFB_CUDA(cudaMemcpy(m.tc_inv_scale, inv_scale.data(), inv_scale.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  m.loaded = true;
  m.n_layers = n_layers;
  m.k_hidden = std::min(h_pad, (hmax + 15) / 16 * 16);
  m.h_pad = h_pad;
  std::copy(dims, dims + n_layers + 1, m.dims);
  return NFB_OK;
}

int nfb_eval_device(nfb_context* ctx, int model_id, int variant, int64_t n, const void* const* in_rows,
                    const int64_t* in_strides, int in_dtype, void* out, int64_t out_ld, int out_dtype,
                    void* stream) {
  int rc = check_eval_args(ctx, model_id, variant, n, in_rows, in_strides, in_dtype, out, out_ld, out_dtype);
  if (rc != NFB_OK || n == 0) return rc;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  return eval_device_impl(ctx, model_id, variant, n, in_rows, in_strides, in_dtype, out, out_ld, out_dtype,
                          static_cast<cudaStream_t>(stream));
}

namespace {
inline int64_t default_chunk(int64_t n) {  // a sixteenth of the batch within [64 Ki, 512 Ki] queries
  return std::min<int64_t>(1 << 19, std::max<int64_t>(1 << 16, (n / 16 + 3) & ~int64_t(3)));
}
}  // namespace

int nfb_eval_host(nfb_context* ctx, int model_id, int variant, int64_t n, const void* const* in_rows,
                  const int64_t* in_strides, int in_dtype, void* out, int64_t out_ld, int out_dtype,
                  int64_t chunk) {
  int rc = check_eval_args(ctx, model_id, variant, n, in_rows, in_strides, in_dtype, out, out_ld, out_dtype);
  if (rc != NFB_OK || n == 0) return rc;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);  // the staging slots below belong to the context
  NFB_CUDA(cudaSetDevice(ctx->device));
  const int thread_cap = ctx->host_thread_cap > 0 ? ctx->host_thread_cap : (1 << 30);
  const size_t isz = in_dtype == NFB_F64 ? 8 : 4, osz = out_dtype == NFB_F64 ? 8 : 4;
  // Small batches (a design optimiser's per-iteration call): pack everything into one pinned block so
  // that there is exactly one H2D copy, one launch and one D2H copy.
  const bool small = n <= 16384;
  // The smallest batches (an optimiser's single query): the kernel reads its inputs from and writes its results to the pinned
  // staging block directly (zero-copy over PCIe: pinned host memory is device-addressable under UVA) -- one launch and one
  // synchronisation, no copy commands at all.
  const bool mapped = small && n <= (ctx->tune_mapped_max >= 0 ? ctx->tune_mapped_max : 8);  // measured (us, zero-copy vs copies): n = 1: 32 vs 37, 64: 51 vs 45, 1024: 188 vs 189
  // Large batches in pageable memory (NumPy arrays): staged through pinned buffers by all host cores, see host_threads().
  const bool stage_out = !small && is_pageable(out);
  bool stage_in = false;
  if (!small)
    for (int r = 0; r < nfb::N_RAW; ++r)
      if (in_strides[r] != 0) {
        stage_in = is_pageable(in_rows[r]);
        break;
      }
  // default chunk: 128 Ki queries when the host copies are the library's own, so that they overlap the GPU sooner; otherwise a
  // sixteenth of the batch within [64 Ki, 512 Ki]: the last chunk's device->host copy is the one nothing hides, and when eight
  // GPUs share the host's 92 GB/s (profiles/r2_pcie_pitched_8gpu.jsonl) a 512 Ki chunk of full outputs is a 36 ms tail on a
  // 115 ms shard (8 GPUs x 1.25 M queries, xxxlarge fp32: 148 ms per pass with 512 Ki chunks, 123 ms with 128 Ki)
  if (chunk <= 0) chunk = stage_out ? (1 << 17) : default_chunk(n);
  chunk = std::min<int64_t>((chunk + 3) & ~int64_t(3), (n + 3) & ~int64_t(3));
  const int out_rows = (variant & NFB_OUTPUT_SCALARS) ? 6 : nfb::N_OUT;
  const size_t in_need = size_t(nfb::N_RAW) * chunk * isz, out_need = size_t(out_rows) * chunk * osz;
  // staged: the GPU writes float32, the copy widens.  Same for float64 results of the larger small batches (their rows
  // are copied out of the pinned block anyway): half the D2H bytes.
  const bool small_widen = small && out_dtype == NFB_F64 && n >= 2048;
  const int dev_out_dtype = (stage_out || small_widen) ? NFB_F32 : out_dtype;
  const size_t dev_osz = dev_out_dtype == NFB_F64 ? 8 : 4;
  const size_t dev_out_need = size_t(out_rows) * chunk * dev_osz;
  struct Landed { HostSlot* slot = nullptr; int64_t c0 = 0, cn = 0; } landed;  // staged chunk whose D2H is in flight
  auto finish_landed = [&]() -> int {
    if (!landed.slot) return NFB_OK;
    NFB_CUDA(cudaStreamSynchronize(landed.slot->stream));
    unstage_output(static_cast<const float*>(landed.slot->h_out), chunk, static_cast<char*>(out) + size_t(landed.c0) * osz,
                   out_ld, out_dtype == NFB_F64, out_rows, landed.cn, thread_cap);
    landed.slot = nullptr;
    return NFB_OK;
  };

  for (int64_t c0 = 0, ci = 0; c0 < n; c0 += chunk, ++ci) {
    HostSlot& s = ctx->slots[ci & 1];
    const int64_t cn = std::min<int64_t>(chunk, n - c0);
    NFB_CUDA(cudaStreamSynchronize(s.stream));  // the slot's previous chunk has fully landed
    if (!mapped) {
      if ((rc = ensure(&s.d_in, &s.in_bytes, in_need, false)) != NFB_OK) return rc;
      if ((rc = ensure(&s.d_out, &s.out_bytes, dev_out_need, false)) != NFB_OK) return rc;
    }

    const void* d_rows[nfb::N_RAW];
    int64_t d_strides[nfb::N_RAW];
    if (small) {
      if ((rc = ensure(&s.h_stage, &s.stage_bytes, in_need + out_need, true)) != NFB_OK) return rc;
      char* hs = static_cast<char*>(s.h_stage);
      for (int r = 0; r < nfb::N_RAW; ++r) {
        char* dst = hs + size_t(r) * chunk * isz;
        const char* src = static_cast<const char*>(in_rows[r]);
        if (in_strides[r] == 1) {
          memcpy(dst, src + size_t(c0) * isz, size_t(cn) * isz);
          d_strides[r] = 1;
        } else if (in_strides[r] == 0) {
          memcpy(dst, src, isz);
          d_strides[r] = 0;
        } else {
          for (int64_t i = 0; i < cn; ++i) memcpy(dst + i * isz, src + size_t(c0 + i) * in_strides[r] * isz, isz);
          d_strides[r] = 1;
        }
        d_rows[r] = (mapped ? hs : static_cast<char*>(s.d_in)) + size_t(r) * chunk * isz;
      }
      if (!mapped) NFB_CUDA(cudaMemcpyAsync(s.d_in, hs, in_need, cudaMemcpyHostToDevice, s.stream));
    } else if (stage_in) {
      if ((rc = ensure(&s.h_in, &s.h_in_bytes, in_need, true)) != NFB_OK) return rc;
      char* hs = static_cast<char*>(s.h_in);
      const int64_t blocks = (cn + STAGE_BLOCK - 1) / STAGE_BLOCK;
      parallel_units(nfb::N_RAW * blocks, [&](int64_t u) {
        const int64_t r = u / blocks, i0 = (u % blocks) * STAGE_BLOCK, i1 = std::min(cn, i0 + STAGE_BLOCK);
        char* dst = hs + size_t(r) * chunk * isz;
        const char* src = static_cast<const char*>(in_rows[r]);
        if (in_strides[r] == 1) {
          memcpy(dst + size_t(i0) * isz, src + size_t(c0 + i0) * isz, size_t(i1 - i0) * isz);
        } else if (in_strides[r] == 0) {
          if (i0 == 0) memcpy(dst, src, isz);
        } else {
          for (int64_t i = i0; i < i1; ++i) memcpy(dst + i * isz, src + size_t(c0 + i) * in_strides[r] * isz, isz);
        }
      }, thread_cap);
      for (int r = 0; r < nfb::N_RAW; ++r) {
        d_rows[r] = static_cast<char*>(s.d_in) + size_t(r) * chunk * isz;
        d_strides[r] = in_strides[r] == 0 ? 0 : 1;
      }
      NFB_CUDA(cudaMemcpyAsync(s.d_in, hs, in_need, cudaMemcpyHostToDevice, s.stream));
    } else {
      for (int r = 0; r < nfb::N_RAW; ++r) {
        char* dst = static_cast<char*>(s.d_in) + size_t(r) * chunk * isz;
        const char* src = static_cast<const char*>(in_rows[r]);
        if (in_strides[r] == 0) {
          NFB_CUDA(cudaMemcpyAsync(dst, src, isz, cudaMemcpyHostToDevice, s.stream));
          d_strides[r] = 0;
        } else if (in_strides[r] == 1) {
          NFB_CUDA(cudaMemcpyAsync(dst, src + size_t(c0) * isz, size_t(cn) * isz, cudaMemcpyHostToDevice, s.stream));
          d_strides[r] = 1;
        } else {
          NFB_CUDA(cudaMemcpy2DAsync(dst, isz, src + size_t(c0) * in_strides[r] * isz, size_t(in_strides[r]) * isz,
                                     isz, size_t(cn), cudaMemcpyHostToDevice, s.stream));
          d_strides[r] = 1;
        }
        d_rows[r] = dst;
      }
    }
    rc = eval_device_impl(ctx, model_id, variant, cn, d_rows, d_strides, in_dtype,
                          mapped ? static_cast<void*>(static_cast<char*>(s.h_stage) + in_need) : s.d_out, chunk, dev_out_dtype, s.stream);
    if (rc != NFB_OK) return rc;
    char* out_c = static_cast<char*>(out) + size_t(c0) * osz;
    if (small) {
      char* hs_out = static_cast<char*>(s.h_stage) + in_need;
      if (!mapped) NFB_CUDA(cudaMemcpyAsync(hs_out, s.d_out, dev_out_need, cudaMemcpyDeviceToHost, s.stream));
      NFB_CUDA(cudaStreamSynchronize(s.stream));
      if (small_widen) {
        unstage_output(reinterpret_cast<const float*>(hs_out), chunk, out_c, out_ld, true, out_rows, cn, thread_cap);
      } else {
        for (int r = 0; r < out_rows; ++r)
          memcpy(out_c + size_t(r) * out_ld * osz, hs_out + size_t(r) * chunk * osz, size_t(cn) * osz);
      }
    } else if (stage_out) {
      if ((rc = ensure(&s.h_out, &s.h_out_bytes, dev_out_need, true)) != NFB_OK) return rc;
      NFB_CUDA(cudaMemcpyAsync(s.h_out, s.d_out, dev_out_need, cudaMemcpyDeviceToHost, s.stream));
      if ((rc = finish_landed()) != NFB_OK) return rc;  // the previous chunk's copy-out runs under this chunk's GPU work
      landed.slot = &s; landed.c0 = c0; landed.cn = cn;
    } else {
      NFB_CUDA(cudaMemcpy2DAsync(out_c, size_t(out_ld) * osz, s.d_out, size_t(chunk) * osz, size_t(cn) * osz,
                                 out_rows, cudaMemcpyDeviceToHost, s.stream));
    }
  }
  if ((rc = finish_landed()) != NFB_OK) return rc;
  for (auto& s : ctx->slots) NFB_CUDA(cudaStreamSynchronize(s.stream));
  return NFB_OK;
}

namespace {
int jacobian_device_impl(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows, const int64_t* in_strides,
                         int in_dtype, void* values, int64_t values_ld, void* jac, const void* cot, int64_t cot_ld,
                         void* grad, int64_t grad_ld, int out_dtype, void* stream) {
  static char dummy_out;  // check_eval_args wants a non-NULL out; `values` is optional here
  int rc = check_eval_args(ctx, model_id, NFB_VARIANT_FP32_FFMA, n, in_rows, in_strides, in_dtype, values ? values : &dummy_out,
                           values ? values_ld : n, out_dtype);
  if (rc != NFB_OK || n == 0) return rc;
  if (!jac && !cot) return fail(NFB_ERR_INVALID_ARGUMENT, "jac is NULL");
  if (cot && (!grad || cot_ld < n || grad_ld < n)) return fail(NFB_ERR_INVALID_ARGUMENT, "VJP: grad is NULL or a leading dimension is smaller than n");
  if (n > (int64_t(1) << 30)) return fail(NFB_ERR_INVALID_ARGUMENT, "n = %lld: the Jacobian kernel takes at most 2^30 queries per call", (long long)n);
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  const Model& m = ctx->models[model_id];
  nfb::JacParams jp;
  fill_eval_params(ctx, m, m.blob, n, in_rows, in_strides, in_dtype, values, values_ld, out_dtype, jp.ev);
  jp.jac = jac;
  jp.cot = cot;
  jp.grad = grad;
  jp.cot_ld = cot_ld;
  jp.grad_ld = grad_ld;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (m.h_pad) {
    case 64: return launch_jac<64>(ctx, jp, st);
    case 128: return launch_jac<128>(ctx, jp, st);
    case 256: return launch_jac<256>(ctx, jp, st);
    default: return launch_jac<512>(ctx, jp, st);
  }
}
}  // namespace

int nfb_eval_jacobian_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows,
                             const int64_t* in_strides, int in_dtype, void* values, int64_t values_ld, void* jac,
                             int out_dtype, void* stream) {
  if (!jac) return fail(NFB_ERR_INVALID_ARGUMENT, "jac is NULL");
  return jacobian_device_impl(ctx, model_id, n, in_rows, in_strides, in_dtype, values, values_ld, jac, nullptr, 0, nullptr, 0,
                              out_dtype, stream);
}

int nfb_eval_vjp_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows, const int64_t* in_strides,
                        int in_dtype, const void* cot, int64_t cot_ld, void* values, int64_t values_ld, void* grad,
                        int64_t grad_ld, int out_dtype, void* stream) {
  if (!cot) return fail(NFB_ERR_INVALID_ARGUMENT, "cot is NULL");
  return jacobian_device_impl(ctx, model_id, n, in_rows, in_strides, in_dtype, values, values_ld, nullptr, cot, cot_ld, grad,
                              grad_ld, out_dtype, stream);
}

namespace {
int grad_device_impl(nfb_context* ctx, bool full, int model_id, int64_t n, const void* const* in_rows, const int64_t* in_strides,
                     int in_dtype, const void* cot, int64_t cot_ld, void* values, int64_t values_ld, void* grad,
                     int64_t grad_ld, int out_dtype, void* stream) {
  static char dummy_out;
  int rc = check_eval_args(ctx, model_id, NFB_VARIANT_FP32_FFMA, n, in_rows, in_strides, in_dtype, values ? values : &dummy_out,
                           values ? values_ld : n, out_dtype);
  if (rc != NFB_OK || n == 0) return rc;
  if (!cot || !grad || cot_ld < n || grad_ld < n)
    return fail(NFB_ERR_INVALID_ARGUMENT, "reverse-mode gradient: cot or grad is NULL, or a leading dimension is smaller than n");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  const Model& m = ctx->models[model_id];
  nfb::GradParams gp;
  fill_eval_params(ctx, m, m.grad_blob, n, in_rows, in_strides, in_dtype, values, values_ld, out_dtype, gp.ev);
  gp.ev.scalars_only = full ? 0 : 1;
  gp.cot = cot;
  gp.grad = grad;
  gp.cot_ld = cot_ld;
  gp.grad_ld = grad_ld;
  gp.scratch = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (m.h_pad) {
    case 64: return full ? launch_grad<Grp64, true>(ctx, gp, st) : launch_grad<Grp64, false>(ctx, gp, st);
    case 128: return full ? launch_grad<Grad128, true>(ctx, gp, st) : launch_grad<Grad128, false>(ctx, gp, st);
    case 256: return full ? launch_grad<Grad256, true>(ctx, gp, st) : launch_grad<Grad256, false>(ctx, gp, st);
    case 512: return full ? launch_grad<Grad512, true>(ctx, gp, st) : launch_grad<Grad512, false>(ctx, gp, st);
  }
  return fail(NFB_ERR_UNSUPPORTED_MODEL, "no kernel for padded width %d", m.h_pad);
}
}  // namespace

int nfb_eval_scalar_grad_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows, const int64_t* in_strides,
                                int in_dtype, const void* cot, int64_t cot_ld, void* values, int64_t values_ld, void* grad,
                                int64_t grad_ld, int out_dtype, void* stream) {
  return grad_device_impl(ctx, false, model_id, n, in_rows, in_strides, in_dtype, cot, cot_ld, values, values_ld, grad, grad_ld,
                          out_dtype, stream);
}

int nfb_eval_vjp_reverse_device(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows, const int64_t* in_strides,
                                int in_dtype, const void* cot, int64_t cot_ld, void* values, int64_t values_ld, void* grad,
                                int64_t grad_ld, int out_dtype, void* stream) {
  return grad_device_impl(ctx, true, model_id, n, in_rows, in_strides, in_dtype, cot, cot_ld, values, values_ld, grad, grad_ld,
                          out_dtype, stream);
}

}  // extern "C"

namespace {
int check_mixed_args(nfb_context* ctx, int model_id, int variant, int64_t n, const void* const* in_rows, const int64_t* in_strides,
                     int in_dtype, float* out_scalars, void* out_bl, int64_t out_ld) {
  int rc = check_eval_args(ctx, model_id, variant, n, in_rows, in_strides, in_dtype, out_scalars, out_ld, NFB_F32);
  if (rc != NFB_OK) return rc;
  if (variant & NFB_OUTPUT_SCALARS) return fail(NFB_ERR_INVALID_ARGUMENT, "NFB_OUTPUT_SCALARS has no boundary-layer rows to write as bfloat16");
  if (n > 0 && !out_bl) return fail(NFB_ERR_INVALID_ARGUMENT, "out_bl is NULL");
  return NFB_OK;
}
}  // namespace

extern "C" {

int nfb_eval_device_mixed(nfb_context* ctx, int model_id, int variant, int64_t n, const void* const* in_rows,
                          const int64_t* in_strides, int in_dtype, float* out_scalars, void* out_bl, int64_t out_ld, void* stream) {
  int rc = check_mixed_args(ctx, model_id, variant, n, in_rows, in_strides, in_dtype, out_scalars, out_bl, out_ld);
  if (rc != NFB_OK || n == 0) return rc;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  return eval_device_impl(ctx, model_id, variant, n, in_rows, in_strides, in_dtype, out_scalars, out_ld, NFB_F32,
                          static_cast<cudaStream_t>(stream), out_bl);
}

int nfb_eval_host_mixed(nfb_context* ctx, int model_id, int variant, int64_t n, const void* const* in_rows,
                        const int64_t* in_strides, int in_dtype, float* out_scalars, void* out_bl, int64_t out_ld, int64_t chunk) {
  int rc = check_mixed_args(ctx, model_id, variant, n, in_rows, in_strides, in_dtype, out_scalars, out_bl, out_ld);
  if (rc != NFB_OK || n == 0) return rc;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  const size_t isz = in_dtype == NFB_F64 ? 8 : 4;
  if (chunk <= 0) chunk = default_chunk(n);
  chunk = std::min<int64_t>((chunk + 3) & ~int64_t(3), (n + 3) & ~int64_t(3));
  constexpr int BL_ROWS = nfb::N_OUT - 6;
  const size_t in_need = size_t(nfb::N_RAW) * chunk * isz;
  const size_t scal_bytes = size_t(6) * chunk * 4, out_need = scal_bytes + size_t(BL_ROWS) * chunk * 2;
  for (int64_t c0 = 0, ci = 0; c0 < n; c0 += chunk, ++ci) {
    HostSlot& s = ctx->slots[ci & 1];
    const int64_t cn = std::min<int64_t>(chunk, n - c0);
    NFB_CUDA(cudaStreamSynchronize(s.stream));  // the slot's previous chunk has fully landed
    if ((rc = ensure(&s.d_in, &s.in_bytes, in_need, false)) != NFB_OK) return rc;
    if ((rc = ensure(&s.d_out, &s.out_bytes, out_need, false)) != NFB_OK) return rc;
    const void* d_rows[nfb::N_RAW];
    int64_t d_strides[nfb::N_RAW];
    for (int r = 0; r < nfb::N_RAW; ++r) {
      char* dst = static_cast<char*>(s.d_in) + size_t(r) * chunk * isz;
      const char* src = static_cast<const char*>(in_rows[r]);
      if (in_strides[r] == 0) {
        NFB_CUDA(cudaMemcpyAsync(dst, src, isz, cudaMemcpyHostToDevice, s.stream));
        d_strides[r] = 0;
      } else if (in_strides[r] == 1) {
        NFB_CUDA(cudaMemcpyAsync(dst, src + size_t(c0) * isz, size_t(cn) * isz, cudaMemcpyHostToDevice, s.stream));
        d_strides[r] = 1;
      } else {
        NFB_CUDA(cudaMemcpy2DAsync(dst, isz, src + size_t(c0) * in_strides[r] * isz, size_t(in_strides[r]) * isz, isz, size_t(cn),
                                   cudaMemcpyHostToDevice, s.stream));
        d_strides[r] = 1;
      }
      d_rows[r] = dst;
    }
    float* const d_scal = static_cast<float*>(s.d_out);
    char* const d_bl = static_cast<char*>(s.d_out) + scal_bytes;
    rc = eval_device_impl(ctx, model_id, variant, cn, d_rows, d_strides, in_dtype, d_scal, chunk, NFB_F32, s.stream, d_bl);
    if (rc != NFB_OK) return rc;
    NFB_CUDA(cudaMemcpy2DAsync(out_scalars + c0, size_t(out_ld) * 4, d_scal, size_t(chunk) * 4, size_t(cn) * 4, 6, cudaMemcpyDeviceToHost, s.stream));
    NFB_CUDA(cudaMemcpy2DAsync(static_cast<char*>(out_bl) + size_t(c0) * 2, size_t(out_ld) * 2, d_bl, size_t(chunk) * 2, size_t(cn) * 2, BL_ROWS,
                               cudaMemcpyDeviceToHost, s.stream));
  }
  for (auto& s : ctx->slots) NFB_CUDA(cudaStreamSynchronize(s.stream));
  return NFB_OK;
}

}  // extern "C"

namespace {
// One training step for one padded width: forward + loss + backward (nfb_train_fwdbwd_kernel), the weight gradients of every
// layer (nfb_train_wgrad_kernel) and the fixed-order reduction of its column splits (nfb_train_reduce_kernel).
template <class Cfg, class WCfg>
int launch_train(nfb_context* ctx, const Model& m, int64_t batch, const float* x, int64_t x_ld, const float* y, int64_t y_ld,
                 const float* loss_weights, float delta, double* loss_out, float* const* grad_w, float* const* grad_b,
                 cudaStream_t stream) {
  constexpr int H = Cfg::H, NC = Cfg::NC, NQ = Cfg::NQ;
  using TC = nfb::TrainCfg<Cfg>;
  const int L = m.n_layers;
  NFB_CONFIGURE(nfb::nfb_train_fwdbwd_kernel<Cfg>, TC::SMEM_BYTES);
  NFB_CONFIGURE(nfb::nfb_train_wgrad_kernel<WCfg>, WCfg::SMEM_BYTES);
  const long long tiles = (batch + NQ - 1) / NQ;
  const long long cols = tiles * NC;
  const int grid = int(std::min<long long>(tiles, ctx->sm_count));

  // ---- jobs of the weight-gradient kernel, one per layer ----
  nfb::WgradParams wp;
  nfb::ReduceParams rp;
  long long part = 0, first_out = 0;
  int ctas = 0;
  for (int l = 0; l < L; ++l) {
    const bool last = (l == L - 1);
    const int m_pad = last ? nfb::TRAIN_LAST_ROWS : H;
    const int n_valid = (l == 0) ? nfb::K0_PAD : H;
    const int n_pad = (n_valid + WCfg::BN - 1) / WCfg::BN * WCfg::BN;
    nfb::WgradJob& j = wp.job[l];
    j.ldz = m_pad;
    j.lda = n_valid;
    j.m_blocks = m_pad / WCfg::BM;
    j.n_blocks = n_pad / WCfg::BN;
    j.n_valid = n_valid;
    j.part_off = part;
    nfb::ReduceJob& r = rp.job[l];
    r.grad_w = grad_w[l];
    r.grad_b = grad_b[l];
    r.out_dim = m.dims[l + 1];
    r.in_dim = m.dims[l];
    r.n_pad = n_pad;
    r.m_pad = m_pad;
    r.last = last ? 1 : 0;
    r.part_off = part;
    r.first = first_out;
    first_out += (long long)r.out_dim * (r.in_dim + 1);
    part += (long long)m_pad * n_pad + m_pad;
    ctas += j.m_blocks * j.n_blocks;
  }
  const long long steps_total = cols / WCfg::BK;
  int k_splits = int(std::max<long long>(1, std::min<long long>({(2LL * ctx->sm_count + ctas - 1) / ctas, steps_total / 4, 64LL})));
  {
    int first = 0;
    for (int l = 0; l < L; ++l) {
      wp.job[l].first_cta = first;
      first += wp.job[l].m_blocks * wp.job[l].n_blocks * k_splits;
    }
    ctas = first;
  }

  // ---- workspace: [scratch of the fused kernel | saved A | saved DZ | partial sums] ----
  auto align = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t scratch_b = align(size_t(grid) * nfb::grad_scratch_floats<H, NC>(L, true) * sizeof(float));
  const size_t a_b = align(nfb::train_saved_a_floats<H>(L, cols) * sizeof(float));
  const size_t dz_b = align(nfb::train_saved_dz_floats<H>(L, cols) * sizeof(float));
  const size_t part_b = align(size_t(k_splits) * size_t(part) * sizeof(float));
  const size_t need = scratch_b + a_b + dz_b + part_b;
  if (need > ctx->train_ws_bytes) {
    // (cudaFree waits for the device: nothing still uses the old block.  Not capturable: warm up with one eager step.)
    int rc = ensure(&ctx->train_ws, &ctx->train_ws_bytes, need, false);
    if (rc != NFB_OK) return rc;
  }
  if (!ctx->packed_row_of) {
    int inv[nfb::N_OUT];
    for (int o = 0; o < nfb::N_OUT; ++o) inv[o] = -1;
    for (int pr = 0; pr < nfb::TRAIN_LAST_ROWS; ++pr) {
      const int o = nfb::pack_last_layer_row(pr);
      if (o >= 0 && o < nfb::N_OUT) inv[o] = pr;
    }
    NFB_CUDA(cudaMalloc(&ctx->packed_row_of, sizeof(inv)));
    NFB_CUDA(cudaMemcpy(ctx->packed_row_of, inv, sizeof(inv), cudaMemcpyHostToDevice));
  }
  char* const ws = static_cast<char*>(ctx->train_ws);
  nfb::TrainParams tp;
  static const void* const no_rows[nfb::N_RAW] = {nullptr};
  static const int64_t no_strides[nfb::N_RAW] = {0};
  fill_eval_params(ctx, m, m.grad_blob, batch, no_rows, no_strides, NFB_F32, nullptr, 0, NFB_F32, tp.ev);
  tp.x = x;
  tp.y = y;
  tp.x_ld = x_ld;
  tp.y_ld = y_ld;
  tp.loss_w = loss_weights;
  tp.inv_batch = 1.0f / float(batch);
  tp.delta = delta;
  tp.loss_out = loss_out;
  tp.scratch = reinterpret_cast<float*>(ws);
  tp.saved_a = reinterpret_cast<float*>(ws + scratch_b);
  tp.saved_dz = reinterpret_cast<float*>(ws + scratch_b + a_b);
  tp.cols = cols;
  for (int l = 0; l < L; ++l) {
    wp.job[l].dz = tp.saved_dz + size_t(l) * H * cols;  // (the last layer's block follows the L - 1 hidden ones)
    wp.job[l].a = tp.saved_a + nfb::train_saved_a_offset<H>(l, cols);
  }
  wp.n_jobs = L;
  wp.k_splits = k_splits;
  wp.cols = cols;
  wp.partial = reinterpret_cast<float*>(ws + scratch_b + a_b + dz_b);
  wp.part_stride = part;
  rp.n_jobs = L;
  rp.k_splits = k_splits;
  rp.partial = wp.partial;
  rp.part_stride = part;
  rp.total = first_out;
  rp.packed_row_of = ctx->packed_row_of;

  NFB_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(double) * nfb::N_OUT, stream));
  nfb::nfb_train_fwdbwd_kernel<Cfg><<<grid, Cfg::THREADS, TC::SMEM_BYTES, stream>>>(tp);
  NFB_CUDA(cudaGetLastError());
  nfb::nfb_train_wgrad_kernel<WCfg><<<ctas, WCfg::THREADS, WCfg::SMEM_BYTES, stream>>>(wp);
  NFB_CUDA(cudaGetLastError());
  nfb::nfb_train_reduce_kernel<<<unsigned((first_out + 255) / 256), 256, 0, stream>>>(rp);
  NFB_CUDA(cudaGetLastError());
  ctx->launches += 3;
  return NFB_OK;
}
}  // namespace

extern "C" {

int nfb_train_step_device(nfb_context* ctx, int model_id, int64_t batch, const float* x, int64_t x_ld, const float* y_data,
                          int64_t y_ld, const float* loss_weights, float huber_delta, double* loss_components,
                          float* const* grad_weights, float* const* grad_biases, void* stream) {
  if (!ctx) return fail(NFB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (model_id < 0 || model_id >= NFB_MAX_MODELS)
    return fail(NFB_ERR_INVALID_ARGUMENT, "model_id %d out of range [0, %d)", model_id, NFB_MAX_MODELS);
  if (!ctx->models[model_id].loaded) return fail(NFB_ERR_NOT_LOADED, "model slot %d is empty", model_id);
  if (!ctx->have_distribution) return fail(NFB_ERR_NOT_LOADED, "nfb_set_distribution has not been called");
  if (batch < 1 || batch > (int64_t(1) << 28)) return fail(NFB_ERR_INVALID_ARGUMENT, "batch = %lld; supported: 1 .. 2^28", (long long)batch);
  if (!x || !y_data || !loss_weights || !loss_components || !grad_weights || !grad_biases)
    return fail(NFB_ERR_INVALID_ARGUMENT, "NULL argument");
  if (x_ld < NFB_N_LATENT_INPUTS || y_ld < NFB_N_OUTPUT_ROWS)
    return fail(NFB_ERR_INVALID_ARGUMENT, "row strides %lld / %lld are smaller than the rows (25 / 198)", (long long)x_ld, (long long)y_ld);
  if (!(huber_delta > 0.0f)) return fail(NFB_ERR_INVALID_ARGUMENT, "huber_delta must be positive");
  const Model& m = ctx->models[model_id];
  for (int l = 0; l < m.n_layers; ++l)
    if (!grad_weights[l] || !grad_biases[l]) return fail(NFB_ERR_INVALID_ARGUMENT, "gradient pointer of layer %d is NULL", l);
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (m.h_pad) {
    case 64: return launch_train<Grp64, nfb::WgradCfg<64, 64>>(ctx, m, batch, x, x_ld, y_data, y_ld, loss_weights, huber_delta, loss_components, grad_weights, grad_biases, st);
    case 128: return launch_train<Grad128, nfb::WgradCfg<128, 128>>(ctx, m, batch, x, x_ld, y_data, y_ld, loss_weights, huber_delta, loss_components, grad_weights, grad_biases, st);
    case 256: return launch_train<Train256, nfb::WgradCfg<128, 128>>(ctx, m, batch, x, x_ld, y_data, y_ld, loss_weights, huber_delta, loss_components, grad_weights, grad_biases, st);
    case 512: return launch_train<Train512, nfb::WgradCfg<128, 128>>(ctx, m, batch, x, x_ld, y_data, y_ld, loss_weights, huber_delta, loss_components, grad_weights, grad_biases, st);
  }
  return fail(NFB_ERR_UNSUPPORTED_MODEL, "no kernel for padded width %d", m.h_pad);
}

int nfb_eval_jacobian_host(nfb_context* ctx, int model_id, int64_t n, const void* const* in_rows,
                           const int64_t* in_strides, int in_dtype, void* values, int64_t values_ld, void* jac,
                           int out_dtype) {
  static char dummy_out;
  int rc = check_eval_args(ctx, model_id, NFB_VARIANT_FP32_FFMA, n, in_rows, in_strides, in_dtype, values ? values : &dummy_out,
                           values ? values_ld : n, out_dtype);
  if (rc != NFB_OK || n == 0) return rc;
  if (!jac) return fail(NFB_ERR_INVALID_ARGUMENT, "jac is NULL");
  if (n > (int64_t(1) << 22)) return fail(NFB_ERR_INVALID_ARGUMENT, "n = %lld: at most 2^22 queries per host call (72 KB of derivatives each)", (long long)n);
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  NFB_CUDA(cudaSetDevice(ctx->device));
  const size_t isz = in_dtype == NFB_F64 ? 8 : 4, osz = out_dtype == NFB_F64 ? 8 : 4;
  const int64_t ld = (n + 3) & ~int64_t(3);
  auto align = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t in_bytes = align(size_t(nfb::N_RAW) * n * isz), val_bytes = align(size_t(nfb::N_OUT) * ld * osz);
  const size_t jac_bytes = size_t(n) * nfb::N_OUT * nfb::JAC_T * osz, total = in_bytes + val_bytes + jac_bytes;
  if ((rc = ensure(&ctx->jac_dev, &ctx->jac_dev_bytes, total, false)) != NFB_OK) return rc;
  char* const d_in = static_cast<char*>(ctx->jac_dev);
  char* const d_val = d_in + in_bytes;
  char* const d_jac = d_val + val_bytes;
  HostSlot& s = ctx->slots[0];
  NFB_CUDA(cudaStreamSynchronize(s.stream));
  const void* d_rows[nfb::N_RAW];
  int64_t d_strides[nfb::N_RAW];
  for (int r = 0; r < nfb::N_RAW; ++r) {
    d_rows[r] = d_in + size_t(r) * n * isz;
    d_strides[r] = 1;
  }
  auto pack_inputs = [&](char* dst_base) {
    for (int r = 0; r < nfb::N_RAW; ++r) {
      char* dst = dst_base + size_t(r) * n * isz;
      const char* src = static_cast<const char*>(in_rows[r]);
      if (in_strides[r] == 1) memcpy(dst, src, size_t(n) * isz);
      else for (int64_t i = 0; i < n; ++i) memcpy(dst + i * isz, src + size_t(i) * in_strides[r] * isz, isz);
    }
  };
  // An optimiser's per-iteration call: everything through one pinned block -- one H2D copy, one launch, one D2H copy.
  const bool small = total <= (size_t(32) << 20);
  if (small) {
    if ((rc = ensure(&ctx->jac_host, &ctx->jac_host_bytes, total, true)) != NFB_OK) return rc;
    char* const h = static_cast<char*>(ctx->jac_host);
    pack_inputs(h);
    NFB_CUDA(cudaMemcpyAsync(d_in, h, in_bytes, cudaMemcpyHostToDevice, s.stream));
    rc = nfb_eval_jacobian_device(ctx, model_id, n, d_rows, d_strides, in_dtype, d_val, ld, d_jac, out_dtype, s.stream);
    if (rc != NFB_OK) return rc;
    NFB_CUDA(cudaMemcpyAsync(h + in_bytes, d_val, val_bytes + jac_bytes, cudaMemcpyDeviceToHost, s.stream));
    NFB_CUDA(cudaStreamSynchronize(s.stream));
    memcpy(jac, h + in_bytes + val_bytes, jac_bytes);
    if (values)
      for (int r = 0; r < nfb::N_OUT; ++r)
        memcpy(static_cast<char*>(values) + size_t(r) * values_ld * osz, h + in_bytes + size_t(r) * ld * osz, size_t(n) * osz);
  } else {
    std::vector<char> h_in(size_t(nfb::N_RAW) * n * isz);
    pack_inputs(h_in.data());
    NFB_CUDA(cudaMemcpyAsync(d_in, h_in.data(), h_in.size(), cudaMemcpyHostToDevice, s.stream));
    rc = nfb_eval_jacobian_device(ctx, model_id, n, d_rows, d_strides, in_dtype, d_val, ld, d_jac, out_dtype, s.stream);
    if (rc != NFB_OK) return rc;
    NFB_CUDA(cudaMemcpyAsync(jac, d_jac, jac_bytes, cudaMemcpyDeviceToHost, s.stream));
    if (values)
      NFB_CUDA(cudaMemcpy2DAsync(values, size_t(values_ld) * osz, d_val, size_t(ld) * osz, size_t(n) * osz, nfb::N_OUT, cudaMemcpyDeviceToHost, s.stream));
    NFB_CUDA(cudaStreamSynchronize(s.stream));
  }
  return NFB_OK;
}

int nfb_set_tuning(nfb_context* ctx, int key, int64_t value) {
  if (!ctx) return fail(NFB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  switch (key) {
    case NFB_TUNE_SMALL_BATCH_MAX: ctx->tune_small_batch_max = value; return NFB_OK;
    case NFB_TUNE_TC2_MAX_GRID: ctx->tune_tc2_max_grid = int(value); return NFB_OK;
    case NFB_TUNE_HOST_THREADS: ctx->host_thread_cap = int(std::max<int64_t>(0, value)); return NFB_OK;
    case NFB_TUNE_MAPPED_MAX: ctx->tune_mapped_max = value; return NFB_OK;
  }
  return fail(NFB_ERR_INVALID_ARGUMENT, "unknown tuning key %d", key);
}

// ---- several devices behind one call ------------------------------------------------------------------------------
int nfb_shard_range(int64_t n, int index, int count, int64_t* begin, int64_t* end) {
  if (count < 1 || index < 0 || index >= count || n < 0 || !begin || !end)
    return fail(NFB_ERR_INVALID_ARGUMENT, "shard %d of %d over %lld queries", index, count, (long long)n);
  const int64_t per = ((n + count - 1) / count + 3) & ~int64_t(3);  // ceil(n / count), rounded up to 4 queries
  *begin = std::min<int64_t>(n, per * index);
  *end = std::min<int64_t>(n, per * (index + 1));
  return NFB_OK;
}

}  // extern "C"

struct nfb_multi {
  std::vector<nfb_context*> ctx;
  std::vector<cpu_set_t> cpus;     // CPUs each device is attached to (empty set = unknown: no binding)
  std::vector<bool> have_cpus;
  std::mutex mu;                   // one multi-device call at a time
};

namespace {
// CPUs local to a CUDA device: /sys/bus/pci/devices/<domain:bus:device.function>/local_cpulist ("0-31,64-95").
bool device_local_cpus(int device, cpu_set_t* set) {
  char bus_id[32] = "";
  if (cudaDeviceGetPCIBusId(bus_id, sizeof(bus_id), device) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  for (char* c = bus_id; *c; ++c) *c = char(tolower(*c));
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus_id);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char line[1024] = "";
  const bool got = fgets(line, sizeof(line), f) != nullptr;
  fclose(f);
  if (!got) return false;
  CPU_ZERO(set);
  int n_set = 0;
  for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k == 1) b = a;
    if (k < 1 || a < 0 || b < a) continue;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) {
      CPU_SET(c, set);
      ++n_set;
    }
  }
  return n_set > 0;
}
}  // namespace

extern "C" {

int nfb_multi_create(const int* devices, int n_devices, nfb_multi** out_multi) {
  if (!out_multi) return fail(NFB_ERR_INVALID_ARGUMENT, "out_multi is NULL");
  *out_multi = nullptr;
  std::vector<int> ids;
  if (devices) {
    if (n_devices < 1) return fail(NFB_ERR_INVALID_ARGUMENT, "n_devices = %d", n_devices);
    ids.assign(devices, devices + n_devices);
    for (size_t i = 0; i < ids.size(); ++i)
      for (size_t j = 0; j < i; ++j)
        if (ids[i] == ids[j]) return fail(NFB_ERR_INVALID_ARGUMENT, "device %d listed twice", ids[i]);
  } else {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      return fail(NFB_ERR_NO_DEVICE, "no CUDA device visible (%s); neuralfoil_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    for (int d = 0; d < count; ++d) {
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ids.push_back(d);
    }
    if (ids.empty()) return fail(NFB_ERR_NO_DEVICE, "none of the %d visible devices is sm_100", count);
  }
  nfb_multi* m = new nfb_multi();
  const int per_device_threads = std::max(2, host_threads() / int(ids.size()));
  for (int d : ids) {
    nfb_context* c = nullptr;
    const int rc = nfb_create(d, &c);
    if (rc != NFB_OK) {
      nfb_multi_destroy(m);
      return rc;
    }
    c->host_thread_cap = per_device_threads;
    m->ctx.push_back(c);
    cpu_set_t set;
    const bool ok = device_local_cpus(d, &set);
    m->cpus.push_back(set);
    m->have_cpus.push_back(ok);
  }
  *out_multi = m;
  return NFB_OK;
}

void nfb_multi_destroy(nfb_multi* multi) {
  if (!multi) return;
  for (nfb_context* c : multi->ctx) nfb_destroy(c);
  delete multi;
}

int nfb_multi_device_count(const nfb_multi* multi) { return multi ? int(multi->ctx.size()) : 0; }

nfb_context* nfb_multi_context(nfb_multi* multi, int index) {
  if (!multi || index < 0 || index >= int(multi->ctx.size())) return nullptr;
  return multi->ctx[index];
}

int nfb_multi_set_distribution(nfb_multi* multi, const double* mean25, const double* inv_cov) {
  if (!multi) return fail(NFB_ERR_INVALID_ARGUMENT, "multi is NULL");
  for (nfb_context* c : multi->ctx) {
    const int rc = nfb_set_distribution(c, mean25, inv_cov);
    if (rc != NFB_OK) return rc;
  }
  return NFB_OK;
}

int nfb_multi_load_model(nfb_multi* multi, int model_id, int n_layers, const int* dims, const float* const* weights,
                         const float* const* biases) {
  if (!multi) return fail(NFB_ERR_INVALID_ARGUMENT, "multi is NULL");
  for (nfb_context* c : multi->ctx) {
    const int rc = nfb_load_model(c, model_id, n_layers, dims, weights, biases);
    if (rc != NFB_OK) return rc;
  }
  return NFB_OK;
}

int nfb_multi_eval_host(nfb_multi* multi, int model_id, int variant, int64_t n, const void* const* in_rows,
                        const int64_t* in_strides, int in_dtype, void* out, int64_t out_ld, int out_dtype, int64_t chunk) {
  if (!multi || multi->ctx.empty()) return fail(NFB_ERR_INVALID_ARGUMENT, "multi is NULL");
  // argument errors are reported once, by the first device's checks, before any thread starts
  int rc = check_eval_args(multi->ctx[0], model_id, variant, n, in_rows, in_strides, in_dtype, out, out_ld, out_dtype);
  if (rc != NFB_OK || n == 0) return rc;
  std::lock_guard<std::mutex> lock(multi->mu);
  const int G = int(multi->ctx.size());
  const size_t isz = in_dtype == NFB_F64 ? 8 : 4, osz = out_dtype == NFB_F64 ? 8 : 4;
  std::vector<int> codes(G, NFB_OK);
  std::vector<std::string> messages(G);
  auto run_shard = [&](int g) {
    int64_t b = 0, e = 0;
    nfb_shard_range(n, g, G, &b, &e);
    if (e <= b) return;
    if (multi->have_cpus[g]) sched_setaffinity(0, sizeof(cpu_set_t), &multi->cpus[g]);  // this thread only; best effort
    const void* rows[nfb::N_RAW];
    for (int r = 0; r < nfb::N_RAW; ++r)
      rows[r] = static_cast<const char*>(in_rows[r]) + size_t(b) * size_t(in_strides[r]) * isz;
    codes[g] = nfb_eval_host(multi->ctx[g], model_id, variant, e - b, rows, in_strides, in_dtype,
                             static_cast<char*>(out) + size_t(b) * osz, out_ld, out_dtype, chunk);
    if (codes[g] != NFB_OK) messages[g] = g_err;  // the error text is thread-local: carry it to the caller's thread
  };
  {
    std::vector<std::thread> workers;
    workers.reserve(G);
    for (int g = 0; g < G; ++g) workers.emplace_back(run_shard, g);  // fresh threads: the caller's affinity is untouched
    for (auto& t : workers) t.join();
  }
  for (int g = 0; g < G; ++g)
    if (codes[g] != NFB_OK) return fail(codes[g], "device %d: %s", multi->ctx[g]->device, messages[g].c_str());
  return NFB_OK;
}

int64_t nfb_multi_launch_count(const nfb_multi* multi) {
  int64_t total = 0;
  if (multi)
    for (const nfb_context* c : multi->ctx) total += c->launches.load();
  return total;
}

int nfb_host_register(void* ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return fail(NFB_ERR_INVALID_ARGUMENT, "bad argument");
  NFB_CUDA(cudaHostRegister(ptr, size_t(bytes), cudaHostRegisterPortable));
  return NFB_OK;
}

int nfb_host_unregister(void* ptr) {
  if (ptr) NFB_CUDA(cudaHostUnregister(ptr));
  return NFB_OK;
}

int nfb_host_alloc(void** out_ptr, int64_t bytes) {
  if (!out_ptr || bytes < 0) return fail(NFB_ERR_INVALID_ARGUMENT, "bad argument");
  *out_ptr = nullptr;
  if (bytes == 0) return NFB_OK;
  NFB_CUDA(cudaMallocHost(out_ptr, size_t(bytes)));
  return NFB_OK;
}

int nfb_host_free(void* ptr) {
  if (ptr) NFB_CUDA(cudaFreeHost(ptr));
  return NFB_OK;
}

int64_t nfb_launch_count(const nfb_context* ctx) { return ctx ? ctx->launches.load() : 0; }

int nfb_debug_tc_stamps(nfb_context* ctx, int enable, long long* out64) {
  if (!ctx) return fail(NFB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  NFB_CUDA(cudaSetDevice(ctx->device));
  if (enable && !ctx->tc_debug) {
    NFB_CUDA(cudaMalloc(&ctx->tc_debug, 64 * sizeof(long long)));
    NFB_CUDA(cudaMemset(ctx->tc_debug, 0, 64 * sizeof(long long)));
  }
  if (out64 && ctx->tc_debug) {
    NFB_CUDA(cudaDeviceSynchronize());
    NFB_CUDA(cudaMemcpy(out64, ctx->tc_debug, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  if (!enable && ctx->tc_debug) {
    NFB_CUDA(cudaFree(ctx->tc_debug));
    ctx->tc_debug = nullptr;
  }
  return NFB_OK;
}

int nfb_measure_fp32_peak(nfb_context* ctx, int iters, double* out_tflops) {
  if (!ctx || !out_tflops) return fail(NFB_ERR_INVALID_ARGUMENT, "NULL argument");
  NFB_CUDA(cudaSetDevice(ctx->device));
  float* sink = nullptr;
  NFB_CUDA(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  NFB_CUDA(cudaEventCreate(&e0));
  NFB_CUDA(cudaEventCreate(&e1));
  const int inner = 2048, blocks = ctx->sm_count * 2;
  const double flops = 2.0 * 16 * 32 * double(inner) * 1024.0 * blocks;
  double best = 0.0;
  if (iters < 1) iters = 1;
  for (int i = 0; i < iters + 2; ++i) {  // two warm-up launches
    NFB_CUDA(cudaEventRecord(e0, 0));
    nfb::nfb_ffma_peak_kernel<<<blocks, 1024>>>(sink, inner);
    NFB_CUDA(cudaEventRecord(e1, 0));
    NFB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    NFB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    ctx->launches += 1;
    if (i >= 2) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  NFB_CUDA(cudaEventDestroy(e0));
  NFB_CUDA(cudaEventDestroy(e1));
  NFB_CUDA(cudaFree(sink));
  *out_tflops = best;
  return NFB_OK;
}

}  // extern "C"
