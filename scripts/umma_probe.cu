// tcgen05 building-block probe for the tensor-core variant (run on the GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe scripts/umma_probe.cu && /tmp/umma_probe
// One "layer": D[512 x NB] = W[512 x 512] * X[512 x NB] with kind::tf32, W as the A operand (four M = 128 blocks,
// streamed as 16 KB K-slices through a TMA ring), X as the B operand (K-major, no swizzle, padded LBO), D in TMEM.
// Checks the result against a CPU reference with tf32-truncated operands and reports cycles per MMA
//   (a) with the slices re-used from shared memory (pure tensor-pipe + smem-operand rate),
//   (b) with every slice streamed from L2 by all 148 SMs at once (the fused kernel's situation).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef NB
#define NB 64  // batch columns (MMA N)
#endif
constexpr int M_TOT = 512, K_TOT = 512, KS = 32, STAGES = 4;
constexpr uint32_t A_SLICE_BYTES = 128 * KS * 4;  // 16 KB: 128 rows x 32 k
#ifndef INTERLEAVE
#define INTERLEAVE 1  // 1: a slice is ONE k-step (K = 8) of all four M blocks, MMAs round-robin over the four accumulators
#endif
#if INTERLEAVE
#define SW128 0
#endif
#ifndef SW128
#define SW128 1  // 1: SWIZZLE_128B K-major atoms (8 rows x 128 B, 16-byte chunk c of row r at c ^ r); 0: no swizzle
#endif
#if SW128
constexpr uint32_t A_LBO = 16, A_SBO = 1024, A_KSTEP = 32;        // one MMA (K = 8) advances 32 B inside the 128 B row
constexpr uint32_t B_LBO = 16, B_SBO = 1024, B_KSTEP = 32;
constexpr uint32_t B_ATOM_BYTES = (NB / 8) * 1024;                // one 32-wide K atom of the activations
constexpr uint32_t B_BYTES = (K_TOT / KS) * B_ATOM_BYTES;
constexpr uint64_t DESC_LAYOUT = (uint64_t)2 << 61;
#else
#if INTERLEAVE
constexpr uint32_t A_LBO = 128, A_SBO = 256, A_KSTEP = 0;           // per M block: [row group 16][k chunk 2] x 128 B = 4 KB
#else
constexpr uint32_t A_LBO = 128, A_SBO = 1024, A_KSTEP = 2 * A_LBO;  // core matrices: [row group 16][k chunk 8] x 128 B
#endif
constexpr uint32_t B_SBO = 128, B_LBO = (NB / 8) * 128 + 16, B_KSTEP = 2 * B_LBO;  // [k chunk 128][col group] x 128 B, +16 B pad
constexpr uint32_t B_ATOM_BYTES = (KS / 4) * B_LBO;
constexpr uint32_t B_BYTES = (K_TOT / 4) * B_LBO;
constexpr uint64_t DESC_LAYOUT = 0;
#endif
constexpr int TMEM_COLS = (4 * NB <= 32) ? 32 : (4 * NB <= 64) ? 64 : (4 * NB <= 128) ? 128 : (4 * NB <= 256) ? 256 : 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | DESC_LAYOUT;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
        "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// blob_a: [mb 4][slice 16] x 16 KB in the canonical layout; blob_b: B_BYTES; out: [512][NB] (CTA 0 only)
__global__ void __launch_bounds__(192, 1) probe(const float* blob_a, const float* blob_b, float* out, int reps, int stream_weights, unsigned long long* stats) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sB = smem;                                   // B_BYTES (multiple of 16)
  unsigned char* sA = smem + ((B_BYTES + 1023) & ~1023u);     // STAGES x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + STAGES * A_SLICE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * STAGES, bar_b = bar_empty + 8 * STAGES, bar_done = bar_b + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_b, 1); mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int n_slices = 4 * (K_TOT / KS);  // per rep: 64

  if (warp == 0 && lane == 0) {  // ---- producer ----
    mbar_expect_tx(bar_b, B_BYTES);
    bulk_g2s(smem_u32(sB), blob_b, B_BYTES, bar_b);
    const int total = stream_weights ? reps * n_slices : STAGES;
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      mbar_expect_tx(bar_full + 8 * s, A_SLICE_BYTES);
      bulk_g2s(smem_u32(sA + s * A_SLICE_BYTES), reinterpret_cast<const unsigned char*>(blob_a) + size_t(it % n_slices) * A_SLICE_BYTES, A_SLICE_BYTES, bar_full + 8 * s);
    }
  } else if (warp == 1 && lane == 0) {  // ---- MMA issuer ----
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(bar_b, 0);
    if (!stream_weights) for (int s = 0; s < STAGES; ++s) mbar_wait(bar_full + 8 * s, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned long long c0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
      for (int sl = 0; sl < n_slices; ++sl) {
        const int it = rep * n_slices + sl, s = it % STAGES, ph = (it / STAGES) & 1;
        if (stream_weights) { mbar_wait(bar_full + 8 * s, ph); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        const uint32_t a_base = smem_u32(sA + s * A_SLICE_BYTES);
#if INTERLEAVE
        const int k8 = sl;  // slice sl = k-step sl (K = 8) of all four M blocks
        const uint32_t b_base = smem_u32(sB) + (k8 / 4) * B_ATOM_BYTES + (k8 % 4) * B_KSTEP;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
          umma_tf32(tmem + mb * NB, make_desc(a_base + mb * 4096, A_LBO, A_SBO), make_desc(b_base, B_LBO, B_SBO), idesc, k8 != 0);
#else
        const int mb = sl / (K_TOT / KS), ks = sl % (K_TOT / KS);
        const uint32_t b_base = smem_u32(sB) + ks * B_ATOM_BYTES;
#pragma unroll
        for (int j = 0; j < KS / 8; ++j)  // 4 MMAs of K = 8 (two 16-byte k chunks each)
          umma_tf32(tmem + mb * NB, make_desc(a_base + j * A_KSTEP, A_LBO, A_SBO), make_desc(b_base + j * B_KSTEP, B_LBO, B_SBO), idesc, (ks | j) != 0);
#endif
        if (stream_weights) umma_commit(bar_empty + 8 * s);
      }
    umma_commit(bar_done);
    mbar_wait(bar_done, 0);
    const unsigned long long c1 = clock64();
    if (stats) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = (unsigned long long)reps * n_slices * (KS / 8); }
  } else if (warp >= 2) {  // ---- epilogue: 4 warps, lane quarter = warp % 4 ----
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    for (int mb = 0; mb < 4; ++mb)
      for (int c0 = 0; c0 < NB; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + mb * NB + c0, v);
        if (blockIdx.x == 0)
          for (int c = 0; c < 32; ++c) out[(mb * 128 + q * 32 + lane) * NB + c0 + c] = __uint_as_float(v[c]);
      }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)TMEM_COLS) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> W(M_TOT * K_TOT), X(K_TOT * NB);
  srand(1);
  for (auto& w : W) w = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
  for (auto& x : X) x = (rand() / (float)RAND_MAX - 0.5f) * 4.0f;
  // pack A: [mb][slice] 16 KB blocks; element (r, k) at (r/8)*A_SBO + (k/4)*A_LBO + (r%8)*16 + (k%4)*4
  std::vector<float> blobA(M_TOT * K_TOT), blobB(B_BYTES / 4, 0.f);
  for (int m = 0; m < M_TOT; ++m)
    for (int k = 0; k < K_TOT; ++k) {
      const int mb = m / 128, r = m % 128, sl = k / KS, kk = k % KS;
#if INTERLEAVE
      const size_t off = size_t(k / 8) * A_SLICE_BYTES + mb * 4096 + (r / 8) * A_SBO + ((k % 8) / 4) * A_LBO + (r % 8) * 16 + (k % 4) * 4;
#elif SW128
      const size_t off = size_t(mb * (K_TOT / KS) + sl) * A_SLICE_BYTES + (r / 8) * 1024 + (r % 8) * 128 + (((kk / 4) ^ (r % 8)) * 16) + (kk % 4) * 4;
#else
      const size_t off = size_t(mb * (K_TOT / KS) + sl) * A_SLICE_BYTES + (r / 8) * A_SBO + (kk / 4) * A_LBO + (r % 8) * 16 + (kk % 4) * 4;
#endif
      blobA[off / 4] = W[m * K_TOT + k];
    }
  for (int k = 0; k < K_TOT; ++k)
    for (int n = 0; n < NB; ++n) {
#if SW128
      const size_t off = size_t(k / KS) * B_ATOM_BYTES + (n / 8) * 1024 + (n % 8) * 128 + ((((k % KS) / 4) ^ (n % 8)) * 16) + (k % 4) * 4;
#else
      const size_t off = size_t(k / 4) * B_LBO + (n / 8) * B_SBO + (n % 8) * 16 + (k % 4) * 4;
#endif
      blobB[off / 4] = X[k * NB + n];
    }
  float *dA, *dB, *dOut; unsigned long long* dStats;
  cudaMalloc(&dA, blobA.size() * 4); cudaMalloc(&dB, blobB.size() * 4); cudaMalloc(&dOut, M_TOT * NB * 4); cudaMalloc(&dStats, 148 * 16);
  cudaMemcpy(dA, blobA.data(), blobA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, blobB.data(), blobB.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0, M_TOT * NB * 4);
  const size_t smem = ((B_BYTES + 1023) & ~1023u) + STAGES * A_SLICE_BYTES + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("NB=%d smem=%zu bytes, TMEM cols=%d\n", NB, smem, TMEM_COLS);

  // correctness: one rep, streamed, one CTA
  probe<<<1, 192, smem>>>(dA, dB, dOut, 1, 1, dStats);
  cudaError_t e = cudaDeviceSynchronize();
  printf("correctness launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> out(M_TOT * NB);
  cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_ref = 0, max_err_full = 0;
  for (int m = 0; m < M_TOT; ++m)
    for (int n = 0; n < NB; ++n) {
      double ref = 0, ref_full = 0;
      for (int k = 0; k < K_TOT; ++k) { ref += (double)tf32_trunc(W[m * K_TOT + k]) * tf32_trunc(X[k * NB + n]); ref_full += (double)W[m * K_TOT + k] * X[k * NB + n]; }
      max_err = fmax(max_err, fabs(out[m * NB + n] - ref)); max_ref = fmax(max_ref, fabs(ref)); max_err_full = fmax(max_err_full, fabs(out[m * NB + n] - ref_full));
    }
  printf("max |D - ref(tf32-truncated operands)| = %.3e, vs fp64 product of fp32 operands %.3e (max |ref| %.3f) -> %s\n", max_err, max_err_full, max_ref, max_err < 1e-3 * max_ref ? "OK" : "MISMATCH");

  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {1, 148}) {
      const int reps = 50;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      probe<<<grid, 192, smem>>>(dA, dB, dOut, reps, mode, dStats); cudaDeviceSynchronize();
      cudaEventRecord(e0); probe<<<grid, 192, smem>>>(dA, dB, dOut, reps, mode, dStats); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long h[2]; cudaMemcpy(h, dStats, 16, cudaMemcpyDeviceToHost);
      const double cyc_per_mma = (double)h[0] / (double)h[1];
      printf("mode=%s grid=%3d: %.1f cycles per MMA (M=128,N=%d,K=8 tf32) = %.0f MAC/clk/SM; %.3f ms; L2->smem %.1f B/clk/SM; err=%s\n", mode ? "stream-from-L2" : "smem-resident  ", grid,
             cyc_per_mma, NB, 128.0 * NB * 8 / cyc_per_mma, ms, mode ? A_SLICE_BYTES / (cyc_per_mma * 4) : 0.0, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
