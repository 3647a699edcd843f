// How fast can one SM pull 8 KB bulk copies (TMA, cp.async.bulk) out of L2 into shared memory, alone and while its
// tensor core runs pair-wide MMAs (tcgen05 cta_group::2, M = 256, N = 256, K = 16) out of the same shared memory?
// All 148 SMs run (74 clusters of 2), every CTA streaming the same 2.8 MB buffer in the same order, like the evaluator.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_mma_probe scripts/tma_mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int MAX_RING = 16;
constexpr uint32_t RING_BYTES = 128 * 1024;
#ifndef ROWS
#define ROWS 128   // A rows per CTA: 128 (pair M = 256, 128 cycles per MMA) or 64 (pair M = 128, 64 cycles)
#endif
constexpr uint32_t A_BYTES = ROWS * 128;   // ROWS x 64 fp16, SW128
constexpr uint32_t B_BYTES = 16384;       // my 128 B rows x 64 k as two no-swizzle 32-k slices

// mode bit 0: TMA stream, bit 1: MMA stream.  copies / mmas: how many of each per CTA (leader issues the MMAs).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const unsigned char* src, uint32_t src_bytes, int mode, int copies, int mmas, unsigned long long* stats, int RING, uint32_t SLOT) {
  extern __shared__ unsigned char smem_un[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_un) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sB = smem + A_BYTES;
  unsigned char* ring = sB + B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + RING_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + MAX_RING + 2);
  const uint32_t bar_full = smem_u32(bars), bar_done = bar_full + 8 * MAX_RING;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) mbar_init(bar_full + 8 * s, 1);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 && (mode & 1)) {  // TMA stream: RING copies in flight, slot re-issued as soon as its previous copy has landed
    const unsigned long long c0 = clock64();
    uint32_t s = 0, ph = 0, off = rank * SLOT;
    for (int it = 0; it < copies; ++it) {
      if (it >= RING) mbar_wait(bar_full + 8 * s, ph ^ 1);
      if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_full + 8 * s), "r"(SLOT) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + s * SLOT)), "l"(src + off), "r"(SLOT), "r"(bar_full + 8 * s) : "memory");
      }
      __syncwarp();
      off += 2 * SLOT;
      if (off + SLOT > src_bytes) off = rank * SLOT;
      if (++s == RING) { s = 0; ph ^= 1; }
    }
    for (int k = 0; k < RING && k < copies; ++k) {  // drain
      const int it = copies - 1 - k;
      mbar_wait(bar_full + 8 * (it % RING), (it / RING) & 1);
    }
    const unsigned long long c1 = clock64();
    if (lane == 0) stats[4 * blockIdx.x + 0] = c1 - c0;
  }
  if (warp == 1 && rank == 0 && (mode & 2)) {  // MMA stream
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)((2 * ROWS) >> 4) << 24);
    const uint64_t a_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t b_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t a0 = (smem_u32(sA) >> 4) & 0x3FFF, b0 = (smem_u32(sB) >> 4) & 0x3FFF;
    const unsigned long long c0 = clock64();
    for (int it = 0; it < mmas / 4; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = a_hi | (a0 + ks * 2), bd = b_hi | (b0 + (ks >> 1) * 512 + (ks & 1) * 16);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        }
      }
      __syncwarp();
    }
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_done), "h"((uint16_t)3) : "memory");
    __syncwarp();
    mbar_wait(bar_done, 0);
    const unsigned long long c1 = clock64();
    if (lane == 0) stats[4 * blockIdx.x + 1] = c1 - c0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
  const uint32_t src_bytes = 2800 * 1024;
  unsigned char* src; unsigned long long* stats;
  cudaMalloc(&src, src_bytes); cudaMemset(src, 0, src_bytes); cudaMalloc(&stats, 148 * 4 * 8);
  const size_t smem = A_BYTES + B_BYTES + RING_BYTES + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int copies = 4000, mmas = 8000;
  struct Case { int grid, mode, ring; uint32_t slot; };
  const Case cases[] = {{148, 1, 8, 8192}, {148, 2, 8, 8192}, {148, 3, 8, 8192}, {2, 1, 8, 8192}, {148, 1, 4, 8192}, {148, 1, 16, 8192},
                        {148, 1, 4, 16384}, {148, 1, 8, 16384}, {148, 1, 8, 4096}, {148, 1, 16, 4096}, {148, 1, 16, 2048}, {148, 3, 16, 8192}};
  for (const Case& c : cases) {
    cudaMemset(stats, 0, 148 * 4 * 8);
    probe<<<c.grid, 128, smem>>>(src, src_bytes, c.mode, copies, mmas, stats, c.ring, c.slot);  // warm
    cudaDeviceSynchronize();
    probe<<<c.grid, 128, smem>>>(src, src_bytes, c.mode, copies, mmas, stats, c.ring, c.slot);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned long long> h(148 * 4);
    cudaMemcpy(h.data(), stats, 148 * 4 * 8, cudaMemcpyDeviceToHost);
    double tma = 0, mma = 0; int nt = 0, nm = 0;
    for (int b = 0; b < c.grid; ++b) { if (h[4 * b]) { tma += h[4 * b]; ++nt; } if (h[4 * b + 1]) { mma += h[4 * b + 1]; ++nm; } }
    printf("M %d  grid %3d  %s%s  %2d x %5u B in flight:", 2 * ROWS, c.grid, c.mode & 1 ? "TMA" : "   ", c.mode & 2 ? "+MMA" : "    ", c.ring, c.slot);
    if (nt) printf("  %.0f cycles per copy = %.1f B/clk/SM (chip %.0f B/clk)", tma / nt / copies, double(c.slot) * copies / (tma / nt), double(c.grid) * c.slot * copies / (tma / nt));
    if (nm) printf("  MMA %.1f cycles each", mma / nm / mmas);
    printf("  [%s]\n", cudaGetErrorString(e));
  }
  return 0;
}
