// tcgen05 kind::f16 rate probe for the shape the tensor-core kernel uses: M = 128 (SW128 K-major A), N = 256, K = 16.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe_f16 scripts/umma_probe_f16.cu
// Reports cycles per MMA with operands resident in shared memory, for B stored (a) as no-swizzle core matrices (what
// nfb_tc.cuh streams) and (b) as SWIZZLE_128B atoms, and with/without concurrent TMA refills of B.
#include <cstdio>
#include <cstdlib>
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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// mode & 3: B layout (0 = no swizzle, 16 KB slices; 1 = SWIZZLE_128B atoms, 32 KB slots (2 of them); 2 = SWIZZLE_64B, 16 KB)
// mode & 4: stream B through the ring with TMA while computing, else keep it resident
// nmma_n: MMA N (64, 128 or 256)
__global__ void __launch_bounds__(128, 1) probe(const unsigned char* gsrc, int reps, int mode, int nmma_n, unsigned long long* stats) {
  extern __shared__ unsigned char smem_un[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_un) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;               // 8 atoms x 16 KB = 128 KB (garbage data is fine: we only time)
  unsigned char* sB = smem + 131072;      // 4 x 16 KB ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 32, bar_done = bar_empty + 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blayout = mode & 3;
  const bool stream = mode & 4;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (131072 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int slices = reps * 32;  // one "layer" = 32 slices of 2 MMAs

  if (warp == 0 && lane == 0 && stream) {
    for (int it = 0; it < slices; ++it) {
      const int s = it & 3, ph = (it >> 2) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      mbar_expect_tx(bar_full + 8 * s, 16384);
      bulk_g2s(smem_u32(sB + s * 16384), gsrc + size_t(it % 182) * 16384, 16384, bar_full + 8 * s);
    }
  } else if (warp == 1 && (mode & 8)) {
    // whole warp converged; descriptors are warp-uniform; one elected lane issues (the CUTLASS pattern)
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nmma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    const uint64_t b_hi = (blayout == 2) ? (((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61) | ((uint64_t)1 << 16))
                                         : (((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 16));
    const uint32_t b_step = (blayout == 2) ? 2 : 16;  // (32 B or 256 B) >> 4 per K = 16 step
    const uint32_t a0 = smem_u32(sA) >> 4, b0 = smem_u32(sB) >> 4;
    const unsigned long long c0 = clock64();
    for (int it = 0; it < slices; ++it) {
      const int s = it & 3, ph = (it >> 2) & 1;
      if (stream) { mbar_wait(bar_full + 8 * s, ph); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      const int ks = (it >> 1) & 15, nb = it & 1;
      const uint32_t a_lo = a0 + (ks >> 1) * 1024 + (ks & 1) * 4;
      const uint32_t b_lo = b0 + s * 1024;
      if (elect_one()) {
        umma_f16(tmem + nb * 256, a_hi | a_lo, b_hi | b_lo, idesc, ks != 0);
        umma_f16(tmem + nb * 256, a_hi | (a_lo + 2), b_hi | (b_lo + b_step), idesc, 1);
        if (stream) umma_commit(bar_empty + 8 * s);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
    mbar_wait(bar_done, 0);
    const unsigned long long c1 = clock64();
    if (lane == 0) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = (unsigned long long)slices * 2; }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nmma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const unsigned long long c0 = clock64();
    for (int it = 0; it < slices; ++it) {
      const int s = it & 3, ph = (it >> 2) & 1;
      if (stream) { mbar_wait(bar_full + 8 * s, ph); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      const int ks = (it >> 1) & 15, nb = it & 1;
      const uint32_t a_base = smem_u32(sA) + (ks >> 1) * 16384 + (ks & 1) * 64;
      const uint32_t b_base = (blayout == 1) ? smem_u32(sB + (s & 1) * 32768) : smem_u32(sB + s * 16384);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint64_t bdesc = (blayout == 1)   ? make_desc(b_base + j * 32, 16, 1024, 2)   // 256 rows x 128 B, SW128
                               : (blayout == 2) ? make_desc(b_base + j * 32, 16, 512, 4)    // 256 rows x 64 B, SW64
                                                : make_desc(b_base + j * 256, 128, 512, 0); // core matrices, no swizzle
        umma_f16(tmem + nb * 256, make_desc(a_base + j * 32, 16, 1024, 2), bdesc, idesc, (ks | j) != 0);
      }
      if (stream) umma_commit(bar_empty + 8 * s);
    }
    umma_commit(bar_done);
    mbar_wait(bar_done, 0);
    const unsigned long long c1 = clock64();
    stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = (unsigned long long)slices * 2;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  unsigned char* dsrc; unsigned long long* dStats;
  cudaMalloc(&dsrc, 182 * 16384); cudaMemset(dsrc, 0, 182 * 16384); cudaMalloc(&dStats, 148 * 16);
  const size_t smem = 131072 + 65536 + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int n : {256, 128, 64})
    for (int mode : {0, 8, 10, 12, 14})
      for (int grid : {148}) {
        probe<<<grid, 128, smem>>>(dsrc, 20, mode, n, dStats); cudaDeviceSynchronize();
        probe<<<grid, 128, smem>>>(dsrc, 20, mode, n, dStats);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s -- stopping\n", cudaGetErrorString(e)); return 1; }
        unsigned long long h[2]; cudaMemcpy(h, dStats, 16, cudaMemcpyDeviceToHost);
        printf("N=%3d B-layout=%-10s %-26s grid=%3d: %.1f cycles per MMA (M=128,K=16 f16) = %.0f MAC/clk/SM  [%s]\n", n, (mode & 3) == 1 ? "SW128" : (mode & 3) == 2 ? "SW64" : "no-swizzle",
               (mode & 4) ? ((mode & 8) ? "TMA-streamed/warp-issue" : "TMA-streamed") : ((mode & 8) ? "smem-resident/warp-issue" : "smem-resident"), grid, (double)h[0] / h[1], 128.0 * n * 16 * h[1] / h[0], cudaGetErrorString(e));
      }
  return 0;
}
