// tcgen05 cta_group::2 probe (kind::f16): one MMA over a CTA pair, M = 128 (64 rows of A per CTA), N = 256 with B split
// (128 rows per CTA), K = 16.  Checks where D lands in each CTA's tensor memory and what an MMA costs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe_2sm scripts/umma_probe_2sm.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
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
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
        "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int KTOT = 64;
#ifndef ROWS
#define ROWS 64   // A rows per CTA: 64 (pair M = 128) or 128 (pair M = 256)
#endif
constexpr int A_BYTES = ROWS * 128;      // SW128 atom: ROWS x 64 fp16
constexpr int OUT_COLS = (ROWS == 64) ? 128 : 256;
// per CTA: A 64 rows x 64 k (SW128 atom, 8 KB); B half: 128 rows x 64 k as two 32-k no-swizzle slices of 8 KB
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
probe(const unsigned char* gA /*[2][8 KB]*/, const unsigned char* gB /*[2][16 KB]*/, float* out /*[2][128 lanes][128 cols]*/, int reps, unsigned long long* stats) {
  extern __shared__ unsigned char smem_un[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_un) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;           // 8 KB
  unsigned char* sB = smem + A_BYTES;    // 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar_done = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  for (int i = threadIdx.x; i < A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(gA + rank * A_BYTES)[i];
  for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(gB + rank * 16384)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { mbar_init(bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // both CTAs' operands and barriers are in place
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 1 && rank == 0) {  // the leader CTA issues for the pair
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)((2 * ROWS) >> 4) << 24);   // M = 2 x ROWS over two CTAs, N = 256
    const uint64_t a_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t b_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t a0 = (smem_u32(sA) >> 4) & 0x3FFF, b0 = (smem_u32(sB) >> 4) & 0x3FFF;
    const unsigned long long c0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16_2sm(tmem, a_hi | (a0 + ks * 2), b_hi | (b0 + (ks >> 1) * 512 + (ks & 1) * 16), idesc, ks != 0);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit_2sm(bar_done, (uint16_t)3);
    __syncwarp();
    mbar_wait(bar_done, 0);
    const unsigned long long c1 = clock64();
    if (lane == 0) { stats[2 * pair] = c1 - c0; stats[2 * pair + 1] = (unsigned long long)reps * 4; }
  } else if (warp >= 2) {
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    if (pair == 0)
      for (int c0 = 0; c0 < OUT_COLS; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
        for (int c = 0; c < 32; ++c) out[(rank * 128 + q * 32 + lane) * OUT_COLS + c0 + c] = __uint_as_float(v[c]);
      }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  std::vector<float> A(2 * ROWS * KTOT), B(256 * KTOT);
  srand(5);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 0.5f;
  std::vector<__half> hA(2 * A_BYTES / 2), hB(2 * 16384 / 2);
  for (int c = 0; c < 2; ++c)
    for (int r = 0; r < ROWS; ++r)
      for (int k = 0; k < KTOT; ++k)
        hA[c * (A_BYTES / 2) + ((r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2) / 2] = __float2half_rn(A[(c * ROWS + r) * KTOT + k]);
  for (int n = 0; n < 256; ++n)
    for (int k = 0; k < KTOT; ++k) {
      const int c = n / 128, r = n % 128;
      const size_t off = size_t(c) * 16384 + size_t(k / 32) * 8192 + (r / 8) * 512 + ((k % 32) / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
      hB[off / 2] = __float2half_rn(B[n * KTOT + k]);
    }
  unsigned char *dA, *dB; float* dOut; unsigned long long* dStats;
  cudaMalloc(&dA, 2 * A_BYTES); cudaMalloc(&dB, 2 * 16384); cudaMalloc(&dOut, 2 * 128 * OUT_COLS * 4); cudaMalloc(&dStats, 74 * 16);
  cudaMemcpy(dA, hA.data(), 2 * A_BYTES, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), 2 * 16384, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0, 2 * 128 * OUT_COLS * 4);
  const size_t smem = A_BYTES + 16384 + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<2, 192, smem>>>(dA, dB, dOut, 1, dStats);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> out(2 * 128 * OUT_COLS);
  cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
  auto ref = [&](int c, int r, int n) { double s = 0; for (int k = 0; k < KTOT; ++k) s += (double)__half2float(__float2half_rn(A[(c * ROWS + r) * KTOT + k])) * __half2float(__float2half_rn(B[n * KTOT + k])); return s; };
  // hypothesis (CUTLASS "2x2" layout): CTA c, row r, column n -> lane r + 64 * (n / 128), TMEM column n % 128
  int bad = 0; double worst = 0;
  for (int c = 0; c < 2; ++c) for (int r = 0; r < ROWS; ++r) for (int n = 0; n < 256; ++n) {
    const int lane_ = (ROWS == 64) ? r + 64 * (n / 128) : r, col_ = (ROWS == 64) ? n % 128 : n;
    const double d = fabs(out[(c * 128 + lane_) * OUT_COLS + col_] - ref(c, r, n));
    worst = fmax(worst, d); bad += d > 1e-3;
  }
  printf("ROWS %d per CTA; layout '%s': max |err| %.3e, mismatches %d of %d -> %s\n", ROWS,
         ROWS == 64 ? "CTA c row r col n -> lane r + 64 (n / 128), column n %% 128" : "CTA c row r col n -> lane r, column n", worst, bad, 2 * ROWS * 256, bad ? "WRONG" : "CONFIRMED");
  if (bad) {
    for (int c = 0; c < 2; ++c) for (int r : {0, 5, 40}) for (int n : {0, 130})
      for (int cc = 0; cc < 2; ++cc) for (int lane = 0; lane < 128; ++lane) for (int col = 0; col < OUT_COLS; ++col)
        if (fabs(out[(cc * 128 + lane) * OUT_COLS + col] - ref(c, r, n)) < 1e-4 && fabs(ref(c, r, n)) > 0.05) printf("  A-cta %d row %d col %d found in cta %d lane %d column %d\n", c, r, n, cc, lane, col);
  }
  probe<<<148, 192, smem>>>(dA, dB, dOut, 2000, dStats); cudaDeviceSynchronize();
  probe<<<148, 192, smem>>>(dA, dB, dOut, 2000, dStats);
  e = cudaDeviceSynchronize();
  unsigned long long h[2]; cudaMemcpy(h, dStats, 16, cudaMemcpyDeviceToHost);
  printf("cta_group::2 M=%d (%d rows per CTA) N=256 K=16 f16: %.1f cycles per MMA = %.0f MAC/clk per SM [%s]\n", 2 * ROWS, ROWS, (double)h[0] / h[1], double(ROWS) * 256 * 16 * h[1] / h[0], cudaGetErrorString(e));
  return 0;
}
