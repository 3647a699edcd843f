// tcgen05 M = 64 probe (kind::f16, cta_group::1): where do the 64 accumulator rows land in tensor memory, can a second
// M = 64 tile be interleaved into lanes 16..31 of each quadrant, and what does an M = 64, N = 256, K = 16 MMA cost?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_probe_m64 scripts/umma_probe_m64.cu
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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

constexpr int MROWS = 64, NCOLS = 256, KTOT = 64;
// A: 64 rows x 64 k fp16, SWIZZLE_128B K-major atom (8 row groups x 1024 B = 8 KB)
// B0, B1: 256 rows x 64 k as two 32-k no-swizzle slices of 16 KB each (element (r, kk) at (r/8)*512 + (kk/8)*128 + (r%8)*16 + (kk%8)*2)
__global__ void __launch_bounds__(192, 1) probe(const unsigned char* gA, const unsigned char* gB0, const unsigned char* gB1, float* out /*[128 lanes][256]*/,
                                                int reps, unsigned long long* stats) {
  extern __shared__ unsigned char smem_un[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_un) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                 // 8 KB
  unsigned char* sB0 = smem + 8192;         // 32 KB
  unsigned char* sB1 = sB0 + 32768;         // 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB1 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar_done = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(gA)[i];
  for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) { reinterpret_cast<uint4*>(sB0)[i] = reinterpret_cast<const uint4*>(gB0)[i]; reinterpret_cast<uint4*>(sB1)[i] = reinterpret_cast<const uint4*>(gB1)[i]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { mbar_init(bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NCOLS >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);   // M = 64
    const uint64_t a_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t b_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t a0 = (smem_u32(sA) >> 4) & 0x3FFF, b00 = (smem_u32(sB0) >> 4) & 0x3FFF, b10 = (smem_u32(sB1) >> 4) & 0x3FFF;
    const unsigned long long c0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // K = 64 in four steps of 16: slice ks/2, step ks%2
          umma_f16(tmem, a_hi | (a0 + ks * 2), b_hi | (b00 + (ks >> 1) * 1024 + (ks & 1) * 16), idesc, ks != 0);                       // tile 0 -> lanes 0..15 of each quadrant
          umma_f16(tmem + (16u << 16), a_hi | (a0 + ks * 2), b_hi | (b10 + (ks >> 1) * 1024 + (ks & 1) * 16), idesc, ks != 0);         // tile 1 -> lanes 16..31 (interleaved)
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
    mbar_wait(bar_done, 0);
    const unsigned long long c1 = clock64();
    if (lane == 0) { stats[0] = c1 - c0; stats[1] = (unsigned long long)reps * 8; }
  } else if (warp >= 2) {
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    for (int c0 = 0; c0 < NCOLS; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      for (int c = 0; c < 32; ++c) out[(q * 32 + lane) * NCOLS + c0 + c] = __uint_as_float(v[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  std::vector<float> A(MROWS * KTOT), B0(NCOLS * KTOT), B1(NCOLS * KTOT);
  srand(3);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
  for (auto& x : B0) x = (rand() / (float)RAND_MAX - 0.5f) * 0.5f;
  for (auto& x : B1) x = (rand() / (float)RAND_MAX - 0.5f) * 0.5f;
  std::vector<__half> hA(8192 / 2), hB0(32768 / 2), hB1(32768 / 2);
  for (int r = 0; r < MROWS; ++r)
    for (int k = 0; k < KTOT; ++k)
      hA[((r / 8) * 1024 + (r % 8) * 128 + (((k / 8) ^ (r % 8)) * 16) + (k % 8) * 2) / 2] = __float2half_rn(A[r * KTOT + k]);
  for (int r = 0; r < NCOLS; ++r)
    for (int k = 0; k < KTOT; ++k) {
      const size_t off = size_t(k / 32) * 16384 + (r / 8) * 512 + ((k % 32) / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
      hB0[off / 2] = __float2half_rn(B0[r * KTOT + k]);
      hB1[off / 2] = __float2half_rn(B1[r * KTOT + k]);
    }
  unsigned char *dA, *dB0, *dB1; float* dOut; unsigned long long* dStats;
  cudaMalloc(&dA, 8192); cudaMalloc(&dB0, 32768); cudaMalloc(&dB1, 32768); cudaMalloc(&dOut, 128 * NCOLS * 4); cudaMalloc(&dStats, 16);
  cudaMemcpy(dA, hA.data(), 8192, cudaMemcpyHostToDevice); cudaMemcpy(dB0, hB0.data(), 32768, cudaMemcpyHostToDevice); cudaMemcpy(dB1, hB1.data(), 32768, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0, 128 * NCOLS * 4);
  const size_t smem = 8192 + 65536 + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 192, smem>>>(dA, dB0, dB1, dOut, 1, dStats);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> out(128 * NCOLS);
  cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
  auto ref = [&](const std::vector<float>& B, int r, int n) { double s = 0; for (int k = 0; k < KTOT; ++k) s += (double)__half2float(__float2half_rn(A[r * KTOT + k])) * __half2float(__float2half_rn(B[n * KTOT + k])); return s; };
  // expected: tile t (0/1), row r -> lane 32*(r/16) + 16*t + r%16
  double worst = 0; int bad = 0;
  for (int t = 0; t < 2; ++t)
    for (int r = 0; r < MROWS; ++r)
      for (int n = 0; n < NCOLS; ++n) {
        const int lane = 32 * (r / 16) + 16 * t + (r % 16);
        const double d = fabs(out[lane * NCOLS + n] - ref(t ? B1 : B0, r, n));
        worst = fmax(worst, d); bad += d > 1e-3;
      }
  printf("layout 'row r of tile t -> lane 32*(r/16) + 16*t + r%%16': max |err| %.3e, mismatches %d of %d -> %s\n", worst, bad, 2 * MROWS * NCOLS, bad ? "WRONG LAYOUT" : "CONFIRMED");
  if (bad) {  // brute-force search where row 0 / row 17 of tile 0 landed
    for (int r : {0, 17, 40}) for (int lane = 0; lane < 128; ++lane) { bool ok = true; for (int n = 0; n < 8; ++n) ok &= fabs(out[lane * NCOLS + n] - ref(B0, r, n)) < 1e-3; if (ok) printf("  tile 0 row %d found at lane %d\n", r, lane); }
  }
  for (int rep : {2000}) {
    probe<<<148, 192, smem>>>(dA, dB0, dB1, dOut, rep, dStats); cudaDeviceSynchronize();
    probe<<<148, 192, smem>>>(dA, dB0, dB1, dOut, rep, dStats);
    e = cudaDeviceSynchronize();
    unsigned long long h[2]; cudaMemcpy(h, dStats, 16, cudaMemcpyDeviceToHost);
    printf("M=64 N=256 K=16 f16: %.1f cycles per MMA (two interleaved accumulator tiles), %.0f MAC/clk/SM [%s]\n", (double)h[0] / h[1], 64.0 * 256 * 16 * h[1] / h[0], cudaGetErrorString(e));
  }
  return 0;
}
