// FFMA issue-rate microbenchmark (build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma_peak scripts/ffma_peak.cu)
// Reports FFMA lanes per clock per SM (cycle-based, clock independent) and the SM clock the kernel ran at.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// V1: 8x8 outer product, the GEMM inner-loop pattern
template <int UNUSED> __global__ void __launch_bounds__(512, 1) k_outer(float* sink, int iters, unsigned long long* stats) {
  float a[8], b[8], acc[8][8];
  for (int i = 0; i < 8; ++i) { a[i] = 1.0f + 1e-3f * (threadIdx.x + i); b[i] = 1.0f - 1e-3f * (threadIdx.x * 3 + i); for (int j = 0; j < 8; ++j) acc[i][j] = 0.f; }
  __syncthreads();
  unsigned long long c0 = clock64(), t0 = gtime();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      a[u] = -a[u];
    }
  }
  unsigned long long c1 = clock64(), t1 = gtime();
  float s = 0.f; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = t1 - t0; }
}

// V2: N independent chains x = x * a + b (two shared source registers)
template <int NCHAIN> __global__ void __launch_bounds__(1024, 1) k_chain(float* sink, int iters, unsigned long long* stats) {
  float x[NCHAIN]; float a = 1.0f + 1e-6f * threadIdx.x, b = 1e-3f;
  for (int i = 0; i < NCHAIN; ++i) x[i] = i;
  __syncthreads();
  unsigned long long c0 = clock64(), t0 = gtime();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 512 / NCHAIN; ++u)
#pragma unroll
      for (int i = 0; i < NCHAIN; ++i) x[i] = fmaf(x[i], a, b);
  }
  unsigned long long c1 = clock64(), t1 = gtime();
  float s = 0.f; for (int i = 0; i < NCHAIN; ++i) s += x[i];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = t1 - t0; }
}

// V3: 8x8 outer product with packed FFMA2 (fma.rn.f32x2): column pairs, the row operand duplicated into a pair
template <int MODE> __global__ void __launch_bounds__(512, 1) k_outer2(float* sink, int iters, unsigned long long* stats) {
  float a[8]; float2 b2[4], acc2[8][4];
  for (int i = 0; i < 8; ++i) { a[i] = 1.0f + 1e-3f * (threadIdx.x + i); }
  for (int j = 0; j < 4; ++j) b2[j] = make_float2(1.0f - 1e-3f * (threadIdx.x * 3 + j), 1.0f - 2e-3f * (threadIdx.x + j));
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
  __syncthreads();
  unsigned long long c0 = clock64(), t0 = gtime();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ap = make_float2(a[i], a[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[i][j] = __ffma2_rn(ap, b2[j], acc2[i][j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 ap = make_float2(a[i], a[i]);
            acc2[i][j] = __ffma2_rn(ap, b2[j], acc2[i][j]);
          }
      }
      a[u] = -a[u];
    }
  }
  unsigned long long c1 = clock64(), t1 = gtime();
  float s = 0.f; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc2[i][j].x + acc2[i][j].y;
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = t1 - t0; }
}

// V4: FFMA2 with the pair operand pre-duplicated (as if loaded duplicated from shared memory): no MOVs
template <int UNUSED> __global__ void __launch_bounds__(512, 1) k_outer2_predup(float* sink, int iters, unsigned long long* stats) {
  float2 a2[8], b2[4], acc2[8][4];
  for (int i = 0; i < 8; ++i) { float v = 1.0f + 1e-3f * (threadIdx.x + i); a2[i] = make_float2(v, v); }
  for (int j = 0; j < 4; ++j) b2[j] = make_float2(1.0f - 1e-3f * (threadIdx.x * 3 + j), 1.0f - 2e-3f * (threadIdx.x + j));
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);
  __syncthreads();
  unsigned long long c0 = clock64(), t0 = gtime();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = __ffma2_rn(a2[i], b2[j], acc2[i][j]);
      a2[u].x = -a2[u].x; a2[u].y = -a2[u].y;
    }
  }
  unsigned long long c1 = clock64(), t1 = gtime();
  float s = 0.f; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc2[i][j].x + acc2[i][j].y;
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = t1 - t0; }
}

// V5: 8x16 FFMA2 tile (128 accumulators), pre-duplicated row operand
template <int UNUSED> __global__ void __launch_bounds__(256, 1) k_outer2_8x16(float* sink, int iters, unsigned long long* stats) {
  float2 a2[8], b2[8], acc2[8][8];
  for (int i = 0; i < 8; ++i) { float v = 1.0f + 1e-3f * (threadIdx.x + i); a2[i] = make_float2(v, v); }
  for (int j = 0; j < 8; ++j) b2[j] = make_float2(1.0f - 1e-3f * (threadIdx.x * 3 + j), 1.0f - 2e-3f * (threadIdx.x + j));
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) acc2[i][j] = make_float2(0.f, 0.f);
  __syncthreads();
  unsigned long long c0 = clock64(), t0 = gtime();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc2[i][j] = __ffma2_rn(a2[i], b2[j], acc2[i][j]);
      a2[u].x = -a2[u].x; a2[u].y = -a2[u].y;
    }
  }
  unsigned long long c1 = clock64(), t1 = gtime();
  float s = 0.f; for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc2[i][j].x + acc2[i][j].y;
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = t1 - t0; }
}

template <class K> void run(const char* name, K kern, int threads, int blocks_per_sm, int sms, int iters, float* sink, unsigned long long* stats) {
  int blocks = sms * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(sink, iters, stats); cudaDeviceSynchronize();
  cudaEventRecord(e0); kern<<<blocks, threads>>>(sink, iters, stats); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[2]; cudaMemcpy(h, stats, sizeof(h), cudaMemcpyDeviceToHost);
  double ffma_per_thread = 512.0 * iters;
  double lanes_per_clk_sm = ffma_per_thread * threads * blocks_per_sm / double(h[0]);
  double ghz = double(h[0]) / double(h[1]);
  double tflops = 2.0 * ffma_per_thread * threads * blocks / (ms * 1e-3) / 1e12;
  printf("%-28s thr=%4d blk/SM=%d  FFMA lanes/clk/SM=%6.1f  SM clock=%.3f GHz  event-timed %.1f TFLOP/s (%.2f ms) err=%s\n", name, threads, blocks_per_sm,
         lanes_per_clk_sm, ghz, tflops, ms, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  float* sink; cudaMalloc(&sink, 4);
  unsigned long long* stats; cudaMalloc(&stats, 16 * sms * 8);
  printf("%s, %d SMs, clockRate %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
  for (int rep = 0; rep < 2; ++rep) {
    run("outer8x8", k_outer<0>, 512, 1, sms, iters, sink, stats);
    run("outer8x8", k_outer<0>, 256, 1, sms, iters, sink, stats);
    run("outer8x8", k_outer<0>, 128, 1, sms, iters, sink, stats);
    run("outer8x8 ffma2 mov i-outer", k_outer2<0>, 512, 1, sms, iters, sink, stats);
    run("outer8x8 ffma2 mov i-outer", k_outer2<0>, 256, 1, sms, iters, sink, stats);
    run("outer8x8 ffma2 mov j-outer", k_outer2<1>, 512, 1, sms, iters, sink, stats);
    run("outer8x8 ffma2 predup", k_outer2_predup<0>, 512, 1, sms, iters, sink, stats);
    run("outer8x8 ffma2 predup", k_outer2_predup<0>, 256, 1, sms, iters, sink, stats);
    run("outer8x16 ffma2 predup", k_outer2_8x16<0>, 256, 1, sms, iters, sink, stats);
    run("outer8x16 ffma2 predup", k_outer2_8x16<0>, 128, 1, sms, iters, sink, stats);
    
    run("chain16", k_chain<16>, 1024, 1, sms, iters, sink, stats);
    run("chain8", k_chain<8>, 1024, 2, sms, iters, sink, stats);
    run("chain16 1 SM-quarter", k_chain<16>, 128, 1, sms, iters, sink, stats);
  }
  // a long run to let the power governor settle
  run("outer8x8 long", k_outer<0>, 512, 1, sms, iters * 20, sink, stats);
  run("chain16 long", k_chain<16>, 1024, 1, sms, iters * 20, sink, stats);
  return 0;
}
