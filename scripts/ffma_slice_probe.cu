// Inner-loop probe of the fp32 kernels: the K-slice loop of nfb_ffma_groups_kernel (fragments from shared memory +
// FFMA2 block) in isolation -- no ring, no barriers, no epilogue -- so that the FMA-pipe rate of the loop body itself
// can be read off, variant by variant.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I include -I neuralfoil_b200/csrc \
//          -o /tmp/ffma_slice_probe scripts/ffma_slice_probe.cu
// Prints FFMA lanes per clock per SM (128 = nominal) per variant.  VARIANT:
//   0  ffma_slice as shipped (j outer, i inner; ptxas places everything)
//   1  snake order: the row index runs 0..7, 7..0, ... so that one source operand (the weight) carries over every
//      change of the column pair
//   2  snake order pinned with asm volatile
//   3  i outer snake: the weight is the reused operand, column pairs stream past it
//   4  as 2, fragments double-buffered by hand (loads of k + 1 issued before the block of k)
//   5  transposed pairing: accumulator pair = two rows of one column; weight pair reused, column scalars stream
//   6  transposed pairing, other nesting (column scalar reused, weight pairs stream)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "nfb_ffma.cuh"

using namespace nfb;

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ void ffma2_pin(float2& acc, float w, const float2& a) {
  asm volatile(
      "{\n\t.reg .b64 ww, aa, cc;\n\t"
      "mov.b64 ww, {%2, %2};\n\t"
      "mov.b64 aa, {%3, %4};\n\t"
      "mov.b64 cc, {%0, %1};\n\t"
      "fma.rn.f32x2 cc, ww, aa, cc;\n\t"
      "mov.b64 {%0, %1}, cc;\n\t}"
      : "+f"(acc.x), "+f"(acc.y)
      : "f"(w), "f"(a.x), "f"(a.y));
}

template <int VAR, int TM, int TN>
__device__ __forceinline__ void block(const float (&wv)[TM], const float2 (&av)[TN / 2], float2 (&acc)[TM][TN / 2]) {
  if constexpr (VAR == 1) {
#pragma unroll
    for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
      for (int ii = 0; ii < TM; ++ii) {
        const int i = (jj & 1) ? TM - 1 - ii : ii;
        acc[i][jj] = __ffma2_rn(make_float2(wv[i], wv[i]), av[jj], acc[i][jj]);
      }
  } else if constexpr (VAR == 2 || VAR == 4) {
#pragma unroll
    for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
      for (int ii = 0; ii < TM; ++ii) {
        const int i = (jj & 1) ? TM - 1 - ii : ii;
        ffma2_pin(acc[i][jj], wv[i], av[jj]);
      }
  } else if constexpr (VAR == 3) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j2 = 0; j2 < TN / 2; ++j2) {
        const int jj = (i & 1) ? TN / 2 - 1 - j2 : j2;
        ffma2_pin(acc[i][jj], wv[i], av[jj]);
      }
  }
}

// Transposed pairing: an accumulator pair is (row 2 ip, row 2 ip + 1) of ONE column; the weight pair is the 64-bit operand that
// stays in the reuse cache while the TN column scalars stream past it (VAR 5), or the other nesting (VAR 6, calibration).
template <int VAR, int TM, int TN, int MP, int NC, int KC>
__device__ __forceinline__ void slice_t(const float* __restrict__ w, const float* __restrict__ a, float2 (&acc)[TM / 2][TN]) {
#pragma unroll
  for (int kk = 0; kk < KC; ++kk) {
    float2 wp[TM / 2];
    float av[TN];
    {
      const float4 t = *reinterpret_cast<const float4*>(w + kk * MP);
      wp[0] = make_float2(t.x, t.y); wp[1] = make_float2(t.z, t.w);
      const float4 u = *reinterpret_cast<const float4*>(w + kk * MP + MP / 2);
      wp[2] = make_float2(u.x, u.y); wp[3] = make_float2(u.z, u.w);
    }
#pragma unroll
    for (int v = 0; v < TN / 8; ++v) {
      const float4 t = *reinterpret_cast<const float4*>(a + kk * NC + 4 * v);
      av[4 * v] = t.x; av[4 * v + 1] = t.y; av[4 * v + 2] = t.z; av[4 * v + 3] = t.w;
      const float4 u = *reinterpret_cast<const float4*>(a + kk * NC + NC / 2 + 4 * v);
      av[TN / 2 + 4 * v] = u.x; av[TN / 2 + 4 * v + 1] = u.y; av[TN / 2 + 4 * v + 2] = u.z; av[TN / 2 + 4 * v + 3] = u.w;
    }
    if constexpr (VAR == 5) {
#pragma unroll
      for (int ip = 0; ip < TM / 2; ++ip)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[ip][j] = __ffma2_rn(wp[ip], make_float2(av[j], av[j]), acc[ip][j]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int ip = 0; ip < TM / 2; ++ip) acc[ip][j] = __ffma2_rn(wp[ip], make_float2(av[j], av[j]), acc[ip][j]);
    }
  }
}

template <int VAR, int TM, int TN, int MP, int NC, int KC>
__device__ __forceinline__ void slice(const float* __restrict__ w, const float* __restrict__ a, float2 (&acc)[TM][TN / 2]) {
  if constexpr (VAR == 0) {
    ffma_slice<TM, TN, MP, NC, KC>(w, a, acc);
  } else if constexpr (VAR == 7 || VAR == 8) {  // no loads inside the loop: one fragment set per slice, registers only
    float wv[TM];
    float2 av[TN / 2];
    load_fragments<TM, TN, MP, NC>(w, a, wv, av);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
#pragma unroll
      for (int jj = 0; jj < TN / 2; ++jj)
#pragma unroll
        for (int ii = 0; ii < TM; ++ii) {
          const int i = (VAR == 8 && (jj & 1)) ? TM - 1 - ii : ii;
          acc[i][jj] = __ffma2_rn(make_float2(wv[i], wv[i]), av[jj], acc[i][jj]);
        }
      wv[kk % TM] = -wv[kk % TM];
    }
  } else if constexpr (VAR == 4) {
    float wv[2][TM];
    float2 av[2][TN / 2];
    load_fragments<TM, TN, MP, NC>(w, a, wv[0], av[0]);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const int cur = kk & 1;
      if (kk + 1 < KC) load_fragments<TM, TN, MP, NC>(w + (kk + 1) * MP, a + (kk + 1) * NC, wv[cur ^ 1], av[cur ^ 1]);
      block<VAR, TM, TN>(wv[cur], av[cur], acc);
    }
  } else {
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float wv[TM];
      float2 av[TN / 2];
      load_fragments<TM, TN, MP, NC>(w + kk * MP, a + kk * NC, wv, av);
      block<VAR, TM, TN>(wv, av, acc);
    }
  }
}

template <int VAR, int TN>
__global__ void __launch_bounds__(128 / TN * 32, 1) probe(float* sink, int iters, unsigned long long* stats) {
  constexpr int H = 512, NC = 64, KC = 8, QG = 16, QT = TN / 2;
  constexpr int LN = 2 * QG / TN, LM = 32 / LN, WM = H / (8 * LM);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const act = reinterpret_cast<float*>(smem_raw);  // [H][NC]
  float* const ring = act + H * NC;                        // [2][KC * H]
  for (int i = threadIdx.x; i < H * NC + 2 * KC * H; i += blockDim.x) act[i] = 1e-3f * float((i * 37) % 101 - 50);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lm = lane / LN, ln = lane % LN, wm = warp % WM, wn = warp / WM;
  const int h_row = (wm * LM + lm) * 4, h_col = wn * QG + ln * QT;
  float2 acc[8][TN / 2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);
  __syncthreads();
  const unsigned long long c0 = clock64(), t0 = gtime();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int c = 0; c < H / KC; ++c) {
      if constexpr (VAR == 5 || VAR == 6)
        slice_t<VAR, 8, TN, H, NC, KC>(ring + (c & 1) * KC * H + h_row, act + (c * KC) * NC + h_col, reinterpret_cast<float2(&)[4][TN]>(acc));
      else
        slice<VAR, 8, TN, H, NC, KC>(ring + (c & 1) * KC * H + h_row, act + (c * KC) * NC + h_col, acc);
    }
  }
  const unsigned long long c1 = clock64(), t1 = gtime();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) s += acc[i][j].x + acc[i][j].y;
  atomicAdd(sink + 1, s);
  if (s == 123.456f) sink[0] = s;
  __syncthreads();
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = c1 - c0; stats[2 * blockIdx.x + 1] = t1 - t0; }
}

template <int VAR, int TN>
void run(int sms, int iters, float* sink, unsigned long long* stats) {
  constexpr int threads = 128 / TN * 32;
  const size_t smem = sizeof(float) * (512 * 64 + 2 * 8 * 512);
  cudaFuncSetAttribute(probe<VAR, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<VAR, TN><<<sms, threads, smem>>>(sink, iters, stats); cudaDeviceSynchronize();
  cudaEventRecord(e0); probe<VAR, TN><<<sms, threads, smem>>>(sink, iters, stats); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[2]; cudaMemcpy(h, stats, sizeof(h), cudaMemcpyDeviceToHost);
  float chk; cudaMemcpy(&chk, sink + 1, 4, cudaMemcpyDeviceToHost); cudaMemset(sink, 0, 8);
  printf("[checksum %.9e] ", chk);
  const double ffma_per_thread = 2.0 * 8 * (TN / 2) * 512.0 * iters;  // scalar FMAs: 2 per FFMA2
  const double lanes = ffma_per_thread * threads / double(h[0]);
  printf("variant %d  tile 8x%-2d  thr=%3d  FFMA lanes/clk/SM=%6.1f (%.1f %% of 128)  SM clock %.3f GHz  %.1f TFLOP/s  err=%s\n", VAR, TN, threads, lanes,
         lanes / 1.28, double(h[0]) / double(h[1]), 2.0 * ffma_per_thread * threads * sms / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  float* sink; cudaMalloc(&sink, 8); cudaMemset(sink, 0, 8);
  unsigned long long* stats; cudaMalloc(&stats, 16 * sms * 8);
  printf("%s, %d SMs\n", p.name, sms);
  for (int rep = 0; rep < 1; ++rep) {
    run<0, 16>(sms, iters, sink, stats);
    run<1, 16>(sms, iters, sink, stats);
    run<2, 16>(sms, iters, sink, stats);
    run<3, 16>(sms, iters, sink, stats);
    run<4, 16>(sms, iters, sink, stats);
    run<5, 16>(sms, iters, sink, stats);
    run<6, 16>(sms, iters, sink, stats);
    run<7, 16>(sms, iters, sink, stats);
    run<8, 16>(sms, iters, sink, stats);
    run<0, 8>(sms, iters, sink, stats);
    run<1, 8>(sms, iters, sink, stats);
    run<2, 8>(sms, iters, sink, stats);
    run<3, 8>(sms, iters, sink, stats);
    run<4, 8>(sms, iters, sink, stats);
    run<5, 8>(sms, iters, sink, stats);
    run<6, 8>(sms, iters, sink, stats);
    run<7, 8>(sms, iters, sink, stats);
    run<8, 8>(sms, iters, sink, stats);
  }
  return 0;
}
