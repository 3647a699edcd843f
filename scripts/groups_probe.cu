// Timing harness of nfb_ffma_groups_kernel<Grp512> outside the library (synthetic weights and inputs, no parity check):
// compiled with -DNFB_EXP=<bits> it runs the kernel with phases switched off, to see what each phase costs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I include -I neuralfoil_b200/csrc [-DNFB_EXP=n] \
//        -o probe_bin/groups_probe_n scripts/groups_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "nfb_ffma_groups.cuh"

#ifndef PROBE_STAGES
#define PROBE_STAGES 5
#endif
#ifndef PROBE_TN
#define PROBE_TN 16
#endif
#ifndef PROBE_LEAD
#define PROBE_LEAD 0
#endif
#ifndef PROBE_H
#define PROBE_H 512
#endif
#ifndef PROBE_KC
#define PROBE_KC 8
#endif
#ifndef PROBE_L
#define PROBE_L 7   // affine layers: xxxlarge 7, xxlarge / xlarge 6, large / medium 5, small / xsmall 4, xxsmall 3
#endif
#ifndef PROBE_KH
#define PROBE_KH PROBE_H  // real hidden width rounded up to 16 (48 for xsmall / xxsmall)
#endif
#if PROBE_H == 512
using Cfg = nfb::GroupCfg<512, PROBE_KC, PROBE_STAGES, PROBE_TN, 16, PROBE_LEAD>;
#else
using Cfg = nfb::GroupCfg<PROBE_H, 8, PROBE_STAGES, PROBE_TN, 32, PROBE_LEAD>;
#endif

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 1000000;
  const int L = PROBE_L;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const size_t blob_floats = nfb::SweepBlob<PROBE_H>::total(L);
  std::vector<float> hb(blob_floats);
  unsigned s = 12345u;
  for (auto& v : hb) { s = s * 1664525u + 1013904223u; v = (float((s >> 8) & 0xffff) / 65536.0f - 0.5f) * 0.08f; }
  float* blob; cudaMalloc(&blob, blob_floats * 4); cudaMemcpy(blob, hb.data(), blob_floats * 4, cudaMemcpyHostToDevice);
  const long long ld = (n + 3) & ~3ll;
  float *in, *out; cudaMalloc(&in, 23 * ld * 4); cudaMalloc(&out, 198 * ld * 4);
  std::vector<float> hi(23 * ld);
  for (long long q = 0; q < ld; ++q) {
    for (int r = 0; r < 18; ++r) { s = s * 1664525u + 1013904223u; hi[r * ld + q] = (float((s >> 8) & 0xffff) / 65536.0f - 0.5f) * 0.4f; }
    hi[18 * ld + q] = float(q % 20) - 10.f; hi[19 * ld + q] = 1e6f; hi[20 * ld + q] = 9.f; hi[21 * ld + q] = 1.f; hi[22 * ld + q] = 1.f;
  }
  cudaMemcpy(in, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice);
  nfb::EvalParams p{};
  for (int r = 0; r < 23; ++r) { p.in[r] = in + r * ld; p.in_stride[r] = 1; }
  p.out = out; p.out_ld = ld; p.n = n; p.blob = blob; p.n_layers = L; p.in_f64 = 0; p.out_f64 = 0; p.scalars_only = 0; p.k_hidden = PROBE_KH;
  for (int i = 0; i < 25; ++i) p.mean[i] = 0.0;
  for (int i = 0; i < 325; ++i) p.quad[i] = 0.0;
  p.out_bl = nullptr;
  cudaFuncSetAttribute(nfb::nfb_ffma_groups_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM_BYTES));
  const long long tiles = (n + Cfg::NQ - 1) / Cfg::NQ;
  const int grid = int(tiles < prop.multiProcessorCount ? tiles : prop.multiProcessorCount);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    nfb::nfb_ffma_groups_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES>>>(p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  // output hash: patched and unpatched builds of the same source must agree bit for bit
  std::vector<uint32_t> ho(size_t(198) * ld);
  cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost);
  uint64_t hash = 1469598103934665603ull;
  for (uint32_t v : ho) { hash ^= v; hash *= 1099511628211ull; }
  printf("[output hash %016llx] ", (unsigned long long)hash);
  const double flop = 4.0 * (25.0 * PROBE_KH + double(L - 2) * PROBE_KH * PROBE_KH + 198.0 * PROBE_KH) * double(n);
  printf("H=%d L=%d NFB_EXP=%d stages=%d tn=%d lead=%d  n=%lld  %.3f ms  %.2f M q/s  %.1f TFLOP/s (%.1f %% of 74.4)  err=%s\n", PROBE_H, L, NFB_EXP, PROBE_STAGES, PROBE_TN, PROBE_LEAD, n, best,
         n / best / 1e3, flop / best / 1e9, flop / best / 1e9 / 74.4 * 100, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
