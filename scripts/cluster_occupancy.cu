// How many clusters of 2 / 4 / 8 CTAs (1 CTA per SM, ~220 KB smem each) can be resident on this GPU?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/cluster_occupancy scripts/cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int smem = 226 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension; attr.val.clusterDim.x = cs; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %d (= %d CTAs) [%s]\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
