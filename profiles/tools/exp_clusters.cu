// Experiment: how many thread-block clusters of size C (CTA = `threads` threads, `smem` bytes of dynamic
// shared memory) can be co-resident on this GPU. Build: nvcc -arch=sm_100a -o exp_clusters exp_clusters.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void dummy(float* p) {
  extern __shared__ float s[];
  s[threadIdx.x] = p ? p[threadIdx.x] : 0.f;
  __syncthreads();
  if (p) p[threadIdx.x] = s[(threadIdx.x + 1) % blockDim.x];
}

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 640;
  const int smem = argc > 2 ? atoi(argv[2]) : 141312;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs; CTA %d threads, %d B smem\n", prop.name, prop.multiProcessorCount, threads, smem);
  for (int C = 1; C <= 16; ++C) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C * 64);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = C;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("  cluster %2d: max active clusters %3d  (%3d SMs)  %s\n", C, n, n * C, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
