// How many thread-block clusters of 2 / 4 / 8 CTAs with one CTA per SM (200 KB of dynamic shared memory, the footprint of the
// CTA-pair GEMM kernels) can be co-resident on this GPU?  A cluster must sit inside one GPC, so a cluster size that does not
// divide a GPC's SM count leaves SMs idle: this bounds what weight multicast across two CTA pairs could gain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu && ./cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: %3d clusters co-resident = %3d SMs busy (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
