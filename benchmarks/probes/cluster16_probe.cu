// Probe: how many clusters of 8 / 16 CTAs (1 CTA per SM, ~140 KB of shared memory) are co-resident on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster16_probe cluster16_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(640, 1) probe_kernel(unsigned* counter, unsigned* smids, unsigned target, long long budget) {
  extern __shared__ unsigned char smem[];
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) {
    smids[blockIdx.x] = smid;
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    // co-residency check: every CTA waits until all CTAs of the grid have arrived (or the budget runs out)
    while (atomicAdd(counter, 0u) < target && clock64() - t0 < budget) { }
    smids[1024 + blockIdx.x] = atomicAdd(counter, 0u);
  }
  smem[threadIdx.x] = (unsigned char)smid;
  __syncthreads();
}

int main() {
  unsigned *counter, *smids;
  cudaMalloc(&counter, 4);
  cudaMalloc(&smids, 2048 * 4);
  for (int smem_kb : {100, 140, 200}) {
    for (int cs : {2, 4, 8, 16}) {
      cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
      if (cs > 8) cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg{};
      cfg.blockDim = dim3(640);
      cfg.gridDim = dim3(cs);
      cfg.dynamicSmemBytes = smem_kb * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
      printf("smem %d KB cluster %2d: max active clusters %d (%s)\n", smem_kb, cs, n, cudaGetErrorString(e));
      if (e != cudaSuccess) { cudaGetLastError(); continue; }
      // launch n clusters and n + 1 clusters: do all CTAs see each other?
      for (int extra = 0; extra <= 1; ++extra) {
        const int grid = (n + extra) * cs;
        if (grid > 1024) continue;
        cudaMemset(counter, 0, 4);
        cudaMemset(smids, 0xff, 2048 * 4);
        cfg.gridDim = dim3(grid);
        e = cudaLaunchKernelEx(&cfg, probe_kernel, counter, smids, (unsigned)grid, 200000000LL);
        cudaError_t e2 = cudaDeviceSynchronize();
        unsigned h[2048];
        cudaMemcpy(h, smids, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned min_seen = 1u << 30;
        for (int i = 0; i < grid; ++i) if (h[1024 + i] < min_seen) min_seen = h[1024 + i];
        printf("   grid %4d CTAs (%d clusters): launch %s / %s, min arrivals seen %u -> %s\n", grid, n + extra,
               cudaGetErrorString(e), cudaGetErrorString(e2), min_seen, min_seen >= (unsigned)grid ? "all co-resident" : "NOT co-resident");
      }
    }
  }
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("SMs %d\n", p.multiProcessorCount);
  return 0;
}
