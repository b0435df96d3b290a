// Probe: issue and completion time of tcgen05.mma chains as the resident kernel issues them (M = 128, fp32 accumulate,
// bf16 operands, K = 16 per instruction, 32 instructions = one half-step of a 512-unit prior), for N = 64 / 128 / 256,
// A operand in tensor memory (TS form) or in shared memory (SS form).  One CTA, one elected lane issues; cycles are SM clocks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I divae_b200/csrc -o ts_mma_probe benchmarks/probes/ts_mma_probe.cu -lcuda
#include <cstdio>
#include "umma_resident.cuh"

using namespace rbm;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_store[2];
  const uint32_t bar = smem_u32(&bar_store[0]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  {
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = 0x3F803F80u;
    for (int c = 256; c < 512; c += 32) tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
    tmem_st_wait();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(BLOCK_M, N, false);
    const uint64_t bdesc = make_smem_desc(smem_base, 16, 1024);
    const uint64_t adesc = make_smem_desc(smem_base + 96 * 1024, 16, 1024);
    for (int rep = 0; rep < 4; ++rep) {
      long long t0 = clock64(), t1 = 0;
      if (elect_one_sync()) {
#pragma unroll 1
        for (int kb = 0; kb < n_mma / 4; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (TS) umma_bf16_ts(tmem_base, tmem_base + 256 + (uint32_t)((kb & 7) * 32 + k * 8), bdesc + (uint64_t)((kb & 3) * (N * 128 >> 4) + k * 2), idesc, (kb | k) != 0);
            else umma_bf16(tmem_base, adesc + (uint64_t)((kb & 3) * 1024 + k * 2), bdesc + (uint64_t)((kb & 3) * (N * 128 >> 4) + k * 2), idesc, (kb | k) != 0);
          }
        }
        t1 = clock64();
        umma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      long long t2 = clock64();
      if (lane == 0) { out[rep * 2] = 0; }
      t1 = __shfl_sync(0xffffffffu, t1, 0) | 0;
      long long m = 0;
      for (int l = 0; l < 32; ++l) { long long v = __shfl_sync(0xffffffffu, t1, l); if (v > m) m = v; }
      if (lane == 0) { out[rep * 2] = m - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int N, bool TS>
void run(const char* name, long long* d_out) {
  cudaFuncSetAttribute(probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int n_mma : {4, 32}) {
    probe<N, TS><<<1, 128, 200 * 1024>>>(d_out, n_mma);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-8s N=%3d  %2d MMAs: issue %5lld cycles, complete %5lld cycles (last rep; first rep %lld / %lld)  %s\n", name, N, n_mma,
           h[6], h[7], h[0], h[1], cudaGetErrorString(e));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  run<64, true>("TS", d_out);
  run<128, true>("TS", d_out);
  run<256, true>("TS", d_out);
  run<64, false>("SS", d_out);
  run<128, false>("SS", d_out);
  run<256, false>("SS", d_out);
  return 0;
}
// stubs for the library's host-side error helpers (this probe links only the headers)
namespace rbm {
int cuda_fail(cudaError_t, const char*) { return 1; }
void set_error(const char*, ...) {}
int get_device_info(DeviceInfo*) { return 1; }
}
