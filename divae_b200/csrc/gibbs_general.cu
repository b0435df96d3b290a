// General fp32 SIMT kernels: any shape, fp32 W, one launch per half-step.
//
// These are the exact-fp32 restatement of the reference's per-half-step op chain
//   matmul -> +bias -> sigmoid -> rand -> >= -> float   (models/samplers/pcd.py:32-35,46-49)
// fused into one kernel, plus the mean-field variant without the draw
// (models/samplers/gibbsSampler.py:20-28), the per-sample energy (models/rbm/rbm.py:56-72)
// and the negative-phase statistics (models/autoencoders/gumbolt.py:109-134 via autograd).
// They serve odd shapes (200x200, 784x128), real-valued inputs, injected uniforms and
// small batches; the SMEM and UMMA variants are the fast paths for binary chains.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "philox.cuh"
#include "variants.h"

namespace rbm {

constexpr int GBM = 64, GBN = 64, GBK = 16;  // CTA tile: 64 rows x 64 units, 256 threads, 4x4 per thread

// acc[i][j] = sum_k in[row0 + ty*T + i][k] * Wop[k][n0 + tx*T + j]      (CTA tile 16T x 16T, 256 threads)
// TRANS = false: Wop[k][n] = W[k*ldw + n]   (v -> h, K = nV, N = nH)
// TRANS = true : Wop[k][n] = W[n*ldw + k]   (h -> v, K = nH, N = nV)
// T = 4: 64x64 tiles (throughput); T = 2: 32x32 tiles for small problems, 4x the CTAs and a 4x shorter
// K-serial loop per CTA (the CD-1 / PCD-batch regime is latency-bound, see profiles/r01_launches_cd1.csv).
template <bool TRANS, int T>
__device__ __forceinline__ void sgemm_tile(const float* __restrict__ in, const float* __restrict__ W,
                                           int64_t B, int K, int N, int ldw, int64_t row0, int n0,
                                           float (&acc)[T][T]) {
  constexpr int BT = 16 * T;
  __shared__ float As[GBK][BT + 4];
  __shared__ float Bs[GBK][BT + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) acc[i][j] = 0.f;

  // global -> registers for one K block (the NEXT block is fetched while the current one is multiplied,
  // so the global-load latency is off the critical path of this K-serial loop)
  float ra[T], rb[T];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int l = 0; l < T; ++l) {
      const int idx = tid + l * 256;
      const int r = idx >> 4, kk = idx & 15;
      const int64_t row = row0 + r;
      const int k = k0 + kk;
      ra[l] = (row < B && k < K) ? in[row * (int64_t)K + k] : 0.f;
      if (!TRANS) {
        const int kb = idx / BT, n = idx % BT;
        const int kx = k0 + kb, col = n0 + n;
        rb[l] = (kx < K && col < N) ? W[(int64_t)kx * ldw + col] : 0.f;
      } else {
        const int n = idx >> 4, kb = idx & 15;
        const int kx = k0 + kb, col = n0 + n;
        rb[l] = (kx < K && col < N) ? W[(int64_t)col * ldw + kx] : 0.f;
      }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += GBK) {
#pragma unroll
    for (int l = 0; l < T; ++l) {
      const int idx = tid + l * 256;
      As[idx & 15][idx >> 4] = ra[l];
      if (!TRANS) Bs[idx / BT][idx % BT] = rb[l];
      else Bs[idx & 15][idx >> 4] = rb[l];
    }
    __syncthreads();
    if (k0 + GBK < K) fetch(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[T], b[T];
#pragma unroll
      for (int i = 0; i < T; ++i) a[i] = As[kk][ty * T + i];
#pragma unroll
      for (int j = 0; j < T; ++j) b[j] = Bs[kk][tx * T + j];
#pragma unroll
      for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

template <bool TRANS, int MODE, int T>
__global__ void __launch_bounds__(256)
half_step_kernel(const float* __restrict__ in, const float* __restrict__ W,
                 const float* __restrict__ bias, float* __restrict__ out,
                 const float* __restrict__ U, int64_t B, int K, int N, int ldw,
                 DrawCtx ctx, uint64_t chain_offset, int64_t u_row_stride, float* __restrict__ probs_out) {
  constexpr int BT = 16 * T;
  const int64_t row0 = (int64_t)blockIdx.x * BT;
  const int n0 = blockIdx.y * BT;
  float acc[T][T];
  sgemm_tile<TRANS, T>(in, W, B, K, N, ldw, row0, n0, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < T; ++i) {
    const int64_t row = row0 + ty * T + i;
    if (row >= B) continue;
    const uint32_t chain = (uint32_t)(chain_offset + (uint64_t)row);
#pragma unroll
    for (int jp = 0; jp < T / 2; ++jp) {  // column pairs share one Philox block
      const int col = n0 + tx * T + jp * 2;
      if (col >= N) continue;
      uint4 w = make_uint4(0, 0, 0, 0);
      if (MODE == RBM_B200_MODE_SAMPLE) w = draw_block(ctx, unit_block_of(col), chain);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = col + e;
        if (c >= N) continue;
        const float x = acc[i][jp * 2 + e] + bias[c];
        float r;
        if (MODE == RBM_B200_MODE_PROBS) {
          r = sigmoid_precise(x);
        } else if (MODE == RBM_B200_MODE_INJECTED) {
          r = (sigmoid_precise(x) >= U[row * u_row_stride + c]) ? 1.f : 0.f;
        } else {
          const uint32_t slot = unit_slot_of(c);
          const uint32_t blk = unit_block_of(c);
          r = bernoulli_from_preact(x, field16(w, slot), [&]() {
                return field16(draw_block_extra(ctx, blk, chain), slot);
              }) ? 1.f : 0.f;
        }
        out[row * (int64_t)N + c] = r;
        if (MODE != RBM_B200_MODE_PROBS && probs_out) probs_out[row * (int64_t)N + c] = sigmoid_precise(x);
      }
    }
  }
}

__global__ void init_chains_kernel(float* __restrict__ v, int64_t B, int nV, DrawCtx ctx,
                                   uint64_t chain_offset) {
  // one thread per (chain, unit_block): eight units per Philox call
  const int nblk = ceil_div(nV, 32) * 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * nblk) return;
  const int64_t row = idx / nblk;
  const int blk = (int)(idx % nblk);
  const uint4 w = draw_block(ctx, (uint32_t)blk, (uint32_t)(chain_offset + (uint64_t)row));
  const int base = (blk >> 2) * 32 + (blk & 3) * 2;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int c = base + (s >> 1) * 8 + (s & 1);
    if (c < nV) v[row * (int64_t)nV + c] = (field16(w, s) >> 15) ? 0.f : 1.f;
  }
}

// ---- energy: E[b] = -(v.a + h.c + sum_j (vW)_j h_j), deterministic two-pass reduction -------------
__global__ void __launch_bounds__(256)
energy_partial_kernel(const float* __restrict__ v, const float* __restrict__ h,
                      const float* __restrict__ W, float* __restrict__ part, int64_t B, int nV,
                      int nH, int nparts) {
  const int64_t row0 = (int64_t)blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  float acc[4][4];
  sgemm_tile<false, 4>(v, W, B, nV, nH, nH, row0, n0, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    float s = 0.f;
    if (row < B) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = n0 + tx * 4 + j;
        if (c < nH) s = fmaf(acc[i][j], h[row * (int64_t)nH + c], s);
      }
    }
    // fixed-order tree over the 16 tx lanes (a half-warp)
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off, 16);
    if (tx == 0 && row < B) part[row * (int64_t)nparts + blockIdx.y] = s;
  }
}

__global__ void energy_finalize_kernel(const float* __restrict__ v, const float* __restrict__ h,
                                       const float* __restrict__ a, const float* __restrict__ c,
                                       const float* __restrict__ part, float* __restrict__ E,
                                       int64_t B, int nV, int nH, int nparts) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.f;
  for (int i = lane; i < nV; i += 32) s = fmaf(v[row * (int64_t)nV + i], a[i], s);
  for (int j = lane; j < nH; j += 32) s = fmaf(h[row * (int64_t)nH + j], c[j], s);
  for (int q = lane; q < nparts; q += 32) s += part[row * (int64_t)nparts + q];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) E[row] = -s;
}

// ---- negative-phase statistics: S_vh = scale * V^T H, S_v, S_h -----------------------------------
// 64x64 output tile per CTA, split over the chain dimension (gridDim.z); partial sums are
// accumulated with atomicAdd.  For binary states the partials are integers < 2^24, so the
// result is exact and independent of the accumulation order.
__global__ void __launch_bounds__(256)
stats_kernel(const float* __restrict__ v, const float* __restrict__ h, int64_t B, int nV, int nH,
             int64_t rows_per_split, float* __restrict__ S) {
  __shared__ float Vs[GBK][GBM + 4];
  __shared__ float Hs[GBK][GBN + 4];
  const int i0 = blockIdx.x * GBM, j0 = blockIdx.y * GBN;
  const int64_t b_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t b_end = min(B, b_begin + rows_per_split);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int64_t b0 = b_begin; b0 < b_end; b0 += GBK) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      int idx = tid + l * 256;
      int kk = idx >> 6, n = idx & 63;
      int64_t b = b0 + kk;
      Vs[kk][n] = (b < b_end && i0 + n < nV) ? v[b * (int64_t)nV + i0 + n] : 0.f;
      Hs[kk][n] = (b < b_end && j0 + n < nH) ? h[b * (int64_t)nH + j0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Vs[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Hs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = i0 + ty * 4 + i, c = j0 + tx * 4 + j;
      if (r < nV && c < nH) atomicAdd(&S[(int64_t)r * nH + c], acc[i][j]);
    }
}

// column sums: out[c] = sum_b x[b][c]; grid (ceil(n/256), splits)
__global__ void colsum_kernel(const float* __restrict__ x, int64_t B, int n, int64_t rows_per_split,
                              float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int64_t b_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t b_end = min(B, b_begin + rows_per_split);
  float s = 0.f;
  for (int64_t b = b_begin; b < b_end; ++b) s += x[b * (int64_t)n + c];
  atomicAdd(&out[c], s);
}

// one launch over the three statistics arrays: set to zero (scale == 0) or multiply by scale
__global__ void stats_scale3_kernel(float* __restrict__ S, float* __restrict__ sv, float* __restrict__ sh, int64_t nS,
                                    int nV, int nH, float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float* p = i < nS ? S + i : (i < nS + nV ? sv + (i - nS) : (i < nS + nV + nH ? sh + (i - nS - nV) : nullptr));
  if (p) *p = scale == 0.f ? 0.f : *p * scale;
}

// ------------------------------------------------------------------------------- host launchers
template <int MODE>
static int launch_half_step(const rbm_b200_params* p, int direction, const float* in, float* out,
                            int64_t B, const float* U, DrawCtx ctx, uint64_t chain_offset,
                            cudaStream_t st, bool broadcast_u = false, float* probs_out = nullptr) {
  const int K = direction == 0 ? p->n_visible : p->n_hidden;
  const int N = direction == 0 ? p->n_hidden : p->n_visible;
  const float* bias = direction == 0 ? p->hbias : p->vbias;
  // small problems are latency-bound: 32x32 tiles give 4x the CTAs and a 4x shorter K loop per CTA
  const bool small = ceil_div64(B, GBM) * ceil_div(N, GBN) < 74;
  const int64_t urs = broadcast_u ? 0 : (int64_t)N;
  if (small) {
    dim3 grid((unsigned)ceil_div64(B, 32), (unsigned)ceil_div(N, 32));
    if (direction == 0)
      half_step_kernel<false, MODE, 2><<<grid, 256, 0, st>>>(in, p->W, bias, out, U, B, K, N, p->n_hidden, ctx, chain_offset, urs, probs_out);
    else
      half_step_kernel<true, MODE, 2><<<grid, 256, 0, st>>>(in, p->W, bias, out, U, B, K, N, p->n_hidden, ctx, chain_offset, urs, probs_out);
  } else {
    dim3 grid((unsigned)ceil_div64(B, GBM), (unsigned)ceil_div(N, GBN));
    if (direction == 0)
      half_step_kernel<false, MODE, 4><<<grid, 256, 0, st>>>(in, p->W, bias, out, U, B, K, N, p->n_hidden, ctx, chain_offset, urs, probs_out);
    else
      half_step_kernel<true, MODE, 4><<<grid, 256, 0, st>>>(in, p->W, bias, out, U, B, K, N, p->n_hidden, ctx, chain_offset, urs, probs_out);
  }
  return check_launch("half_step_kernel");
}

int general_half_step(const rbm_b200_params* p, int direction, const float* in, float* out, int64_t B,
                      int mode, const float* U, uint64_t seed, uint64_t chain_offset,
                      uint64_t half_step, uint32_t tag, cudaStream_t st, bool broadcast_u, float* probs_out) {
  if (B == 0) return 0;
  DrawCtx ctx = make_draw_ctx(seed, half_step, tag);
  switch (mode) {
    case RBM_B200_MODE_SAMPLE: return launch_half_step<RBM_B200_MODE_SAMPLE>(p, direction, in, out, B, U, ctx, chain_offset, st, false, probs_out);
    case RBM_B200_MODE_INJECTED: return launch_half_step<RBM_B200_MODE_INJECTED>(p, direction, in, out, B, U, ctx, chain_offset, st, broadcast_u, probs_out);
    case RBM_B200_MODE_PROBS: return launch_half_step<RBM_B200_MODE_PROBS>(p, direction, in, out, B, U, ctx, chain_offset, st);
  }
  set_error("unknown half-step mode %d", mode);
  return RBM_B200_ERR_INVALID;
}

int general_init_chains(float* v, int64_t B, int nV, uint64_t seed, uint64_t chain_offset,
                        uint64_t step_offset, cudaStream_t st) {
  if (B == 0) return 0;
  const int nblk = ceil_div(nV, 32) * 4;
  const int64_t total = B * nblk;
  init_chains_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(
      v, B, nV, make_draw_ctx(seed, step_offset, TAG_INIT), chain_offset);
  return check_launch("init_chains_kernel");
}

int general_energy(const rbm_b200_params* p, const float* v, const float* h, float* E, int64_t B,
                   float* part, cudaStream_t st) {
  if (B == 0) return 0;
  const int nparts = ceil_div(p->n_hidden, GBN);
  dim3 grid((unsigned)ceil_div64(B, GBM), (unsigned)nparts);
  energy_partial_kernel<<<grid, 256, 0, st>>>(v, h, p->W, part, B, p->n_visible, p->n_hidden, nparts);
  int rc = check_launch("energy_partial_kernel");
  if (rc) return rc;
  energy_finalize_kernel<<<(unsigned)ceil_div64(B, 8), 256, 0, st>>>(v, h, p->vbias, p->hbias, part, E, B,
                                                                     p->n_visible, p->n_hidden, nparts);
  return check_launch("energy_finalize_kernel");
}

// E = -(<W, S_vh> + a.S_v + c.S_h): the mean energy of the batch whose (scaled) statistics are S_*.
// One cluster of eight CTAs: every CTA reduces a fixed, interleaved share of <W, S_vh> in a fixed order,
// rank 0 adds the eight partial sums (read over distributed shared memory) in rank order - deterministic,
// no scratch buffer, and eight SMs' worth of load bandwidth instead of one (165 -> ~15 us at 1024 x 1024).
constexpr int kEnergyCluster = 8;

__global__ void __launch_bounds__(1024)
stats_energy_kernel(const float* __restrict__ W, const float* __restrict__ a, const float* __restrict__ c,
                    const float* __restrict__ S, const float* __restrict__ sv, const float* __restrict__ sh,
                    int nV, int nH, float* __restrict__ E) {
  __shared__ float red[32];
  __shared__ float cta_sum;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int64_t n = (int64_t)nV * nH;
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(S)) & 15) == 0;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (vec) {
    const float4* W4 = reinterpret_cast<const float4*>(W);
    const float4* S4 = reinterpret_cast<const float4*>(S);
    for (int64_t i = (int64_t)rank * blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)kEnergyCluster * blockDim.x) {
      const float4 w = __ldg(W4 + i), x = __ldg(S4 + i);
      s0 = fmaf(w.x, x.x, s0); s1 = fmaf(w.y, x.y, s1); s2 = fmaf(w.z, x.z, s2); s3 = fmaf(w.w, x.w, s3);
    }
  } else {
    for (int64_t i = (int64_t)rank * blockDim.x + threadIdx.x; i < n; i += (int64_t)kEnergyCluster * blockDim.x)
      s0 = fmaf(W[i], S[i], s0);
  }
  float s = (s0 + s1) + (s2 + s3);
  if (rank == 0) {
    for (int i = threadIdx.x; i < nV; i += blockDim.x) s = fmaf(a[i], sv[i], s);
    for (int i = threadIdx.x; i < nH; i += blockDim.x) s = fmaf(c[i], sh[i], s);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (threadIdx.x == 0) cta_sum = s;
  }
  // partial sums visible cluster-wide
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && threadIdx.x == 0) {
    float total = 0.f;
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&cta_sum);
    for (uint32_t r = 0; r < (uint32_t)kEnergyCluster; ++r) {
      uint32_t remote;
      float part;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(part) : "r"(remote) : "memory");
      total += part;
    }
    *E = -total;
  }
  // no CTA may exit while rank 0 still reads its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int stats_energy(const rbm_b200_params* p, const float* S_vh, const float* S_v, const float* S_h, float* E, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kEnergyCluster);
  cfg.blockDim = dim3(1024);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kEnergyCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RBM_CUDA_TRY(cudaLaunchKernelEx(&cfg, stats_energy_kernel, (const float*)p->W, (const float*)p->vbias,
                                  (const float*)p->hbias, S_vh, S_v, S_h, (int)p->n_visible, (int)p->n_hidden, E));
  return check_launch("stats_energy_kernel");
}

int stats_zero(float* S_vh, float* S_v, float* S_h, int nV, int nH, cudaStream_t st) {
  const int64_t n = (int64_t)nV * nH + nV + nH;
  stats_scale3_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(S_vh, S_v, S_h, (int64_t)nV * nH, nV, nH, 0.f);
  return check_launch("stats_scale3_kernel(zero)");
}

int stats_scale(float* S_vh, float* S_v, float* S_h, int nV, int nH, float scale, cudaStream_t st) {
  if (scale == 1.f) return 0;
  if (scale == 0.f) return stats_zero(S_vh, S_v, S_h, nV, nH, st);
  const int64_t n = (int64_t)nV * nH + nV + nH;
  stats_scale3_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(S_vh, S_v, S_h, (int64_t)nV * nH, nV, nH, scale);
  return check_launch("stats_scale3_kernel");
}

int general_negative_phase(const float* v, const float* h, int64_t B, int nV, int nH, float scale,
                           float* S_vh, float* S_v, float* S_h, cudaStream_t st) {
  int rc0 = stats_zero(S_vh, S_v, S_h, nV, nH, st);
  if (rc0) return rc0;
  if (B == 0) return 0;
  const int tiles = ceil_div(nV, GBM) * ceil_div(nH, GBN);
  int splits = (int)std::min<int64_t>(ceil_div64(B, 256), std::max(1, 592 / tiles));
  int64_t rows_per_split = ceil_div64(ceil_div64(B, splits), GBK) * GBK;
  splits = (int)ceil_div64(B, rows_per_split);
  dim3 grid((unsigned)ceil_div(nV, GBM), (unsigned)ceil_div(nH, GBN), (unsigned)splits);
  stats_kernel<<<grid, 256, 0, st>>>(v, h, B, nV, nH, rows_per_split, S_vh);
  int rc = check_launch("stats_kernel");
  if (rc) return rc;
  int csplits = (int)std::min<int64_t>(ceil_div64(B, 64), 64);
  int64_t crows = ceil_div64(B, csplits);
  csplits = (int)ceil_div64(B, crows);
  colsum_kernel<<<dim3((unsigned)ceil_div(nV, 256), (unsigned)csplits), 256, 0, st>>>(v, B, nV, crows, S_v);
  colsum_kernel<<<dim3((unsigned)ceil_div(nH, 256), (unsigned)csplits), 256, 0, st>>>(h, B, nH, crows, S_h);
  rc = check_launch("colsum_kernel");
  if (rc) return rc;
  // raw sums are exact integers for binary states; scale once at the end
  return stats_scale(S_vh, S_v, S_h, nV, nH, scale, st);
}

// ------------------------------------------------------------------------------- CD-k update
// dW <- momentum * dW + (v0^T h0 - pv^T ph) - W * weight_cost;  W <- W + lr * dW
// (sandbox/rbm_standalone.py:86,91,95-103,113).  One 64x64 tile of W per CTA, the batch reduced inside
// the CTA in a fixed order: deterministic, no atomics.
__global__ void __launch_bounds__(256)
cd_weight_update_kernel(const float* __restrict__ v0, const float* __restrict__ h0, const float* __restrict__ pv,
                        const float* __restrict__ ph, int64_t B, int nV, int nH, float* __restrict__ W,
                        float* __restrict__ dW, float lr, float momentum, float weight_cost) {
  __shared__ float Vs[GBK][GBM + 4], Hs[GBK][GBN + 4], PVs[GBK][GBM + 4], PHs[GBK][GBN + 4];
  const int i0 = blockIdx.x * GBM, j0 = blockIdx.y * GBN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int64_t b0 = 0; b0 < B; b0 += GBK) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int idx = tid + l * 256;
      const int kk = idx >> 6, n = idx & 63;
      const int64_t b = b0 + kk;
      const bool vi = b < B && i0 + n < nV, hj = b < B && j0 + n < nH;
      Vs[kk][n] = vi ? v0[b * nV + i0 + n] : 0.f;
      PVs[kk][n] = vi ? pv[b * nV + i0 + n] : 0.f;
      Hs[kk][n] = hj ? h0[b * nH + j0 + n] : 0.f;
      PHs[kk][n] = hj ? ph[b * nH + j0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float a[4], bb[4], pa[4], pb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Vs[kk][ty * 4 + i]; pa[i] = PVs[kk][ty * 4 + i]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { bb[j] = Hs[kk][tx * 4 + j]; pb[j] = PHs[kk][tx * 4 + j]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], fmaf(-pa[i], pb[j], acc[i][j]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = i0 + ty * 4 + i, c = j0 + tx * 4 + j;
      if (r < nV && c < nH) {
        const int64_t idx = (int64_t)r * nH + c;
        const float w = W[idx];
        const float u = dW[idx] * momentum + acc[i][j] - w * weight_cost;
        dW[idx] = u;
        W[idx] = w + u * lr;
      }
    }
}

// bias updates (rbm_standalone.py:105-115) and the summed squared reconstruction error (:119-120):
//   x = data side, y = reconstruction side; db <- momentum*db + sum_b (x - y); bias += lr*db
__global__ void cd_bias_update_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t B, int n,
                                      float* __restrict__ bias, float* __restrict__ db, float lr, float momentum,
                                      float* __restrict__ loss) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f, sq = 0.f;
  if (c < n) {
#pragma unroll 8
    for (int64_t b = 0; b < B; ++b) {
      const float d = x[b * n + c] - y[b * n + c];
      s += d;
      sq = fmaf(d, d, sq);
    }
    const float u = db[c] * momentum + s;
    db[c] = u;
    bias[c] += u * lr;
  }
  if (loss) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if ((threadIdx.x & 31) == 0 && sq != 0.f) atomicAdd(loss, sq);
  }
}


// ------------------------------------------------------------------ CD-k in ONE launch (small batches)
// The reference's CD-1 on MNIST-shape data (BASELINE config 1: 784 x 128, batch 128) is ~0.13 GFLOP: as separate
// launches it is 3k + 4 dependent launch latencies with 16-CTA grids.  Here the whole update is one cooperative
// kernel: every phase is cut into ~100-130 work items spread over the SMs, phases are separated by grid barriers
//   P1  ph0 = sigma(v0 W + c), h0 = [ph0 >= u_0]                              (rbm_standalone.py:79-84)
//   k x P2  pv = sigma(h W^T + a)      P3  ph = sigma(pv W + c) [, h = [ph >= u_s]]     (:66-73)
//   P4  dW, W, da, a, dc, c, loss with the OLD parameters' statistics          (:86-120)
// Same arithmetic as the multi-launch path (fp32 FMA accumulation, sigmoid_precise, the same draw contract).
struct CdFusedArgs {
  float *W, *a, *c, *dW, *da, *dc;
  int nV, nH;
  const float* v0;
  int B, k;
  float lr, momentum, weight_cost;
  const float* uniforms;   // [k + 1, nH] rows broadcast over the batch, or NULL (Philox draws)
  uint64_t seed, draw_offset;
  float* loss;
  float *ph0, *h0, *hcur, *ph, *pv;
  unsigned int* barrier;   // zeroed before the launch; counts CTA arrivals over all grid barriers of the launch
};

constexpr int kCdThreads = 256;
constexpr int kCdSmemFloats = 4 * 32 * 33;   // the largest phase (P4: four [32][33] tiles)

// out[r, j] = sigma(sum_i in[r, i] W[i, j] + c[j]); item = 4 rows x 32 hidden units, the 8 warps split the visible units
__device__ void cd_hidden_phase(const CdFusedArgs& A, const float* in, float* probs, float* sample, int draw_idx, float* sm) {
  constexpr int RB = 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_items = ceil_div(A.nH, 32);
  const int n_items = ceil_div(A.B, RB) * col_items;
  const int n_gran = ceil_div(A.nV, 32);
  float (*red)[RB][32] = reinterpret_cast<float (*)[RB][32]>(sm);   // [8][RB][32]
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int r0 = (item / col_items) * RB, j = (item % col_items) * 32 + lane;
    float acc[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) acc[rb] = 0.f;
    for (int g = warp; g < n_gran; g += 8) {
      const int i0 = g * 32;
      float x[RB];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        const int r = r0 + rb;
        x[rb] = (r < A.B && i0 + lane < A.nV) ? in[(int64_t)r * A.nV + i0 + lane] : 0.f;
      }
      // all 32 W loads of the granule are issued before the first FMA (one L2 round trip per granule)
      float w[32];
#pragma unroll
      for (int kk = 0; kk < 32; ++kk)
        w[kk] = (j < A.nH && i0 + kk < A.nV) ? A.W[(int64_t)(i0 + kk) * A.nH + j] : 0.f;
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) {
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) acc[rb] = fmaf(__shfl_sync(0xffffffffu, x[rb], kk), w[kk], acc[rb]);
      }
    }
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) red[warp][rb][lane] = acc[rb];
    __syncthreads();
    if (warp < RB) {
      const int r = r0 + warp;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += red[w][warp][lane];
      if (r < A.B && j < A.nH) {
        const float xx = sum + A.c[j];
        const float pr = sigmoid_precise(xx);
        if (probs) probs[(int64_t)r * A.nH + j] = pr;
        if (sample) {
          bool one;
          if (A.uniforms) {
            one = pr >= A.uniforms[(int64_t)draw_idx * A.nH + j];
          } else {
            const DrawCtx ctx = make_draw_ctx(A.seed, A.draw_offset + (uint64_t)draw_idx, TAG_GIBBS);
            const uint32_t blk = unit_block_of((uint32_t)j), slot = unit_slot_of((uint32_t)j);
            const uint4 wd = draw_block(ctx, blk, (uint32_t)r);
            one = bernoulli_from_preact(xx, field16(wd, slot), [&]() { return field16(draw_block_extra(ctx, blk, (uint32_t)r), slot); });
          }
          sample[(int64_t)r * A.nH + j] = one ? 1.f : 0.f;
        }
      }
    }
    __syncthreads();
  }
}

// pv[r, i] = sigma(sum_j h[r, j] W[i, j] + a[i]); item = 32 rows x 32 visible units, hidden units staged 64 at a time
__device__ void cd_visible_phase(const CdFusedArgs& A, const float* hin, float* pv, float* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float (*hs)[33] = reinterpret_cast<float (*)[33]>(sm);             // [64][33]: hs[j][r]
  float (*ws)[65] = reinterpret_cast<float (*)[65]>(sm + 64 * 33);   // [32][65]: ws[i][j]
  const int col_items = ceil_div(A.nV, 32);
  const int n_items = ceil_div(A.B, 32) * col_items;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int r0 = (item / col_items) * 32, i0 = (item % col_items) * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < A.nH; j0 += 64) {
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const int idx = threadIdx.x + l * kCdThreads;
        const int x = idx >> 6, jj = idx & 63;    // x: row (for hs) or unit (for ws)
        const bool jok = j0 + jj < A.nH;
        hs[jj][x] = (jok && r0 + x < A.B) ? hin[(int64_t)(r0 + x) * A.nH + j0 + jj] : 0.f;
        ws[x][jj] = (jok && i0 + x < A.nV) ? A.W[(int64_t)(i0 + x) * A.nH + j0 + jj] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int jj = 0; jj < 64; ++jj) {
        const float hv = hs[jj][lane];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fmaf(hv, ws[warp * 4 + q][jj], acc[q]);
      }
      __syncthreads();
    }
    const int r = r0 + lane;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + warp * 4 + q;
      if (r < A.B && i < A.nV) pv[(int64_t)r * A.nV + i] = sigmoid_precise(acc[q] + A.a[i]);
    }
  }
}

// P4: weights (32 x 32 tiles, batch staged 32 chains at a time), biases and the reconstruction error
__device__ void cd_update_phase(const CdFusedArgs& A, float* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float (*v0s)[33] = reinterpret_cast<float (*)[33]>(sm);
  float (*pvs)[33] = reinterpret_cast<float (*)[33]>(sm + 32 * 33);
  float (*h0s)[33] = reinterpret_cast<float (*)[33]>(sm + 2 * 32 * 33);
  float (*phs)[33] = reinterpret_cast<float (*)[33]>(sm + 3 * 32 * 33);
  const int col_items = ceil_div(A.nH, 32);
  const int n_items = ceil_div(A.nV, 32) * col_items;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int i0 = (item / col_items) * 32, j0 = (item % col_items) * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // the tiles of the first tile column also own the visible-bias sums of their 32 units (warp 0), the tiles
    // of the first tile row the hidden-bias sums of theirs (warp 1): the operands are already in shared memory
    float bsum = 0.f, bsq = 0.f;
    for (int b0 = 0; b0 < A.B; b0 += 32) {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const int idx = threadIdx.x + l * kCdThreads;
        const int b = idx >> 5, x = idx & 31;
        const bool bok = b0 + b < A.B;
        const bool iok = bok && i0 + x < A.nV, jok = bok && j0 + x < A.nH;
        v0s[b][x] = iok ? A.v0[(int64_t)(b0 + b) * A.nV + i0 + x] : 0.f;
        pvs[b][x] = iok ? A.pv[(int64_t)(b0 + b) * A.nV + i0 + x] : 0.f;
        h0s[b][x] = jok ? A.h0[(int64_t)(b0 + b) * A.nH + j0 + x] : 0.f;
        phs[b][x] = jok ? A.ph[(int64_t)(b0 + b) * A.nH + j0 + x] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int b = 0; b < 32; ++b) {
        const float hj = h0s[b][lane], pj = phs[b][lane];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          acc[q] = fmaf(v0s[b][warp * 4 + q], hj, fmaf(-pvs[b][warp * 4 + q], pj, acc[q]));
      }
      if (j0 == 0 && warp == 0) {
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {
          const float d = v0s[b][lane] - pvs[b][lane];   // zero padded outside the batch / the units
          bsum += d;
          bsq = fmaf(d, d, bsq);
        }
      }
      if (i0 == 0 && warp == 1 && j0 + lane < A.nH) {
        float p0[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) p0[b] = b0 + b < A.B ? A.ph0[(int64_t)(b0 + b) * A.nH + j0 + lane] : 0.f;
#pragma unroll
        for (int b = 0; b < 32; ++b) bsum += p0[b] - phs[b][lane];
      }
      __syncthreads();
    }
    if (j0 == 0 && warp == 0) {
      const int u = i0 + lane;
      if (u < A.nV) {
        const float up = A.da[u] * A.momentum + bsum;
        A.da[u] = up;
        A.a[u] += up * A.lr;
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) bsq += __shfl_xor_sync(0xffffffffu, bsq, off);
      if (lane == 0 && bsq != 0.f) atomicAdd(A.loss, bsq);
    }
    if (i0 == 0 && warp == 1) {
      const int u = j0 + lane;
      if (u < A.nH) {
        const float up = A.dc[u] * A.momentum + bsum;
        A.dc[u] = up;
        A.c[u] += up * A.lr;
      }
    }
    const int j = j0 + lane;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + warp * 4 + q;
      if (i < A.nV && j < A.nH) {
        const int64_t idx = (int64_t)i * A.nH + j;
        const float w = A.W[idx];
        const float u = A.dW[idx] * A.momentum + acc[q] - w * A.weight_cost;
        A.dW[idx] = u;
        A.W[idx] = w + u * A.lr;
      }
    }
  }
}

// Grid barrier number `n` (1, 2, ...) of this launch: every CTA adds one arrival to a counter that starts at zero
// and waits until n * gridDim.x arrivals are in.  The launch is cooperative, so all CTAs are resident.  (The
// cooperative-groups grid.sync() spent ~20 us per barrier here: its poll loop runs a gpu-scope fence with an L1
// invalidation per iteration; an acquire load per iteration is enough.)
__device__ __forceinline__ void cd_grid_barrier(unsigned int* counter, unsigned int n) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = n * gridDim.x;
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCdThreads)
cd_fused_kernel(const CdFusedArgs A) {
  __shared__ float sm[kCdSmemFloats];
  unsigned int nbar = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *A.loss = 0.f;   // added to in P4, three grid barriers later
  cd_hidden_phase(A, A.v0, A.ph0, A.h0, 0, sm);
  cd_grid_barrier(A.barrier, ++nbar);
  const float* hin = A.h0;
  for (int s = 0; s < A.k; ++s) {
    cd_visible_phase(A, hin, A.pv, sm);
    cd_grid_barrier(A.barrier, ++nbar);
    // the draw after the last sweep is discarded by the reference
    cd_hidden_phase(A, A.pv, A.ph, s + 1 < A.k ? A.hcur : nullptr, s + 1, sm);
    cd_grid_barrier(A.barrier, ++nbar);
    hin = A.hcur;
  }
  cd_update_phase(A, sm);
}

// true if the fused kernel was launched
int cd_step_fused(const CdFusedArgs& A, cudaStream_t st, bool* launched) {
  *launched = false;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  // per device: 0 = not queried yet, 1 = usable, -1 = no cooperative launch / kernel does not fit
  static int usable[64] = {0};
  const int dev = di.device >= 0 && di.device < 64 ? di.device : 0;
  if (usable[dev] == 0) {
    int coop = 0, per_sm = 0;
    RBM_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, di.device));
    RBM_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cd_fused_kernel, kCdThreads, 0));
    usable[dev] = (coop && per_sm >= 1) ? 1 : -1;
  }
  if (usable[dev] < 0) return 0;
  const int items = std::max({ceil_div(A.B, 4) * ceil_div(A.nH, 32), ceil_div(A.B, 32) * ceil_div(A.nV, 32),
                              ceil_div(A.nV, 32) * ceil_div(A.nH, 32)});
  const int grid = std::max(1, std::min(items, di.sm_count));
  RBM_CUDA_TRY(cudaMemsetAsync(A.barrier, 0, sizeof(unsigned int), st));
  void* params[] = {const_cast<CdFusedArgs*>(&A)};
  RBM_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)cd_fused_kernel, dim3((unsigned)grid), dim3(kCdThreads), params, 0, st));
  *launched = true;
  return check_launch("cd_fused_kernel");
}

size_t general_cd_workspace(int nV, int nH, int64_t B) {
  return sizeof(float) * (size_t)B * (size_t)(4 * nH + nV) + 256;   // + the fused kernel's grid-barrier counter
}

// One CD-k update (sandbox/rbm_standalone.py:75-121), 3k + 4 launches, everything in fp32.
int general_cd_step(float* W, float* a, float* c, float* dW, float* da, float* dc, int nV, int nH, const float* v0,
                    int64_t B, int k, float lr, float momentum, float weight_cost, const float* uniforms,
                    uint64_t seed, uint64_t draw_offset, float* loss, float* ws, cudaStream_t st) {
  float* ph0 = ws;
  float* h0 = ph0 + (size_t)B * nH;
  float* hcur = h0 + (size_t)B * nH;
  float* ph = hcur + (size_t)B * nH;
  float* pv = ph + (size_t)B * nH;
  int rc;
  // small batches (the reference's CD-1 on 128 digits): one cooperative launch instead of 3k + 4 launches
  static const int fused_mode = [] { const char* e = getenv("RBM_B200_CD_FUSED"); return e ? atoi(e) : -1; }();
  if (fused_mode != 0 && B <= 1024) {
    CdFusedArgs fa{};
    fa.W = W; fa.a = a; fa.c = c; fa.dW = dW; fa.da = da; fa.dc = dc; fa.nV = nV; fa.nH = nH; fa.v0 = v0; fa.B = (int)B;
    fa.k = k; fa.lr = lr; fa.momentum = momentum; fa.weight_cost = weight_cost; fa.uniforms = uniforms; fa.seed = seed;
    fa.draw_offset = draw_offset; fa.loss = loss; fa.ph0 = ph0; fa.h0 = h0; fa.hcur = hcur; fa.ph = ph; fa.pv = pv;
    fa.barrier = reinterpret_cast<unsigned int*>(pv + (size_t)B * nV);
    bool launched = false;
    if ((rc = cd_step_fused(fa, st, &launched))) return rc;
    if (launched) return 0;
  }
  rbm_b200_params p{};
  p.n_visible = nV; p.n_hidden = nH; p.W = W; p.vbias = a; p.hbias = c;
  const int draw_mode = uniforms ? RBM_B200_MODE_INJECTED : RBM_B200_MODE_SAMPLE;
  auto u_row = [&](int i) { return uniforms ? uniforms + (size_t)i * nH : nullptr; };
  // positive phase: ph0 = sigma(v0 W + c);  h0 = (ph0 >= u_0)                              :79-84
  // (one kernel: the draw and the probabilities come from the same pre-activation)
  if ((rc = general_half_step(&p, 0, v0, h0, B, draw_mode, u_row(0), seed, 0, draw_offset, TAG_GIBBS, st, true, ph0))) return rc;
  // negative phase: k x { pv = sigma(h W^T + a); ph = sigma(pv W + c); h = (ph >= u_{s+1}) }   :66-73
  const float* hin = h0;
  for (int s = 0; s < k; ++s) {
    if ((rc = general_half_step(&p, 1, hin, pv, B, RBM_B200_MODE_PROBS, nullptr, 0, 0, 0, TAG_GIBBS, st, false))) return rc;
    if ((rc = general_half_step(&p, 0, pv, ph, B, RBM_B200_MODE_PROBS, nullptr, 0, 0, 0, TAG_GIBBS, st, false))) return rc;
    if (s + 1 < k) {  // the draw after the last sweep is discarded by the reference
      if ((rc = general_half_step(&p, 0, pv, hcur, B, draw_mode, u_row(s + 1), seed, 0, draw_offset + s + 1, TAG_GIBBS, st, true))) return rc;
      hin = hcur;
    }
  }
  // updates: all statistics above were taken with the OLD parameters                           :95-120
  RBM_CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), st));
  dim3 grid((unsigned)ceil_div(nV, GBM), (unsigned)ceil_div(nH, GBN));
  cd_weight_update_kernel<<<grid, 256, 0, st>>>(v0, h0, pv, ph, B, nV, nH, W, dW, lr, momentum, weight_cost);
  cd_bias_update_kernel<<<ceil_div(nV, 128), 128, 0, st>>>(v0, pv, B, nV, a, da, lr, momentum, loss);
  cd_bias_update_kernel<<<ceil_div(nH, 128), 128, 0, st>>>(ph0, ph, B, nH, c, dc, lr, momentum, nullptr);
  return check_launch("cd update kernels");
}

}  // namespace rbm
