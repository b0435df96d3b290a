// UMMA variant: one Gibbs half-step = one tcgen05/TMEM GEMM fed by TMA, with the
// bias + sigmoid + Philox + Bernoulli epilogue fused behind the accumulator.
//
//   H = [ sigmoid(V W + c) >= u ]      (models/samplers/pcd.py:32-35)   B operand MN-major
//   V = [ sigmoid(H W^T + a) >= u ]    (models/samplers/pcd.py:46-49)   B operand K-major
//
// States live in HBM/L2 between half-steps as {0,1} bf16 tiles [chains_pad, units_pad]
// (exact in bf16, so the tensor-core product with bf16 W is exact and the fp32 TMEM
// accumulation differs from the oracle only in summation order).  W (bf16) is read from
// ONE padded copy by both directions through two tensor maps.
//
// Kernel (umma_common.cuh: umma_half_step_kernel): persistent, warp-specialised, 1 CTA per SM, 128 x BLOCK_N output tile.
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16)
//   warp 2      TMEM allocator (2 accumulator stages x BLOCK_N fp32 columns)
//   warps 4-19  epilogue       (tcgen05.ld 32x32b.x16 -> ex2 + Philox + compare -> bf16 store)
// The MMA of tile i+1 overlaps the epilogue of tile i through the two TMEM stages.
// This file: the multi-step cluster kernel, the statistics GEMM, staging kernels and the host drivers
// (umma_block_gibbs, umma_ais_run).
#include <cmath>
#include <vector>

#include "umma_chain.cuh"
#include "umma_resident.cuh"

namespace rbm {
namespace {

// ------------------------------------------------------------- helper kernels
// fp32 [B, n] (row pitch n) -> bf16 [rows_pad, ld] zero padded
__global__ void f32_to_bf16_padded(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t B, int n,
                                   int64_t rows_pad, int ld) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_pad * ld) return;
  const int64_t r = idx / ld;
  const int c = (int)(idx % ld);
  dst[idx] = __float2bfloat16((r < B && c < n) ? src[r * n + c] : 0.f);
}
// bf16 [rows_pad, ld] -> fp32 [B, n]
__global__ void bf16_to_f32_crop(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t B, int n, int ld) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * n) return;
  const int64_t r = idx / n;
  const int c = (int)(idx % n);
  dst[idx] = __bfloat162float(src[r * ld + c]);
}
// same, eight units per thread (n % 8 == 0): one 16-byte load, two 16-byte stores
__global__ void bf16_to_f32_crop8(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t B, int n, int ld) {
  const int n8 = n >> 3;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * n8) return;
  const int64_t r = idx / n8;
  const int c = (int)(idx % n8) * 8;
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + r * ld + c));
  float4* d = reinterpret_cast<float4*>(dst + r * n + c);
  d[0] = make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xFFFF0000u),
                     __uint_as_float(q.y << 16), __uint_as_float(q.y & 0xFFFF0000u));
  d[1] = make_float4(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xFFFF0000u),
                     __uint_as_float(q.w << 16), __uint_as_float(q.w & 0xFFFF0000u));
}
__global__ void pad_w_kernel(const uint16_t* __restrict__ W, uint16_t* __restrict__ Wp, int nV, int nH, int nVp, int nHp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nVp * nHp) return;
  const int i = (int)(idx / nHp), j = (int)(idx % nHp);
  Wp[idx] = (i < nV && j < nH) ? W[(int64_t)i * nH + j] : (uint16_t)0;
}
// fp32 W -> padded bf16 (round to nearest even): the rounding lives in the library, so there is no bf16 copy of the
// parameters outside the workspace that could go stale
__global__ void pad_w_f32_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Wp, int nV, int nH, int nVp, int nHp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nVp * nHp) return;
  const int i = (int)(idx / nHp), j = (int)(idx % nHp);
  Wp[idx] = __float2bfloat16_rn((i < nV && j < nH) ? W[(int64_t)i * nH + j] : 0.f);
}
// split weights: W = high + low with high = bf16(W), low = bf16(W - high); both zero padded
__global__ void pad_w_split_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ Whi, __nv_bfloat16* __restrict__ Wlo,
                                   int nV, int nH, int nVp, int nHp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nVp * nHp) return;
  const int i = (int)(idx / nHp), j = (int)(idx % nHp);
  const float w = (i < nV && j < nH) ? W[(int64_t)i * nH + j] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  Whi[idx] = hi;
  Wlo[idx] = __float2bfloat16_rn(w - __bfloat162float(hi));
}
__global__ void nbias_kernel(const float* __restrict__ bias, float* __restrict__ nb, int n, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) nb[i] = i < n ? -1.4426950408889634f * bias[i] : 0.f;
}
// Bernoulli(1/2) initial state straight into the bf16 state buffer (contract: 1 - (u16 >> 15), TAG_INIT)
__global__ void init_chains_bf16_kernel(__nv_bfloat16* __restrict__ v, int64_t rows, int nV, int ld, DrawCtx ctx,
                                        uint64_t chain_offset) {
  // one 32-unit group (four Philox blocks, philox.cuh slot mapping) per thread: 64 contiguous bytes of the row
  const int ngrp = ld / 32;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * ngrp) return;
  const int64_t row = idx / ngrp;
  const int grp = (int)(idx % ngrp);
  const uint32_t chain = (uint32_t)(chain_offset + (uint64_t)row);
  uint32_t w[4][4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint4 x = draw_block(ctx, (uint32_t)(grp * 4 + b), chain);
    w[b][0] = x.x; w[b][1] = x.y; w[b][2] = x.z; w[b][3] = x.w;
  }
  uint32_t pk[16];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const int b = (c & 7) >> 1, slot = 2 * (c >> 3) + (c & 1);
    const uint32_t word = w[b][slot >> 1];
    const uint32_t k16 = (slot & 1) ? (word >> 16) : (word & 0xFFFFu);
    const uint32_t one = (grp * 32 + c < nV && !(k16 >> 15)) ? 0x3F80u : 0u;   // v0 = 1 - (u16 >> 15)
    if (c & 1) pk[c >> 1] |= one << 16; else pk[c >> 1] = one;
  }
  uint4* dst = reinterpret_cast<uint4*>(v + row * (int64_t)ld + grp * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// AIS initial state v ~ p_A = Bernoulli(sigmoid(a)) (base-rate model with a_A = a), TAG_AIS_INIT
__global__ void ais_init_kernel(__nv_bfloat16* __restrict__ v, int64_t rows, int nV, int ld,
                                const float* __restrict__ nbias_v, DrawCtx ctx, uint64_t chain_offset) {
  const int nblk = ld / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * nblk) return;
  const int64_t row = idx / nblk;
  const int blk = (int)(idx % nblk);
  const uint32_t chain = (uint32_t)(chain_offset + (uint64_t)row);
  const uint4 w = draw_block(ctx, (uint32_t)blk, chain);
  const int base = (blk >> 2) * 32 + (blk & 3) * 2;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int c = base + (s >> 1) * 8 + (s & 1);
    if (c >= ld) continue;
    uint32_t one = 0;
    if (c < nV) one = resolve_exact(nbias_v[c], field16(w, s), ctx, (uint32_t)blk, chain, (uint32_t)s);
    v[row * (int64_t)ld + c] = __float2bfloat16(one ? 1.f : 0.f);
  }
}
// logw[b] = ln2 * sum_slots wpart[b][slot]   (fp64 sum of the fp32 running partials)
__global__ void ais_reduce_kernel(const float* __restrict__ wpart, int ld, int64_t B, double* __restrict__ logw) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int i = 0; i < ld; ++i) s += (double)wpart[b * ld + i];
  logw[b] = s * 0.6931471805599453;
}


// ---------------------------------------------------------------------------------------------------
// Negative-phase statistics on the tensor cores:  S_vh += V^T H  over the chains of one K split
// (the W-gradient of GumBolt.energy_exp at the sampler's states, models/autoencoders/gumbolt.py:109-134,
// dvaepp.py:141-159).
//   D[i, j] = sum_b V[b, i] * H[b, j]      M = visible units, N = hidden units, K = chains
// Both operands are read straight from the bf16 state tiles [chains, units] the sampler leaves in the
// workspace: a [64 chains x 64 units] TMA box IS one MN-major SWIZZLE_128B block, for A (V^T) and for B (H)
// alike, so there is no transpose pass.  The tensor maps are built with B rows: chains of the padding
// are zero-filled by TMA.  Counts are exact in fp32 (B <= 2^24 chains); the split-K partial tiles are added
// to the zeroed S_vh with red.global.add.f32, and integer-valued sums do not depend on the order.
constexpr int kStatsThreads = 256;   // warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue
constexpr int kStatsStages = 4;

struct StatsArgs {
  float* S_vh;         // [nV, nH] fp32, raw counts are ADDED
  int nV, nH;
  int tiles_n;         // N tiles; blockIdx.x = split * tiles + m_blk * tiles_n + n_blk
  int tiles;           // M tiles * N tiles
  int k_blocks;        // ceil(B / 64)
  int kb_per_split;
};

template <int BLOCK_N>
constexpr int stats_smem_bytes() { return kStatsStages * (kABytes + BLOCK_N * BLOCK_K * 2) + 256 + 1024; }

template <int BLOCK_N>
__global__ void __launch_bounds__(kStatsThreads, 1)
umma_stats_kernel(const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_h,
                  const StatsArgs args) {
  constexpr int kStageBytes = kABytes + BLOCK_N * BLOCK_K * 2;
  constexpr int kBoxBytes = BLOCK_K * 128;   // one [64 chains x 64 units] box
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + kStatsStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStatsStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStatsStages);
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kStatsStages * kStageBytes + 8 * (2 * kStatsStages + 1));

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = (int)blockIdx.x % args.tiles, split = (int)blockIdx.x / args.tiles;
  const int m_blk = tile / args.tiles_n, n_blk = tile % args.tiles_n;
  const int kb0 = split * args.kb_per_split;
  const int nkb = min(args.k_blocks, kb0 + args.kb_per_split) - kb0;   // >= 1 (host)

  if (warp_idx == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_v) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_h) : "memory");
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < kStatsStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp_idx == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"((uint32_t)BLOCK_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // TMA producer (warp-uniform loop, one elected lane issues)
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t sa = smem_base + stage * kStageBytes;
      const uint32_t sb = sa + kABytes;
      const int chain0 = (kb0 + i) * BLOCK_K;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
#pragma unroll
        for (int mb = 0; mb < BLOCK_M / 64; ++mb)
          tma_load_2d(sa + mb * kBoxBytes, &tmap_v, full_bar(stage), m_blk * BLOCK_M + mb * 64, chain0);
#pragma unroll
        for (int nb = 0; nb < BLOCK_N / 64; ++nb)
          tma_load_2d(sb + nb * kBoxBytes, &tmap_h, full_bar(stage), n_blk * BLOCK_N + nb * 64, chain0);
      }
      __syncwarp();
      if (++stage == kStatsStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp_idx == 1) {
    // MMA issuer: A and B both MN-major (64-unit blocks LBO = 8 KB apart, 8-chain groups SBO = 1 KB apart)
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, true, true);
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(full_bar(stage), phase);
      tcgen05_fence_after();
      const uint32_t sa = smem_base + stage * kStageBytes;
      const uint32_t sb = sa + kABytes;
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint64_t adesc = make_smem_desc(sa + k * (UMMA_K * 128), kBoxBytes, 1024);
          const uint64_t bdesc = make_smem_desc(sb + k * (UMMA_K * 128), kBoxBytes, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      if (++stage == kStatsStages) { stage = 0; phase ^= 1u; }
    }
    if (elect_one_sync()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else if (warp_idx >= 4) {
    // epilogue: lane = visible unit (TMEM lane), registers = 16 consecutive hidden units
    const int quarter = warp_idx & 3;
    const int row = m_blk * BLOCK_M + quarter * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    float* dst_row = args.S_vh + (size_t)min(row, args.nV - 1) * args.nH;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 16; ++c) {
      const int col0 = n_blk * BLOCK_N + c * 16;
      if (col0 >= args.nH) break;   // warp-uniform
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 16), r);
      tmem_ld_wait();
      if (row < args.nV) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < args.nH && r[j] != 0u) atomicAdd(dst_row + col0 + j, __uint_as_float(r[j]));
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N) : "memory");
  }
}

// S[c] += sum_b X[b, c] over this block's rows; X is a bf16 {0,1} state tile [rows, ld].
// A warp reads 256 consecutive units of one chain (16 B per lane); the eight warps of a block take
// different chains; integer-valued partial sums are combined in shared memory, then one atomic per unit.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, int64_t B, int ld, int n, int64_t rows_per_block,
                   float* __restrict__ S) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < B ? r0 + rows_per_block : B;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < ld) {
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(X + r * ld + c0));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < n) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (s != 0.f) atomicAdd(S + c, s);
  }
}

// ------------------------------------------------------------------ host side
constexpr int kMaxDeviceBetas = 1 << 18;   // longer AIS schedules fall back to one launch per half-step

// Tile shape of the chain kernel: one BLOCK_N for both directions, single CTAs or CTA pairs (256-wide tiles only)
struct ChainCfg { int bn, ctas; };

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Measured on B200 (profiles/r02_chain_tile_sweep.txt): wide tiles and CTA pairs stage the fewest shared-memory bytes
// per flop and win once every CTA has several items per half-step in flight; with few chains narrower tiles
// keep all SMs busy and shorten the per-half-step dependency chain (TMA store -> flag -> TMA load -> MMA -> epilogue).
ChainCfg pick_chain_cfg(int nVp, int nHp, int64_t rows_pad, int sm_count) {
  static const int force_bn = env_int("RBM_B200_CHAIN_BN", 0);
  static const int force_pairs = env_int("RBM_B200_CHAIN_PAIRS", -1);
  auto cols = [&](int bn) { return round_up(nVp, bn) + round_up(nHp, bn); };
  auto items = [&](int bn) { return (rows_pad / BLOCK_M) * (int64_t)std::min(ceil_div(nVp, bn), ceil_div(nHp, bn)); };
  ChainCfg c{64, 1};
  if (force_bn == 64 || force_bn == 128 || force_bn == 256) {
    c.bn = force_bn;
  } else {
    // fewest padded columns first; then the widest tile that still gives every SM an item per half-step
    // (r02_chain_sweep1: at 1024^2 256-wide pair tiles win from 8192 chains up - 1022 vs 782 TFLOP/s for 128-wide)
    const int best = std::min(cols(64), std::min(cols(128), cols(256)));
    for (int bn : {256, 128, 64}) {
      if (cols(bn) != best) continue;
      // small priors (K <= 512: a 256-wide tile is only 8 K blocks deep) with few items run better on 128-wide tiles:
      // 512^2 at 5,120 - 8,192 chains 8.1 - 8.8 us per half-step against 12.3 (profiles/r02_chain_sweep_512.txt)
      // (at 16,384 chains, 256 items, the 256-wide pair tiles win again: 18.2 vs 20.1 us for AIS)
      if (bn == 256 && std::max(nVp, nHp) <= 512 && items(256) < (int64_t)sm_count) continue;
      c.bn = bn;
      if (items(bn) >= (int64_t)sm_count / 2) break;
    }
  }
  c.ctas = 1;
  if (c.bn == 256 && pairs_enabled()) {
    const bool many = items(256) >= (int64_t)sm_count;
    if (force_pairs == 1 || (force_pairs == -1 && many)) c.ctas = 2;
  }
  return c;
}

size_t chain_flag_bytes(int nVp, int nHp, int64_t rows_pad) {
  return (size_t)2 * (size_t)(rows_pad / BLOCK_M) * (size_t)(round_up(std::max(nVp, nHp), 256) / 64) * 4;
}

struct Plan {
  int nV, nH, nVp, nHp;      // padded to multiples of 64
  int64_t rows_pad;          // chains padded to a multiple of 256 (one CTA-pair tile)
  int bn_h, bn_v;            // BLOCK_N of the per-half-step launches (v->h, h->v)
  size_t off_W, off_Wlo, off_nbh, off_nbv, off_V, off_H, off_wpart, off_betas, off_flags, total;
  int wpart_ld;              // AIS log-weight slots per chain: one per 32 hidden units
};

Plan make_plan(int nV, int nH, int64_t B) {
  Plan p;
  p.nV = nV; p.nH = nH;
  p.nVp = round_up(nV, 64); p.nHp = round_up(nH, 64);
  p.rows_pad = ceil_div64(B, 2 * BLOCK_M) * (2 * BLOCK_M);   // whole CTA-pair tiles (256 chains)
  p.bn_h = pick_block_n(p.nHp, p.rows_pad); p.bn_v = pick_block_n(p.nVp, p.rows_pad);
  auto align = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  size_t o = 0;
  p.off_W = o; o = align(o + (size_t)p.nVp * p.nHp * 2);
  p.off_Wlo = o; o = align(o + (size_t)p.nVp * p.nHp * 2);       // low part of split weights (W - bf16(W))
  p.off_nbh = o; o = align(o + (size_t)round_up(p.nHp, 256) * 4);
  p.off_nbv = o; o = align(o + (size_t)round_up(p.nVp, 256) * 4);
  p.off_V = o; o = align(o + (size_t)p.rows_pad * p.nVp * 2);
  p.off_H = o; o = align(o + (size_t)p.rows_pad * p.nHp * 2);
  p.wpart_ld = p.nHp / 32;
  p.off_wpart = o; o = align(o + (size_t)p.rows_pad * p.wpart_ld * 4);
  p.off_betas = o; o = align(o + (size_t)kMaxDeviceBetas * 4);   // AIS schedule for the chain kernel
  // chain kernel: arrival counters per (direction, 128-chain row block, 64-column chunk or wider)
  p.off_flags = o; o = align(o + chain_flag_bytes(p.nVp, p.nHp, p.rows_pad));
  p.total = o + 1024;  // slack to align the caller's pointer
  return p;
}

// Dataflow chain kernel (umma_chain.cuh, all half-steps in ONE launch) or one launch per half-step?
//   RBM_B200_CHAIN = 1: always the chain kernel; 0: always per-half-step launches; unset: by shape.
// Measured at 1024 x 1024, sustained (profiles/r02_chain_vs_launches.txt): the chain kernel wins while a half-step
// has few tiles per cluster - no partial last wave, no launch gap: 17.3 vs 23.0 us per half-step at 8,192 chains,
// 65.3 vs 72.6 at 32,768 - and loses ~3 % at 65,536 chains (13.8 tiles per cluster; 130.6 vs 126.6 us), where the
// launch tail is amortised and the per-tile dependency bookkeeping is not.
bool chain_enabled(int nVp, int nHp, int64_t rows_pad, int sm_count) {
  static const int mode = env_int("RBM_B200_CHAIN", -1);
  if (mode >= 0) return mode != 0;
  static const int max_items = env_int("RBM_B200_CHAIN_MAX_ITEMS", 10);
  const ChainCfg cfg = pick_chain_cfg(nVp, nHp, rows_pad, sm_count);
  const double items = (double)(rows_pad / (BLOCK_M * cfg.ctas)) * std::min(ceil_div(nVp, cfg.bn), ceil_div(nHp, cfg.bn));
  // short K loops (priors below 512 units): an item is only 4 - 6 K blocks long and the dependency bookkeeping per item
  // shows - 2.81 vs 2.26 ms per 100 steps at 256^2 x 16,384 chains, 6.09 vs 4.21 at 384^2 x 32,768 - so the chain kernel
  // is kept to the few-items regime there (profiles/r02_chain_vs_launches.txt)
  const double limit = std::max(nVp, nHp) < 512 ? 2.0 : (double)max_items;
  return items < limit * std::max(1, sm_count / cfg.ctas);
}

// Launches the chain kernel over `steps` half-steps starting with direction 0.
//   W_lo != nullptr: split weights (bf16 high + bf16 low part), K axis doubled
template <int EPI0, int EPI1>
int run_chain(const Plan& pl, uint16_t* Wp, uint16_t* Wlo, const float* nbh, const float* nbv, __nv_bfloat16* Vs,
              __nv_bfloat16* Hs, uint32_t* flags, int steps, uint64_t seed, uint64_t chain_offset, uint64_t step_offset,
              uint32_t tag, const float* betas_dev, float* wpart, int sm_count, cudaStream_t st) {
  const ChainCfg cfg = pick_chain_cfg(pl.nVp, pl.nHp, pl.rows_pad, sm_count);
  const bool split = Wlo != nullptr;
  int rc;
  CUtensorMap tm[8];
  if ((rc = make_tmap(&tm[0], Vs, (uint64_t)pl.rows_pad, (uint64_t)pl.nVp, (uint64_t)pl.nVp, BLOCK_M))) return rc;
  if ((rc = make_tmap(&tm[1], Hs, (uint64_t)pl.rows_pad, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_M))) return rc;
  // W is read MN-major (v->h: [64 k x 64 n] boxes) and K-major (h->v: [n rows x 64] boxes; a CTA of a pair stages half)
  const uint32_t kbox = (uint32_t)(cfg.bn / cfg.ctas);
  if ((rc = make_tmap(&tm[2], Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_K))) return rc;
  if ((rc = make_tmap(&tm[3], Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, kbox))) return rc;
  if ((rc = make_tmap(&tm[4], split ? Wlo : Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_K))) return rc;
  if ((rc = make_tmap(&tm[5], split ? Wlo : Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, kbox))) return rc;
  // epilogue stores: [32 chains x 32 units] boxes, 64-byte swizzle
  if ((rc = make_tmap(&tm[6], Hs, (uint64_t)pl.rows_pad, (uint64_t)pl.nHp, (uint64_t)pl.nHp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap(&tm[7], Vs, (uint64_t)pl.rows_pad, (uint64_t)pl.nVp, (uint64_t)pl.nVp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  ChainArgs a{};
  a.nbias[0] = nbh; a.nbias[1] = nbv; a.ldo[0] = pl.nHp; a.ldo[1] = pl.nVp;
  a.n_valid[0] = pl.nH; a.n_valid[1] = pl.nV;
  a.n_tiles[0] = ceil_div(pl.nHp, cfg.bn); a.n_tiles[1] = ceil_div(pl.nVp, cfg.bn);
  a.k_blocks[0] = pl.nVp / BLOCK_K; a.k_blocks[1] = pl.nHp / BLOCK_K;
  a.m_groups = (int)(pl.rows_pad / (BLOCK_M * cfg.ctas));
  a.steps = steps; a.first_dir = 0; a.seed = seed; a.chain_offset = chain_offset; a.step_offset = step_offset;
  a.tag = tag; a.betas = betas_dev; a.wpart = wpart; a.wpart_ld = pl.wpart_ld; a.kf_magic = kKfMagic; a.flags = flags;
  a.row_blocks = (int)(pl.rows_pad / BLOCK_M);
  // Publication policy (measured, profiles/r02_chain_opts.txt): a dependency warp polls ahead of the producer and the
  // writer side needs no proxy fence in every case.  With fewer than ~2.5 items per cluster and half-step the call
  // is bound by the half-step to half-step latency: chunks are published one by one, half a group after their
  // store; with more items the consumers are far behind and a tile's chunks are published together.
  static const int chain_opts_env = env_int("RBM_B200_CHAIN_OPTS", -1);
  {
    const double items_per_cluster = (double)a.m_groups * std::min(a.n_tiles[0], a.n_tiles[1]) / std::max(1, sm_count / cfg.ctas);
    // (deferring a tile's last publication past the accumulator wait bought nothing: r02_chain_sweep5)
    // (one gpu-scope release per chunk and CTA instead of one per epilogue warp: 15.7 vs 16.7 us per half-step at 8,192 chains)
    const int base = CHAIN_OPT_DEPWARP | CHAIN_OPT_CHUNKED | CHAIN_OPT_NO_WFENCE | CHAIN_OPT_NO_DEFER | CHAIN_OPT_CTA_ARRIVE;
    a.opts = chain_opts_env >= 0 ? chain_opts_env
                                 : (items_per_cluster < 2.5 ? (base | CHAIN_OPT_MIDARRIVE) : (base | CHAIN_OPT_TILE_ARRIVE));
  }
  // chunks of 32 * kSets columns (EpiSplit): 128 for tiles >= 128 wide, 64 for 64-wide tiles
  const int chunk_w = cfg.bn >= 128 ? 128 : 64;
  a.chunks_ld = std::max(a.n_tiles[0], a.n_tiles[1]) * (cfg.bn / chunk_w);
  RBM_REQUIRE((size_t)2 * a.row_blocks * a.chunks_ld * 4 <= chain_flag_bytes(pl.nVp, pl.nHp, pl.rows_pad),
              RBM_B200_ERR_WORKSPACE, "chain kernel: counter region too small");
  RBM_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)2 * a.row_blocks * a.chunks_ld * 4, st));
  // Super-blocks: all half-steps for as many chains as keep both state tiles L2-resident, then the next chains.
  // RBM_B200_CHAIN_SB = chains per super-block (0: automatic, -1: one super-block)
  {
    static const int sb_chains_env = env_int("RBM_B200_CHAIN_SB", 0);
    DeviceInfo di;
    if ((rc = get_device_info(&di))) return rc;
    const int64_t group_rows = (int64_t)BLOCK_M * cfg.ctas;
    const double bytes_per_chain = 2.0 * (pl.nVp + pl.nHp);
    int64_t n_sb = 1;
    if (sb_chains_env > 0) n_sb = ceil_div64(pl.rows_pad, sb_chains_env);
    else if (sb_chains_env == 0) n_sb = (int64_t)std::ceil(pl.rows_pad * bytes_per_chain / (0.55 * (double)di.l2_bytes));
    n_sb = std::max<int64_t>(1, std::min<int64_t>(n_sb, a.m_groups));
    a.sb_groups = (int)ceil_div64(ceil_div64(pl.rows_pad, n_sb), group_rows);
    // never so small that a half-step of a super-block has fewer than ~3 items per cluster (dependency slack)
    const int min_groups = sb_chains_env > 0 ? 1 : ceil_div(3 * (sm_count / cfg.ctas), std::min(a.n_tiles[0], a.n_tiles[1]));
    a.sb_groups = std::min(a.m_groups, std::max(a.sb_groups, min_groups));
  }
  // every cluster must be resident (items wait for lower-numbered items of other clusters)
  static const int grid_cap = env_int("RBM_B200_CHAIN_GRID", 0);
  int64_t total_items = 0;
  for (int t = 0; t < std::min(steps, 2); ++t) total_items += (int64_t)a.m_groups * a.n_tiles[t & 1];
  if (steps > 2) total_items = (int64_t)1 << 40;
  int grid = sm_count / cfg.ctas;
  // Fewer items per half-step than clusters: one cluster per item, so that the items a cluster takes one after the other
  // are always one half-step apart (with 120 - 128 items dealt over 148 CTAs a cluster's next item can sit in the same
  // half-step as items it depends on, queued behind unfinished work elsewhere: 25 - 35 us per half-step instead of 8
  // at 512^2, profiles/r02_chain_sweep_512.txt).
  {
    const int64_t items_hs = (int64_t)std::min(a.sb_groups, a.m_groups) * std::max(a.n_tiles[0], a.n_tiles[1]);
    if (items_hs < grid) grid = (int)items_hs;
    static const int align_grid = env_int("RBM_B200_CHAIN_ALIGN", 0);   // experiment: grid = items / ceil(items / grid)
    if (align_grid && items_hs < 4 * (int64_t)grid) grid = (int)(items_hs / ceil_div64(items_hs, grid));
  }
  if (grid_cap > 0) grid = std::min(grid, grid_cap);
  grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid, total_items));
  return launch_chain<EPI0, EPI1>(cfg.bn, cfg.ctas, split, tm, a, grid, st);
}

// Resident kernel (umma_resident.cuh): W in shared memory, states exchanged as bit words inside a cluster.
//   RBM_B200_RESIDENT = 1: whenever the shape fits; 0: never; unset: by shape.
// Measured on B200 (profiles/r02_resident_vs_chain.txt, us per half-step): 512 x 512 Gibbs 3.96 vs 7.13 at 1,024 chains
// and 5.55 vs 7.91 at 3,840; AIS 4.76 vs 8.2 and 6.7 vs 9.3; 256 x 256 2.87 vs 5.97.  Its time grows in steps of one
// "round" (two row blocks on every resident cluster: 30 x 128 chains), so beyond one round the dataflow chain kernel,
// which fills all SMs, takes over.
bool resident_enabled(int nVp, int nHp, int64_t rows_pad, bool split_w, int sm_count) {
  static const int mode = env_int("RBM_B200_RESIDENT", -1);
  if (split_w || !resident_supported(nVp, nHp)) return false;
  if (mode >= 0) return mode != 0;
  const int cluster = resident_cluster_size(nHp / kResBN, nVp / kResBN);
  int max_clusters = 0;
  if (resident_max_clusters<EPI_GIBBS, EPI_GIBBS>(cluster, resident_smem(nHp / kResBN, nVp / kResBN).total, sm_count, &max_clusters))
    return false;
  return rows_pad / BLOCK_M <= 2 * (int64_t)max_clusters;
}

template <int EPI0, int EPI1>
int run_resident(const Plan& pl, uint16_t* Wp, const float* nbh, const float* nbv, __nv_bfloat16* Vs, __nv_bfloat16* Hs,
                 int steps, uint64_t seed, uint64_t chain_offset, uint64_t step_offset, uint32_t tag,
                 const float* betas_dev, float* wpart, int sm_count, cudaStream_t st) {
  int rc;
  CUtensorMap tm_w;
  // one [64 x 64] box serves both directions: (v rows = k, h cols = n) MN-major, (v rows = n, h cols = k) K-major
  if ((rc = make_tmap(&tm_w, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_K))) return rc;
  ResidentArgs a{};
  a.nbias[0] = nbh; a.nbias[1] = nbv;
  a.n_valid[0] = pl.nH; a.n_valid[1] = pl.nV;
  a.n_tiles[0] = pl.nHp / kResBN; a.n_tiles[1] = pl.nVp / kResBN;
  a.ld[0] = pl.nHp; a.ld[1] = pl.nVp;
  a.row_blocks = (int)(pl.rows_pad / BLOCK_M);
  a.steps = steps; a.seed = seed; a.chain_offset = chain_offset; a.step_offset = step_offset; a.tag = tag;
  a.betas = betas_dev; a.wpart = wpart; a.wpart_ld = pl.wpart_ld; a.kf_magic = kKfMagic;
  a.v_init = Vs; a.out[0] = Hs; a.out[1] = Vs;
  return launch_resident<EPI0, EPI1>(tm_w, a, resident_cluster_size(a.n_tiles[0], a.n_tiles[1]), sm_count, st);
}

template <int BLOCK_N>
int launch_stats(const CUtensorMap& tv, const CUtensorMap& th, const StatsArgs& a, int grid, cudaStream_t st) {
  auto kern = umma_stats_kernel<BLOCK_N>;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if ((rc = ensure_dynamic_smem(kern, di.device, stats_smem_bytes<BLOCK_N>()))) return rc;
  kern<<<(unsigned)grid, kStatsThreads, stats_smem_bytes<BLOCK_N>(), st>>>(tv, th, a);
  return check_launch("umma_stats_kernel");
}

// C[M, N] += A^T B over `kdim` rows: A is [kdim, >= M] bf16 with row pitch lda (elements), B is [kdim, >= N] bf16 with row
// pitch ldb; both operands are read MN-major straight from row-major storage ([64 rows x 64 columns] TMA boxes), split-K
// over CTAs, fp32 partial tiles ADDED with red.global.  C has row pitch N.
int launch_tn_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t kdim, int M, int N, float* C, int sm_count,
                   cudaStream_t st) {
  int rc;
  CUtensorMap tv, th;
  // kdim rows: rows beyond the end read as zeros
  if ((rc = make_tmap(&tv, A, (uint64_t)kdim, (uint64_t)round_up(M, 64), (uint64_t)lda, BLOCK_K))) return rc;
  if ((rc = make_tmap(&th, B, (uint64_t)kdim, (uint64_t)round_up(N, 64), (uint64_t)ldb, BLOCK_K))) return rc;
  const int Np = round_up(N, 64);
  const int bn = round_up(Np, 128) < round_up(Np, 256) ? 128 : 256;
  StatsArgs a{};
  a.S_vh = C; a.nV = M; a.nH = N;
  const int tiles_m = ceil_div(M, BLOCK_M);
  a.tiles_n = ceil_div(N, bn);
  a.tiles = tiles_m * a.tiles_n;
  a.k_blocks = (int)ceil_div64(kdim, BLOCK_K);
  // split the rows so that one wave of CTAs covers the GPU; every split owns at least one K block
  int splits = std::max(1, std::min(a.k_blocks, sm_count / a.tiles));
  a.kb_per_split = ceil_div(a.k_blocks, splits);
  splits = ceil_div(a.k_blocks, a.kb_per_split);
  return bn == 256 ? launch_stats<256>(tv, th, a, a.tiles * splits, st) : launch_stats<128>(tv, th, a, a.tiles * splits, st);
}

// Raw counts of the final (v, h) pair of this call's chains, ADDED to S_vh [nV, nH], S_v [nV], S_h [nH].
int umma_stats(const Plan& pl, const __nv_bfloat16* Vs, const __nv_bfloat16* Hs, int64_t B, float* S_vh, float* S_v,
               float* S_h, int sm_count, cudaStream_t st) {
  RBM_REQUIRE(B <= (int64_t)1 << 24, RBM_B200_ERR_INVALID,
              "fused statistics count in fp32: at most 2^24 chains per call (got %lld)", (long long)B);
  // B rows: the padding chains read as zeros
  int rc = launch_tn_gemm(Vs, pl.nVp, Hs, pl.nHp, B, pl.nV, pl.nH, S_vh, sm_count, st);
  if (rc) return rc;
  const int64_t rows_per_block = std::max<int64_t>(64, ceil_div64(B, 4 * sm_count));
  const unsigned row_blocks = (unsigned)ceil_div64(B, rows_per_block);
  colsum_bf16_kernel<<<dim3((unsigned)ceil_div(pl.nVp, 256), row_blocks), 256, 0, st>>>(Vs, B, pl.nVp, pl.nV, rows_per_block, S_v);
  colsum_bf16_kernel<<<dim3((unsigned)ceil_div(pl.nHp, 256), row_blocks), 256, 0, st>>>(Hs, B, pl.nHp, pl.nH, rows_per_block, S_h);
  return check_launch("colsum_bf16_kernel");
}

// x [rows, n] fp32 -> [rows, 2 np] bf16 = [x_hi | x_lo] (x = x_hi + x_lo to ~2^-17), columns zero padded to np
__global__ void split_rows_kernel(const float* __restrict__ x, int64_t rows, int n, uint16_t* __restrict__ dst, int np) {
  const int groups = np >> 3;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * groups) return;
  const int64_t r = idx / groups;
  const int c0 = (int)(idx % groups) * 8;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float val = c < n ? __ldg(x + r * n + c) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(val);
    const __nv_bfloat16 l = __float2bfloat16_rn(val - __bfloat162float(h));
    if (j & 1) { hi[j >> 1] |= (uint32_t)__bfloat16_as_ushort(h) << 16; lo[j >> 1] |= (uint32_t)__bfloat16_as_ushort(l) << 16; }
    else { hi[j >> 1] = __bfloat16_as_ushort(h); lo[j >> 1] = __bfloat16_as_ushort(l); }
  }
  uint16_t* row = dst + r * (2 * (int64_t)np);
  *reinterpret_cast<uint4*>(row + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(row + np + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace

#ifdef RBM_RES_TRACE
}  // namespace rbm
extern "C" __attribute__((visibility("default"))) int rbm_b200_debug_resident_trace(unsigned long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, rbm::g_res_trace, sizeof(unsigned long long) * (size_t)n);
}
namespace rbm {
#endif

// C[M, N] = A^T B for fp32 A [kdim, M], B [kdim, N] (the weight gradient dW = dY^T X of a linear layer, the <v h^T>
// gradient of the energy): operands split into bf16 high + low parts, three tensor-core products
// (hi hi + hi lo + lo hi, ~2^-16 relative), partial tiles of the K splits added in fp32.
size_t umma_gemm_tn_workspace_bytes(int64_t kdim, int M, int N) {
  return (size_t)kdim * 2 * (size_t)(round_up(M, 64) + round_up(N, 64)) * 2 + 2048;
}

int umma_gemm_tn(const float* A, const float* B, int64_t kdim, int M, int N, float* C, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  RBM_REQUIRE(ws != nullptr && ws_bytes >= umma_gemm_tn_workspace_bytes(kdim, M, N), RBM_B200_ERR_WORKSPACE,
              "gemm (A^T B) needs %zu bytes of workspace, got %zu", umma_gemm_tn_workspace_bytes(kdim, M, N), ws_bytes);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  RBM_REQUIRE(di.cc_major == 10, RBM_B200_ERR_NO_DEVICE, "the tensor-core GEMM needs an sm_100 device (got cc %d.x)", di.cc_major);
  RBM_CUDA_TRY(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
  if (kdim == 0) return 0;
  const int Mp = round_up(M, 64), Np = round_up(N, 64);
  uint16_t* At = reinterpret_cast<uint16_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  uint16_t* Bt = At + (size_t)kdim * 2 * Mp;
  split_rows_kernel<<<(unsigned)ceil_div64(kdim * (Mp / 8), 256), 256, 0, st>>>(A, kdim, M, At, Mp);
  split_rows_kernel<<<(unsigned)ceil_div64(kdim * (Np / 8), 256), 256, 0, st>>>(B, kdim, N, Bt, Np);
  if ((rc = check_launch("split_rows_kernel"))) return rc;
  // (hi, hi), (hi, lo), (lo, hi)
  if ((rc = launch_tn_gemm(At, 2 * Mp, Bt, 2 * Np, kdim, M, N, C, di.sm_count, st))) return rc;
  if ((rc = launch_tn_gemm(At, 2 * Mp, Bt + Np, 2 * Np, kdim, M, N, C, di.sm_count, st))) return rc;
  return launch_tn_gemm(At + Mp, 2 * Mp, Bt, 2 * Np, kdim, M, N, C, di.sm_count, st);
}

bool umma_variant_supported(int nV, int nH) { return nV >= 1 && nH >= 1 && nV <= 16384 && nH <= 16384; }

size_t umma_workspace_bytes(int nV, int nH, int64_t B) { return make_plan(nV, nH, B).total; }

int umma_stats_tiles(const void* v_tiles, const void* h_tiles, int64_t B, int nV, int nH, float* S_vh, float* S_v,
                     float* S_h, cudaStream_t st) {
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  RBM_REQUIRE(di.cc_major == 10, RBM_B200_ERR_NO_DEVICE, "tensor-core statistics need an sm_100 device (got cc %d.x)", di.cc_major);
  const Plan pl = make_plan(nV, nH, B);
  return umma_stats(pl, reinterpret_cast<const __nv_bfloat16*>(v_tiles), reinterpret_cast<const __nv_bfloat16*>(h_tiles), B,
                    S_vh, S_v, S_h, di.sm_count, st);
}

// which kernel runs the half-steps of a UMMA-variant call of this shape (reporting / tests)
const char* umma_kernel_choice(int nV, int nH, int64_t B, bool split_w) {
  const Plan pl = make_plan(nV, nH, B);
  DeviceInfo di;
  if (get_device_info(&di)) return "unknown (no device)";
  if (resident_enabled(pl.nVp, pl.nHp, pl.rows_pad, split_w, di.sm_count)) return "umma_resident_kernel";
  if (split_w || chain_enabled(pl.nVp, pl.nHp, pl.rows_pad, di.sm_count)) return "umma_chain_kernel";
  return "umma_half_step_kernel";
}

// kernels one umma_block_gibbs call launches (no statistics): 3 parameter staging + 1 initial state + chain + 2 outputs
int64_t umma_launch_count(int nV, int nH, int64_t B, int k, bool split_w) {
  if (k <= 0) return 5;
  const Plan pl = make_plan(nV, nH, B);
  DeviceInfo di;
  if (get_device_info(&di)) return -1;
  const bool chain = split_w || resident_enabled(pl.nVp, pl.nHp, pl.rows_pad, split_w, di.sm_count) ||
                     chain_enabled(pl.nVp, pl.nHp, pl.rows_pad, di.sm_count);
  return 6 + (chain ? 1 : 2 * (int64_t)k);
}

int umma_block_gibbs(const rbm_b200_params* p, float* v, float* h, int64_t B, int k, int init_chains,
                     uint64_t seed, uint64_t chain_offset, uint64_t step_offset, void* ws, size_t ws_bytes,
                     cudaStream_t st, float* S_vh, float* S_v, float* S_h, bool split_w) {
  const Plan pl = make_plan(p->n_visible, p->n_hidden, B);
  RBM_REQUIRE(ws != nullptr && ws_bytes >= pl.total, RBM_B200_ERR_WORKSPACE,
              "UMMA variant needs %zu bytes of workspace, got %zu", pl.total, ws_bytes);
  DeviceInfo di;
  {
    int rc_di = get_device_info(&di);
    if (rc_di) return rc_di;
  }
  const int sm_count = di.sm_count, cc_major = di.cc_major;
  RBM_REQUIRE(cc_major == 10, RBM_B200_ERR_NO_DEVICE, "UMMA variant needs an sm_100 device (got cc %d.x)", cc_major);

  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  uint16_t* Wp = reinterpret_cast<uint16_t*>(base + pl.off_W);
  uint16_t* Wlo = split_w ? reinterpret_cast<uint16_t*>(base + pl.off_Wlo) : nullptr;
  float* nbh = reinterpret_cast<float*>(base + pl.off_nbh);
  float* nbv = reinterpret_cast<float*>(base + pl.off_nbv);
  __nv_bfloat16* Vs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_V);
  __nv_bfloat16* Hs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_H);

  // stage parameters: padded bf16 W (or its high + low parts from the fp32 W), pre-scaled biases
  {
    const int64_t n = (int64_t)pl.nVp * pl.nHp;
    if (split_w)
      pad_w_split_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W, reinterpret_cast<__nv_bfloat16*>(Wp),
                                                                      reinterpret_cast<__nv_bfloat16*>(Wlo), pl.nV, pl.nH, pl.nVp, pl.nHp);
    else if (p->W_bf16 != nullptr)
      pad_w_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W_bf16, Wp, pl.nV, pl.nH, pl.nVp, pl.nHp);
    else
      pad_w_f32_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W, reinterpret_cast<__nv_bfloat16*>(Wp), pl.nV, pl.nH,
                                                                    pl.nVp, pl.nHp);
    const int nh256 = round_up(pl.nHp, 256), nv256 = round_up(pl.nVp, 256);
    nbias_kernel<<<ceil_div(nh256, 256), 256, 0, st>>>(p->hbias, nbh, pl.nH, nh256);
    nbias_kernel<<<ceil_div(nv256, 256), 256, 0, st>>>(p->vbias, nbv, pl.nV, nv256);
  }
  // initial state
  if (init_chains) {
    const int64_t n = pl.rows_pad * (pl.nVp / 32);
    init_chains_bf16_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(Vs, pl.rows_pad, pl.nV, pl.nVp,
                                                                         make_draw_ctx(seed, step_offset, TAG_INIT), chain_offset);
  } else {
    const int64_t n = pl.rows_pad * pl.nVp;
    f32_to_bf16_padded<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(v, Vs, B, pl.nV, pl.rows_pad, pl.nVp);
  }
  int rc = check_launch("umma staging kernels");
  if (rc) return rc;

  if (k > 0 && resident_enabled(pl.nVp, pl.nHp, pl.rows_pad, split_w, sm_count)) {
    // priors up to 512 x 512: W resident in shared memory, states on the chip (umma_resident.cuh)
    if ((rc = run_resident<EPI_GIBBS, EPI_GIBBS>(pl, Wp, nbh, nbv, Vs, Hs, 2 * k, seed, chain_offset, step_offset, TAG_GIBBS,
                                                 nullptr, nullptr, sm_count, st))) return rc;
  } else if (k > 0 && (split_w || chain_enabled(pl.nVp, pl.nHp, pl.rows_pad, sm_count))) {
    // all 2k half-steps in ONE dataflow launch (umma_chain.cuh)
    uint32_t* flags = reinterpret_cast<uint32_t*>(base + pl.off_flags);
    if ((rc = run_chain<EPI_GIBBS, EPI_GIBBS>(pl, Wp, Wlo, nbh, nbv, Vs, Hs, flags, 2 * k, seed, chain_offset, step_offset,
                                              TAG_GIBBS, nullptr, nullptr, sm_count, st))) return rc;
  } else if (k > 0) {
    // round-1 path, kept for A/B measurements (RBM_B200_CHAIN=0): one launch per half-step
    static const int pairs_min_tiles = env_int("RBM_B200_PAIRS_MIN_TILES", 8);
    // tensor maps: states are K-major A operands; W is read MN-major (v->h) and K-major (h->v)
    CUtensorMap tm_w_mn, tm_w_k_single, tm_w_k_pair;
    if ((rc = make_tmap(&tm_w_mn, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_K))) return rc;
    // h->v reads W K-major in [n rows x 64] boxes: a CTA of a pair stages half of the 256 rows
    if ((rc = make_tmap(&tm_w_k_single, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, (uint32_t)pl.bn_v))) return rc;
    if ((rc = make_tmap(&tm_w_k_pair, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, 128))) return rc;
    const int64_t rows = pl.rows_pad;
    // pairs pay off once every cluster loops over enough tiles (profiles/r01_pairs_vs_single.txt)
    const bool pairs = pairs_enabled() &&
                       (rows / (2 * BLOCK_M)) * ceil_div(std::max(pl.nVp, pl.nHp), 256) >= (int64_t)pairs_min_tiles * (sm_count / 2);
    CUtensorMap tm_v, tm_h, to_v, to_h;
    if ((rc = make_tmap(&tm_v, Vs, (uint64_t)rows, (uint64_t)pl.nVp, (uint64_t)pl.nVp, BLOCK_M))) return rc;
    if ((rc = make_tmap(&tm_h, Hs, (uint64_t)rows, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_M))) return rc;
    // epilogue stores: [32 chains x 32 units] boxes, 64-byte swizzle
    if ((rc = make_tmap(&to_v, Vs, (uint64_t)rows, (uint64_t)pl.nVp, (uint64_t)pl.nVp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_tmap(&to_h, Hs, (uint64_t)rows, (uint64_t)pl.nHp, (uint64_t)pl.nHp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    const CUtensorMap& tm_w_k = (pairs && pl.bn_v == 256) ? tm_w_k_pair : tm_w_k_single;

    HalfStepArgs ah{}, av{};
    ah.out = Hs; ah.nbias = nbh; ah.ldo = pl.nHp; ah.m_tiles = (int)(rows / BLOCK_M);
    ah.n_tiles = ceil_div(pl.nHp, pl.bn_h); ah.k_blocks = pl.nVp / BLOCK_K; ah.chain_offset = chain_offset; ah.n_valid = pl.nH;
    av.out = Vs; av.nbias = nbv; av.ldo = pl.nVp; av.m_tiles = ah.m_tiles;
    av.n_tiles = ceil_div(pl.nVp, pl.bn_v); av.k_blocks = pl.nHp / BLOCK_K; av.chain_offset = chain_offset; av.n_valid = pl.nV;

    for (int s = 0; s < k; ++s) {
      ah.ctx = make_draw_ctx(seed, step_offset + 2ull * s, TAG_GIBBS);
      rc = pl.bn_h == 256 ? launch_half_step<256, true>(tm_v, tm_w_mn, to_h, ah, sm_count, pairs, st)
                          : launch_half_step<128, true>(tm_v, tm_w_mn, to_h, ah, sm_count, pairs, st);
      if (rc) return rc;
      av.ctx = make_draw_ctx(seed, step_offset + 2ull * s + 1, TAG_GIBBS);
      rc = pl.bn_v == 256 ? launch_half_step<256, false>(tm_h, tm_w_k, to_v, av, sm_count, pairs, st)
                          : launch_half_step<128, false>(tm_h, tm_w_k, to_v, av, sm_count, pairs, st);
      if (rc) return rc;
    }
  }
  // negative-phase statistics of the final (v, h), on the tensor cores, from the bf16 state tiles
  if (S_vh != nullptr && k > 0) {
    if ((rc = umma_stats(pl, Vs, Hs, B, S_vh, S_v, S_h, sm_count, st))) return rc;
  }
  // back to the reference's dtype
  {
    const int64_t nv = B * pl.nV, nh = B * pl.nH;
    if (pl.nV % 8 == 0 && ((uintptr_t)v & 15) == 0) bf16_to_f32_crop8<<<(unsigned)ceil_div64(nv / 8, 256), 256, 0, st>>>(Vs, v, B, pl.nV, pl.nVp);
    else bf16_to_f32_crop<<<(unsigned)ceil_div64(nv, 256), 256, 0, st>>>(Vs, v, B, pl.nV, pl.nVp);
    if (k > 0) {
      if (pl.nH % 8 == 0 && ((uintptr_t)h & 15) == 0) bf16_to_f32_crop8<<<(unsigned)ceil_div64(nh / 8, 256), 256, 0, st>>>(Hs, h, B, pl.nH, pl.nHp);
      else bf16_to_f32_crop<<<(unsigned)ceil_div64(nh, 256), 256, 0, st>>>(Hs, h, B, pl.nH, pl.nHp);
    }
  }
  return check_launch("umma output conversion");
}

size_t umma_ais_workspace_bytes(int nV, int nH, int64_t M) { return make_plan(nV, nH, M).total; }

// Annealed importance sampling (Salakhutdinov & Murray 2008, RBM form).  The
// reference has no estimator (models/rbm/rbm.py:51-54 returns a literal); the CPU restatement
// is oracle/rbm_oracle.py:ais_logZ.  Per beta step: ONE GEMM x = vW + c feeds both the weight
// increment sum_j softplus(b_k x_j) - softplus(b_{k-1} x_j) and the draw h ~ sigma(b_k x);
// then v ~ sigma(a + b_k h W^T).  logw[b] (nats) is written for this shard's chains.
int umma_ais_run(const rbm_b200_params* p, int64_t M, const double* betas, int n_betas, uint64_t seed,
                 uint64_t chain_offset, double* logw, void* ws, size_t ws_bytes, cudaStream_t st, bool split_w) {
  const Plan pl = make_plan(p->n_visible, p->n_hidden, M);
  RBM_REQUIRE(ws != nullptr && ws_bytes >= pl.total, RBM_B200_ERR_WORKSPACE,
              "AIS needs %zu bytes of workspace, got %zu", pl.total, ws_bytes);
  DeviceInfo di;
  {
    int rc_di = get_device_info(&di);
    if (rc_di) return rc_di;
  }
  const int sm_count = di.sm_count, cc_major = di.cc_major;
  RBM_REQUIRE(cc_major == 10, RBM_B200_ERR_NO_DEVICE, "AIS needs an sm_100 device (got cc %d.x)", cc_major);
  const bool use_resident = n_betas <= kMaxDeviceBetas && resident_enabled(pl.nVp, pl.nHp, pl.rows_pad, split_w, sm_count);
  const bool use_chain = use_resident ||
                         (n_betas <= kMaxDeviceBetas && (split_w || chain_enabled(pl.nVp, pl.nHp, pl.rows_pad, sm_count)));
  RBM_REQUIRE(!split_w || use_chain, RBM_B200_ERR_UNSUPPORTED,
              "AIS with split weights needs the chain kernel (at most %d betas)", kMaxDeviceBetas);

  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  uint16_t* Wp = reinterpret_cast<uint16_t*>(base + pl.off_W);
  uint16_t* Wlo = split_w ? reinterpret_cast<uint16_t*>(base + pl.off_Wlo) : nullptr;
  float* nbh = reinterpret_cast<float*>(base + pl.off_nbh);
  float* nbv = reinterpret_cast<float*>(base + pl.off_nbv);
  __nv_bfloat16* Vs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_V);
  __nv_bfloat16* Hs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_H);
  float* wpart = reinterpret_cast<float*>(base + pl.off_wpart);

  {
    const int64_t n = (int64_t)pl.nVp * pl.nHp;
    if (split_w)
      pad_w_split_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W, reinterpret_cast<__nv_bfloat16*>(Wp),
                                                                      reinterpret_cast<__nv_bfloat16*>(Wlo), pl.nV, pl.nH, pl.nVp, pl.nHp);
    else if (p->W_bf16 != nullptr)
      pad_w_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W_bf16, Wp, pl.nV, pl.nH, pl.nVp, pl.nHp);
    else
      pad_w_f32_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W, reinterpret_cast<__nv_bfloat16*>(Wp), pl.nV, pl.nH,
                                                                    pl.nVp, pl.nHp);
    const int nh256 = round_up(pl.nHp, 256), nv256 = round_up(pl.nVp, 256);
    nbias_kernel<<<ceil_div(nh256, 256), 256, 0, st>>>(p->hbias, nbh, pl.nH, nh256);
    nbias_kernel<<<ceil_div(nv256, 256), 256, 0, st>>>(p->vbias, nbv, pl.nV, nv256);
    RBM_CUDA_TRY(cudaMemsetAsync(wpart, 0, (size_t)pl.rows_pad * pl.wpart_ld * 4, st));
    const int64_t nb = pl.rows_pad * (pl.nVp / 8);
    ais_init_kernel<<<(unsigned)ceil_div64(nb, 256), 256, 0, st>>>(Vs, pl.rows_pad, pl.nV, pl.nVp, nbv,
                                                                  make_draw_ctx(seed, 0, TAG_AIS_INIT), chain_offset);
  }
  int rc = check_launch("ais staging kernels");
  if (rc) return rc;
  const int K = n_betas - 1;

  if (use_chain) {
    // the whole schedule in ONE dataflow launch: K hidden steps, K-1 visible steps (the last transition cannot
    // change the estimate); half-step counter 2k / 2k+1, beta index 1 + t/2
    float* betas_dev = reinterpret_cast<float*>(base + pl.off_betas);
    std::vector<float> bf(n_betas);
    for (int i = 0; i < n_betas; ++i) bf[i] = (float)betas[i];
    // pageable source: the runtime stages the copy before returning, so `bf` may go out of scope
    RBM_CUDA_TRY(cudaMemcpyAsync(betas_dev, bf.data(), sizeof(float) * (size_t)n_betas, cudaMemcpyHostToDevice, st));
    uint32_t* flags = reinterpret_cast<uint32_t*>(base + pl.off_flags);
    if (use_resident) {
      if ((rc = run_resident<EPI_AIS_WEIGHT, EPI_ANNEALED>(pl, Wp, nbh, nbv, Vs, Hs, 2 * K - 1, seed, chain_offset, 2, TAG_AIS,
                                                           betas_dev, wpart, sm_count, st))) return rc;
    } else if ((rc = run_chain<EPI_AIS_WEIGHT, EPI_ANNEALED>(pl, Wp, Wlo, nbh, nbv, Vs, Hs, flags, 2 * K - 1, seed, chain_offset, 2,
                                                             TAG_AIS, betas_dev, wpart, sm_count, st))) return rc;
    ais_reduce_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, st>>>(wpart, pl.wpart_ld, M, logw);
    return check_launch("ais_reduce_kernel");
  }

  // one launch per half-step (schedules longer than kMaxDeviceBetas, or RBM_B200_CHAIN=0)
  CUtensorMap tm_v, tm_h, tm_w_mn, tm_w_k;
  if ((rc = make_tmap(&tm_v, Vs, (uint64_t)pl.rows_pad, (uint64_t)pl.nVp, (uint64_t)pl.nVp, BLOCK_M))) return rc;
  if ((rc = make_tmap(&tm_h, Hs, (uint64_t)pl.rows_pad, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_M))) return rc;
  if ((rc = make_tmap(&tm_w_mn, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_K))) return rc;
  // pairs pay off once every cluster loops over >= 8 tiles (profiles/r01_pairs_vs_single.txt)
  const bool pairs = pairs_enabled() && (pl.rows_pad / (2 * BLOCK_M)) * ceil_div(std::max(pl.nVp, pl.nHp), 256) >= 8 * (sm_count / 2);
  // h->v reads W K-major in [n rows x 64] boxes: a CTA of a pair stages half of the 256 rows
  if ((rc = make_tmap(&tm_w_k, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp,
                      (uint32_t)((pairs && pl.bn_v == 256) ? 128 : pl.bn_v)))) return rc;
  // epilogue stores: [32 chains x 32 units] boxes, 64-byte swizzle
  CUtensorMap to_v, to_h;
  if ((rc = make_tmap(&to_v, Vs, (uint64_t)pl.rows_pad, (uint64_t)pl.nVp, (uint64_t)pl.nVp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  if ((rc = make_tmap(&to_h, Hs, (uint64_t)pl.rows_pad, (uint64_t)pl.nHp, (uint64_t)pl.nHp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;

  HalfStepArgs ah{}, av{};
  ah.out = Hs; ah.nbias = nbh; ah.ldo = pl.nHp; ah.m_tiles = (int)(pl.rows_pad / BLOCK_M);
  ah.n_tiles = ceil_div(pl.nHp, pl.bn_h); ah.k_blocks = pl.nVp / BLOCK_K; ah.chain_offset = chain_offset;
  ah.n_valid = pl.nH; ah.wpart = wpart; ah.wpart_ld = pl.wpart_ld;
  av.out = Vs; av.nbias = nbv; av.ldo = pl.nVp; av.m_tiles = ah.m_tiles;
  av.n_tiles = ceil_div(pl.nVp, pl.bn_v); av.k_blocks = pl.nHp / BLOCK_K; av.chain_offset = chain_offset;
  av.n_valid = pl.nV; av.wpart = wpart; av.wpart_ld = pl.wpart_ld;

  for (int k = 1; k <= K; ++k) {
    const float bk = (float)betas[k], bkm1 = (float)betas[k - 1];
    ah.ctx = make_draw_ctx(seed, 2ull * k, TAG_AIS);
    ah.scale = -1.4426950408889634f * bk; ah.bmul = bk; ah.ratio = bkm1 / bk;
    rc = pl.bn_h == 256 ? launch_half_step<256, true, EPI_AIS_WEIGHT>(tm_v, tm_w_mn, to_h, ah, sm_count, pairs, st)
                        : launch_half_step<128, true, EPI_AIS_WEIGHT>(tm_v, tm_w_mn, to_h, ah, sm_count, pairs, st);
    if (rc) return rc;
    if (k == K) break;  // the last transition cannot change the estimate
    av.ctx = make_draw_ctx(seed, 2ull * k + 1, TAG_AIS);
    av.scale = -1.4426950408889634f * bk; av.bmul = 1.0f; av.ratio = 0.f;
    rc = pl.bn_v == 256 ? launch_half_step<256, false, EPI_ANNEALED>(tm_h, tm_w_k, to_v, av, sm_count, pairs, st)
                        : launch_half_step<128, false, EPI_ANNEALED>(tm_h, tm_w_k, to_v, av, sm_count, pairs, st);
    if (rc) return rc;
  }
  ais_reduce_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, st>>>(wpart, pl.wpart_ld, M, logw);
  return check_launch("ais_reduce_kernel");
}

}  // namespace rbm
