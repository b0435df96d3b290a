// Shared machinery of the tcgen05/TMEM (UMMA) paths: PTX wrappers, shared-memory and instruction descriptors,
// the persistent warp-specialised GEMM kernel with its fused epilogues, tensor-map and launch helpers.
// Included by gibbs_umma.cu (Gibbs half-steps, AIS, statistics) and umma_decoder.cu (decoder layers); everything
// lives in an anonymous namespace, so each translation unit instantiates only the kernels it launches.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "philox.cuh"
#include "variants.h"

namespace rbm {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // one 128-byte swizzle atom of bf16 along K
constexpr int UMMA_K = 16;
#ifndef RBM_EPI_WARPS
#define RBM_EPI_WARPS 16
#endif
constexpr int kNumEpilogueWarps = RBM_EPI_WARPS;  // a multiple of 4: one set per TMEM lane quarter, columns split among sets
constexpr int kEpiSets = kNumEpilogueWarps / 4;
constexpr int kNumThreads = 128 + kNumEpilogueWarps * 32;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int kStoreWarpBytes = 32 * 32 * 2;    // one [32 chains x 32 units] bf16 box per epilogue warp (SWIZZLE_64B)

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    // (a suspend-time hint on try_wait changed nothing: profiles/r01_pairs_vs_single.txt)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// one lane of a converged warp (the warp stays convergent, so descriptors can live in uniform registers)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- cluster-pair (cta_group::2) flavours
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 64-bit shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version field = 1).
//   K-major : rows of 128 B (64 bf16 along K), 8-row groups SBO = 1024 B apart; LBO unused (1).
//   MN-major: 64 MN-elements (128 B) per k-row, 8-k groups SBO = 1024 B apart,
//             64-element MN blocks LBO bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// 32-bit instruction descriptor: D fp32, A/B bf16, M x N, A and B major selectable (0 = K-major).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool b_mn_major, bool a_mn_major = false) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7)                     // a_format = BF16
         | (1u << 10)                    // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)  // a_major
         | ((b_mn_major ? 1u : 0u) << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)    // n_dim
         | ((uint32_t)(M >> 4) << 24);   // m_dim
}

template <int BLOCK_N, int CTAS>
struct SmemLayout {
  // CTAS = 2: the two CTAs of a cluster pair each hold HALF of the B (W) tile; 32 KB stages
#ifndef RBM_PAIR_STAGES
#define RBM_PAIR_STAGES 5
#endif
  static constexpr int kStages = CTAS == 2 ? RBM_PAIR_STAGES : (BLOCK_N >= 256 ? 4 : BLOCK_N == 128 ? 6 : 8);
  static constexpr int kBBytes = (BLOCK_N / CTAS) * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStoreOffset = kStages * kStageBytes;             // epilogue -> TMA-store staging
  static constexpr int kStoreBytes = kNumEpilogueWarps * kStoreWarpBytes;
  static constexpr int kBarOffset = kStoreOffset + kStoreBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // + barriers/tmem ptr + alignment slack
};

constexpr uint32_t kKfMagic = 0x4B000000u;   // float 2^23: OR-ing a 16-bit field into its mantissa gives 2^23 + k exactly

struct HalfStepArgs {
  __nv_bfloat16* out;      // [rows, ldo] destination state (bf16 {0,1}); written through tmap_out, kept for reference
  const float* nbias;      // [>= n_tiles*BLOCK_N]  -bias * log2(e), zero padded
  int ldo;                 // row pitch of out (elements)
  int m_tiles, n_tiles;
  int k_blocks;            // K_pad / 64
  int n_valid;             // real (unpadded) number of output units
  DrawCtx ctx;
  uint64_t chain_offset;
  // annealed (AIS) epilogues only: arg = acc * scale + nbias * bmul
  float scale;             // -log2(e) * beta_k
  float bmul;              // beta_k on the hidden step (bias c is annealed), 1 on the visible step
  float ratio;             // beta_{k-1} / beta_k
  float* wpart;            // [rows_pad, wpart_ld] running log-weight partials (log2 units)
  int wpart_ld;
  int debug_mode;          // profiling only (env RBM_B200_DEBUG_MODE): 1 = skip epilogue math, 2 = skip TMA+MMA
  uint32_t kf_magic;       // kKfMagic, passed at run time (see EpiParams)
  // linear (decoder) epilogues only
  int relu;                // max(0, .) after the bias
  int lo_off;              // EPI_LINEAR_SPLIT: column offset of the low-part tile inside the output buffer
  float* out_f32;          // EPI_LINEAR_F32: [rows_valid, ld_out] fp32 destination
  int ld_out;
  const uint16_t* mask;    // EPI_LINEAR_F32: optional {0,1} bf16 gate [rows, ld_mask] multiplied into the output
  int ld_mask;
  int64_t rows_valid;      // EPI_LINEAR_F32: rows >= rows_valid are padding and are not stored
  int f32_tma;             // EPI_LINEAR_F32: tmap_out describes out_f32 ([32 rows x 16 fp32] boxes, SWIZZLE_64B): staged TMA stores
};

// Epilogue flavours
enum { EPI_GIBBS = 0, EPI_ANNEALED = 1, EPI_AIS_WEIGHT = 2,
       EPI_LINEAR_SPLIT = 3,   // y = act(acc + bias) stored as bf16 high part + bf16 low part (next layer's A operand)
       EPI_LINEAR_F32 = 4 };   // y = act(acc + bias) [* gate] stored as fp32

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// AIS log-weight of the hidden units, FOUR units at a time (the innermost loop of every epilogue runs over 4 consecutive
// units).  Per unit the term is log p*_k(v) - log p*_{k-1}(v) = softplus(b_k x) - softplus(b_{k-1} x), in log2 units
// (a' - a) + lg2(1 + 2^a) - lg2(1 + 2^a') with a = arg = -log2(e) b_k x and a' = a * b_{k-1} / b_k.  Taking the two
// logarithms of every unit made the weight half-steps MUFU-bound (4 MUFU per unit with the sampler's own ex2: 16 lanes per
// clock and SM); here the four (1 + 2^a) and the four (1 + 2^a') are multiplied up and the two logarithms are taken per
// QUAD: 2.5 MUFU per unit (ex2 of a', half a lg2).  a is clamped at 30, where softplus(-x) = lg2(1 + 2^-a) < 1.4e-9 has
// vanished in fp32 anyway, so a product of four factors stays below 2^121.
// Padding units need no guard: their W columns and biases are zero padded, a = a' = 0, both factors are exactly 2 and
// their logarithms cancel exactly.
// (-DRBM_AIS_PER_UNIT_LOG restores the two logarithms per unit for accuracy comparisons: benchmarks/ais_weight_accuracy.py)
struct AisQuad { float p, pm1, dsum; };
__device__ __forceinline__ void ais_quad_add(AisQuad& q, int i4, float arg, float s, float ratio) {
#ifdef RBM_AIS_PER_UNIT_LOG
  const float ak80 = fminf(arg, 80.0f), akm80 = ak80 * ratio;
  const float term = (akm80 - ak80) + lg2_approx(fminf(s, 1.2089258e24f)) - lg2_approx(1.0f + ex2_approx(akm80));
  q.p = q.pm1 = 1.0f;
  if (i4 == 0) q.dsum = term; else q.dsum += term;
  return;
#endif
  const float ak = fminf(arg, 30.0f);
  const float akm1 = ak * ratio;
  const float sk = fminf(s, 1073741824.0f);          // 1 + 2^a, a <= 30 (s = 1 + 2^arg is +inf for arg > 128)
  const float skm1 = 1.0f + ex2_approx(akm1);
  if (i4 == 0) { q.p = sk; q.pm1 = skm1; q.dsum = akm1 - ak; }
  else { q.p *= sk; q.pm1 *= skm1; q.dsum += akm1 - ak; }
}
// (lg2 of the QUOTIENT p * rcp.approx(pm1) would save nothing - rcp is a MUFU operation too - and rcp.approx is biased by
//  about +2^-24 relative: measured +7.5e-6 nats per beta step on every chain's log-weight, 0.075 nats over the 10,000 betas
//  of BASELINE configs[4] (benchmarks/ais_weight_accuracy.py, profiles/r02_ais_weight_accuracy.txt).  In the difference of
//  the two logarithms the approximation errors of lg2 cancel: the arguments are within a factor 1 + O(d beta) of each other.)
__device__ __forceinline__ float ais_quad_term(const AisQuad& q) { return q.dsum + (lg2_approx(q.p) - lg2_approx(q.pm1)); }

// Decision rule shared by the fast path and the exact path (philox.cuh contract, p = 1/s):
//   d = 65536 - k16 * s.   d >= s  <=>  (k16+1) * s <= 65536  <=>  p*65536 >= k16+1  -> 1
//                          d <  0  <=>   k16 * s    >  65536  <=>  p*65536 <  k16    -> 0
//   0 <= d < s: the 16-bit field alone cannot decide (probability 2^-16 per draw).
// For non-negative floats the bit patterns order like the values and negative floats have the
// sign bit set, so "0 <= d < s" is ONE unsigned integer compare of the bit patterns.
struct Decision { float d, s; };
// arg = -log2(e) * pre-activation, so 2^arg = e^-x and s = 1 + e^-x = 1/p
__device__ __forceinline__ Decision decision_terms(float arg, float kf) {
  Decision r;
  r.s = 1.0f + ex2_approx(arg);
  r.d = fmaf(-kf, r.s, 65536.0f);
  return r;
}

// Exact path for a unit of a span in which some unit was ambiguous (rare, kept out of line).
__device__ __noinline__ uint32_t resolve_exact(float arg, uint32_t k16, DrawCtx ctx, uint32_t blk, uint32_t chain,
                                               uint32_t slot) {
  const float kf = (float)k16;
  const Decision q = decision_terms(arg, kf);
  if (q.d >= q.s) return 1u;
  if (!(__float_as_uint(q.d) < __float_as_uint(q.s))) return 0u;
  const float y = __frcp_rn(q.s) * 65536.0f;   // p * 65536, exact scaling
  const float fl = floorf(y);
  if (fl > kf) return 1u;
  if (fl < kf) return 0u;
  const uint32_t e16 = field16(draw_block_extra(ctx, blk, chain), slot);
  return ((y - fl) * 65536.0f >= (float)e16) ? 1u : 0u;
}

// Per-launch epilogue parameters
struct EpiParams {
  const float* nbias;  // -log2(e) * bias, zero padded
  int ldo;             // padded unit count of the destination state
  int n_valid;         // real unit count
  DrawCtx ctx;
  float scale, bmul, ratio;   // annealed flavours (see HalfStepArgs)
  int debug_mode;
  // 0x4B000000 (float 2^23) handed over as a RUN-TIME value: as a literal, ptxas encodes it as PRMT's
  // immediate and then has to materialise the byte selector in a register before every PRMT
  uint32_t kf_magic;
  // linear flavours (see HalfStepArgs)
  int relu, lo_off;
  float* out_f32;
  int ld_out;
  const uint16_t* mask;
  int ld_mask;
  int64_t rows_valid;
  int f32_tma;
};

// Epilogue of ONE 128 x BLOCK_N tile for one warp: TMEM -> decisions -> swizzled smem box -> TMA store.
//   tmem_acc   TMEM address of the accumulator stage (column 0 of the tile, lane 0)
//   wait_full  blocks until the MMA has completed the accumulator;  release  hands the stage back
// Column sets that take part in a BLOCK_N-wide tile: every set needs at least one 32-unit Philox group, so
// tiles narrower than 32 * kEpiSets columns leave the surplus epilogue warps idle (they skip the tile).
template <int BLOCK_N>
struct EpiSplit {
  static constexpr int kSets = (BLOCK_N / 32 < kEpiSets) ? BLOCK_N / 32 : kEpiSets;
  static constexpr int kWarps = 4 * kSets;            // epilogue warps that work on (and arrive for) a tile
  static constexpr int kColsPerWarp = BLOCK_N / kSets;
};

// Per-group callbacks of epilogue_tile (the chain kernel publishes finished column chunks through them)
struct NoGroupHooks {
  __device__ __forceinline__ void mid(int) {}
  __device__ __forceinline__ void before_store(int) {}
  __device__ __forceinline__ void stored(int) {}
};

//   wrow       EPI_AIS_WEIGHT: this chain's log-weight slots, one per 32-unit group of hidden units (log2 units);
//              the group's sum is ADDED with one red per thread, so the result does not depend on the tiling
//   hooks      mid(g) is called half-way through group g, stored(g) after group g's box has been handed to TMA
// Column mapping: the tile's 32-unit groups are dealt round-robin to the column sets, i.e. in round g the kSets
// warps of a lane quarter cover the CONTIGUOUS chunk of columns [g * 32 kSets, (g + 1) * 32 kSets): the tile's
// columns are finished chunk by chunk, not all at the very end.
template <int BLOCK_N, int EPI, class WaitFn, class ReleaseFn, class Hooks>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, const CUtensorMap* tmap_out, uint32_t tmem_acc,
                                              int n_blk, int row_base, uint32_t chain, int quarter, int cq, int lane,
                                              uint32_t stage_warp, float* wrow, WaitFn wait_full, ReleaseFn release,
                                              Hooks& hooks) {
  constexpr int kSets = EpiSplit<BLOCK_N>::kSets;
  constexpr int kColsPerWarp = EpiSplit<BLOCK_N>::kColsPerWarp;
  constexpr int kGroups = kColsPerWarp / 32;   // 32-unit Philox groups per warp per tile
  float wsum = 0.f;
  // staging box of this warp: row `lane` = 64 B, 16-B chunk c stored at c ^ ((lane >> 1) & 3)  (SWIZZLE_64B)
  const uint32_t stage_row = stage_warp + (uint32_t)lane * 64u;
  const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
    // Philox words of one 32-unit group (slot mapping: philox.cuh).  They do not depend on the accumulator,
    // so the first group's words are drawn BEFORE waiting for the MMA and the next group's words right
    // after the current group: the draw latency overlaps the wait / the TMA store.
    uint32_t w[4][4];
    auto draw_group = [&](int g) {
      const uint32_t b0 = (uint32_t)((n_blk * BLOCK_N + (g * kSets + cq) * 32) >> 5) * 4u;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint4 x = draw_block(ep.ctx, b0 + b, chain);
        w[b][0] = x.x; w[b][1] = x.y; w[b][2] = x.z; w[b][3] = x.w;
      }
    };
    draw_group(0);
    // all lanes poll: measured 8 % faster than one polling lane + __syncwarp (profiles/r01_pairs_vs_single.txt)
    wait_full();
    tcgen05_fence_after();
#pragma unroll 1
    for (int g = 0; g < kGroups; ++g) {
      const int col_in_tile = (g * kSets + cq) * 32;
      const int col0 = n_blk * BLOCK_N + col_in_tile;
      const bool live = col0 < ep.ldo;   // tile overhang beyond the padded unit count
      const uint32_t blk0 = (uint32_t)(col0 >> 5) * 4u;
      if (live) {
        // the previous TMA store of this warp must have finished READING the staging box
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (hh == 1) hooks.mid(g);
        uint32_t r[16];
        tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col_in_tile + 16 * hh), r);
        tmem_ld_wait();
        if (g == kGroups - 1 && hh == 1) {
          // all of this warp's TMEM reads for the tile are done: hand the stage back
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) release();
        }
        if (!live || ep.debug_mode == 1) continue;
        uint32_t pk[8];       // 16 units -> 16 bf16 {0,1}, packed in pairs
        bool amb = false;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 nb4 = __ldg(reinterpret_cast<const float4*>(ep.nbias + col0 + 16 * hh) + q4);
          const float nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
          AisQuad quad;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int c16 = 4 * q4 + i4;           // column within this 16-wide half
            const int cc = 16 * hh + c16;          // column within the 32-group
            const int b = (cc & 7) >> 1;           // Philox block
            const int slot = 2 * (cc >> 3) + (cc & 1);
            const uint32_t word = w[b][slot >> 1];
            // float(8388608 + k16) with one byte-permute, minus 2^23 -> exact float(k16)
            const uint32_t mbits = (slot & 1) ? __byte_perm(word, ep.kf_magic, 0x7632) : __byte_perm(word, ep.kf_magic, 0x7610);
            const float kf = __uint_as_float(mbits) - 8388608.0f;
            const float arg = EPI == EPI_GIBBS ? fmaf(__uint_as_float(r[c16]), -1.4426950408889634f, nbv[i4])
                                               : fmaf(__uint_as_float(r[c16]), ep.scale, nbv[i4] * ep.bmul);
            const Decision q = decision_terms(arg, kf);
            if (EPI == EPI_AIS_WEIGHT) {
              // log p*_k(v) - log p*_{k-1}(v), hidden-unit term (ais_quad_add above), summed over the 4 units of q4
              ais_quad_add(quad, i4, arg, q.s, ep.ratio);
              if (i4 == 3) wsum += ais_quad_term(quad);
            }
            // fp32 1.0 / 0.0 straight from the compare (FSET); the top halves of two of them are the bf16 pair
            const uint32_t onef = __float_as_uint(q.d >= q.s ? 1.0f : 0.0f);
            amb = amb || (__float_as_uint(q.d) < __float_as_uint(q.s));
            if (c16 & 1) pk[c16 >> 1] = __byte_perm(pk[c16 >> 1], onef, 0x7632); else pk[c16 >> 1] = onef;
          }
        }
        if (amb) {
          // some unit of this span needs its refinement bits: redo the span on the exact path
#pragma unroll
          for (int c16 = 0; c16 < 16; ++c16) {
            const int cc = 16 * hh + c16;
            const int b = (cc & 7) >> 1, slot = 2 * (cc >> 3) + (cc & 1);
            const uint32_t word = w[b][slot >> 1];
            const uint32_t k16 = (slot & 1) ? (word >> 16) : (word & 0xFFFFu);
            const float nbx = __ldg(ep.nbias + col0 + cc);
            const float arg = EPI == EPI_GIBBS ? fmaf(__uint_as_float(r[c16]), -1.4426950408889634f, nbx)
                                               : fmaf(__uint_as_float(r[c16]), ep.scale, nbx * ep.bmul);
            const uint32_t one = resolve_exact(arg, k16, ep.ctx, blk0 + b, chain, (uint32_t)slot);
            const uint32_t val = one ? ((c16 & 1) ? 0x3F800000u : 0x3F80u) : 0u;
            if (c16 & 1) pk[c16 >> 1] |= val; else pk[c16 >> 1] = val;
          }
        }
        // 16 units = 32 B = two 16-B chunks of this chain's row in the staging box
        st_shared_v4(stage_row + (((uint32_t)(2 * hh) ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
        st_shared_v4(stage_row + (((uint32_t)(2 * hh + 1) ^ sw) << 4), pk[4], pk[5], pk[6], pk[7]);
      }
      if (live && ep.debug_mode != 1) {
        // generic-proxy writes -> visible to the async proxy, then ONE coalesced TMA store of the
        // [32 chains x 32 units] box; columns beyond the padded unit count are clipped by TMA
        hooks.before_store(g);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmap_out, stage_warp, col0, row_base);
          tma_store_commit();
        }
      }
      if (EPI == EPI_AIS_WEIGHT && live) {
        atomicAdd(wrow + (col0 >> 5), wsum);   // one owner per (chain, group) and half-step: the order of adds is fixed
        wsum = 0.f;
      }
      hooks.stored(g);
      if (g + 1 < kGroups) draw_group(g + 1);
    }
}


// Linear-layer epilogues of ONE 128 x BLOCK_N tile for one warp (decoder path):
//   EPI_LINEAR_SPLIT  y = act(acc + bias) -> bf16 high part and bf16 low part (y - high), two TMA-stored boxes;
//                     [high | high | low] x [W_high | W_low | W_high] in the next layer's GEMM keeps fp32-grade accuracy
//   EPI_LINEAR_F32    y = act(acc + bias) [* gate] -> fp32, stored straight from registers (64 B per lane and span)
template <int BLOCK_N, int EPI, class WaitFn, class ReleaseFn>
__device__ __forceinline__ void epilogue_tile_linear(const EpiParams& ep, const CUtensorMap* tmap_out, uint32_t tmem_acc,
                                                     int n_blk, int row_base, int quarter, int cq, int lane,
                                                     uint32_t stage_warp, WaitFn wait_full, ReleaseFn release) {
  constexpr int kColsPerWarp = BLOCK_N / kEpiSets;
  constexpr int kGroups = kColsPerWarp / 32;
  const uint32_t stage_row = stage_warp + (uint32_t)lane * 64u;
  const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
  const int64_t row = (int64_t)row_base + lane;
  wait_full();
  tcgen05_fence_after();
#pragma unroll 1
  for (int g = 0; g < kGroups; ++g) {
    const int col_in_tile = cq * kColsPerWarp + g * 32;
    const int col0 = n_blk * BLOCK_N + col_in_tile;
    const bool live = col0 < ep.ldo;
    uint32_t hi[16], lo[16];   // EPI_LINEAR_SPLIT: the 32 units of this group as bf16 pairs
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t r[16];
      tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col_in_tile + 16 * hh), r);
      tmem_ld_wait();
      if (g == kGroups - 1 && hh == 1) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) release();
      }
      if (!live) continue;
      float y[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.nbias + col0 + 16 * hh) + q4);
        y[4 * q4 + 0] = __uint_as_float(r[4 * q4 + 0]) + b4.x;
        y[4 * q4 + 1] = __uint_as_float(r[4 * q4 + 1]) + b4.y;
        y[4 * q4 + 2] = __uint_as_float(r[4 * q4 + 2]) + b4.z;
        y[4 * q4 + 3] = __uint_as_float(r[4 * q4 + 3]) + b4.w;
      }
      if (ep.relu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
      }
      if (EPI == EPI_LINEAR_SPLIT) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(y[j]), h1 = __float2bfloat16_rn(y[j + 1]);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(y[j] - __bfloat162float(h0));
          const __nv_bfloat16 l1 = __float2bfloat16_rn(y[j + 1] - __bfloat162float(h1));
          hi[8 * hh + (j >> 1)] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lo[8 * hh + (j >> 1)] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
      } else {
        const int cbase = col0 + 16 * hh;
        const bool in_range = row < ep.rows_valid && cbase < ep.n_valid;
        if (in_range && ep.mask != nullptr) {
          const uint4* mp = reinterpret_cast<const uint4*>(ep.mask + row * ep.ld_mask + cbase);
          const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
          const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            y[j] *= __uint_as_float(mw[j >> 1] << 16);
            y[j + 1] *= __uint_as_float(mw[j >> 1] & 0xFFFF0000u);
          }
        }
        if (ep.f32_tma) {
          // [32 rows x 16 fp32] box: a row is 64 B, chunk c at c ^ ((lane >> 1) & 3) (SWIZZLE_64B), one TMA store;
          // rows and columns beyond the real extent are clipped by the tensor map
          if (cbase < ep.n_valid) {   // warp-uniform
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              st_shared_v4(stage_row + (((uint32_t)ch ^ sw) << 4), __float_as_uint(y[4 * ch]), __float_as_uint(y[4 * ch + 1]),
                           __float_as_uint(y[4 * ch + 2]), __float_as_uint(y[4 * ch + 3]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(tmap_out, stage_warp, cbase, row_base);
              tma_store_commit();
            }
          }
        } else if (in_range) {
          float* dst = ep.out_f32 + row * ep.ld_out + cbase;
          if ((ep.ld_out & 3) == 0 && cbase + 16 <= ep.n_valid && (reinterpret_cast<uintptr_t>(ep.out_f32) & 15) == 0) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              reinterpret_cast<float4*>(dst)[q4] = make_float4(y[4 * q4], y[4 * q4 + 1], y[4 * q4 + 2], y[4 * q4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cbase + j < ep.n_valid) dst[j] = y[j];
          }
        }
      }
    }
    if (EPI == EPI_LINEAR_SPLIT && live) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        // the previous TMA store of this warp must have finished READING the staging box
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        const uint32_t* pk = part == 0 ? hi : lo;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          st_shared_v4(stage_row + (((uint32_t)ch ^ sw) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmap_out, stage_warp, (part == 0 ? 0 : ep.lo_off) + col0, row_base);
          tma_store_commit();
        }
      }
    }
  }
}

// SPLIT3: the A operand is [high | low] column blocks read as [high | high | low] (decoder path)
template <int BLOCK_N, bool B_MN_MAJOR, int EPI, int CTAS, bool SPLIT3 = false>
__global__ void __launch_bounds__(kNumThreads, 1)
umma_half_step_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_out, const HalfStepArgs args) {
  using L = SmemLayout<BLOCK_N, CTAS>;
  constexpr int kStages = L::kStages;
  constexpr int kBCols = BLOCK_N / CTAS;        // columns of the B tile this CTA stages
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (same offset in both CTAs of a pair)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::kBarOffset;
  // barrier slots (8 bytes each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOffset + 8 * (2 * kStages + 4));

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTAS = 2: CTA pair (cluster of 2) works on a 256-chain x BLOCK_N tile; rank 0 is the leader that
  // issues the MMAs; the full/tmem_empty barriers that gate the MMAs live in the leader.
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int num_tiles = (args.m_tiles / CTAS) * args.n_tiles;     // m_tiles counts 128-chain tiles
  const int first_tile = (int)blockIdx.x / CTAS, tile_stride = (int)gridDim.x / CTAS;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages (power of two: 256 or 512)

  if (warp_idx == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), CTAS); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), kNumEpilogueWarps * CTAS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp_idx == 2) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // =============================== TMA producer ===============================
    // (whole warp, warp-uniform; one elected lane issues the copies - same reasoning as the MMA issuer)
    if (args.debug_mode != 2) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const int n_blk = tile % args.n_tiles, m_blk = tile / args.n_tiles;
        const int row0 = (m_blk * CTAS + (int)cta_rank) * BLOCK_M;          // this CTA's 128 chains
        const int ncol0 = n_blk * BLOCK_N + (int)cta_rank * kBCols;         // this CTA's share of the W tile
        for (int kb = 0; kb < args.k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * L::kStageBytes;
          const uint32_t sb = sa + kABytes;
          // SPLIT3: K blocks [0,n) and [n,2n) read the high part, [2n,3n) the low part stored behind it
          const int ka = (SPLIT3 && kb >= args.k_blocks / 3) ? kb - args.k_blocks / 3 : kb;
          if (elect_one_sync()) {
            if (CTAS == 1) {
              mbar_arrive_expect_tx(full_bar(stage), L::kStageBytes);
              tma_load_2d(sa, &tmap_a, full_bar(stage), ka * BLOCK_K, row0);
              if (B_MN_MAJOR) {
                // W[k rows][n cols]: one [64 x 64] box per 64-column MN block, 8 KB apart (LBO)
#pragma unroll
                for (int nb = 0; nb < kBCols / 64; ++nb)
                  tma_load_2d(sb + nb * (BLOCK_K * 128), &tmap_b, full_bar(stage), ncol0 + nb * 64, kb * BLOCK_K);
              } else {
                // W[n rows][k cols]: one [kBCols x 64] box, rows of 128 B
                tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BLOCK_K, ncol0);
              }
            } else {
              // both CTAs credit the LEADER's full barrier (2 arrivals + the bytes of both CTAs)
              const uint32_t lead_full = mapa_u32(full_bar(stage), 0);
              if (is_leader) mbar_arrive_expect_tx(full_bar(stage), 2 * L::kStageBytes);
              else mbar_arrive_cluster(lead_full);
              tma_load_2d_pair(sa, &tmap_a, lead_full, ka * BLOCK_K, row0);
              if (B_MN_MAJOR) {
#pragma unroll
                for (int nb = 0; nb < kBCols / 64; ++nb)
                  tma_load_2d_pair(sb + nb * (BLOCK_K * 128), &tmap_b, lead_full, ncol0 + nb * 64, kb * BLOCK_K);
              } else {
                tma_load_2d_pair(sb, &tmap_b, lead_full, kb * BLOCK_K, ncol0);
              }
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ================================
    // The WHOLE warp runs the loop (waits, descriptor arithmetic) so everything stays warp-uniform and in
    // uniform registers; one elected lane issues the tcgen05 instructions.  (With a single-lane loop the
    // compiler emitted ~110 dependent instructions per K block - vector->uniform register moves in ELECT
    // loops - and the tensor pipe idled a third of the time waiting for its issuer: profiles/r01_mma_issue.txt.)
    if (is_leader) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M * CTAS, BLOCK_N, B_MN_MAJOR);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
        const int as = it & 1;
        mbar_wait(tmem_empty_bar(as), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < (args.debug_mode == 2 ? 0 : args.k_blocks); ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * L::kStageBytes;
          const uint32_t sb = sa + kABytes;
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
              const uint64_t bdesc = B_MN_MAJOR ? make_smem_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                                : make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
              if (CTAS == 2) umma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) once these MMAs retire
            if (CTAS == 2) umma_commit_pair(empty_bar(stage)); else umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (elect_one_sync()) {
          if (CTAS == 2) umma_commit_pair(tmem_full_bar(as)); else umma_commit(tmem_full_bar(as));
        }
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4) {
    // ================================= epilogue =================================
    const int quarter = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int cq = (warp_idx - 4) >> 2;          // which share of the tile's columns
    const uint32_t stage_warp = smem_base + L::kStoreOffset + (uint32_t)(warp_idx - 4) * kStoreWarpBytes;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_stride, ++it) {
      const int n_blk = tile % args.n_tiles, m_blk = tile / args.n_tiles;
      const int as = it & 1;
      const int row_base = (m_blk * CTAS + (int)cta_rank) * BLOCK_M + quarter * 32;
      const int row = row_base + lane;
      const uint32_t chain = (uint32_t)(args.chain_offset + (uint64_t)row);
      EpiParams ep;
      ep.nbias = args.nbias; ep.ldo = args.ldo; ep.n_valid = args.n_valid; ep.ctx = args.ctx;
      ep.scale = args.scale; ep.bmul = args.bmul; ep.ratio = args.ratio; ep.debug_mode = args.debug_mode;
      ep.kf_magic = args.kf_magic;
      ep.relu = args.relu; ep.lo_off = args.lo_off; ep.out_f32 = args.out_f32; ep.ld_out = args.ld_out;
      ep.mask = args.mask; ep.ld_mask = args.ld_mask; ep.rows_valid = args.rows_valid; ep.f32_tma = args.f32_tma;
      auto wait_full = [&] { mbar_wait(tmem_full_bar(as), ((uint32_t)it >> 1) & 1u); };
      auto release = [&] {
        if (CTAS == 2) mbar_arrive_cluster(mapa_u32(tmem_empty_bar(as), 0));   // the leader's barrier
        else mbar_arrive(tmem_empty_bar(as));
      };
      NoGroupHooks no_hooks;
      if constexpr (EPI == EPI_LINEAR_SPLIT || EPI == EPI_LINEAR_F32)
        epilogue_tile_linear<BLOCK_N, EPI>(ep, &tmap_out, tmem_base + (uint32_t)(as * BLOCK_N), n_blk, row_base, quarter,
                                           cq, lane, stage_warp, wait_full, release);
      else
        epilogue_tile<BLOCK_N, EPI>(ep, &tmap_out, tmem_base + (uint32_t)(as * BLOCK_N), n_blk, row_base, chain, quarter,
                                    cq, lane, stage_warp, EPI == EPI_AIS_WEIGHT ? args.wpart + (size_t)row * args.wpart_ld : nullptr,
                                    wait_full, release, no_hooks);
    }
    if (lane == 0) tma_store_wait_all();   // state tiles must be complete before the kernel ends
  }

  __syncwarp();   // re-converge the single-lane roles before the block / cluster barrier
  tcgen05_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp_idx == 2) {
    tcgen05_fence_after();
    if (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    RBM_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    RBM_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, RBM_B200_ERR_NO_DEVICE,
                "cuTensorMapEncodeTiled is not available from the driver");
    fn = (EncodeTiledFn)p;
  }
  *out = fn;
  return 0;
}

// bf16 row-major [rows, cols] with row pitch `ld` elements; box = [box_rows, 64 cols], 128B swizzle
int make_tmap(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
              uint32_t box_cols = 64, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = nullptr;
  int rc = get_encode_fn(&fn);
  if (rc) return rc;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RBM_REQUIRE(r == CUDA_SUCCESS, RBM_B200_ERR_INVALID, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// fp32 row-major [rows, cols] with row pitch `ld` elements (ld * 4 bytes must be a multiple of 16)
int make_tmap_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                  uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = nullptr;
  int rc = get_encode_fn(&fn);
  if (rc) return rc;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RBM_REQUIRE(r == CUDA_SUCCESS, RBM_B200_ERR_INVALID, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return 0;
}

// CTA pairs (tcgen05 cta_group::2) are the default for 256-wide tiles; RBM_B200_PAIRS=0 forces single CTAs
bool pairs_enabled() {
  static const bool on = [] { const char* e = getenv("RBM_B200_PAIRS"); return e ? atoi(e) != 0 : true; }();
  return on;
}

int pick_block_n(int n_pad, int64_t rows_pad) {
  // fewest padded columns wins; ties go to the wider tile ...
  const int w256 = round_up(n_pad, 256), w128 = round_up(n_pad, 128);
  if (w128 < w256) return 128;
  // ... unless 256-wide tiles would leave most SMs idle (few chains): narrower tiles double the CTAs
  // and halve the per-launch latency (AIS shards of 2048 chains: profiles/r01_ais_config5_scale.jsonl)
  static const int force = [] { const char* e = getenv("RBM_B200_BLOCK_N"); return e ? atoi(e) : 0; }();
  if (force == 128 || force == 256) return force;
  // (measured, benchmarks/multistep_ab.py: 512x512 -> 128-wide tiles, 7.4 vs 10.7 us per half-step;
  //  1024x1024 -> 256-wide tiles keep the cluster at 4 CTAs, 13.5 vs 19.9 us)
  const int64_t tiles256 = (rows_pad / BLOCK_M) * (w256 / 256);
  return (tiles256 < 74 && n_pad <= 512) ? 128 : 256;
}

template <int BLOCK_N, bool B_MN_MAJOR, int EPI, int CTAS, bool SPLIT3 = false>
int launch_half_step_impl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const HalfStepArgs& a,
                          int sm_count, cudaStream_t st) {
  auto kern = umma_half_step_kernel<BLOCK_N, B_MN_MAJOR, EPI, CTAS, SPLIT3>;
  using L = SmemLayout<BLOCK_N, CTAS>;
  {
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    if ((rc = ensure_dynamic_smem(kern, di.device, L::kTotal))) return rc;   // per device, not per process
  }
  static const int debug_mode = [] { const char* e = getenv("RBM_B200_DEBUG_MODE"); return e ? atoi(e) : 0; }();
  HalfStepArgs a2 = a;
  a2.debug_mode = debug_mode;
  a2.kf_magic = kKfMagic;
  const int tiles = (a.m_tiles / CTAS) * a.n_tiles;
  const int grid = CTAS * std::min(tiles, sm_count / CTAS);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RBM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, a2));
  return check_launch("umma_half_step_kernel");
}

// use_pairs: CTA pairs (tcgen05 cta_group::2) when the tile is 256 wide.  Measured on B200 under the
// 1 kW power cap (profiles/r01_pairs_vs_single.txt): ~4% faster than single CTAs on long persistent
// loops (half the W traffic into shared memory), slower when a CTA only sees a few tiles.
template <int BLOCK_N, bool B_MN_MAJOR, int EPI = EPI_GIBBS>
int launch_half_step(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const HalfStepArgs& a,
                     int sm_count, bool use_pairs, cudaStream_t st) {
  if (BLOCK_N == 256 && use_pairs)
    return launch_half_step_impl<256, B_MN_MAJOR, EPI, 2>(ta, tb, to, a, sm_count, st);
  return launch_half_step_impl<BLOCK_N, B_MN_MAJOR, EPI, 1>(ta, tb, to, a, sm_count, st);
}

}  // namespace
}  // namespace rbm
