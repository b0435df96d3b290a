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
// Kernel: persistent, warp-specialised, 1 CTA per SM, 128 x BLOCK_N output tile.
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16)
//   warp 2      TMEM allocator (2 accumulator stages x BLOCK_N fp32 columns)
//   warps 4-19  epilogue       (tcgen05.ld 32x32b.x16 -> ex2 + Philox + compare -> bf16 store)
// The MMA of tile i+1 overlaps the epilogue of tile i through the two TMEM stages.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

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
#ifdef RBM_TRYWAIT_HINT
    // suspend-time hint: the thread may sleep in hardware up to this many ns before try_wait returns false
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"((uint32_t)RBM_TRYWAIT_HINT)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
#endif
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
  static constexpr int kStages = CTAS == 2 ? RBM_PAIR_STAGES : 4;
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
};

// Epilogue of ONE 128 x BLOCK_N tile for one warp: TMEM -> decisions -> swizzled smem box -> TMA store.
//   tmem_acc   TMEM address of the accumulator stage (column 0 of the tile, lane 0)
//   wait_full  blocks until the MMA has completed the accumulator;  release  hands the stage back
template <int BLOCK_N, int EPI, class WaitFn, class ReleaseFn>
__device__ __forceinline__ void epilogue_tile(const EpiParams& ep, const CUtensorMap* tmap_out, uint32_t tmem_acc,
                                              int n_blk, int row_base, uint32_t chain, int quarter, int cq, int lane,
                                              uint32_t stage_warp, float& wsum, WaitFn wait_full, ReleaseFn release) {
  constexpr int kColsPerWarp = BLOCK_N / kEpiSets;
  constexpr int kGroups = kColsPerWarp / 32;   // 32-unit Philox groups per warp per tile
  // staging box of this warp: row `lane` = 64 B, 16-B chunk c stored at c ^ ((lane >> 1) & 3)  (SWIZZLE_64B)
  const uint32_t stage_row = stage_warp + (uint32_t)lane * 64u;
  const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
    // Philox words of one 32-unit group (slot mapping: philox.cuh).  They do not depend on the accumulator,
    // so the first group's words are drawn BEFORE waiting for the MMA and the next group's words right
    // after the current group: the draw latency overlaps the wait / the TMA store.
    uint32_t w[4][4];
    auto draw_group = [&](int g) {
      const uint32_t b0 = (uint32_t)((n_blk * BLOCK_N + cq * kColsPerWarp + g * 32) >> 5) * 4u;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint4 x = draw_block(ep.ctx, b0 + b, chain);
        w[b][0] = x.x; w[b][1] = x.y; w[b][2] = x.z; w[b][3] = x.w;
      }
    };
#ifndef RBM_LATE_PHILOX
    draw_group(0);
#endif
    // all lanes poll: measured 8 % faster than one polling lane + __syncwarp (profiles/r01_pairs_vs_single.txt)
    wait_full();
#ifdef RBM_LATE_PHILOX
    draw_group(0);
#endif
    tcgen05_fence_after();
#pragma unroll 1
    for (int g = 0; g < kGroups; ++g) {
      const int col_in_tile = cq * kColsPerWarp + g * 32;
      const int col0 = n_blk * BLOCK_N + col_in_tile;
      const bool live = col0 < ep.ldo;   // tile overhang beyond the padded unit count
      const uint32_t blk0 = (uint32_t)(col0 >> 5) * 4u;
#ifndef RBM_DEBUG_NO_WAIT_READ
      if (live) {
        // the previous TMA store of this warp must have finished READING the staging box
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#endif
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
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
              // log p*_k(v) - log p*_{k-1}(v), hidden-unit term: softplus(b_k x) - softplus(b_{k-1} x),
              // in log2 units: softplus(bx)/ln2 = -arg + lg2(1 + 2^arg); clamped so 2^arg stays finite
              const float ak = fminf(arg, 80.0f);
              const float akm1 = ak * ep.ratio;
              if (col0 + cc < ep.n_valid)
                wsum += (akm1 - ak) + lg2_approx(fminf(q.s, 1.2089258e24f)) - lg2_approx(1.0f + ex2_approx(akm1));
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
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(tmap_out, stage_warp, col0, row_base);
          tma_store_commit();
        }
      }
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
        if (row < ep.rows_valid && cbase < ep.n_valid) {
          if (ep.mask != nullptr) {
            const uint4* mp = reinterpret_cast<const uint4*>(ep.mask + row * ep.ld_mask + cbase);
            const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              y[j] *= __uint_as_float(mw[j >> 1] << 16);
              y[j + 1] *= __uint_as_float(mw[j >> 1] & 0xFFFF0000u);
            }
          }
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
      float wsum = 0.f;   // EPI_AIS_WEIGHT: this thread's log-weight increment over its columns (log2 units)
      EpiParams ep;
      ep.nbias = args.nbias; ep.ldo = args.ldo; ep.n_valid = args.n_valid; ep.ctx = args.ctx;
      ep.scale = args.scale; ep.bmul = args.bmul; ep.ratio = args.ratio; ep.debug_mode = args.debug_mode;
      ep.kf_magic = args.kf_magic;
      ep.relu = args.relu; ep.lo_off = args.lo_off; ep.out_f32 = args.out_f32; ep.ld_out = args.ld_out;
      ep.mask = args.mask; ep.ld_mask = args.ld_mask; ep.rows_valid = args.rows_valid;
      auto wait_full = [&] { mbar_wait(tmem_full_bar(as), ((uint32_t)it >> 1) & 1u); };
      auto release = [&] {
        if (CTAS == 2) mbar_arrive_cluster(mapa_u32(tmem_empty_bar(as), 0));   // the leader's barrier
        else mbar_arrive(tmem_empty_bar(as));
      };
      if constexpr (EPI == EPI_LINEAR_SPLIT || EPI == EPI_LINEAR_F32)
        epilogue_tile_linear<BLOCK_N, EPI>(ep, &tmap_out, tmem_base + (uint32_t)(as * BLOCK_N), n_blk, row_base, quarter,
                                           cq, lane, stage_warp, wait_full, release);
      else
        epilogue_tile<BLOCK_N, EPI>(ep, &tmap_out, tmem_base + (uint32_t)(as * BLOCK_N), n_blk, row_base, chain, quarter,
                                    cq, lane, stage_warp, wsum, wait_full, release);
      if (EPI == EPI_AIS_WEIGHT) {
        // exactly one thread owns (row, slot) in a launch: plain read-modify-write, deterministic
        float* slot_ptr = args.wpart + (size_t)row * args.wpart_ld + (n_blk * kEpiSets + cq);
        *slot_ptr += wsum;
      }
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

// ---------------------------------------------------------------------------------------------------
// Multi-step variant: ALL half-steps of a chain block in ONE launch, for jobs whose per-half-step grid
// would not fill the GPU (few chains: PCD batches, AIS shards).  A cluster of n_tiles CTAs owns one
// 128-chain row block for the whole run; CTA `n_blk` computes output tile n_blk of every half-step.
// Between half-steps only that row block's CTAs have to agree (rows are independent), so the only
// synchronisation is ONE cluster barrier per half-step after the state tile has been stored - no
// relaunch, no grid-wide sync, and the draw context / annealing schedule advance in registers.
struct MultiArgs {
  const float* nbias[2];   // [dir]: dir 0 = v->h (hidden biases), dir 1 = h->v
  int ldo[2];              // padded unit count of the state WRITTEN by dir
  int n_valid[2];
  int n_tiles[2];
  int k_blocks[2];
  int steps;               // number of half-steps in this launch
  int first_dir;
  uint64_t seed, chain_offset, step_offset;   // half-step t draws with counter step_offset + t
  uint32_t tag;
  const float* betas;      // annealed flavours: beta index of half-step t is 1 + t/2
  float* wpart;            // AIS: [rows, wpart_ld] log-weight partials (log2 units), added to at the end
  int wpart_ld;
  uint32_t kf_magic;       // kKfMagic (see EpiParams)
};

template <int BLOCK_N, int EPI0, int EPI1>
__global__ void __launch_bounds__(kNumThreads, 1)
umma_multistep_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                      const __grid_constant__ CUtensorMap tm_b0, const __grid_constant__ CUtensorMap tm_b1,
                      const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1,
                      const MultiArgs args) {
  using L = SmemLayout<BLOCK_N, 1>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOffset + 8 * (2 * kStages + 4));

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blk = (int)cluster_ctarank();                     // output tile of this CTA
  uint32_t cs;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
  const int m_blk = (int)(blockIdx.x / cs);                      // the cluster's 128-chain row block
  constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;    // one accumulator (steps are serialised)

  if (warp_idx == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o1) : "memory");
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp_idx == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int quarter = warp_idx & 3;
  const int cq = (warp_idx - 4) >> 2;
  const uint32_t stage_warp = smem_base + L::kStoreOffset + (uint32_t)(warp_idx >= 4 ? warp_idx - 4 : 0) * kStoreWarpBytes;
  const int row_base = m_blk * BLOCK_M + quarter * 32;
  const uint32_t chain = (uint32_t)(args.chain_offset + (uint64_t)(row_base + lane));
  int stage = 0; uint32_t phase = 0;     // smem ring position (producer lane and MMA lane each keep their own copy)
  int it = 0;                            // number of tiles this CTA has processed (accumulator barrier parity)
  float wsum = 0.f;

  for (int t = 0; t < args.steps; ++t) {
    const int dir = (args.first_dir + t) & 1;
    const bool active = n_blk < args.n_tiles[dir];
    if (active) {
      const CUtensorMap* tm_a = dir ? &tm_a1 : &tm_a0;
      const CUtensorMap* tm_b = dir ? &tm_b1 : &tm_b0;
      if (warp_idx == 0) {
        {
          // ---- TMA producer: this row block's state (A) and tile n_blk of W (B: MN-major for dir 0, K-major for dir 1)
          for (int kb = 0; kb < args.k_blocks[dir]; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * L::kStageBytes;
            const uint32_t sb = sa + kABytes;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(full_bar(stage), L::kStageBytes);
              tma_load_2d(sa, tm_a, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
              if (dir == 0) {
#pragma unroll
                for (int nb = 0; nb < BLOCK_N / 64; ++nb)
                  tma_load_2d(sb + nb * (BLOCK_K * 128), tm_b, full_bar(stage), n_blk * BLOCK_N + nb * 64, kb * BLOCK_K);
              } else {
                tma_load_2d(sb, tm_b, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      } else if (warp_idx == 1) {
        {
          // ---- MMA issuer (the accumulator is free: the previous half-step ended with a cluster barrier).
          // Whole warp, warp-uniform; one elected lane issues (see umma_half_step_kernel).
          const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, dir == 0);
          for (int kb = 0; kb < args.k_blocks[dir]; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_base + stage * L::kStageBytes;
            const uint32_t sb = sa + kABytes;
            if (elect_one_sync()) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                const uint64_t adesc = make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
                const uint64_t bdesc = dir == 0 ? make_smem_desc(sb + k * (UMMA_K * 128), BLOCK_K * 128, 1024)
                                                : make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
                umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              }
              umma_commit(empty_bar(stage));
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          if (elect_one_sync()) umma_commit(tmem_full_bar);
          __syncwarp();
        }
      } else if (warp_idx >= 4) {
        // ---- epilogue
        EpiParams ep;
        ep.nbias = args.nbias[dir]; ep.ldo = args.ldo[dir]; ep.n_valid = args.n_valid[dir];
        ep.ctx = make_draw_ctx(args.seed, args.step_offset + (uint64_t)t, args.tag);
        ep.scale = -1.4426950408889634f; ep.bmul = 1.f; ep.ratio = 0.f; ep.debug_mode = 0;
        ep.kf_magic = args.kf_magic;
        if (EPI0 != EPI_GIBBS) {
          const float bk = __ldg(args.betas + 1 + (t >> 1)), bkm1 = __ldg(args.betas + (t >> 1));
          ep.scale = -1.4426950408889634f * bk;
          ep.bmul = dir == 0 ? bk : 1.f;
          ep.ratio = bkm1 / bk;
        }
        auto wait_full = [&] { mbar_wait(tmem_full_bar, (uint32_t)it & 1u); };
        auto release = [] {};
        const CUtensorMap* tm_o = dir ? &tm_o1 : &tm_o0;
        if (dir == 0) epilogue_tile<BLOCK_N, EPI0>(ep, tm_o, tmem_base, n_blk, row_base, chain, quarter, cq, lane, stage_warp, wsum, wait_full, release);
        else epilogue_tile<BLOCK_N, EPI1>(ep, tm_o, tmem_base, n_blk, row_base, chain, quarter, cq, lane, stage_warp, wsum, wait_full, release);
        if (lane == 0) tma_store_wait_all();       // this CTA's part of the new state is in global memory
      }
      ++it;
    }
    // ---- end of the half-step: every CTA of the row block has stored its tile before anyone reads the new state
    __syncwarp();
    asm volatile("fence.proxy.async;" ::: "memory");
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  if (EPI0 == EPI_AIS_WEIGHT && warp_idx >= 4 && n_blk < args.n_tiles[0]) {
    float* slot_ptr = args.wpart + (size_t)(row_base + lane) * args.wpart_ld + (n_blk * kEpiSets + cq);
    *slot_ptr += wsum;
  }
  __syncthreads();
  if (warp_idx == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

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
  EncodeTiledFn fn;
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

// CTA pairs (tcgen05 cta_group::2) are the default for 256-wide tiles; RBM_B200_PAIRS=0 forces single CTAs
bool pairs_enabled() {
  static const bool on = [] { const char* e = getenv("RBM_B200_PAIRS"); return e ? atoi(e) != 0 : true; }();
  return on;
}

constexpr int kMaxDeviceBetas = 1 << 18;   // longer AIS schedules fall back to one launch per half-step

struct Plan {
  int nV, nH, nVp, nHp;      // padded to multiples of 64
  int64_t rows_pad;          // chains padded to a multiple of 256 (one CTA-pair tile)
  int bn_h, bn_v;            // BLOCK_N for the v->h and h->v half-steps
  size_t off_W, off_nbh, off_nbv, off_V, off_H, off_wpart, off_betas, total;
  int wpart_ld;              // AIS log-weight partial slots per chain: 4 per hidden-side N tile
};

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

Plan make_plan(int nV, int nH, int64_t B) {
  Plan p;
  p.nV = nV; p.nH = nH;
  p.nVp = round_up(nV, 64); p.nHp = round_up(nH, 64);
  p.rows_pad = ceil_div64(B, 2 * BLOCK_M) * (2 * BLOCK_M);   // whole CTA-pair tiles (256 chains)
  p.bn_h = pick_block_n(p.nHp, p.rows_pad); p.bn_v = pick_block_n(p.nVp, p.rows_pad);
  auto align = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  size_t o = 0;
  p.off_W = o; o = align(o + (size_t)p.nVp * p.nHp * 2);
  p.off_nbh = o; o = align(o + (size_t)round_up(p.nHp, 256) * 4);
  p.off_nbv = o; o = align(o + (size_t)round_up(p.nVp, 256) * 4);
  p.off_V = o; o = align(o + (size_t)p.rows_pad * p.nVp * 2);
  p.off_H = o; o = align(o + (size_t)p.rows_pad * p.nHp * 2);
  p.wpart_ld = kEpiSets * ceil_div(p.nHp, p.bn_h);
  p.off_wpart = o; o = align(o + (size_t)p.rows_pad * p.wpart_ld * 4);
  p.off_betas = o; o = align(o + (size_t)kMaxDeviceBetas * 4);   // AIS schedule for the multi-step kernel
  p.total = o + 1024;  // slack to align the caller's pointer
  return p;
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

// cluster sizes the multi-step kernel is launched with (one CTA per output tile of a row block)
bool multistep_cluster_ok(int cs) { return cs == 1 || cs == 2 || cs == 4 || cs == 8; }

// -1: automatic (one wave of CTAs), 0: never, 1: whenever the shape allows
int multistep_mode() {
  static const int mode = [] { const char* e = getenv("RBM_B200_MULTISTEP"); return e ? atoi(e) : -1; }();
  return mode;
}

template <int BLOCK_N, int EPI0, int EPI1>
int launch_multistep(const CUtensorMap* tm, const MultiArgs& a, int m_tiles, int cs, cudaStream_t st) {
  auto kern = umma_multistep_kernel<BLOCK_N, EPI0, EPI1>;
  using L = SmemLayout<BLOCK_N, 1>;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if ((rc = ensure_dynamic_smem(kern, di.device, L::kTotal))) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(m_tiles * cs));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RBM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], a));
  return check_launch("umma_multistep_kernel");
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

// Raw counts of the final (v, h) pair of this call's chains, ADDED to S_vh [nV, nH], S_v [nV], S_h [nH].
int umma_stats(const Plan& pl, const __nv_bfloat16* Vs, const __nv_bfloat16* Hs, int64_t B, float* S_vh, float* S_v,
               float* S_h, int sm_count, cudaStream_t st) {
  RBM_REQUIRE(B <= (int64_t)1 << 24, RBM_B200_ERR_INVALID,
              "fused statistics count in fp32: at most 2^24 chains per call (got %lld)", (long long)B);
  int rc;
  CUtensorMap tv, th;
  // B rows: the padding chains read as zeros
  if ((rc = make_tmap(&tv, Vs, (uint64_t)B, (uint64_t)pl.nVp, (uint64_t)pl.nVp, BLOCK_K))) return rc;
  if ((rc = make_tmap(&th, Hs, (uint64_t)B, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_K))) return rc;
  const int bn = round_up(pl.nHp, 128) < round_up(pl.nHp, 256) ? 128 : 256;
  StatsArgs a{};
  a.S_vh = S_vh; a.nV = pl.nV; a.nH = pl.nH;
  const int tiles_m = ceil_div(pl.nV, BLOCK_M);
  a.tiles_n = ceil_div(pl.nH, bn);
  a.tiles = tiles_m * a.tiles_n;
  a.k_blocks = (int)ceil_div64(B, BLOCK_K);
  // split the chains so that one wave of CTAs covers the GPU; every split owns at least one K block
  int splits = std::max(1, std::min(a.k_blocks, sm_count / a.tiles));
  a.kb_per_split = ceil_div(a.k_blocks, splits);
  splits = ceil_div(a.k_blocks, a.kb_per_split);
  rc = bn == 256 ? launch_stats<256>(tv, th, a, a.tiles * splits, st) : launch_stats<128>(tv, th, a, a.tiles * splits, st);
  if (rc) return rc;
  const int64_t rows_per_block = std::max<int64_t>(64, ceil_div64(B, 4 * sm_count));
  const unsigned row_blocks = (unsigned)ceil_div64(B, rows_per_block);
  colsum_bf16_kernel<<<dim3((unsigned)ceil_div(pl.nVp, 256), row_blocks), 256, 0, st>>>(Vs, B, pl.nVp, pl.nV, rows_per_block, S_v);
  colsum_bf16_kernel<<<dim3((unsigned)ceil_div(pl.nHp, 256), row_blocks), 256, 0, st>>>(Hs, B, pl.nHp, pl.nH, rows_per_block, S_h);
  return check_launch("colsum_bf16_kernel");
}


// ===================================================================================================
// Decoder MLP on generated prior samples: the step right after the sampler in
// shower / image generation (models/autoencoders/gumboltCaloV5.py:78-96, dvaepp.py:214-225) -
//   BasicDecoder   (models/networks/basicCoders.py:28-42): chain of nn.Linear, ReLU between, last layer linear
//   BasicDecoderV3 (basicCoders.py:63-84): shared trunk (_layers, all ReLU), two heads (_layers2 hits, _layers3
//                  activations), last head layer linear
//   shower tail    sample = relu(activations) * heaviside(hits + logit(rho)) (gumboltCaloV5.py:91-93,
//                  utils/dists/gumbelmod.py:18-24) = relu(activations) * [sigmoid(hits) >= u], u = 1 - rho
// Every layer is one launch of the SAME tcgen05 kernel as a Gibbs half-step (A = activations K-major, B = the
// nn.Linear weight [out, in], which already is K-major).  fp32-grade accuracy on bf16 tensor cores: activations
// and weights are split x = x_hi + x_lo (both bf16) and the GEMM runs over the tripled K axis
//   [x_hi | x_hi | x_lo] . [W_hi | W_lo | W_hi]^T = x_hi W_hi + x_hi W_lo + x_lo W_hi      (error ~2^-16 relative),
// accumulated in fp32 in TMEM.  Hidden layers store [high | low] bf16 tiles for the next layer (EPI_LINEAR_SPLIT),
// the hits head ends in the Gibbs epilogue itself (sigmoid + Philox + compare, TAG_HITS) and the activation head
// multiplies that gate into relu(acc + bias) on its way to the fp32 output (EPI_LINEAR_F32).
constexpr uint32_t TAG_HITS = 4;   // Philox tag of the hit draws (oracle/philox.py)

__global__ void split_input_kernel(const float* __restrict__ x, int64_t B, int n, uint16_t* __restrict__ dst,
                                   int64_t rows_pad, int Kp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_pad * Kp) return;
  const int64_t r = idx / Kp;
  const int c = (int)(idx % Kp);
  const float val = (r < B && c < n) ? x[r * n + c] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(val);
  const __nv_bfloat16 lo = __float2bfloat16_rn(val - __bfloat162float(hi));
  dst[r * (2 * (int64_t)Kp) + c] = __bfloat16_as_ushort(hi);
  dst[r * (2 * (int64_t)Kp) + Kp + c] = __bfloat16_as_ushort(lo);
}
// W [N, K] fp32 -> [Np, 3 Kp] bf16 = [W_hi | W_lo | W_hi], zero padded
__global__ void split_weight_kernel(const float* __restrict__ W, int N, int K, uint16_t* __restrict__ dst, int Np, int Kp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Np * Kp) return;
  const int r = (int)(idx / Kp), c = (int)(idx % Kp);
  const float val = (r < N && c < K) ? W[(int64_t)r * K + c] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(val);
  const __nv_bfloat16 lo = __float2bfloat16_rn(val - __bfloat162float(hi));
  uint16_t* row = dst + (int64_t)r * (3 * Kp);
  row[c] = __bfloat16_as_ushort(hi);
  row[Kp + c] = __bfloat16_as_ushort(lo);
  row[2 * Kp + c] = __bfloat16_as_ushort(hi);
}
__global__ void pad_bias_kernel(const float* __restrict__ bias, float* __restrict__ dst, int n, int n_pad, float mul) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = i < n ? mul * bias[i] : 0.f;
}

struct LayerPlan {
  const rbm_b200_linear* lin;
  int Kp, Np, bn;
  size_t off_W, off_bias, off_out;   // off_out: [rows_pad, 2 Np] bf16 (hidden layers)
};
struct DecoderPlan {
  std::vector<LayerPlan> trunk, hits, act;
  int64_t rows_pad;
  int n_out, Np_out;
  size_t off_in, off_mask, off_nbias_hits, total;
};

int make_decoder_plan(const rbm_b200_decoder* d, int64_t B, DecoderPlan* out) {
  RBM_REQUIRE(d != nullptr && d->n_trunk >= 1 && d->trunk != nullptr, RBM_B200_ERR_INVALID, "decoder needs at least one trunk layer");
  RBM_REQUIRE(d->n_head >= 0 && (d->n_head == 0 || d->head_act != nullptr), RBM_B200_ERR_INVALID, "decoder head description is incomplete");
  DecoderPlan& pl = *out;
  pl.rows_pad = ceil_div64(std::max<int64_t>(B, 1), 2 * BLOCK_M) * (2 * BLOCK_M);
  auto align = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  size_t o = 0;
  auto add_chain = [&](const rbm_b200_linear* ls, int n, int in_features, std::vector<LayerPlan>& dst) -> int {
    int prev = in_features;
    for (int i = 0; i < n; ++i) {
      const rbm_b200_linear& l = ls[i];
      RBM_REQUIRE(l.in_features >= 1 && l.out_features >= 1 && l.in_features <= 16384 && l.out_features <= 16384,
                  RBM_B200_ERR_INVALID, "layer shape %d -> %d out of range", l.in_features, l.out_features);
      RBM_REQUIRE(l.in_features == prev, RBM_B200_ERR_INVALID, "size mismatch: layer expects %d inputs, previous layer has %d",
                  l.in_features, prev);
      RBM_REQUIRE(l.weight != nullptr && l.bias != nullptr, RBM_B200_ERR_INVALID, "layer parameter pointer is NULL");
      LayerPlan lp;
      lp.lin = &l;
      lp.Kp = round_up(l.in_features, 64); lp.Np = round_up(l.out_features, 64);
      lp.bn = pick_block_n(lp.Np, pl.rows_pad);
      // staged parameters first: their offsets do not depend on B, so a caller may keep them across calls
      lp.off_W = o; o = align(o + (size_t)lp.Np * 3 * lp.Kp * 2);
      lp.off_bias = o; o = align(o + (size_t)round_up(lp.Np, 256) * 4);
      lp.off_out = 0;
      dst.push_back(lp);
      prev = l.out_features;
    }
    return 0;
  };
  int rc = add_chain(d->trunk, d->n_trunk, d->trunk[0].in_features, pl.trunk);
  if (rc) return rc;
  const int trunk_out = d->trunk[d->n_trunk - 1].out_features;
  pl.n_out = trunk_out;
  if (d->n_head > 0) {
    if (d->head_hits && (rc = add_chain(d->head_hits, d->n_head, trunk_out, pl.hits))) return rc;
    if ((rc = add_chain(d->head_act, d->n_head, trunk_out, pl.act))) return rc;
    pl.n_out = d->head_act[d->n_head - 1].out_features;
    RBM_REQUIRE(!d->head_hits || d->head_hits[d->n_head - 1].out_features == pl.n_out, RBM_B200_ERR_INVALID,
                "size mismatch: the two heads end in %d and %d units", d->head_hits[d->n_head - 1].out_features, pl.n_out);
  }
  pl.Np_out = round_up(pl.n_out, 64);
  const int K0p = round_up(d->trunk[0].in_features, 64);
  pl.off_nbias_hits = o; o = align(o + (size_t)round_up(pl.Np_out, 256) * 4);
  // activations ([rows_pad, 2 Np] bf16 per hidden layer), input tile, hit gate
  for (std::vector<LayerPlan>* ch : {&pl.trunk, &pl.hits, &pl.act})
    for (LayerPlan& lp : *ch) { lp.off_out = o; o = align(o + (size_t)pl.rows_pad * 2 * lp.Np * 2); }
  pl.off_in = o; o = align(o + (size_t)pl.rows_pad * 2 * K0p * 2);
  pl.off_mask = o; o = align(o + (size_t)pl.rows_pad * pl.Np_out * 2);
  pl.total = o + 1024;
  return 0;
}

template <int EPI>
int launch_linear(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const HalfStepArgs& a,
                  int sm_count, cudaStream_t st) {
  return bn == 256 ? launch_half_step_impl<256, false, EPI, 1, true>(ta, tb, to, a, sm_count, st)
                   : launch_half_step_impl<128, false, EPI, 1, true>(ta, tb, to, a, sm_count, st);
}

}  // namespace

size_t umma_decoder_workspace_bytes(const rbm_b200_decoder* d, int64_t B) {
  DecoderPlan pl;
  if (make_decoder_plan(d, B, &pl)) return 0;
  return pl.total;
}

int umma_decoder_forward(const rbm_b200_decoder* d, const float* x, int64_t B, float* out_hits, float* out_act,
                         float* samples, uint64_t seed, uint64_t chain_offset, uint64_t step, int params_staged,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
  DecoderPlan pl;
  int rc = make_decoder_plan(d, B, &pl);
  if (rc) return rc;
  RBM_REQUIRE(ws != nullptr && ws_bytes >= pl.total, RBM_B200_ERR_WORKSPACE,
              "decoder needs %zu bytes of workspace, got %zu", pl.total, ws_bytes);
  RBM_REQUIRE(samples == nullptr || !pl.hits.empty(), RBM_B200_ERR_INVALID,
              "shower samples need both heads (hits and activations)");
  RBM_REQUIRE(out_hits == nullptr || !pl.hits.empty(), RBM_B200_ERR_INVALID, "there is no hits head to output");
  if (B == 0) return 0;
  RBM_REQUIRE(x != nullptr, RBM_B200_ERR_INVALID, "input pointer is NULL");
  DeviceInfo di;
  if ((rc = get_device_info(&di))) return rc;
  RBM_REQUIRE(di.cc_major == 10, RBM_B200_ERR_NO_DEVICE, "the decoder path needs an sm_100 device (got cc %d.x)", di.cc_major);
  const int sm_count = di.sm_count;
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  const int in_features = d->trunk[0].in_features;
  const int K0p = round_up(in_features, 64);
  uint16_t* A0 = reinterpret_cast<uint16_t*>(base + pl.off_in);
  uint16_t* mask = reinterpret_cast<uint16_t*>(base + pl.off_mask);
  float* nbias_hits = reinterpret_cast<float*>(base + pl.off_nbias_hits);

  // stage: split input, split weights, padded biases
  {
    const int64_t n = pl.rows_pad * K0p;
    split_input_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(x, B, in_features, A0, pl.rows_pad, K0p);
  }
  auto stage_chain = [&](const std::vector<LayerPlan>& ch) {
    for (const LayerPlan& lp : ch) {
      const int64_t n = (int64_t)lp.Np * lp.Kp;
      split_weight_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(lp.lin->weight, lp.lin->out_features, lp.lin->in_features,
                                                                     reinterpret_cast<uint16_t*>(base + lp.off_W), lp.Np, lp.Kp);
      const int np256 = round_up(lp.Np, 256);
      pad_bias_kernel<<<ceil_div(np256, 256), 256, 0, st>>>(lp.lin->bias, reinterpret_cast<float*>(base + lp.off_bias),
                                                           lp.lin->out_features, np256, 1.f);
    }
  };
  if (!params_staged) { stage_chain(pl.trunk); stage_chain(pl.hits); stage_chain(pl.act); }
  if (!params_staged && !pl.hits.empty()) {
    const LayerPlan& lp = pl.hits.back();
    const int np256 = round_up(lp.Np, 256);
    pad_bias_kernel<<<ceil_div(np256, 256), 256, 0, st>>>(lp.lin->bias, nbias_hits, lp.lin->out_features, np256,
                                                         -1.4426950408889634f);
  }
  if ((rc = check_launch("decoder staging kernels"))) return rc;

  // one layer = one launch.  mode: 0 hidden (split bf16 out), 1 fp32 out, 2 hit gate (Gibbs epilogue), 3 gated shower
  auto run_layer = [&](const LayerPlan& lp, const uint16_t* A, int Kp_in, int mode, float* dst_f32) -> int {
    CUtensorMap ta, tb, to;
    int r;
    if ((r = make_tmap(&ta, A, (uint64_t)pl.rows_pad, (uint64_t)(2 * Kp_in), (uint64_t)(2 * Kp_in), BLOCK_M))) return r;
    if ((r = make_tmap(&tb, base + lp.off_W, (uint64_t)lp.Np, (uint64_t)(3 * lp.Kp), (uint64_t)(3 * lp.Kp), (uint32_t)lp.bn))) return r;
    HalfStepArgs a{};
    a.nbias = reinterpret_cast<const float*>(base + lp.off_bias);
    a.ldo = lp.Np; a.m_tiles = (int)(pl.rows_pad / BLOCK_M); a.n_tiles = ceil_div(lp.Np, lp.bn);
    a.k_blocks = 3 * lp.Kp / BLOCK_K; a.n_valid = lp.lin->out_features; a.chain_offset = chain_offset;
    a.rows_valid = B;
    if (mode == 0) {
      uint16_t* out = reinterpret_cast<uint16_t*>(base + lp.off_out);
      if ((r = make_tmap(&to, out, (uint64_t)pl.rows_pad, (uint64_t)(2 * lp.Np), (uint64_t)(2 * lp.Np), 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
      a.relu = 1; a.lo_off = lp.Np;
      return launch_linear<EPI_LINEAR_SPLIT>(lp.bn, ta, tb, to, a, sm_count, st);
    }
    if (mode == 2) {
      if ((r = make_tmap(&to, mask, (uint64_t)pl.rows_pad, (uint64_t)lp.Np, (uint64_t)lp.Np, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return r;
      a.nbias = nbias_hits;
      a.ctx = make_draw_ctx(seed, step, TAG_HITS);
      return launch_linear<EPI_GIBBS>(lp.bn, ta, tb, to, a, sm_count, st);
    }
    a.out_f32 = dst_f32; a.ld_out = lp.lin->out_features;
    a.relu = mode == 3 ? 1 : 0;
    a.mask = mode == 3 ? mask : nullptr; a.ld_mask = lp.Np;
    return launch_linear<EPI_LINEAR_F32>(lp.bn, ta, tb, ta, a, sm_count, st);   // no TMA store: any valid map
  };

  // trunk
  const uint16_t* cur = A0;
  int cur_Kp = K0p;
  const bool single = pl.act.empty();
  for (size_t i = 0; i < pl.trunk.size(); ++i) {
    const LayerPlan& lp = pl.trunk[i];
    if (single && i + 1 == pl.trunk.size()) {
      // BasicDecoder: the last layer is linear (output_activation_fct = Identity, basicCoders.py:38-39)
      if (out_act != nullptr && (rc = run_layer(lp, cur, cur_Kp, 1, out_act))) return rc;
      return 0;
    }
    if ((rc = run_layer(lp, cur, cur_Kp, 0, nullptr))) return rc;
    cur = reinterpret_cast<const uint16_t*>(base + lp.off_out); cur_Kp = lp.Np;
  }
  // heads
  auto run_head = [&](const std::vector<LayerPlan>& ch, const uint16_t** last_in, int* last_Kp) -> int {
    const uint16_t* a_in = cur; int kp = cur_Kp;
    for (size_t i = 0; i + 1 < ch.size(); ++i) {
      int r = run_layer(ch[i], a_in, kp, 0, nullptr);
      if (r) return r;
      a_in = reinterpret_cast<const uint16_t*>(base + ch[i].off_out); kp = ch[i].Np;
    }
    *last_in = a_in; *last_Kp = kp;
    return 0;
  };
  const uint16_t* in_hits = nullptr; const uint16_t* in_act = nullptr;
  int kp_hits = 0, kp_act = 0;
  if (!pl.hits.empty() && (out_hits != nullptr || samples != nullptr)) {
    if ((rc = run_head(pl.hits, &in_hits, &kp_hits))) return rc;
    if (out_hits != nullptr && (rc = run_layer(pl.hits.back(), in_hits, kp_hits, 1, out_hits))) return rc;
    if (samples != nullptr && (rc = run_layer(pl.hits.back(), in_hits, kp_hits, 2, nullptr))) return rc;
  }
  if (out_act != nullptr || samples != nullptr) {
    if ((rc = run_head(pl.act, &in_act, &kp_act))) return rc;
    if (out_act != nullptr && (rc = run_layer(pl.act.back(), in_act, kp_act, 1, out_act))) return rc;
    if (samples != nullptr && (rc = run_layer(pl.act.back(), in_act, kp_act, 3, samples))) return rc;
  }
  return 0;
}

bool umma_variant_supported(int nV, int nH) { return nV >= 1 && nH >= 1 && nV <= 16384 && nH <= 16384; }

size_t umma_workspace_bytes(int nV, int nH, int64_t B) { return make_plan(nV, nH, B).total; }

int umma_block_gibbs(const rbm_b200_params* p, float* v, float* h, int64_t B, int k, int init_chains,
                     uint64_t seed, uint64_t chain_offset, uint64_t step_offset, void* ws, size_t ws_bytes,
                     cudaStream_t st, float* S_vh, float* S_v, float* S_h) {
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
  float* nbh = reinterpret_cast<float*>(base + pl.off_nbh);
  float* nbv = reinterpret_cast<float*>(base + pl.off_nbv);
  __nv_bfloat16* Vs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_V);
  __nv_bfloat16* Hs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_H);

  // stage parameters: padded bf16 W, pre-scaled biases
  {
    const int64_t n = (int64_t)pl.nVp * pl.nHp;
    pad_w_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W_bf16, Wp, pl.nV, pl.nH, pl.nVp, pl.nHp);
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

  // Chains are independent, so the k-step loop MAY run chunk by chunk (RBM_B200_CHUNK_ROWS), e.g. with chunks
  // whose two state tiles fit in L2 so that half-steps stop round-tripping through HBM.  Measured on B200
  // (profiles/r01_pairs_vs_single.txt): at config 4 it is a LOSS (158 vs 141 us per launch-equivalent for
  // 18,944-chain chunks) - the per-launch overheads of 3.5x more, 3.5x smaller launches outweigh the HBM
  // energy saved - so the default is one chunk.
  int64_t chunk_rows = pl.rows_pad;
  {
    static const long env_chunk = [] { const char* e = getenv("RBM_B200_CHUNK_ROWS"); return e ? atol(e) : 0L; }();
    if (env_chunk > 0) chunk_rows = std::min<int64_t>(pl.rows_pad, ceil_div64(env_chunk, 2 * BLOCK_M) * 2 * BLOCK_M);
  }
  static const int pairs_min_tiles = [] { const char* e = getenv("RBM_B200_PAIRS_MIN_TILES"); return e ? atoi(e) : 8; }();

  // tensor maps: states are K-major A operands; W is read MN-major (v->h) and K-major (h->v)
  CUtensorMap tm_w_mn, tm_w_k_single, tm_w_k_pair;
  if ((rc = make_tmap(&tm_w_mn, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_K))) return rc;
  // h->v reads W K-major in [n rows x 64] boxes: a CTA of a pair stages half of the 256 rows
  if ((rc = make_tmap(&tm_w_k_single, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, (uint32_t)pl.bn_v))) return rc;
  if ((rc = make_tmap(&tm_w_k_pair, Wp, (uint64_t)pl.nVp, (uint64_t)pl.nHp, (uint64_t)pl.nHp, 128))) return rc;

  // Few chains: one launch for all 2k half-steps (cluster per 128-chain row block, see umma_multistep_kernel)
  bool ran_multistep = false;
  {
    const int nt_h = ceil_div(pl.nHp, pl.bn_h), nt_v = ceil_div(pl.nVp, pl.bn_v);
    const int cs = std::max(nt_h, nt_v);
    const int m_tiles = (int)(pl.rows_pad / BLOCK_M);
    const int mode = multistep_mode();
    const bool shape_ok = k >= 1 && pl.bn_h == pl.bn_v && multistep_cluster_ok(cs);
    if (shape_ok && (mode == 1 || (mode == -1 && (int64_t)m_tiles * cs <= sm_count))) {
      CUtensorMap tm[6];
      if ((rc = make_tmap(&tm[0], Vs, (uint64_t)pl.rows_pad, (uint64_t)pl.nVp, (uint64_t)pl.nVp, BLOCK_M))) return rc;
      if ((rc = make_tmap(&tm[1], Hs, (uint64_t)pl.rows_pad, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_M))) return rc;
      tm[2] = tm_w_mn;
      tm[3] = tm_w_k_single;
      if ((rc = make_tmap(&tm[4], Hs, (uint64_t)pl.rows_pad, (uint64_t)pl.nHp, (uint64_t)pl.nHp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
      if ((rc = make_tmap(&tm[5], Vs, (uint64_t)pl.rows_pad, (uint64_t)pl.nVp, (uint64_t)pl.nVp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
      MultiArgs ma{};
      ma.nbias[0] = nbh; ma.nbias[1] = nbv; ma.ldo[0] = pl.nHp; ma.ldo[1] = pl.nVp;
      ma.n_valid[0] = pl.nH; ma.n_valid[1] = pl.nV; ma.n_tiles[0] = nt_h; ma.n_tiles[1] = nt_v;
      ma.k_blocks[0] = pl.nVp / BLOCK_K; ma.k_blocks[1] = pl.nHp / BLOCK_K;
      ma.steps = 2 * k; ma.first_dir = 0; ma.seed = seed; ma.chain_offset = chain_offset; ma.step_offset = step_offset;
      ma.tag = TAG_GIBBS; ma.kf_magic = kKfMagic;
      rc = pl.bn_h == 256 ? launch_multistep<256, EPI_GIBBS, EPI_GIBBS>(tm, ma, m_tiles, cs, st)
                          : launch_multistep<128, EPI_GIBBS, EPI_GIBBS>(tm, ma, m_tiles, cs, st);
      if (rc) return rc;
      ran_multistep = true;
    }
  }

  for (int64_t r0 = 0; !ran_multistep && r0 < pl.rows_pad; r0 += chunk_rows) {
    const int64_t rows = std::min(chunk_rows, pl.rows_pad - r0);
    __nv_bfloat16* Vc = Vs + r0 * pl.nVp;
    __nv_bfloat16* Hc = Hs + r0 * pl.nHp;
    // pairs pay off once every cluster loops over enough tiles (profiles/r01_pairs_vs_single.txt)
    const bool pairs = pairs_enabled() &&
                       (rows / (2 * BLOCK_M)) * ceil_div(std::max(pl.nVp, pl.nHp), 256) >= (int64_t)pairs_min_tiles * (sm_count / 2);
    CUtensorMap tm_v, tm_h, to_v, to_h;
    if ((rc = make_tmap(&tm_v, Vc, (uint64_t)rows, (uint64_t)pl.nVp, (uint64_t)pl.nVp, BLOCK_M))) return rc;
    if ((rc = make_tmap(&tm_h, Hc, (uint64_t)rows, (uint64_t)pl.nHp, (uint64_t)pl.nHp, BLOCK_M))) return rc;
    // epilogue stores: [32 chains x 32 units] boxes, 64-byte swizzle
    if ((rc = make_tmap(&to_v, Vc, (uint64_t)rows, (uint64_t)pl.nVp, (uint64_t)pl.nVp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_tmap(&to_h, Hc, (uint64_t)rows, (uint64_t)pl.nHp, (uint64_t)pl.nHp, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    const CUtensorMap& tm_w_k = (pairs && pl.bn_v == 256) ? tm_w_k_pair : tm_w_k_single;

    HalfStepArgs ah{}, av{};
    ah.out = Hc; ah.nbias = nbh; ah.ldo = pl.nHp; ah.m_tiles = (int)(rows / BLOCK_M);
    ah.n_tiles = ceil_div(pl.nHp, pl.bn_h); ah.k_blocks = pl.nVp / BLOCK_K; ah.chain_offset = chain_offset + (uint64_t)r0; ah.n_valid = pl.nH;
    av.out = Vc; av.nbias = nbv; av.ldo = pl.nVp; av.m_tiles = ah.m_tiles;
    av.n_tiles = ceil_div(pl.nVp, pl.bn_v); av.k_blocks = pl.nHp / BLOCK_K; av.chain_offset = chain_offset + (uint64_t)r0; av.n_valid = pl.nV;

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
                 uint64_t chain_offset, double* logw, void* ws, size_t ws_bytes, cudaStream_t st) {
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

  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  uint16_t* Wp = reinterpret_cast<uint16_t*>(base + pl.off_W);
  float* nbh = reinterpret_cast<float*>(base + pl.off_nbh);
  float* nbv = reinterpret_cast<float*>(base + pl.off_nbv);
  __nv_bfloat16* Vs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_V);
  __nv_bfloat16* Hs = reinterpret_cast<__nv_bfloat16*>(base + pl.off_H);
  float* wpart = reinterpret_cast<float*>(base + pl.off_wpart);

  {
    const int64_t n = (int64_t)pl.nVp * pl.nHp;
    pad_w_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(p->W_bf16, Wp, pl.nV, pl.nH, pl.nVp, pl.nHp);
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

  const int K = n_betas - 1;
  // Few chains per GPU (e.g. config 5 sharded over 8 GPUs: 2048 chains): the whole schedule in ONE launch
  {
    const int nt_h = ceil_div(pl.nHp, pl.bn_h), nt_v = ceil_div(pl.nVp, pl.bn_v);
    const int cs = std::max(nt_h, nt_v);
    const int m_tiles = (int)(pl.rows_pad / BLOCK_M);
    const int mode = multistep_mode();
    const bool shape_ok = pl.bn_h == pl.bn_v && multistep_cluster_ok(cs) && n_betas <= kMaxDeviceBetas && !pairs;
    if (shape_ok && (mode == 1 || (mode == -1 && (int64_t)m_tiles * cs <= sm_count))) {
      float* betas_dev = reinterpret_cast<float*>(base + pl.off_betas);
      std::vector<float> bf(n_betas);
      for (int i = 0; i < n_betas; ++i) bf[i] = (float)betas[i];
      // pageable source: the runtime stages the copy before returning, so `bf` may go out of scope
      RBM_CUDA_TRY(cudaMemcpyAsync(betas_dev, bf.data(), sizeof(float) * (size_t)n_betas, cudaMemcpyHostToDevice, st));
      CUtensorMap tm[6];
      tm[0] = tm_v; tm[1] = tm_h; tm[2] = tm_w_mn; tm[3] = tm_w_k; tm[4] = to_h; tm[5] = to_v;
      MultiArgs ma{};
      ma.nbias[0] = nbh; ma.nbias[1] = nbv; ma.ldo[0] = pl.nHp; ma.ldo[1] = pl.nVp;
      ma.n_valid[0] = pl.nH; ma.n_valid[1] = pl.nV; ma.n_tiles[0] = nt_h; ma.n_tiles[1] = nt_v;
      ma.k_blocks[0] = pl.nVp / BLOCK_K; ma.k_blocks[1] = pl.nHp / BLOCK_K;
      ma.steps = 2 * K - 1;            // K hidden steps, K-1 visible steps (the last transition is skipped)
      ma.first_dir = 0; ma.seed = seed; ma.chain_offset = chain_offset; ma.step_offset = 2;   // half-step counter 2k / 2k+1
      ma.tag = TAG_AIS; ma.kf_magic = kKfMagic; ma.betas = betas_dev; ma.wpart = wpart; ma.wpart_ld = pl.wpart_ld;
      rc = pl.bn_h == 256 ? launch_multistep<256, EPI_AIS_WEIGHT, EPI_ANNEALED>(tm, ma, m_tiles, cs, st)
                          : launch_multistep<128, EPI_AIS_WEIGHT, EPI_ANNEALED>(tm, ma, m_tiles, cs, st);
      if (rc) return rc;
      ais_reduce_kernel<<<(unsigned)ceil_div64(M, 256), 256, 0, st>>>(wpart, pl.wpart_ld, M, logw);
      return check_launch("ais_reduce_kernel");
    }
  }
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
