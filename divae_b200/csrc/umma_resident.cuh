// Resident kernel: priors up to 512 x 512 with W held in shared memory and the chain states kept on the chip for ALL
// half-steps of a block-Gibbs (or AIS) call.  (BASELINE configs[4]: AIS of a 512 x 512 prior; also few chains on
// 256 x 256 ... 512 x 512 priors, where a half-step is bound by its dependency latency, not by the tensor pipe.)
//
// A thread-block CLUSTER of C = units / 64 CTAs (<= 8) owns one 128-chain row block - or two, ping-pong - for the whole
// schedule.  CTA j computes the 64 output units [64 j, 64 j + 64) of every half-step:
//   * its two W slices stay in shared memory for the life of the kernel (loaded once by TMA):
//       direction 0 (v -> h): W[:, 64j : 64j+64]  as K blocks of [64 k x 64 n], MN-major B operand
//       direction 1 (h -> v): W[64j : 64j+64, :]  as K blocks of [64 n x 64 k], K-major B operand
//     (2 x 64 KB at 512 x 512 - there is no W traffic at all after the prologue);
//   * the states travel between the CTAs as BIT WORDS over distributed shared memory: an epilogue lane owns one chain
//     (one TMEM lane) and 32 units, i.e. exactly one 32-bit word; a warp stages its 32 words in local shared memory and
//     issues ONE 128-byte cp.async.bulk.shared::cluster per consumer CTA (... mbarrier::complete_tx): 8 KB per CTA and
//     half-step instead of the 128 KB bf16 tile, no global memory, no flags - the consumer has one mbarrier per SOURCE
//     CTA (= per K block of its next half-step), which completes when that CTA's words have landed, so a late CTA delays
//     only its own K block;
//   * expander warps turn the words of one K block (128 chains x 64 units) into bf16 pairs in registers - a set unit
//     becomes 2.0 = 0x4000, a single bit of the high byte: shift + AND per unit-byte, one PRMT per pair = ONE ALU
//     instruction per unit; the epilogue halves its scale, both exact - and store them straight into TENSOR MEMORY
//     (tcgen05.st: a thread owns one chain = one TMEM lane, a K block is 32 columns): the A operand of the MMAs is read
//     from TMEM (tcgen05.mma with A in tensor memory), so the state never passes through shared memory as bf16 - no
//     swizzled stores, no proxy fence, no A reads on the shared-memory port, and all 8 K blocks of a half-step have
//     their own TMEM columns (a ring of 8 x 32 columns next to the two 64-column accumulators);
//     the MMA warp issues tcgen05.mma (M = 128, N = 64, B = the resident W slice) two K blocks per trip as they appear;
//   * the epilogue (tcgen05.ld -> bias, sigmoid, Philox, Bernoulli compare: the same arithmetic and draw contract as
//     umma_common.cuh:epilogue_tile, so every kernel draws the same chains) produces the next words.
// With two row blocks per cluster (ping-pong, two TMEM accumulators) the MMAs and the expansion of one block overlap
// the epilogue of the other; with one the kernel runs at the latency of a single dependent half-step.
// Clusters are independent of each other (chains are independent), so they need not be co-resident.
#pragma once
#include <mutex>
#include <type_traits>

#include "umma_chain.cuh"

namespace rbm {
namespace {

constexpr int kResBN = 64;                 // output units per CTA and direction
constexpr int kResStages = 8;              // A ring in tensor memory: 32 columns (64 bf16 units) per stage
constexpr int kResAccCols = kResBN;        // one 64-column accumulator per row block
constexpr int kResAColumn = 2 * kResAccCols;   // first TMEM column of the A ring (behind the accumulators of both slots)
constexpr int kResEpiWarps = 8;            // warps 4..11: 4 TMEM lane quarters x 2 column sets of 32 units
#ifndef RBM_RES_EXP_WARPS
#define RBM_RES_EXP_WARPS 8
#endif
constexpr int kResExpWarps = RBM_RES_EXP_WARPS;   // warps 12..: groups of four (one per TMEM lane quarter) take K blocks in turn
constexpr int kResExpGroups = kResExpWarps / 4;
constexpr int kResThreads = 32 * (4 + kResEpiWarps + kResExpWarps);
constexpr int kResMaxTiles = 8;            // portable cluster size
constexpr int kResWBlockBytes = BLOCK_K * kResBN * 2;        // 8 KB: one K block of a W slice
constexpr int kResBitsBytes = kResMaxTiles * 2 * BLOCK_M * 4;   // 8 KB: the words of one (slot, parity)
constexpr int kResTableBytes = 256 * 16;

struct ResidentArgs {
  const float* nbias[2];   // [dir]: -log2(e) * bias of the units WRITTEN by dir (0: hidden, 1: visible), zero padded
  int n_valid[2];          // real unit count written by dir
  int n_tiles[2];          // 64-unit tiles written by dir (= K blocks READ by the other direction)
  int ld[2];               // padded unit count (row pitch) of the state written by dir
  int row_blocks;          // 128-chain blocks
  int slots;               // 1 or 2 row blocks per cluster in flight
  int steps;               // half-steps, starting with direction 0
  uint64_t seed, chain_offset, step_offset;
  uint32_t tag;
  const float* betas;      // annealed flavours: beta index of half-step t is 1 + t/2
  float* wpart;            // AIS: [rows, wpart_ld] log-weight partials (log2 units), one slot per 32 hidden units
  int wpart_ld;
  uint32_t kf_magic;
  const __nv_bfloat16* v_init;   // [rows_pad, ld[1]] initial visible state ({0,1} bf16)
  __nv_bfloat16* out[2];         // final states: out[0] = hidden [rows_pad, ld[0]], out[1] = visible [rows_pad, ld[1]]
};

constexpr int kResStageBytes = 2 * 2 * 1024;   // [slot][use parity][word 0 / 1 of this CTA][128 chains]
constexpr int kResPrologueBytes = 2 * 8 * 128;       // [slot][expander warp < 8][32 words]

// Latency tracing (debug builds only: RBM_B200_NVCC_DEFS=-DRBM_RES_TRACE): SM clock of the key events of a few
// half-steps of CTA 0, read back through rbm_b200_debug_resident_trace().
#ifdef RBM_RES_TRACE
__device__ unsigned long long g_res_trace[64 * 16];
#define RES_TRACE(ev, t, cond)                                                                       \
  do {                                                                                               \
    if (blockIdx.x == 0 && (cond) && (t) >= 200 && (t) < 216) g_res_trace[((t) - 200) * 64 + (ev)] = clock64(); \
  } while (0)
#else
#define RES_TRACE(ev, t, cond) do { } while (0)
#endif

struct ResidentSmem {
  uint32_t w0, w1, ring, table, bits, stage, pro, bars;
  int total;
};
__host__ __device__ inline ResidentSmem resident_smem(int tiles_h, int tiles_v) {
  ResidentSmem s;
  s.w0 = 0;                                                  // K blocks of direction 0 = visible tiles
  s.w1 = s.w0 + (uint32_t)tiles_v * kResWBlockBytes;          // K blocks of direction 1 = hidden tiles
  s.ring = s.w1 + (uint32_t)tiles_h * kResWBlockBytes;
  s.table = s.ring;                                          // (the A ring lives in tensor memory)
  s.bits = s.table + kResTableBytes;
  s.stage = s.bits + 4 * kResBitsBytes;                       // bits: [slot][parity]
  s.pro = s.stage + kResStageBytes;
  s.bars = s.pro + kResPrologueBytes;
  s.total = (int)s.bars + 512 + 1024;
  return s;
}

// bulk copy shared::cta -> shared::cluster (any CTA of the cluster, this one included); the destination CTA's mbarrier
// receives ONE complete_tx of `bytes` when the data has landed
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t cluster_dst, uint32_t cta_src, uint32_t bytes, uint32_t cluster_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(cluster_dst), "r"(cta_src), "r"(bytes), "r"(cluster_bar) : "memory");
}
// generic-mode PRMT: selector nibble bit 3 replicates the sign bit of the selected byte over the whole byte
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// 32 consecutive 32-bit columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] B[smem]: the A operand (M = 128 lanes x 8 columns = 16 bf16 of K per lane) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// mbarrier wait with a suspend-time hint: the waits of this kernel are long (a dependent half-step away) and every
// spin of a waiting warp takes an issue slot from the expander / epilogue warps
// RBM_RES_HINT: which waits carry the hint (bit 0: the epilogue's wait for the accumulator, bit 1: the MMA warp's waits,
// bit 2: the expanders' waits).  In the single-block trace the epilogue sees the accumulator ~700 cycles after the commit
// was issued, of which the tensor pipe accounts for ~300; un-hinted waits on the MMA and epilogue side (RBM_RES_HINT=4)
// measured the same within noise (512 x 512: 2.81 vs 2.83 us per half-step at 1,024 chains, 4.03 vs 4.06 at 2,048), so
// the wake-up of a hinted wait is not where that time goes.
#ifndef RBM_RES_HINT
#define RBM_RES_HINT 7
#endif
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
  } while (!done);
}
template <int BIT>
__device__ __forceinline__ void mbar_wait_role(uint32_t bar, uint32_t parity) {
  if ((RBM_RES_HINT >> BIT) & 1) mbar_wait_long(bar, parity); else mbar_wait(bar, parity);
}


// Long waits of a whole role (8 epilogue warps for an accumulator, 8 expander warps for the words of a half-step): only the
// role's first warp polls the mbarrier, the others block on a named barrier.  A warp sleeping in try_wait is woken by
// every mbarrier event of the SM (~160 times per half-step and row block here: bulk-copy completions, ring arrivals) and
// re-checks - 10 % of all issued instructions of the ping-pong profile (profiles/r02_ncu_resident_v8_*), taken from the
// warps that work; a warp blocked in bar.sync issues nothing.  MEASURED SLOWER (Gibbs 512 x 512: 3.10 vs 3.06 us per
// half-step at 1,024 chains, 4.28 vs 4.13 at 2,048: the role then moves at the pace of its slowest warp) - off by default.
#ifndef RBM_RES_LEADER_WAIT
#define RBM_RES_LEADER_WAIT 0
#endif
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void role_wait(uint32_t bar, uint32_t parity, bool leader, int bar_id, int threads) {
#if RBM_RES_LEADER_WAIT
  if (leader) mbar_wait_long(bar, parity);
  named_bar_sync(bar_id, threads);
#else
  mbar_wait_role<0>(bar, parity);
#endif
}

// One epilogue warp, one half-step of one row block: 32 chains (lanes) x 32 units -> one 32-bit word per lane.
// Same arithmetic as epilogue_tile (umma_common.cuh): arg = acc * scale + nbias * bmul, s = 1 + 2^arg,
// d = 65536 - k16 s; fire iff d >= s; 0 <= d < s is re-done on the exact path.
template <int EPI, class WaitFn, class ReleaseFn>
__device__ __forceinline__ uint32_t resident_epilogue(const EpiParams& ep, uint32_t tmem_acc, int col0, uint32_t chain,
                                                      int quarter, int cq, int lane, float& wsum, WaitFn wait_full,
                                                      ReleaseFn release) {
  uint32_t w[4][4];
  const uint32_t blk0 = (uint32_t)(col0 >> 5) * 4u;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const uint4 x = draw_block(ep.ctx, blk0 + b, chain);
    w[b][0] = x.x; w[b][1] = x.y; w[b][2] = x.z; w[b][3] = x.w;
  }
  wait_full();
  tcgen05_fence_after();
  uint32_t bits = 0;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    uint32_t r[16];
    tmem_ld16(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cq * 32 + 16 * hh), r);
    tmem_ld_wait();
    if (hh == 1) {
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) release();
    }
    // ambq[q4]: some unit of the 4-unit group q4 cannot be decided by its 16-bit field alone (probability 2^-16 per draw)
    bool ambq[4] = {false, false, false, false};
    uint32_t half_bits = 0;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 nb4 = __ldg(reinterpret_cast<const float4*>(ep.nbias + col0 + 16 * hh) + q4);
      const float nbv[4] = {nb4.x, nb4.y, nb4.z, nb4.w};
      AisQuad quad;
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const int c16 = 4 * q4 + i4;
        const int cc = 16 * hh + c16;
        const int b = (cc & 7) >> 1;
        const int slot = 2 * (cc >> 3) + (cc & 1);
        const uint32_t word = w[b][slot >> 1];
        const uint32_t mbits = (slot & 1) ? __byte_perm(word, ep.kf_magic, 0x7632) : __byte_perm(word, ep.kf_magic, 0x7610);
        const float kf = __uint_as_float(mbits) - 8388608.0f;
        // (the A operand holds 2.0 for a set unit, see the expanders: the accumulator is exactly twice the
        //  pre-activation and the scale is halved - both are exact, so arg is bit for bit what the other kernels compute)
        const float arg = EPI == EPI_GIBBS ? fmaf(__uint_as_float(r[c16]), -0.5f * 1.4426950408889634f, nbv[i4])
                          : EPI == EPI_ANNEALED ? fmaf(__uint_as_float(r[c16]), ep.scale, nbv[i4])
                                                : fmaf(__uint_as_float(r[c16]), ep.scale, nbv[i4] * ep.bmul);
        const Decision q = decision_terms(arg, kf);
        if (EPI == EPI_AIS_WEIGHT) {
          // log-weight terms of the 4 units of q4 at a time (umma_common.cuh:ais_quad_add - the same arithmetic in the same
          // order as the other kernels; padding units contribute exactly 0)
          ais_quad_add(quad, i4, arg, q.s, ep.ratio);
          if (i4 == 3) wsum += ais_quad_term(quad);
        }
        // the unit's bit: one FSETP and one predicated OR (this kernel is bound by issued instructions)
        asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"
            : "+r"(half_bits) : "f"(q.d), "f"(q.s), "r"(1u << c16));
        ambq[q4] = ambq[q4] || (__float_as_uint(q.d) < __float_as_uint(q.s));
      }
    }
    if (ambq[0] || ambq[1] || ambq[2] || ambq[3]) {
      // Only the ambiguous units go through the exact path (for every other unit it returns what the fast path
      // decided), and only the 4-unit groups that flagged are looked at: in this kernel a whole cluster waits for its
      // slowest warp every half-step and SOME warp of a 512-unit cluster takes this path in 63 % of the half-steps, so
      // the rare path must stay short (it was ~1,800 cycles when it re-examined all 16 units: res_trace_*_v6.txt).
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        if (!ambq[q4]) continue;
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int c16 = 4 * q4 + i4;
          const int cc = 16 * hh + c16;
          const int b = (cc & 7) >> 1, slot = 2 * (cc >> 3) + (cc & 1);
          const uint32_t word = w[b][slot >> 1];
          const uint32_t k16 = (slot & 1) ? (word >> 16) : (word & 0xFFFFu);
          const float nbx = __ldg(ep.nbias + col0 + cc);
          const float arg = EPI == EPI_GIBBS ? fmaf(__uint_as_float(r[c16]), -0.5f * 1.4426950408889634f, nbx)
                            : EPI == EPI_ANNEALED ? fmaf(__uint_as_float(r[c16]), ep.scale, nbx)
                                                  : fmaf(__uint_as_float(r[c16]), ep.scale, nbx * ep.bmul);
          const Decision q = decision_terms(arg, (float)k16);
          if (__float_as_uint(q.d) < __float_as_uint(q.s)) {
            const uint32_t one = resolve_exact(arg, k16, ep.ctx, blk0 + b, chain, (uint32_t)slot);
            half_bits = (half_bits & ~(1u << c16)) | (one << c16);
          }
        }
      }
    }
    bits |= half_bits << (16 * hh);
  }
  return bits;
}

template <int EPI0, int EPI1>
__global__ void __launch_bounds__(kResThreads, 1)
umma_resident_kernel(const __grid_constant__ CUtensorMap tm_w, const ResidentArgs args) {
  static_assert(EPI0 != EPI_ANNEALED, "EPI_ANNEALED is the direction-1 flavour here (bias multiplier 1)");
  static_assert(kResStages == 8, "ring indexing assumes 8 stages");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const ResidentSmem L = resident_smem(args.n_tiles[0], args.n_tiles[1]);
  const uint32_t bar_base = smem_base + L.bars;
  const uint32_t w_full = bar_base;
  auto ring_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto ring_empty = [&](int s) { return bar_base + 8u * (1 + kResStages + s); };
  auto tmem_full = [&](int s) { return bar_base + 8u * (1 + 2 * kResStages + s); };
  auto tmem_empty = [&](int s) { return bar_base + 8u * (3 + 2 * kResStages + s); };
  // words of one K block = the 64 units CTA kb of the cluster wrote: one barrier per (slot, parity, source CTA), so a late
  // CTA delays only its own K block (the whole cluster used to wait for its slowest warp before ANY expansion started)
  auto bits_full = [&](int s, int par, int kb) { return bar_base + 8u * (10 + 2 * kResStages + (2 * s + par) * kResMaxTiles + kb); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L.bars + 8 * (9 + 2 * kResStages));
  auto bits_addr = [&](int s, int par, int word, int row) {
    return smem_base + L.bits + (uint32_t)(((s * 2 + par) * 16 + word) * BLOCK_M + row) * 4u;
  };

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t csize;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int C = (int)csize;
  const int j = (int)cluster_ctarank();
  const int cluster_id = (int)blockIdx.x / C, n_clusters = (int)gridDim.x / C;
  // (scalars and selects instead of arrays indexed by the direction: run-time indexing would put them, and the
  //  whole argument struct, into local memory)
  const bool act0 = j < args.n_tiles[0], act1 = j < args.n_tiles[1];     // this CTA computes units of direction 0 / 1
  const int nkb0 = args.n_tiles[1], nkb1 = args.n_tiles[0];              // K blocks read by direction 0 / 1
  constexpr uint32_t kTmemCols = 512;   // 2 x 64 accumulator columns + 8 x 32 columns of A ring (power of two)
  const int n_pairs = ceil_div(ceil_div(args.row_blocks, args.slots), n_clusters);   // upper bound, same for every cluster

  if (warp_idx == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  if (warp_idx == 1 && lane == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < kResStages; ++s) { mbar_init(ring_full(s), 4); mbar_init(ring_empty(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tmem_full(s), 1); mbar_init(tmem_empty(s), kResEpiWarps);
      for (int par = 0; par < 2; ++par) {
        for (int kb = 0; kb < (par ? nkb1 : nkb0); ++kb) {
          mbar_init(bits_full(s, par, kb), 1);
          // armed for its first use: the two words per chain of K block kb
          mbar_arrive_expect_tx(bits_full(s, par, kb), 2u * BLOCK_M * 4u);
        }
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp_idx == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 256) {
    // byte pattern -> eight bf16 {0,1}: unit i of the byte is element i of the 16-byte chunk
    const uint32_t b = threadIdx.x;
    uint32_t e[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) e[p] = (((b >> (2 * p)) & 1u) ? 0x3F80u : 0u) | (((b >> (2 * p + 1)) & 1u) ? 0x3F800000u : 0u);
    st_shared_v4(smem_base + L.table + b * 16u, e[0], e[1], e[2], e[3]);
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // row block of (pair, slot) for this cluster, or -1
  auto row_block = [&](int pair, int s) {
    const int item = (pair * n_clusters + cluster_id) * args.slots + s;
    return (s < args.slots && item < args.row_blocks) ? item : -1;
  };

  if (warp_idx == 0) {
    // ============================ W slices: loaded once ============================
    if (lane == 0) {
      const uint32_t bytes = (act0 ? (uint32_t)nkb0 : 0u) * kResWBlockBytes + (act1 ? (uint32_t)nkb1 : 0u) * kResWBlockBytes;
      if (bytes != 0u) {
        mbar_arrive_expect_tx(w_full, bytes);
        if (act0)
          for (int kb = 0; kb < nkb0; ++kb)     // W[64 kb .. +64 (v, k)][64 j .. +64 (h, n)]
            tma_load_2d(smem_base + L.w0 + kb * kResWBlockBytes, &tm_w, w_full, j * kResBN, kb * BLOCK_K);
        if (act1)
          for (int kb = 0; kb < nkb1; ++kb)     // W[64 j .. +64 (v, n)][64 kb .. +64 (h, k)]
            tma_load_2d(smem_base + L.w1 + kb * kResWBlockBytes, &tm_w, w_full, kb * BLOCK_K, j * kResBN);
      } else {
        mbar_arrive(w_full);
      }
    }
    __syncwarp();
    for (int pair = 0; pair < n_pairs; ++pair) cluster_sync_all();
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ================================
    mbar_wait(w_full, 0);
    uint32_t ring_pos = 0;                  // K blocks issued so far: block i of the sequence uses stage i % 8, phase (i / 8) & 1
    uint32_t uses[2] = {0u, 0u};
    const uint64_t bdesc0_base = make_smem_desc(smem_base + L.w0, BLOCK_K * 128, 1024);  // dir 0: MN-major
    const uint64_t bdesc1_base = make_smem_desc(smem_base + L.w1, 16, 1024);             // dir 1: K-major
    // One half-step of one row block.  With 64-wide tiles a K block is only ~130 cycles of tensor time, so the issue loop
    // must not be longer: the direction is a compile-time constant and the loop over the (at most 8) K blocks is fully
    // unrolled, so every descriptor is a uniform base plus an immediate (descriptors advance by ADDING to the 14-bit
    // address field: all shared-memory addresses are below 256 KB, nothing carries out of it).  The rolled loop with
    // run-time strides took ~45 instructions / ~340 cycles per K block (13 vector -> uniform register moves) and was the
    // longest single stretch of a half-step (gpurun_out/res_trace_*_v6.txt).
    auto issue = [&](auto dir_c, int s, int t) {
      constexpr int DIR = decltype(dir_c)::value;
      constexpr uint32_t idesc = make_idesc(BLOCK_M, kResBN, DIR == 0);
      constexpr uint64_t bstep = DIR ? (uint64_t)((UMMA_K * 2) >> 4) : (uint64_t)((UMMA_K * 128) >> 4);
      const int nkb = DIR ? nkb1 : nkb0;
      const uint64_t bdesc = DIR ? bdesc1_base : bdesc0_base;
      mbar_wait_role<1>(tmem_empty(s), (uses[s] & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(s * kResAccCols);
      // K blocks are issued two at a time: the tensor pipe needs ~150 cycles for the four N = 64 MMAs of a K block
      // (benchmarks/probes/ts_mma_probe.cu: 32 of them complete in 1,200 cycles) but one trip through this loop (barrier
      // wait, elect, descriptors into uniform registers) costs ~190 cycles that do not overlap them, and the expanders
      // run well ahead of the MMAs anyway (two groups finish K blocks 0 and 1 together).
#pragma unroll
      for (int kb = 0; kb < kResMaxTiles; kb += 2) {
        if (kb < nkb) {
          const bool two = kb + 1 < nkb;
          const uint32_t idx = ring_pos + (uint32_t)kb;
          const uint32_t stage0 = idx & (uint32_t)(kResStages - 1), stage1 = (idx + 1u) & (uint32_t)(kResStages - 1);
          mbar_wait_role<1>(ring_full((int)stage0), (idx >> 3) & 1u);
          if (two) mbar_wait_role<1>(ring_full((int)stage1), ((idx + 1u) >> 3) & 1u);
          tcgen05_fence_after();
          RES_TRACE(32 * s + 10 + kb, t, lane == 0);
          if (elect_one_sync()) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 0 || two) {
                const uint32_t stage = h ? stage1 : stage0;
                const uint32_t a_tmem = tmem_base + (uint32_t)kResAColumn + stage * 32u;   // 4 x 8 columns: K = 16 each
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                  // (splitting the accumulate chain over two accumulators - even / odd K steps - bought nothing:
                  //  4.61 vs 4.45 us per half-step, profiles/r02_resident_vs_chain.txt)
                  umma_bf16_ts(tmem_d, a_tmem + (uint32_t)(k * 8),
                               bdesc + (uint64_t)((kb + h) * (kResWBlockBytes >> 4)) + (uint64_t)k * bstep, idesc,
                               (kb | h | k) != 0 ? 1u : 0u);
                umma_commit(ring_empty((int)stage));
              }
            }
          }
          __syncwarp();
        }
      }
      if (elect_one_sync()) umma_commit(tmem_full(s));
      __syncwarp();
      RES_TRACE(32 * s + 18, t, lane == 0);
      ring_pos += (uint32_t)nkb;
      ++uses[s];
    };
    for (int pair = 0; pair < n_pairs; ++pair) {
      for (int t = 0; t < args.steps; ++t) {
        const int dir = t & 1;
        if (dir ? act1 : act0) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            if (row_block(pair, s) < 0) continue;
            if (dir) issue(std::integral_constant<int, 1>{}, s, t);
            else issue(std::integral_constant<int, 0>{}, s, t);
          }
        }
      }
      cluster_sync_all();
    }
  } else if (warp_idx >= 4 && warp_idx < 4 + kResEpiWarps) {
    // ================================= epilogue =================================
    const int quarter = warp_idx & 3;
    const int cq = (warp_idx - 4) >> 2;
    const int row_in_blk = quarter * 32 + lane;
    uint32_t uses[2] = {0u, 0u};
    for (int pair = 0; pair < n_pairs; ++pair) {
      float wsum[2] = {0.f, 0.f};
      for (int t = 0; t < args.steps; ++t) {
        const int dir = t & 1;
        if (dir ? act1 : act0) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int rb = row_block(pair, s);
            if (rb < 0) continue;
            const int row = rb * BLOCK_M + row_in_blk;
            const uint32_t chain = (uint32_t)(args.chain_offset + (uint64_t)row);
            const int col0 = j * kResBN + cq * 32;
            EpiParams ep;
            ep.nbias = dir ? args.nbias[1] : args.nbias[0]; ep.ldo = dir ? args.ld[1] : args.ld[0];
            ep.n_valid = dir ? args.n_valid[1] : args.n_valid[0];
            ep.ctx = make_draw_ctx(args.seed, args.step_offset + (uint64_t)t, args.tag);
            ep.scale = -1.4426950408889634f; ep.bmul = 1.f; ep.ratio = 0.f; ep.debug_mode = 0;
            ep.kf_magic = args.kf_magic;
            if (EPI0 != EPI_GIBBS) {
              const float bk = __ldg(args.betas + 1 + (t >> 1)), bkm1 = __ldg(args.betas + (t >> 1));
              ep.scale = 0.5f * (-1.4426950408889634f * bk);     // accumulators are doubled (A = 2.0 per set unit)
              ep.bmul = dir == 0 ? bk : 1.f;                     // (EPI_ANNEALED = direction 1 here: bias multiplier 1)
              ep.ratio = bkm1 / bk;
            }
            const uint32_t par_full = uses[s] & 1u;
            auto wait_full = [&] { role_wait(tmem_full(s), par_full, warp_idx == 4, 1, 32 * kResEpiWarps); RES_TRACE(32 * s + 20 + 4 * (warp_idx == 11), t, lane == 0 && (warp_idx == 4 || warp_idx == 11)); };
            auto release = [&] { mbar_arrive(tmem_empty(s)); RES_TRACE(32 * s + 21 + 4 * (warp_idx == 11), t, (warp_idx == 4 || warp_idx == 11)); };
            const uint32_t tmem_acc = tmem_base + (uint32_t)(s * kResAccCols);
            uint32_t bits;
            float wstep = 0.f;      // this half-step's 32-unit sum is added as ONE term, like the chain kernel's red
            if (dir == 0) bits = resident_epilogue<EPI0>(ep, tmem_acc, col0, chain, quarter, cq, lane, wstep, wait_full, release);
            else bits = resident_epilogue<EPI1>(ep, tmem_acc, col0, chain, quarter, cq, lane, wstep, wait_full, release);
            wsum[s] += wstep;
            __syncwarp();
            RES_TRACE(32 * s + 22 + 4 * (warp_idx == 11), t, lane == 0 && (warp_idx == 4 || warp_idx == 11));
            if (t + 1 < args.steps) {
              // The CTA's words of this half-step (128 chains x 2 words = 1 KB, laid out like the consumer's buffer) are
              // staged in local shared memory by the eight epilogue warps, which meet on a named barrier; ONE warp then
              // issues one 1 KB bulk copy per destination CTA (one complete_tx per copy on the destination's barrier for
              // this source CTA).  (128-byte copies per warp: 64 instead of 8 copies per CTA and half-step, ~100 issued
              // instructions per warp; per-lane st.async: 256 barrier updates per CTA pair and half-step.)
              // The staging block alternates with the use count: the copies issued two uses ago have been consumed.
              const int par = (t + 1) & 1;
              const uint32_t stg = smem_base + L.stage + (uint32_t)((s * 2 + (int)(uses[s] & 1u)) * 1024);
              st_shared_u32(stg + (uint32_t)(cq * 512 + row_in_blk * 4), bits);
              fence_proxy_async_smem();
              named_bar_sync(3, 32 * kResEpiWarps);
              if (warp_idx == 4) {
                const uint32_t dst = bits_addr(s, par, 2 * j, 0);
                const uint32_t bar = bits_full(s, par, j);      // this CTA is K block j of its consumers
                const int n_dst = dir ? args.n_tiles[0] : args.n_tiles[1];
                if (elect_one_sync()) {
#pragma unroll
                  for (int d = 0; d < kResMaxTiles; ++d)
                    if (d < n_dst) dsmem_bulk_copy(mapa_u32(dst, (uint32_t)d), stg, 1024u, mapa_u32(bar, (uint32_t)d));
                }
                __syncwarp();
              }
              RES_TRACE(32 * s + 23 + 4 * (warp_idx == 11), t, lane == 0 && (warp_idx == 4 || warp_idx == 11));
            }
            ++uses[s];
            if (t + 2 >= args.steps) {
              // final (v, h) of the call: bf16 {0,1} into the state tiles the rest of the library reads
              uint4* o = reinterpret_cast<uint4*>((dir ? args.out[1] : args.out[0]) + (size_t)row * (dir ? args.ld[1] : args.ld[0]) + col0);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t by = (bits >> (8 * c)) & 0xFFu;
                o[c] = ld_shared_v4(smem_base + L.table + by * 16u);
              }
            }
          }
        }
      }
      if (EPI0 == EPI_AIS_WEIGHT && act0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int rb = row_block(pair, s);
          if (rb < 0) continue;
          const int row = rb * BLOCK_M + row_in_blk;
          atomicAdd(args.wpart + (size_t)row * args.wpart_ld + ((j * kResBN + cq * 32) >> 5), wsum[s]);
        }
      }
      __syncwarp();
      cluster_sync_all();
    }
  } else if (warp_idx >= 4 + kResEpiWarps) {
    // ================================ expanders ================================
    const int e = warp_idx - 4 - kResEpiWarps;
    const int row = (e & 3) * 32 + lane;       // chain within the row block
    const int half = e >> 2;                   // prologue (warps e < 8): which 32 units of this CTA's 64
    const int grp = e >> 2;                    // main loop: expander group
    uint32_t ring_pos = 0;                     // K blocks expanded so far: block i of the ring sequence uses stage i % 8, phase (i / 8) & 1
    uint32_t bphase = 0;                       // bit (2 s + par): phase parity of bits_full(s, par, *) to wait for
    for (int pair = 0; pair < n_pairs; ++pair) {
      // prologue: this CTA's 64 units of the initial visible state -> words for the CTAs that compute direction 0
      if (act1 && e < 8) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int rb = row_block(pair, s);
          if (rb < 0) continue;
          const uint4* src = reinterpret_cast<const uint4*>(args.v_init + (size_t)(rb * BLOCK_M + row) * args.ld[1] + j * kResBN + half * 32);
          uint32_t bits = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 q = __ldg(src + c);
            const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              if (wds[p] & 0x0000FFFFu) bits |= 1u << (8 * c + 2 * p);
              if (wds[p] & 0xFFFF0000u) bits |= 1u << (8 * c + 2 * p + 1);
            }
          }
          const uint32_t stg = smem_base + L.pro + (uint32_t)((s * 8 + e) * 128);
          st_shared_u32(stg + (uint32_t)lane * 4u, bits);
          fence_proxy_async_smem();
          __syncwarp();
          {
            const uint32_t dst = bits_addr(s, 0, 2 * j + half, (e & 3) * 32);
            const uint32_t bar = bits_full(s, 0, j);
            if (lane < args.n_tiles[0]) dsmem_bulk_copy(mapa_u32(dst, (uint32_t)lane), stg, 128u, mapa_u32(bar, (uint32_t)lane));
          }
          __syncwarp();
        }
      }
      for (int t = 0; t < args.steps; ++t) {
        const int dir = t & 1, par = t & 1;
        const int nkb = dir ? nkb1 : nkb0;
        if (dir ? act1 : act0) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            if (row_block(pair, s) < 0) continue;
            const uint32_t bit = 1u << (2 * s + par);
            // The ring slots of this half-step (nkb <= 8 distinct stages) are normally free long before its words arrive:
            // wait for this group's slots first, off the critical path.
            for (int kb = grp; kb < nkb; kb += kResExpGroups) {
              const uint32_t idx = ring_pos + (uint32_t)kb;
              mbar_wait_role<2>(ring_empty((int)(idx & (kResStages - 1))), ((idx >> 3) & 1u) ^ 1u);
            }
            const uint32_t bpar = (bphase & bit) ? 1u : 0u;
            bphase ^= bit;
            tcgen05_fence_after();          // the MMAs that read these stages before have completed (commit -> ring_empty)
            // The groups of four warps expand alternate K blocks concurrently; a warp covers 32 chains x 64 units = both
            // words of its rows.
            for (int kb = grp; kb < nkb; kb += kResExpGroups) {
              mbar_wait_role<2>(bits_full(s, par, kb), bpar);
              RES_TRACE(32 * s + 0, t, e == 0 && lane == 0 && kb == 0);
              // re-arm for the next use of this (slot, parity, K block): the next words cannot arrive before every warp of
              // this CTA has passed this wait (they depend on this half-step's accumulator)
              if ((e & 3) == 0 && lane == 0) mbar_arrive_expect_tx(bits_full(s, par, kb), 2u * BLOCK_M * 4u);
              const uint32_t word0 = ld_shared_u32(bits_addr(s, par, 2 * kb, row));
              const uint32_t word1 = ld_shared_u32(bits_addr(s, par, 2 * kb + 1, row));
              const int stage = (int)((ring_pos + (uint32_t)kb) & (kResStages - 1));
              uint32_t o[32];             // column c of the stage = units 2c, 2c + 1 of the K block
#pragma unroll
              for (int hw = 0; hw < 2; ++hw) {
                // 32 units -> 32 bf16 {0, 2} in registers, ONE ALU instruction per unit.  A set unit becomes 2.0 = 0x4000:
                // a single bit of the high byte, so m[i] = (word << (6 - i)) & 0x40404040 holds the high bytes of units
                // 8c + i (c = 0..3) and one PRMT per bf16 pair picks them (selector nibbles with bit 3 replicate the sign
                // bit of a byte whose top bit is clear: the zero low bytes).  The epilogue halves its scale - exact.
                const uint32_t word = hw ? word1 : word0;
                uint32_t m[8];
#pragma unroll
                for (int i = 0; i < 7; ++i) m[i] = (word << (6 - i)) & 0x40404040u;
                m[7] = (word >> 1) & 0x40404040u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
#pragma unroll
                  for (int pp = 0; pp < 4; ++pp) {
                    const uint32_t sel = (uint32_t)((8 | c) | (c << 4) | ((12 | c) << 8) | ((4 | c) << 12));
                    o[hw * 16 + c * 4 + pp] = prmt_b32(m[2 * pp], m[2 * pp + 1], sel);
                  }
                }
              }
              tmem_st32(tmem_base + ((uint32_t)((e & 3) * 32) << 16) + (uint32_t)(kResAColumn + stage * 32), o);
              tmem_st_wait();
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(ring_full(stage));
              RES_TRACE(32 * s + 1 + kb, t, (e & 3) == 0 && lane == 0);
            }
            ring_pos += (uint32_t)nkb;
          }
        }
      }
      __syncwarp();
      cluster_sync_all();
    }
  } else {
    // warps 2, 3: only the pair barriers
    for (int pair = 0; pair < n_pairs; ++pair) cluster_sync_all();
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp_idx == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

inline int res_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Shapes the resident kernel covers: both sides fit 8 tiles of 64 units (W slices 2 x 64 KB at most).
inline bool resident_supported(int nVp, int nHp) { return nVp <= kResMaxTiles * kResBN && nHp <= kResMaxTiles * kResBN; }

// Clusters of `cluster` CTAs of the resident kernel that fit on the device at once (GPC boundaries make this less than
// sm_count / cluster: 15 clusters of 8 on a 148-SM B200).  Queried once per (flavour, cluster size, shared-memory size).
template <int EPI0, int EPI1>
int resident_max_clusters(int cluster, int smem_bytes, int sm_count, int* out) {
  static int cache[64][9][8] = {};   // [device][cluster size][smem size in 32 KB steps], 0 = not queried
  static std::mutex mu;
  auto kern = umma_resident_kernel<EPI0, EPI1>;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if ((rc = ensure_dynamic_smem(kern, di.device, smem_bytes))) return rc;
  std::lock_guard<std::mutex> lock(mu);
  int& slot = cache[di.device & 63][cluster][std::min(7, smem_bytes >> 15)];
  if (slot == 0) {
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kResThreads);
    cfg.gridDim = dim3((unsigned)cluster);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      n = std::max(1, sm_count / cluster / 2);
    }
    slot = n;
  }
  *out = slot;
  return 0;
}

inline int resident_cluster_size(int tiles_h, int tiles_v) {
  const int tiles = std::max(tiles_h, tiles_v);
  return tiles <= 1 ? 1 : tiles <= 2 ? 2 : tiles <= 4 ? 4 : 8;
}

template <int EPI0, int EPI1>
int launch_resident(const CUtensorMap& tm_w, ResidentArgs a, int cluster, int sm_count, cudaStream_t st) {
  auto kern = umma_resident_kernel<EPI0, EPI1>;
  const ResidentSmem L = resident_smem(a.n_tiles[0], a.n_tiles[1]);
  int max_clusters = 0;
  int rc = resident_max_clusters<EPI0, EPI1>(cluster, L.total, sm_count, &max_clusters);
  if (rc) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kResThreads);
  cfg.dynamicSmemBytes = (size_t)L.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int force_slots = res_env_int("RBM_B200_RESIDENT_SLOTS", 0);
  // one row block per cluster while there are clusters to spare (latency), two in ping-pong otherwise (throughput)
  a.slots = force_slots == 1 || force_slots == 2 ? force_slots : (a.row_blocks > max_clusters ? 2 : 1);
  const int items = ceil_div(a.row_blocks, a.slots);
  const int rounds = ceil_div(items, max_clusters);
  const int n_clusters = ceil_div(items, rounds);
  cfg.gridDim = dim3((unsigned)(n_clusters * cluster));
  RBM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm_w, a));
  return check_launch("umma_resident_kernel");
}

}  // namespace
}  // namespace rbm
