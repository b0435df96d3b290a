// Chain kernel: ALL half-steps of a block-Gibbs (or AIS) call in ONE persistent launch, scheduled as a dataflow.
//
// A work item is one output tile of one half-step: (super-block, t, row group, n tile).  Items are numbered in that
// order and dealt round-robin to the resident CTAs (or CTA pairs), which run the usual warp-specialised pipeline over
// them (TMA producer -> tcgen05.mma -> fused sampling epilogue -> TMA store; two TMEM accumulator stages, so the
// epilogue of an item overlaps the MMAs of the next one).
//
// Dependencies.  Chains are independent, so item (t, m, n) depends only on the items (t-1, m, *) that wrote row block m
// of its A operand - and on them chunk by chunk: the epilogue finishes a tile in CHUNKS of 32 * kSets columns
// (umma_common.cuh: epilogue_tile), every (direction, 128-chain row block, chunk) has an arrival counter in global
// memory, an epilogue warp adds 1 (red.release.gpu) once its box of a chunk is complete in global memory, and the
// consumer reads the K blocks of its A operand in the order in which its producers finish them (first chunks of all
// producer tiles, then second chunks): its MMAs start while the producers are still sampling their last columns.
// A dependency warp (warp 3) runs ahead through the item list, polls the counters (ld.acquire.gpu, up to 32 chunks
// per round trip) and publishes the number of satisfied chunks in shared memory, so the TMA producer never waits on
// a global round trip; W (the B operand) does not depend on anything and is fetched before the wait.
// There is no relaunch, no grid-wide or cluster barrier, and no wave quantisation: a CTA that finishes its tile of
// half-step t early starts on a tile of half-step t+1 whose row block is already complete.  Deadlock-free because
// every item only waits for lower-numbered items and every CTA takes its items in increasing order (all CTAs are
// co-resident: grid <= number of SMs, 1 CTA per SM).  Write-after-read on the two state buffers is implied: an item
// stores only after its whole K loop, i.e. after every chunk of (t-1, m, *) has been stored, i.e. after those items'
// MMAs - the readers of the buffer it overwrites - have completed.
//
// Super-blocks.  With many chains the two state tiles exceed the L2 and every half-step would round-trip through HBM
// (230 MB per half-step at BASELINE config 4; the chip runs at its power cap, so those bytes cost throughput).  The
// item order therefore runs ALL half-steps for one super-block of chains whose state tiles fit in L2 before it moves
// to the next one (the first half-step of a super-block has no dependencies, so there is no bubble at the seam).
//
// Replaces the per-half-step launches of round 1 (2k launches per call, each with its own tail and a partial last
// wave: 256 tiles on 148 SMs at BASELINE config 4 sharded over 8 GPUs) and the cluster-per-row-block multi-step
// kernel (one cluster barrier per half-step, one TMEM stage, SMs idle when row blocks x tiles < 148).
#pragma once
#include "umma_common.cuh"

namespace rbm {
namespace {

struct ChainArgs {
  const float* nbias[2];   // [dir]: dir 0 = v->h (hidden biases), dir 1 = h->v; -log2(e) * bias, zero padded
  int ldo[2];              // padded unit count of the state WRITTEN by dir
  int n_valid[2];          // real unit count of the state written by dir
  int n_tiles[2];          // output tiles per row group
  int k_blocks[2];         // K / 64 of the state READ by dir (one pass; split weights run two passes)
  int m_groups;            // row groups of CTAS * 128 chains
  int sb_groups;           // row groups per super-block
  int steps;               // half-steps in this launch
  int first_dir;
  uint64_t seed, chain_offset, step_offset;   // half-step t draws with counter step_offset + t
  uint32_t tag;
  const float* betas;      // annealed flavours: beta index of half-step t is 1 + t/2
  float* wpart;            // AIS: [rows, wpart_ld] log-weight partials (log2 units), one slot per 32 hidden units
  int wpart_ld;
  uint32_t kf_magic;       // kKfMagic (see EpiParams)
  uint32_t* flags;         // [2 dirs][row blocks][chunks_ld] arrival counters, zeroed before the launch
  int row_blocks;          // m_groups * CTAS
  int chunks_ld;
  int opts;                // tuning switches (CHAIN_OPT_*), see run_chain
};
enum { CHAIN_OPT_DEPWARP = 1,     // a dedicated warp polls the counters ahead of the producer
       CHAIN_OPT_MIDARRIVE = 2,   // publish a chunk half-way through the next group (else: before the next group's store)
       CHAIN_OPT_CHUNKED = 4,     // wait chunk by chunk (else: for the whole row block before the first state load)
       CHAIN_OPT_NO_WFENCE = 8,   // no fence.proxy.async on the writer side (wait_group already orders the stores)
       CHAIN_OPT_TILE_ARRIVE = 16,  // publish all chunks of a tile together when the tile is done
       CHAIN_OPT_ACQ_SPIN = 32,     // spin with ld.acquire (else relaxed loads + one final acquire)
       CHAIN_OPT_DEBUG_NOWAIT = 64, // TIMING EXPERIMENT ONLY (results invalid): publish without waiting for the stores
       CHAIN_OPT_NO_DEFER = 128,    // always publish a tile's last chunk at the accumulator wait (never defer it)
       CHAIN_OPT_CTA_ARRIVE = 256 };  // the epilogue warps of a CTA meet on a shared-memory counter and the LAST one publishes
                                      // the chunk for all of them: one gpu-scope release per chunk and CTA instead of one per warp

// Position in the item sequence of one cluster: items c, c + G, c + 2G, ... of the (super-block, t, tile) enumeration.
struct ItemCursor {
  int sb, t, dir, tile;
  __device__ __forceinline__ int groups_in_sb(const ChainArgs& a) const {
    const int left = a.m_groups - sb * a.sb_groups;
    return left < a.sb_groups ? left : a.sb_groups;
  }
  __device__ __forceinline__ void normalise(const ChainArgs& a) {
    while (sb * a.sb_groups < a.m_groups && tile >= groups_in_sb(a) * a.n_tiles[dir]) {
      tile -= groups_in_sb(a) * a.n_tiles[dir];
      ++t; dir ^= 1;
      if (t == a.steps) { t = 0; dir = a.first_dir & 1; ++sb; }
    }
  }
  __device__ __forceinline__ void init(const ChainArgs& a, int first) {
    sb = 0; t = 0; dir = a.first_dir & 1; tile = first;
    normalise(a);
  }
  __device__ __forceinline__ void advance(const ChainArgs& a, int stride) { tile += stride; normalise(a); }
  __device__ __forceinline__ bool valid(const ChainArgs& a) const { return sb * a.sb_groups < a.m_groups; }
  __device__ __forceinline__ int m_grp(const ChainArgs& a) const { return sb * a.sb_groups + tile / a.n_tiles[dir]; }
  __device__ __forceinline__ int n_blk(const ChainArgs& a) const { return tile % a.n_tiles[dir]; }
  // arrivals per warp-slot a chunk of the state READ at half-step t must have collected: the number of earlier
  // half-steps of this launch that wrote that state (direction dir ^ 1)
  __device__ __forceinline__ uint32_t writes_before(const ChainArgs& a) const {
    return (uint32_t)((t + (((dir ^ 1) == (a.first_dir & 1)) ? 1 : 0)) >> 1);
  }
};

// Order in which an item reads the K blocks of its A operand: chunk round g of every producer tile before round g+1.
//   BLOCK_N-wide producer tiles, kGroups chunk rounds per tile, kCB K blocks (of 64) per chunk
template <int BLOCK_N>
struct KOrder {
  static constexpr int kCW = 32 * EpiSplit<BLOCK_N>::kSets;    // chunk width in columns (64 or 128)
  static constexpr int kCB = kCW / BLOCK_K;                    // K blocks per chunk
  static constexpr int kGroups = BLOCK_N / kCW;                // chunk rounds per producer tile
  static constexpr int kTB = BLOCK_N / BLOCK_K;                // K blocks per producer tile
  // slot q of an item = (round g, producer tile pt); its chunk index within the row block
  __device__ __forceinline__ static int slots(int nk) { return ceil_div(nk, kTB) * kGroups; }
  __device__ __forceinline__ static int chunk_of_slot(int q, int nk) {
    const int nt = ceil_div(nk, kTB);
    const int g = q / nt, pt = q % nt;
    return pt * kGroups + g;
  }
  __device__ __forceinline__ static bool slot_exists(int q, int nk) { return chunk_of_slot(q, nk) * kCB < nk; }
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin with relaxed loads (no L1 invalidation per iteration), then ONE acquire once the count is there
__device__ __forceinline__ void wait_counter(const uint32_t* p, uint32_t need, bool acq_spin) {
  if (acq_spin) {
    while (ld_acquire_gpu(p) < need) { }
  } else {
    while (ld_relaxed_gpu(p) < need) { __nanosleep(64); }
    (void)ld_acquire_gpu(p);
  }
}
__device__ __forceinline__ void red_relaxed_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Epilogue hooks of the chain kernel: publish every finished chunk of the new state.
// A chunk is published by waiting for the warp's TMA stores to be COMPLETE in global memory (~1 us after the issue)
// and adding 1 to its counter.  The wait is placed where the store is already old: the previous group of the same
// tile is published while the next group is computed; the LAST group of a tile is published while the warp waits
// for its next accumulator - or, if that accumulator is already there (many items per CTA: the consumers are a whole
// half-step away), half-way through the next tile's first group, so the epilogue never idles on a store.
struct ChainHooks {
  uint32_t* pending;       // first counter of the chunks whose boxes this warp has handed to TMA but not yet published
  int n_pending;
  uint32_t pending_slot;   // CTA_ARRIVE: shared-memory counter of the pending publication
  uint32_t* item_flags;    // counters of this item's first chunk
  int cur_item;            // per-CTA index of the item item_flags belongs to
  int lane, n_groups;
  bool at_mid, tile_arrive, wfence, nowait, cta_arrive;
  uint32_t cnt_base;       // CTA_ARRIVE: 8 shared-memory counters (4 items x 2 chunks in flight at most)
  uint32_t arrivers;       // epilogue warps of this CTA that publish every chunk
  bool carried;            // `pending` belongs to the previous item (deferred at the accumulator wait)
  __device__ __forceinline__ void arrive() {
    if (pending != nullptr) {
      __syncwarp();                            // the other lanes' log-weight reds are ordered before the release
      if (lane == 0) {
        if (!nowait) tma_store_wait_all();     // this warp's boxes of the new state are complete in global memory
        if (wfence) asm volatile("fence.proxy.async;" ::: "memory");
        bool publish = true;
        if (cta_arrive) {
          // release this warp's completed stores to the CTA, acquire the earlier arrivers': the last warp's gpu-scope
          // release below then covers the boxes of all of them (release/acquire chains are cumulative)
          uint32_t old;
          const uint32_t addr = cnt_base + 4u * pending_slot;
          asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
          publish = old == arrivers - 1u;
          if (publish) asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(0u) : "memory");
        }
        if (publish) {
          if (n_pending == 1) {
            red_release_gpu_add(pending, 1u);
          } else {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            for (int i = 0; i < n_pending; ++i) red_relaxed_gpu_add(pending + i, 1u);
          }
        }
      }
      pending = nullptr;
    }
    carried = false;
  }
  // before the wait for the next accumulator: publish now if there is time to kill, else defer
  __device__ __forceinline__ void at_acc_wait(bool acc_ready) {
    if (pending != nullptr) {
      if (acc_ready) carried = true; else arrive();
    }
  }
  // half-way through the next group the previous group's store has had > 1 us to complete: publish it
  __device__ __forceinline__ void mid(int) { if (pending != nullptr && (carried || (at_mid && !tile_arrive))) arrive(); }
  // (else the previous group is published right before this group's store is issued: its own store is ~2 us old)
  __device__ __forceinline__ void before_store(int) { if (pending != nullptr && (carried || (!at_mid && !tile_arrive))) arrive(); }
  __device__ __forceinline__ void stored(int g) {
    if (tile_arrive) {
      if (pending != nullptr && pending != item_flags) arrive();   // (a carried tile is normally published at mid(0))
      pending = item_flags; n_pending = g + 1;
      pending_slot = (uint32_t)((cur_item & 3) * 2);
    } else {
      // groups beyond the padded unit count skip before_store(): never overwrite an unpublished chunk
      if (pending != nullptr) arrive();
      pending = item_flags + g; n_pending = 1;
      pending_slot = (uint32_t)((cur_item & 3) * 2 + g);
    }
  }
};

// one non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

// SPLITW: W = W_high + W_low (both bf16).  The GEMM runs over a doubled K axis, [x | x] . [W_high ; W_low]: the
// producer reads the state tile twice and the low part of W through a second tensor map.  States are {0,1}, so the
// products are exact and the pre-activation carries W to ~2^-17 relative instead of 2^-9.
template <int BLOCK_N, int EPI0, int EPI1, int CTAS, bool SPLITW>
__global__ void __launch_bounds__(kNumThreads, 1)
umma_chain_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                  const __grid_constant__ CUtensorMap tm_b0, const __grid_constant__ CUtensorMap tm_b1,
                  const __grid_constant__ CUtensorMap tm_b0_lo, const __grid_constant__ CUtensorMap tm_b1_lo,
                  const __grid_constant__ CUtensorMap tm_o0, const __grid_constant__ CUtensorMap tm_o1,
                  const ChainArgs args) {
  using L = SmemLayout<BLOCK_N, CTAS>;
  using ES = EpiSplit<BLOCK_N>;
  using KO = KOrder<BLOCK_N>;
  constexpr int kStages = L::kStages;
  constexpr int kBCols = BLOCK_N / CTAS;        // columns of the W tile this CTA stages
  constexpr int kArr = ES::kWarps;              // arrivals per chunk and half-step
  constexpr int kPasses = SPLITW ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOffset + 8 * (2 * kStages + 4));
  // number of dependency slots (in item order) the dependency warp has seen satisfied
  const uint32_t ready_seq = bar_base + 8u * (2 * kStages + 4) + 8u;
  const uint32_t arrive_cnt = ready_seq + 8u;      // CTA_ARRIVE: 8 counters of 4 bytes
  const bool cta_arrive = (args.opts & CHAIN_OPT_CTA_ARRIVE) != 0;
  const uint32_t arrivals_per_write = cta_arrive ? 1u : (uint32_t)kArr;   // counter increments per chunk and half-step

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int first_item = (int)blockIdx.x / CTAS, item_stride = (int)gridDim.x / CTAS;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;   // two accumulator stages (power of two)

  if (warp_idx == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_o1) : "memory");
    if (SPLITW) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b0_lo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b1_lo) : "memory");
    }
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), CTAS); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), kArr * CTAS); }
    st_release_cta_shared(ready_seq, 0u);
    for (int i = 0; i < 8; ++i) st_release_cta_shared(arrive_cnt + 4u * i, 0u);
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
    // whole warp, warp-uniform; one elected lane issues the copies
    int stage = 0; uint32_t phase = 0;
    uint32_t seq_base = 0;      // dependency slots of the items before this one
    uint32_t seen = 0;          // last value read from ready_seq
    const bool use_depwarp = (args.opts & CHAIN_OPT_DEPWARP) != 0, chunked = (args.opts & CHAIN_OPT_CHUNKED) != 0;
    const bool acq_spin = (args.opts & CHAIN_OPT_ACQ_SPIN) != 0;
    ItemCursor it;
    for (it.init(args, first_item); it.valid(args); it.advance(args, item_stride)) {
      const int dir = it.dir;
      const int n_blk = it.n_blk(args);
      const int row_blk = it.m_grp(args) * CTAS + (int)cta_rank;
      const int row0 = row_blk * BLOCK_M;                                 // this CTA's 128 chains
      const int ncol0 = n_blk * BLOCK_N + (int)cta_rank * kBCols;         // this CTA's share of the W tile
      const CUtensorMap* tm_a = dir ? &tm_a1 : &tm_a0;
      const int nk = args.k_blocks[dir];
      const int nslots = KO::slots(nk);
      const int nt_prod = ceil_div(nk, KO::kTB);
      const uint32_t need = it.writes_before(args) * arrivals_per_write;
      const uint32_t* dep_flags = args.flags + ((size_t)(dir ^ 1) * args.row_blocks + row_blk) * args.chunks_ld;
      const uint32_t lead_full_base = CTAS == 2 ? mapa_u32(full_bar(0), 0) : 0u;
      auto issue_b = [&](int kb, bool lo, int s) {
        // arm the barrier and fetch the W block: dir 0 reads W[k rows][n cols] as [64 x 64] MN-major boxes,
        // dir 1 reads W[n rows][k cols] as one [kBCols x 64] K-major box
        const uint32_t sb = smem_base + s * L::kStageBytes + kABytes;
        const CUtensorMap* tm_b = dir ? (lo ? &tm_b1_lo : &tm_b1) : (lo ? &tm_b0_lo : &tm_b0);
        if (CTAS == 1) {
          mbar_arrive_expect_tx(full_bar(s), L::kStageBytes);
          if (dir == 0) {
#pragma unroll
            for (int nb = 0; nb < kBCols / 64; ++nb)
              tma_load_2d(sb + nb * (BLOCK_K * 128), tm_b, full_bar(s), ncol0 + nb * 64, kb * BLOCK_K);
          } else {
            tma_load_2d(sb, tm_b, full_bar(s), kb * BLOCK_K, ncol0);
          }
        } else {
          // both CTAs credit the LEADER's full barrier (2 arrivals + the bytes of both CTAs)
          const uint32_t lead_full = lead_full_base + 8u * s;
          if (is_leader) mbar_arrive_expect_tx(full_bar(s), 2 * L::kStageBytes);
          else mbar_arrive_cluster(lead_full);
          if (dir == 0) {
#pragma unroll
            for (int nb = 0; nb < kBCols / 64; ++nb)
              tma_load_2d_pair(sb + nb * (BLOCK_K * 128), tm_b, lead_full, ncol0 + nb * 64, kb * BLOCK_K);
          } else {
            tma_load_2d_pair(sb, tm_b, lead_full, kb * BLOCK_K, ncol0);
          }
        }
      };
      auto issue_a = [&](int kb, int s) {
        const uint32_t sa = smem_base + s * L::kStageBytes;
        if (CTAS == 1) tma_load_2d(sa, tm_a, full_bar(s), kb * BLOCK_K, row0);
        else tma_load_2d_pair(sa, tm_a, lead_full_base + 8u * s, kb * BLOCK_K, row0);
      };
      // K blocks in the order their producers finish them (KOrder): i-th block of a pass -> (slot, kb)
      int pre_left = kStages < nk * kPasses ? kStages : nk * kPasses;    // W blocks fetched ahead of the dependency wait
      {
        // 1. W blocks of the first stages: no dependency, issued before any wait
        int s = stage; uint32_t ph = phase; int left = pre_left;
        for (int pass = 0; pass < kPasses && left > 0; ++pass)
          for (int q = 0; q < nslots && left > 0; ++q) {
            const int kb0 = (q % nt_prod) * KO::kTB + (q / nt_prod) * KO::kCB;
            for (int j = 0; j < KO::kCB && left > 0; ++j) {
              const int kb = kb0 + j;
              if (kb >= nk) break;
              mbar_wait(empty_bar(s), ph ^ 1u);
              if (elect_one_sync()) issue_b(kb, pass == 1, s);
              __syncwarp();
              if (++s == kStages) { s = 0; ph ^= 1u; }
              --left;
            }
          }
      }
      // 2. state blocks, each chunk as soon as the dependency warp has seen it complete; then the remaining W blocks
      for (int pass = 0; pass < kPasses; ++pass)
        for (int q = 0; q < nslots; ++q) {
          const int kb0 = (q % nt_prod) * KO::kTB + (q / nt_prod) * KO::kCB;
          if (kb0 >= nk) continue;
          if (pass == 0) {
            if (use_depwarp) {
              // all slots up to the last one when the dependencies are not chunked
              const uint32_t want = seq_base + (chunked ? (uint32_t)q + 1u : (uint32_t)nslots);
              if (seen < want) {
                do { seen = ld_acquire_cta_shared(ready_seq); } while (seen < want);
                asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy acquire -> async-proxy (TMA) reads
              }
            } else if (need != 0u && (chunked || q == 0)) {
              // the producer polls the counters itself
              const int q_end = chunked ? q + 1 : nslots;
              for (int qq = q; qq < q_end; ++qq)
                if (KO::slot_exists(qq, nk)) wait_counter(dep_flags + KO::chunk_of_slot(qq, nk), need, acq_spin);
              asm volatile("fence.proxy.async.global;" ::: "memory");
            }
          }
          for (int j = 0; j < KO::kCB; ++j) {
            const int kb = kb0 + j;
            if (kb >= nk) break;
            const bool armed = pre_left > 0;
            if (!armed) mbar_wait(empty_bar(stage), phase ^ 1u);
            if (elect_one_sync()) {
              if (!armed) issue_b(kb, pass == 1, stage);
              issue_a(kb, stage);
            }
            __syncwarp();
            if (armed) --pre_left;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      seq_base += (uint32_t)nslots;
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ================================
    // whole warp runs the loop (everything stays in uniform registers); one elected lane issues
    if (is_leader) {
      constexpr uint32_t idesc0 = make_idesc(BLOCK_M * CTAS, BLOCK_N, true);    // dir 0: B MN-major
      constexpr uint32_t idesc1 = make_idesc(BLOCK_M * CTAS, BLOCK_N, false);   // dir 1: B K-major
      int stage = 0; uint32_t phase = 0;
      int n = 0;
      ItemCursor it;
      for (it.init(args, first_item); it.valid(args); it.advance(args, item_stride), ++n) {
        const int dir = it.dir;
        const int as = n & 1;
        mbar_wait(tmem_empty_bar(as), (((uint32_t)n >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        const uint32_t idesc = dir ? idesc1 : idesc0;
        const int nk = args.k_blocks[dir] * kPasses;
        for (int kb = 0; kb < nk; ++kb) {
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
              if (CTAS == 2) umma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if (CTAS == 2) umma_commit_pair(empty_bar(stage)); else umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (elect_one_sync()) {
          if (CTAS == 2) umma_commit_pair(tmem_full_bar(as)); else umma_commit(tmem_full_bar(as));
        }
        __syncwarp();
      }
    }
  } else if (warp_idx == 3 && (args.opts & CHAIN_OPT_DEPWARP) != 0) {
    // ============================== dependency warp ==============================
    // Runs ahead of the producer through this CTA's items: lane l polls the counter of the l-th dependency slot
    // of a batch; the count of slots satisfied IN ORDER is published in shared memory.
    uint32_t seq_base = 0;
    ItemCursor it;
    for (it.init(args, first_item); it.valid(args); it.advance(args, item_stride)) {
      const int dir = it.dir;
      const int row_blk = it.m_grp(args) * CTAS + (int)cta_rank;
      const int nk = args.k_blocks[dir];
      const int nslots = KO::slots(nk);
      const uint32_t need = it.writes_before(args) * arrivals_per_write;
      // counters of the state this item READS: written by direction dir ^ 1
      const uint32_t* flags = args.flags + ((size_t)(dir ^ 1) * args.row_blocks + row_blk) * args.chunks_ld;
      for (int q0 = 0; q0 < nslots; q0 += 32) {
        const int q = q0 + lane;
        const int batch = nslots - q0 < 32 ? nslots - q0 : 32;
        bool ok = q >= nslots || need == 0u || !KO::slot_exists(q, nk);
        const uint32_t* f = flags + (ok ? 0 : KO::chunk_of_slot(q, nk));
        int published = 0;
        while (true) {
          if (!ok) {
            ok = ld_relaxed_gpu(f) >= need;
            if (ok) (void)ld_acquire_gpu(f); else __nanosleep(64);
          }
          const uint32_t mask = __ballot_sync(0xffffffffu, ok);
          const int ready = mask == 0xffffffffu ? 32 : __ffs((int)~mask) - 1;      // leading satisfied slots
          const int upto = ready < batch ? ready : batch;
          if (upto > published) {
            published = upto;
            // the ballot ordered every lane's acquire before this store (warp-level synchronisation)
            if (lane == 0) st_release_cta_shared(ready_seq, seq_base + (uint32_t)(q0 + published));
          }
          if (published == batch) break;
        }
      }
      seq_base += (uint32_t)nslots;
    }
  } else if (warp_idx >= 4 && warp_idx < 4 + kArr) {
    // ================================= epilogue =================================
    const int quarter = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int cq = (warp_idx - 4) >> 2;          // which share of the tile's columns
    const uint32_t stage_warp = smem_base + L::kStoreOffset + (uint32_t)(warp_idx - 4) * kStoreWarpBytes;
    ChainHooks hooks;
    hooks.pending = nullptr; hooks.n_pending = 0; hooks.item_flags = nullptr; hooks.lane = lane; hooks.n_groups = KO::kGroups;
    hooks.at_mid = (args.opts & CHAIN_OPT_MIDARRIVE) != 0;
    hooks.tile_arrive = (args.opts & CHAIN_OPT_TILE_ARRIVE) != 0;
    hooks.wfence = (args.opts & CHAIN_OPT_NO_WFENCE) == 0;
    hooks.nowait = (args.opts & CHAIN_OPT_DEBUG_NOWAIT) != 0;
    hooks.carried = false;
    hooks.cta_arrive = cta_arrive; hooks.cnt_base = arrive_cnt; hooks.arrivers = (uint32_t)kArr; hooks.cur_item = 0;
    hooks.pending_slot = 0;
    const bool defer = (args.opts & CHAIN_OPT_NO_DEFER) == 0;
    int n = 0;
    ItemCursor it;
    for (it.init(args, first_item); it.valid(args); it.advance(args, item_stride), ++n) {
      const int dir = it.dir;
      const int n_blk = it.n_blk(args);
      const int row_blk = it.m_grp(args) * CTAS + (int)cta_rank;
      const int as = n & 1;
      const int row_base = row_blk * BLOCK_M + quarter * 32;
      const int row = row_base + lane;
      const uint32_t chain = (uint32_t)(args.chain_offset + (uint64_t)row);
      EpiParams ep;
      ep.nbias = args.nbias[dir]; ep.ldo = args.ldo[dir]; ep.n_valid = args.n_valid[dir];
      ep.ctx = make_draw_ctx(args.seed, args.step_offset + (uint64_t)it.t, args.tag);
      ep.scale = -1.4426950408889634f; ep.bmul = 1.f; ep.ratio = 0.f; ep.debug_mode = 0;
      ep.kf_magic = args.kf_magic;
      if (EPI0 != EPI_GIBBS) {
        const float bk = __ldg(args.betas + 1 + (it.t >> 1)), bkm1 = __ldg(args.betas + (it.t >> 1));
        ep.scale = -1.4426950408889634f * bk;
        ep.bmul = dir == 0 ? bk : 1.f;
        ep.ratio = bkm1 / bk;
      }
      // the arrival for the previous item's last chunk is sent after this item's first Philox words have been
      // drawn (they hide most of the store latency) and before the wait for the accumulator
      auto wait_full = [&] {
        const uint32_t par = ((uint32_t)n >> 1) & 1u;
        // lane 0's probe decides for the warp (the hooks synchronise the warp)
        const bool ready = defer && __shfl_sync(0xffffffffu, mbar_test(tmem_full_bar(as), par) ? 1 : 0, 0) != 0;
        hooks.at_acc_wait(ready);
        mbar_wait(tmem_full_bar(as), par);
      };
      auto release = [&] {
        if (CTAS == 2) mbar_arrive_cluster(mapa_u32(tmem_empty_bar(as), 0));   // the leader's barrier
        else mbar_arrive(tmem_empty_bar(as));
      };
      // counters of the state this item WRITES (direction dir), first chunk of this tile
      uint32_t* item_flags = args.flags + ((size_t)dir * args.row_blocks + row_blk) * args.chunks_ld + n_blk * KO::kGroups;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BLOCK_N);
      // (a deferred tile keeps its own counters in hooks.pending; item_flags is only used for the chunks stored from now on)
      auto wait_full2 = [&] { wait_full(); hooks.item_flags = item_flags; hooks.cur_item = n; };
      if (dir == 0)
        epilogue_tile<BLOCK_N, EPI0>(ep, &tm_o0, tmem_acc, n_blk, row_base, chain, quarter, cq, lane, stage_warp,
                                     EPI0 == EPI_AIS_WEIGHT ? args.wpart + (size_t)row * args.wpart_ld : nullptr,
                                     wait_full2, release, hooks);
      else
        epilogue_tile<BLOCK_N, EPI1>(ep, &tm_o1, tmem_acc, n_blk, row_base, chain, quarter, cq, lane, stage_warp,
                                     nullptr, wait_full2, release, hooks);
    }
    hooks.arrive();
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

template <int BLOCK_N, int EPI0, int EPI1, int CTAS, bool SPLITW>
int launch_chain_impl(const CUtensorMap* tm, const ChainArgs& a, int grid_clusters, cudaStream_t st) {
  auto kern = umma_chain_kernel<BLOCK_N, EPI0, EPI1, CTAS, SPLITW>;
  using L = SmemLayout<BLOCK_N, CTAS>;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if ((rc = ensure_dynamic_smem(kern, di.device, L::kTotal))) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(grid_clusters * CTAS));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RBM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7], a));
  return check_launch("umma_chain_kernel");
}

// tm: { A dir 0 (V), A dir 1 (H), W MN-major, W K-major, W_low MN-major, W_low K-major, out dir 0 (H), out dir 1 (V) }
template <int EPI0, int EPI1>
int launch_chain(int block_n, int ctas, bool split_w, const CUtensorMap* tm, const ChainArgs& a, int grid_clusters,
                 cudaStream_t st) {
  if (split_w) {
    if (block_n == 256 && ctas == 2) return launch_chain_impl<256, EPI0, EPI1, 2, true>(tm, a, grid_clusters, st);
    if (block_n == 256) return launch_chain_impl<256, EPI0, EPI1, 1, true>(tm, a, grid_clusters, st);
    if (block_n == 128) return launch_chain_impl<128, EPI0, EPI1, 1, true>(tm, a, grid_clusters, st);
    return launch_chain_impl<64, EPI0, EPI1, 1, true>(tm, a, grid_clusters, st);
  }
  if (block_n == 256 && ctas == 2) return launch_chain_impl<256, EPI0, EPI1, 2, false>(tm, a, grid_clusters, st);
  if (block_n == 256) return launch_chain_impl<256, EPI0, EPI1, 1, false>(tm, a, grid_clusters, st);
  if (block_n == 128) return launch_chain_impl<128, EPI0, EPI1, 1, false>(tm, a, grid_clusters, st);
  return launch_chain_impl<64, EPI0, EPI1, 1, false>(tm, a, grid_clusters, st);
}

}  // namespace
}  // namespace rbm
