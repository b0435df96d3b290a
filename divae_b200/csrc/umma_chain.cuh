// Chain kernel: ALL half-steps of a block-Gibbs (or AIS) call in ONE persistent launch, scheduled as a dataflow.
//
// A work item is one output tile of one half-step: (t, row group, n tile).  Items are numbered in that order and dealt
// round-robin to the resident CTAs (or CTA pairs), which run the usual warp-specialised pipeline over them
// (TMA producer -> tcgen05.mma -> fused sampling epilogue -> TMA store; two TMEM accumulator stages, so the
// epilogue of an item overlaps the MMAs of the next one).  Chains are independent, so item (t, m, n) depends only on
// the items (t-1, m, *) that wrote row block m of its A operand.  Every 128-chain row block has one arrival
// counter in global memory: an epilogue warp adds 1 (red.release.gpu) once its part of the new state tile is
// complete in global memory, and the producer warp polls (ld.acquire.gpu) for the cumulative count of half-step
// t-1 before it issues the first A load of an item of half-step t.  The W (B operand) loads of an item do not
// depend on anything and are issued BEFORE the wait.  There is no relaunch, no grid-wide or cluster barrier, and
// no wave quantisation: a CTA that finishes its tile of half-step t early starts on a tile of half-step t+1 whose
// row block is already complete.  Deadlock-free because every item only waits for lower-numbered items and every
// CTA takes its items in increasing order (all CTAs are co-resident: grid <= number of SMs, 1 CTA per SM).
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
  int k_blocks[2];         // K / 64 of the GEMM of dir (x2 with split weights: [high; low] stacked along K)
  int split_k[2];          // split weights: K blocks >= split_k[dir] read the low-part tensor map (else k_blocks[dir])
  int m_groups;            // row groups of CTAS * 128 chains
  int steps;               // half-steps in this launch
  int first_dir;
  uint64_t seed, chain_offset, step_offset;   // half-step t draws with counter step_offset + t
  uint32_t tag;
  const float* betas;      // annealed flavours: beta index of half-step t is 1 + t/2
  float* wpart;            // AIS: [rows, wpart_ld] log-weight partials (log2 units), one slot per 32 hidden units
  int wpart_ld;
  uint32_t kf_magic;       // kKfMagic (see EpiParams)
  uint32_t* flags;         // [m_groups * CTAS] arrival counters of the 128-chain row blocks, zeroed before the launch
};

// Position in the item sequence of one cluster: items c, c + G, c + 2G, ... of the (t, tile) enumeration.
struct ItemCursor {
  int t, dir, tile;
  uint32_t need;           // arrivals a row block must have collected before an item of half-step t may read it
  __device__ __forceinline__ void init(const ChainArgs& a, int first, int arrivals_per_tile) {
    t = 0; dir = a.first_dir & 1; tile = first; need = 0;
    normalise(a, arrivals_per_tile);
  }
  __device__ __forceinline__ void normalise(const ChainArgs& a, int arrivals_per_tile) {
    while (t < a.steps && tile >= a.m_groups * a.n_tiles[dir]) {
      tile -= a.m_groups * a.n_tiles[dir];
      need += (uint32_t)(a.n_tiles[dir] * arrivals_per_tile);
      ++t; dir ^= 1;
    }
  }
  __device__ __forceinline__ void advance(const ChainArgs& a, int stride, int arrivals_per_tile) {
    tile += stride;
    normalise(a, arrivals_per_tile);
  }
  __device__ __forceinline__ bool valid(const ChainArgs& a) const { return t < a.steps; }
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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
  constexpr int kStages = L::kStages;
  constexpr int kBCols = BLOCK_N / CTAS;        // columns of the W tile this CTA stages
  constexpr int kArr = ES::kWarps;              // arrivals per tile and 128-chain row block
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::kBarOffset + 8 * (2 * kStages + 4));

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
    ItemCursor it;
    for (it.init(args, first_item, kArr); it.valid(args); it.advance(args, item_stride, kArr)) {
      const int dir = it.dir;
      const int n_blk = it.tile % args.n_tiles[dir], m_grp = it.tile / args.n_tiles[dir];
      const int row_blk = m_grp * CTAS + (int)cta_rank;
      const int row0 = row_blk * BLOCK_M;                                 // this CTA's 128 chains
      const int ncol0 = n_blk * BLOCK_N + (int)cta_rank * kBCols;         // this CTA's share of the W tile
      const CUtensorMap* tm_a = dir ? &tm_a1 : &tm_a0;
      const int nk = args.k_blocks[dir], ksplit = args.split_k[dir];
      const uint32_t lead_full_base = CTAS == 2 ? mapa_u32(full_bar(0), 0) : 0u;
      auto issue_b = [&](int kb, int s) {
        // arm the barrier and fetch the W block: dir 0 reads W[k rows][n cols] as [64 x 64] MN-major boxes,
        // dir 1 reads W[n rows][k cols] as one [kBCols x 64] K-major box
        const uint32_t sb = smem_base + s * L::kStageBytes + kABytes;
        const bool lo = SPLITW && kb >= ksplit;
        const int kw = lo ? kb - ksplit : kb;
        const CUtensorMap* tm_b = dir ? (lo ? &tm_b1_lo : &tm_b1) : (lo ? &tm_b0_lo : &tm_b0);
        if (CTAS == 1) {
          mbar_arrive_expect_tx(full_bar(s), L::kStageBytes);
          if (dir == 0) {
#pragma unroll
            for (int nb = 0; nb < kBCols / 64; ++nb)
              tma_load_2d(sb + nb * (BLOCK_K * 128), tm_b, full_bar(s), ncol0 + nb * 64, kw * BLOCK_K);
          } else {
            tma_load_2d(sb, tm_b, full_bar(s), kw * BLOCK_K, ncol0);
          }
        } else {
          // both CTAs credit the LEADER's full barrier (2 arrivals + the bytes of both CTAs)
          const uint32_t lead_full = lead_full_base + 8u * s;
          if (is_leader) mbar_arrive_expect_tx(full_bar(s), 2 * L::kStageBytes);
          else mbar_arrive_cluster(lead_full);
          if (dir == 0) {
#pragma unroll
            for (int nb = 0; nb < kBCols / 64; ++nb)
              tma_load_2d_pair(sb + nb * (BLOCK_K * 128), tm_b, lead_full, ncol0 + nb * 64, kw * BLOCK_K);
          } else {
            tma_load_2d_pair(sb, tm_b, lead_full, kw * BLOCK_K, ncol0);
          }
        }
      };
      auto issue_a = [&](int kb, int s) {
        const uint32_t sa = smem_base + s * L::kStageBytes;
        const int ka = (SPLITW && kb >= ksplit) ? kb - ksplit : kb;      // the state tile is read once per W part
        if (CTAS == 1) tma_load_2d(sa, tm_a, full_bar(s), ka * BLOCK_K, row0);
        else tma_load_2d_pair(sa, tm_a, lead_full_base + 8u * s, ka * BLOCK_K, row0);
      };
      // 1. W blocks of the first stages: no dependency, issued before the wait
      const int pre = it.t > 0 ? (nk < kStages ? nk : kStages) : 0;
      {
        int s = stage; uint32_t ph = phase;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (elect_one_sync()) issue_b(kb, s);
          __syncwarp();
          if (++s == kStages) { s = 0; ph ^= 1u; }
        }
      }
      // 2. the row block's previous half-step must be complete in global memory
      if (it.t > 0) {
        const uint32_t* flag = args.flags + row_blk;
        while (ld_acquire_gpu(flag) < it.need) { }
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire -> async-proxy (TMA) reads
      }
      // 3. state blocks of the pre-armed stages, then the rest of the K loop
      for (int kb = 0; kb < nk; ++kb) {
        if (kb >= pre) mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one_sync()) {
          if (kb >= pre) issue_b(kb, stage);
          issue_a(kb, stage);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
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
      for (it.init(args, first_item, kArr); it.valid(args); it.advance(args, item_stride, kArr), ++n) {
        const int dir = it.dir;
        const int as = n & 1;
        mbar_wait(tmem_empty_bar(as), (((uint32_t)n >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        const uint32_t idesc = dir ? idesc1 : idesc0;
        const int nk = args.k_blocks[dir];
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
  } else if (warp_idx >= 4 && warp_idx < 4 + kArr) {
    // ================================= epilogue =================================
    const int quarter = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int cq = (warp_idx - 4) >> 2;          // which share of the tile's columns
    const uint32_t stage_warp = smem_base + L::kStoreOffset + (uint32_t)(warp_idx - 4) * kStoreWarpBytes;
    uint32_t* pending = nullptr;                 // row-block counter the previous item still has to arrive on
    auto arrive_pending = [&] {
      if (pending != nullptr) {
        __syncwarp();                            // the other lanes' log-weight reds are ordered before the release
        if (lane == 0) {
          tma_store_wait_all();                  // this warp's boxes of the new state are complete in global memory
          asm volatile("fence.proxy.async;" ::: "memory");
          red_release_gpu_add(pending, 1u);
        }
        pending = nullptr;
      }
    };
    int n = 0;
    ItemCursor it;
    for (it.init(args, first_item, kArr); it.valid(args); it.advance(args, item_stride, kArr), ++n) {
      const int dir = it.dir;
      const int n_blk = it.tile % args.n_tiles[dir], m_grp = it.tile / args.n_tiles[dir];
      const int row_blk = m_grp * CTAS + (int)cta_rank;
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
      // the arrival for the previous item is sent after this item's first Philox words have been drawn (they
      // hide most of the store latency) and before the wait for the accumulator
      auto wait_full = [&] { arrive_pending(); mbar_wait(tmem_full_bar(as), ((uint32_t)n >> 1) & 1u); };
      auto release = [&] {
        if (CTAS == 2) mbar_arrive_cluster(mapa_u32(tmem_empty_bar(as), 0));   // the leader's barrier
        else mbar_arrive(tmem_empty_bar(as));
      };
      const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BLOCK_N);
      if (dir == 0)
        epilogue_tile<BLOCK_N, EPI0>(ep, &tm_o0, tmem_acc, n_blk, row_base, chain, quarter, cq, lane, stage_warp,
                                     EPI0 == EPI_AIS_WEIGHT ? args.wpart + (size_t)row * args.wpart_ld : nullptr,
                                     wait_full, release);
      else
        epilogue_tile<BLOCK_N, EPI1>(ep, &tm_o1, tmem_acc, n_blk, row_base, chain, quarter, cq, lane, stage_warp,
                                     nullptr, wait_full, release);
      pending = args.flags + row_blk;
    }
    arrive_pending();
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
