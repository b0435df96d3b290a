// SMEM variant: persistent k-step block Gibbs for small priors (nV, nH <= 256).
//
// One launch runs ALL 2k half-steps of PCD.block_gibbs_sampling (models/samplers/pcd.py:51-69):
//   * W (bf16) and the pre-scaled biases are loaded into shared memory once per CTA;
//   * chain states never leave the SM: they live as packed bits in shared memory (4 KB per
//     128 chains) and, while being multiplied, as bf16 {0,1} MMA A-fragments in registers;
//   * each half-step is  acc = state x W (mma.sync m16n8k16 bf16, fp32 accumulate, B fragments
//     by ldmatrix / ldmatrix.trans from the ONE row-major copy of W),  then per unit:
//     bias add, sigmoid via one ex2, Philox4x32-10 field, Bernoulli compare, bit pack.
// The accumulator fragment layout of two adjacent n-tiles IS the A-fragment layout of the next
// half-step's k-tile, so new states go registers -> bits -> registers with no transposition;
// the Philox slot mapping (philox.cuh) gives every thread exactly its own 8 units per call.
//
// A chain group = 16 chains.  WN warps cooperate on one group (each owns a quarter/half/all
// of the output units, 64-unit words) and exchange the new state through shared memory; a CTA
// holds GROUPS groups.  WN=1 maximises throughput for many chains, WN=2/4 cuts the latency of
// one half-step when there are few chains (BASELINE configs 2-3 are latency-bound).
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "philox.cuh"
#include "variants.h"

namespace rbm {
namespace {

constexpr int kMaxUnits = 256;
constexpr int kMaxWords = kMaxUnits / 64;
constexpr int kSlots = 2 * kMaxWords;       // state slots per side: words, or half-words in HALF mode
constexpr float kNegLog2e = -1.4426950408889634f;

struct SmemArgs {
  const uint16_t* W;     // [nV, nH] bf16 row-major (global), or nullptr: W32 is rounded to bf16 (RNE) while staging
  const float* W32;      // [nV, nH] fp32 row-major (global)
  const float* vbias;
  const float* hbias;
  float* v;              // [B, nV] fp32 in/out
  float* h;              // [B, nH] fp32 out
  int64_t B;
  int nV, nH;
  int k;
  int init_chains;
  uint64_t seed, chain_offset, step_offset;
  int wn;                // warps per chain group (1, 2 or 4)
  int groups;            // chain groups per CTA
  int64_t n_batches;     // ceil(B / (16 * groups))
  // fused negative-phase statistics of the FINAL (v, h) pair (raw counts, atomically added):
  // only for priors of at most 64 x 64 units (one state word per side); else nullptr
  float* S_vh;           // [nV, nH]
  float* S_v;            // [nV]
  float* S_h;            // [nH]
  // optional bf16 {0,1} copies of the final states, [B, ld16v] / [B, ld16h] (ld = units rounded up to 64): written from the
  // on-chip state words in the final sweep and read by the tensor-core statistics GEMM for priors of 65 - 256 units
  uint16_t* v16;
  uint16_t* h16;
  int ld16v, ld16h;
};

constexpr int kStatLd = 72;   // bf16 row pitch of the per-warp [16 chains x 64 units] statistics tiles (+16 B pad)

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  if (nthreads == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Exact decision for the rare ambiguous draw (same rule as gibbs_umma.cu / philox.cuh).
__device__ __noinline__ uint32_t smem_resolve_exact(float acc, float nb, uint32_t k16, DrawCtx ctx, uint32_t blk,
                                                    uint32_t chain, uint32_t slot) {
  const float kf = (float)k16;
  const float s = 1.0f + ex2_approx(fmaf(acc, kNegLog2e, nb));
  const float d = fmaf(-kf, s, 65536.0f);
  if (d >= s) return 1u;
  if (!(__float_as_uint(d) < __float_as_uint(s))) return 0u;
  const float y = __frcp_rn(s) * 65536.0f;
  const float fl = floorf(y);
  if (fl > kf) return 1u;
  if (fl < kf) return 0u;
  const uint32_t e16 = field16(draw_block_extra(ctx, blk, chain), slot);
  return ((y - fl) * 65536.0f >= (float)e16) ? 1u : 0u;
}

// State word layout (one 32-bit word per lane per 64 units): for k-tile q (16 units) of the
// word and A-fragment register r (0..3) the two bf16 halves come from bits (4q + r) and
// (4q + r + 16), so   a[q][r] = ((word >> (4q + r)) & 0x00010001) * 0x3F80   (bf16 1.0 = 0x3F80).
// Thread (g = lane/4, t = lane%4) owns rows g, g+8 and, in n-tile j of the word (8 units),
// units 8j + 2t + e:  q = j/2,  r = 2*(j%2) + (row half),  e = unit parity.
__device__ __forceinline__ int state_bit(int j, int rowhalf, int e) { return 4 * (j >> 1) + 2 * (j & 1) + rowhalf + 16 * e; }

// One half-step for this warp: out words [ow_begin, ow_end) of  f(state_in x Wop + bias).
//   TRANS_B = false: v -> h, Wop[k][n] = W[k][n]   (ldmatrix.trans on rows k)
//   TRANS_B = true : h -> v, Wop[k][n] = W[n][k]   (ldmatrix on rows n)
// FULL: both sides are whole 64-unit words (k_tiles == 4*MAXW, every word has 4 live column pairs):
// all tile guards fold away, leaving a branch-free ldmatrix/mma loop.
// HALF (needs FULL): the team has twice as many warps as the side has words; a warp owns HALF-words
// (32 units = one Philox group).  [ow_begin, ow_end) then counts half-words and the state buffers hold
// one slot per half-word (bits of the other half are zero), OR-ed together by the reader.
template <int MAXW, bool TRANS_B, bool FULL, bool HALF>
__device__ __forceinline__ void half_step(const uint32_t* __restrict__ in_words, uint32_t* __restrict__ out_words,
                                          int k_tiles, int n_units, int ow_begin, int ow_end, uint32_t w_smem,
                                          int ldw_bytes, const float* __restrict__ nbias, const DrawCtx ctx,
                                          uint32_t chain_lo, uint32_t chain_hi, int lane) {
  const int t = lane & 3;
  // expand the input state bits into bf16 A fragments (registers)
  uint32_t a[MAXW * 4][4];
#pragma unroll
  for (int w = 0; w < MAXW; ++w) {
    const uint32_t word = HALF ? (in_words[(2 * w) * 32 + lane] | in_words[(2 * w + 1) * 32 + lane])
                               : ((FULL || w * 4 < k_tiles) ? in_words[w * 32 + lane] : 0u);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < 4; ++r) a[w * 4 + q][r] = ((word >> (4 * q + r)) & 0x00010001u) * 0x3F80u;
  }
  // per-lane ldmatrix row/column offsets (matrix m = lane/8 of the x4 group)
  const int m = lane >> 3, rr = lane & 7;
  // TRANS_B=false: rows are k: row = k0 + (m&1)*8 + rr, col = n0 + (m>>1)*8
  // TRANS_B=true : rows are n: row = n0 + (m>>1)*8 + rr, col = k0 + (m&1)*8
  const int lrow = TRANS_B ? ((m >> 1) * 8 + rr) : ((m & 1) * 8 + rr);
  const int lcol = TRANS_B ? ((m & 1) * 8) : ((m >> 1) * 8);

  constexpr int NP = HALF ? 2 : 4;          // 16-unit column pairs per pass
  for (int oi = ow_begin; oi < ow_end; ++oi) {
    const int ow = HALF ? (oi >> 1) : oi;
    const int hsel = HALF ? (oi & 1) : 0;   // which half of the word this pass covers (HALF only)
    const int n_base = ow * 64;
    const int n_tiles16 = FULL ? 4 : min(4, (n_units - n_base + 15) >> 4);  // live 16-unit column pairs in this word
    float acc[2 * NP][4];
#pragma unroll
    for (int j = 0; j < 2 * NP; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < MAXW * 4; ++kt) {
      if (FULL || kt < k_tiles) {
#pragma unroll
        for (int pi = 0; pi < NP; ++pi) {       // 16-unit column pair p -> n-tiles 2p, 2p+1
          const int p = HALF ? 2 * hsel + pi : pi;
          if (FULL || p < n_tiles16) {
            uint32_t b0, b1, b2, b3;
            const int row = TRANS_B ? (n_base + p * 16 + lrow) : (kt * 16 + lrow);
            const int col = TRANS_B ? (kt * 16 + lcol) : (n_base + p * 16 + lcol);
            const uint32_t addr = w_smem + (uint32_t)(row * ldw_bytes + col * 2);
            if (TRANS_B) ldmatrix_x4(addr, b0, b1, b2, b3);
            else ldmatrix_x4_trans(addr, b0, b1, b2, b3);
            mma_bf16_16816(acc[2 * pi], a[kt], b0, b1);
            mma_bf16_16816(acc[2 * pi + 1], a[kt], b2, b3);
          }
        }
      }
    }
    // epilogue: halves of 32 units; per half one Philox call per owned row
    uint32_t word = 0;
#pragma unroll
    for (int hfi = 0; hfi < (HALF ? 1 : 2); ++hfi) {
      const int hf = HALF ? hsel : hfi;
      if (FULL || hf * 2 < n_tiles16) {
        const uint32_t blk = (uint32_t)((ow * 2 + hf) * 4 + t);
        const uint4 x0 = draw_block(ctx, blk, chain_lo);
        const uint4 x1 = draw_block(ctx, blk, chain_hi);
        const uint32_t wlo[4] = {x0.x, x0.y, x0.z, x0.w};
        const uint32_t whi[4] = {x1.x, x1.y, x1.z, x1.w};
        bool amb = false;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = hf * 4 + jj;
          const float2 nb = *reinterpret_cast<const float2*>(nbias + n_base + 8 * j + 2 * t);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = i & 1, rowhalf = i >> 1;
            const uint32_t wd = rowhalf ? whi[jj] : wlo[jj];
            const uint32_t mbits = e ? __byte_perm(wd, 0x4B000000u, 0x7632) : __byte_perm(wd, 0x4B000000u, 0x7610);
            const float kf = __uint_as_float(mbits) - 8388608.0f;
            const float s = 1.0f + ex2_approx(fmaf(acc[HALF ? jj : j][i], kNegLog2e, e ? nb.y : nb.x));
            const float d = fmaf(-kf, s, 65536.0f);
            amb = amb || (__float_as_uint(d) < __float_as_uint(s));
            word |= (d >= s) ? (1u << state_bit(j, rowhalf, e)) : 0u;
          }
        }
        if (amb) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = hf * 4 + jj;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = i & 1, rowhalf = i >> 1;
              const uint32_t wd = rowhalf ? whi[jj] : wlo[jj];
              const uint32_t k16 = e ? (wd >> 16) : (wd & 0xFFFFu);
              const uint32_t one = smem_resolve_exact(acc[HALF ? jj : j][i], nbias[n_base + 8 * j + 2 * t + e], k16, ctx, blk,
                                                      rowhalf ? chain_hi : chain_lo, (uint32_t)(2 * jj + e));
              const uint32_t bit = 1u << state_bit(j, rowhalf, e);
              word = one ? (word | bit) : (word & ~bit);
            }
          }
        }
      }
    }
    // units beyond n_units inside the last live 16-pair must stay 0 (they are K padding next half-step)
    if (!FULL && n_base + 64 > n_units) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (n_base + 8 * j + 2 * t + e >= n_units) word &= ~((1u << state_bit(j, 0, e)) | (1u << state_bit(j, 1, e)));
    }
    out_words[oi * 32 + lane] = word;
  }
}

template <int MAXW, bool FULL, bool HALF>
__global__ void __launch_bounds__(256, 1) smem_gibbs_kernel(const SmemArgs args) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int nV = args.nV, nH = args.nH;
  const int nVp = round_up(nV, 16), nHp = round_up(nH, 16);
  const int ldw = nHp + 8;                   // +16 B row pad: conflict-free ldmatrix in both directions
  const int ldw_bytes = ldw * 2;
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem);
  float* nbh = reinterpret_cast<float*>(smem + (size_t)nVp * ldw_bytes);
  float* nbv = nbh + kMaxUnits;
  uint32_t* vbits = reinterpret_cast<uint32_t*>(nbv + kMaxUnits);   // [groups][kSlots][32]
  uint32_t* hbits = vbits + args.groups * kSlots * 32;
  // fused statistics (S_vh != nullptr): CTA accumulator [64][64] + column sums, per-warp bf16 tiles
  float* S_s = reinterpret_cast<float*>(hbits + args.groups * kSlots * 32);
  float* sv_s = S_s + 64 * 64;
  float* sh_s = sv_s + 64;
  __nv_bfloat16* stat_tiles = reinterpret_cast<__nv_bfloat16*>(sh_s + 64);   // [groups][2][16][kStatLd]
  const bool fuse_stats = args.S_vh != nullptr;

  const int tid = threadIdx.x;
  // ---- stage W (zero padded) and the pre-scaled biases once per CTA
  if (args.W == nullptr) {
    // fp32 parameters, rounded here: the sampler never sees a bf16 copy that could be stale
    if ((nH & 3) == 0 && (reinterpret_cast<uintptr_t>(args.W32) & 15) == 0) {
      const int chunks = nH >> 2;                     // float4 chunks per row
      for (int idx = tid; idx < nV * chunks; idx += blockDim.x) {
        const int i = idx / chunks, c4 = idx % chunks;
        const float4 w = __ldg(reinterpret_cast<const float4*>(args.W32 + (size_t)i * nH) + c4);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(w.x, w.y), hi = __floats2bfloat162_rn(w.z, w.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(Ws + (size_t)i * ldw + 4 * c4) = pk;
      }
      const int padc = ldw - nH;
      for (int idx = tid; idx < nV * padc; idx += blockDim.x)
        reinterpret_cast<uint16_t*>(Ws)[(idx / padc) * ldw + nH + idx % padc] = 0;
      for (int idx = tid + nV * ldw; idx < nVp * ldw; idx += blockDim.x) reinterpret_cast<uint16_t*>(Ws)[idx] = 0;
    } else {
      for (int idx = tid; idx < nVp * ldw; idx += blockDim.x) {
        const int i = idx / ldw, j = idx % ldw;
        Ws[idx] = __float2bfloat16_rn((i < nV && j < nH) ? args.W32[(size_t)i * nH + j] : 0.f);
      }
    }
  } else if ((nH & 7) == 0 && (reinterpret_cast<uintptr_t>(args.W) & 15) == 0) {
    const int chunks = nH >> 3;                       // 16-byte chunks per row
    // asynchronous 16-byte copies global -> shared (LDGSTS): all of a thread's copies are in flight at once
    for (int idx = tid; idx < nV * chunks; idx += blockDim.x) {
      const int i = idx / chunks, c8 = idx % chunks;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Ws + (size_t)i * ldw + 8 * c8);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(args.W + (size_t)i * nH + 8 * c8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int padc = ldw - nH;                        // zero the column pad and the row pad
    for (int idx = tid; idx < nV * padc; idx += blockDim.x)
      reinterpret_cast<uint16_t*>(Ws)[(idx / padc) * ldw + nH + idx % padc] = 0;
    for (int idx = tid + nV * ldw; idx < nVp * ldw; idx += blockDim.x) reinterpret_cast<uint16_t*>(Ws)[idx] = 0;
  } else {
    for (int idx = tid; idx < nVp * ldw; idx += blockDim.x) {
      const int i = idx / ldw, j = idx % ldw;
      uint16_t val = 0;
      if (i < nV && j < nH) val = args.W[(size_t)i * nH + j];
      reinterpret_cast<uint16_t*>(Ws)[idx] = val;
    }
  }
  for (int i = tid; i < kMaxUnits; i += blockDim.x) {
    nbh[i] = i < nH ? kNegLog2e * args.hbias[i] : 0.f;
    nbv[i] = i < nV ? kNegLog2e * args.vbias[i] : 0.f;
  }
  if (fuse_stats)
    for (int i = tid; i < 64 * 64 + 128; i += blockDim.x) S_s[i] = 0.f;
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int wn = args.wn;
  const int grp = warp / wn;             // chain group of this warp inside the CTA
  const int wsub = warp % wn;            // position inside the group's warp team
  const bool active_warp = grp < args.groups;
  const int g = lane >> 2, t = lane & 3;
  const int vwords = (nV + 63) >> 6, hwords = (nH + 63) >> 6;
  const int vt = (nV + 15) >> 4, ht = (nH + 15) >> 4;  // k-tiles per side
  // split of output words among the team
  // ownership in words, or in half-words when HALF
  const int h_own = HALF ? 2 * hwords : hwords, v_own = HALF ? 2 * vwords : vwords;
  const int h_per = (h_own + wn - 1) / wn, v_per = (v_own + wn - 1) / wn;
  const int h_begin = min(h_own, wsub * h_per), h_end = min(h_own, h_begin + h_per);
  const int v_begin = min(v_own, wsub * v_per), v_end = min(v_own, v_begin + v_per);
  const uint32_t w_smem = (uint32_t)__cvta_generic_to_shared(Ws);
  uint32_t* myv = vbits + grp * kSlots * 32;
  uint32_t* myh = hbits + grp * kSlots * 32;
  const int bar_id = 1 + grp, bar_threads = 32 * wn;

  for (int64_t batch = blockIdx.x; batch < args.n_batches; batch += gridDim.x) {
    if (active_warp) {
      const int64_t row0 = (batch * args.groups + grp) * 16;
      const int64_t r_lo = row0 + g, r_hi = row0 + g + 8;
      const uint32_t chain_lo = (uint32_t)(args.chain_offset + (uint64_t)r_lo);
      const uint32_t chain_hi = (uint32_t)(args.chain_offset + (uint64_t)r_hi);
      // ---- initial visible state -> bits (this warp's share of the words)
      for (int oi = v_begin; oi < v_end; ++oi) {
        const int ow = HALF ? (oi >> 1) : oi;
        const int hf_lo = HALF ? (oi & 1) : 0, hf_hi = HALF ? (oi & 1) + 1 : 2;
        uint32_t word = 0;
        if (args.init_chains) {
          const DrawCtx ictx = make_draw_ctx(args.seed, args.step_offset, TAG_INIT);
          for (int hf = hf_lo; hf < hf_hi; ++hf) {
            const uint32_t blk = (uint32_t)((ow * 2 + hf) * 4 + t);
            const uint4 x0 = draw_block(ictx, blk, chain_lo);
            const uint4 x1 = draw_block(ictx, blk, chain_hi);
            const uint32_t wlo[4] = {x0.x, x0.y, x0.z, x0.w}, whi[4] = {x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int e = i & 1, rowhalf = i >> 1, j = hf * 4 + jj;
                const uint32_t wd = rowhalf ? whi[jj] : wlo[jj];
                const uint32_t f = e ? (wd >> 16) : (wd & 0xFFFFu);
                const int unit = ow * 64 + 8 * j + 2 * t + e;
                if (unit < nV && !(f >> 15)) word |= 1u << state_bit(j, rowhalf, e);
              }
          }
        } else {
          for (int j = 4 * hf_lo; j < 4 * hf_hi; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int e = i & 1, rowhalf = i >> 1;
              const int unit = ow * 64 + 8 * j + 2 * t + e;
              const int64_t r = rowhalf ? r_hi : r_lo;
              if (unit < nV && r < args.B && args.v[r * nV + unit] > 0.5f) word |= 1u << state_bit(j, rowhalf, e);
            }
        }
        myv[oi * 32 + lane] = word;
      }
      group_barrier(bar_id, bar_threads);
      // ---- 2k half-steps, all on chip
      for (int s = 0; s < args.k; ++s) {
        const DrawCtx ch = make_draw_ctx(args.seed, args.step_offset + 2ull * s, TAG_GIBBS);
        half_step<MAXW, false, FULL, HALF>(myv, myh, vt, nH, h_begin, h_end, w_smem, ldw_bytes, nbh, ch, chain_lo, chain_hi, lane);
        group_barrier(bar_id, bar_threads);
        const DrawCtx cv = make_draw_ctx(args.seed, args.step_offset + 2ull * s + 1, TAG_GIBBS);
        half_step<MAXW, true, FULL, HALF>(myh, myv, ht, nV, v_begin, v_end, w_smem, ldw_bytes, nbv, cv, chain_lo, chain_hi, lane);
        group_barrier(bar_id, bar_threads);
      }
      // ---- final states back to the reference's dtype (fp32 {0,1})
      for (int side = 0; side < 2; ++side) {
        const uint32_t* bits = side ? myh : myv;
        float* out = side ? args.h : args.v;
        uint16_t* out16 = side ? args.h16 : args.v16;
        const int ld16 = side ? args.ld16h : args.ld16v;
        const int n = side ? nH : nV;
        const int wb = side ? h_begin : v_begin, we = side ? h_end : v_end;
        if (side == 1 && args.k == 0) continue;
        for (int oi = wb; oi < we; ++oi) {
          const int ow = HALF ? (oi >> 1) : oi;
          const uint32_t word = bits[oi * 32 + lane];
          for (int j = HALF ? 4 * (oi & 1) : 0; j < (HALF ? 4 * (oi & 1) + 4 : 8); ++j)
#pragma unroll
            for (int rowhalf = 0; rowhalf < 2; ++rowhalf) {
              const int64_t r = rowhalf ? r_hi : r_lo;
              const int unit = ow * 64 + 8 * j + 2 * t;
              if (r < args.B) {
                const float f0 = (word >> state_bit(j, rowhalf, 0)) & 1u ? 1.f : 0.f;
                const float f1 = (word >> state_bit(j, rowhalf, 1)) & 1u ? 1.f : 0.f;
                if (unit < n) out[r * n + unit] = f0;
                if (unit + 1 < n) out[r * n + unit + 1] = f1;
                if (out16 != nullptr)   // (units beyond n inside the last word are 0: K padding of the half-steps)
                  *reinterpret_cast<uint32_t*>(out16 + r * ld16 + unit) = (f0 != 0.f ? 0x3F80u : 0u) | (f1 != 0.f ? 0x3F800000u : 0u);
              }
            }
        }
      }
      if (fuse_stats && args.k > 0) {
        // ---- fused negative-phase statistics of this group's final (v, h): S += v^T h, column sums.
        // (<= 64 units per side => one word per side; the group's warps - one, or two in HALF mode - each
        //  expand the state bits they own into the group's bf16 tiles and split the MMA row tiles.)
        __nv_bfloat16* vs = stat_tiles + (size_t)grp * 2 * 16 * kStatLd;
        __nv_bfloat16* hs = vs + 16 * kStatLd;
        const int vj0 = HALF ? 4 * (v_begin & 1) : 0, hj0 = HALF ? 4 * (h_begin & 1) : 0;
        const int nj = HALF ? 4 : 8;
        const uint32_t wv = myv[(HALF ? v_begin : 0) * 32 + lane], wh = myh[(HALF ? h_begin : 0) * 32 + lane];
        for (int jj = 0; jj < nj; ++jj)
#pragma unroll
          for (int rowhalf = 0; rowhalf < 2; ++rowhalf) {
            const bool valid = (rowhalf ? r_hi : r_lo) < args.B;    // padded chains contribute nothing
            const int row = g + 8 * rowhalf;
            {
              const int j = vj0 + jj;
              const uint32_t b0 = state_bit(j, rowhalf, 0), b1 = state_bit(j, rowhalf, 1);
              const uint32_t pv = valid ? ((((wv >> b0) & 1u) * 0x3F80u) | (((wv >> b1) & 1u) * 0x3F800000u)) : 0u;
              *reinterpret_cast<uint32_t*>(vs + row * kStatLd + 8 * j + 2 * t) = pv;
            }
            {
              const int j = hj0 + jj;
              const uint32_t b0 = state_bit(j, rowhalf, 0), b1 = state_bit(j, rowhalf, 1);
              const uint32_t ph = valid ? ((((wh >> b0) & 1u) * 0x3F80u) | (((wh >> b1) & 1u) * 0x3F800000u)) : 0u;
              *reinterpret_cast<uint32_t*>(hs + row * kStatLd + 8 * j + 2 * t) = ph;
            }
          }
        group_barrier(bar_id, bar_threads);
        const uint32_t vs_a = (uint32_t)__cvta_generic_to_shared(vs), hs_a = (uint32_t)__cvta_generic_to_shared(hs);
        const int m = lane >> 3, rr = lane & 7;
        // A = v^T (m = visible unit, k = chain): x4.trans of [chains 0-7|8-15] x [units 0-7|8-15] blocks
        // B = h   (k = chain, n = hidden unit): x4.trans of the same block shape
#pragma unroll 1
        for (int mt = wsub; mt < vt; mt += wn) {
          uint32_t a[4];
          ldmatrix_x4_trans(vs_a + (uint32_t)((((m >> 1) * 8 + rr) * kStatLd + mt * 16 + (m & 1) * 8) * 2), a[0], a[1], a[2], a[3]);
          float acc[8][4];
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            if (p < ht) {
              uint32_t b0, b1, b2, b3;
              ldmatrix_x4_trans(hs_a + (uint32_t)((((m & 1) * 8 + rr) * kStatLd + p * 16 + (m >> 1) * 8) * 2), b0, b1, b2, b3);
              mma_bf16_16816(acc[2 * p], a, b0, b1);
              mma_bf16_16816(acc[2 * p + 1], a, b2, b3);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int vi = mt * 16 + g + 8 * (i >> 1), hj = 8 * j + 2 * t + (i & 1);
              if (acc[j][i] != 0.f) atomicAdd(&S_s[vi * 64 + hj], acc[j][i]);
            }
        }
        // column sums: the group's lanes split the 64 units
        for (int u = lane + 32 * wsub; u < 64; u += 32 * wn) {
          float cv = 0.f, chh = 0.f;
#pragma unroll
          for (int row = 0; row < 16; ++row) {
            cv += __bfloat162float(vs[row * kStatLd + u]);
            chh += __bfloat162float(hs[row * kStatLd + u]);
          }
          if (cv != 0.f) atomicAdd(&sv_s[u], cv);
          if (chh != 0.f) atomicAdd(&sh_s[u], chh);
        }
      }
      group_barrier(bar_id, bar_threads);   // bits buffers are reused by the next batch
    }
  }
  if (fuse_stats) {
    // raw counts of this CTA -> global accumulators (integer-valued floats: exact, order-independent)
    __syncthreads();
    for (int idx = tid; idx < nV * nH; idx += blockDim.x) {
      const float x = S_s[(idx / nH) * 64 + idx % nH];
      if (x != 0.f) atomicAdd(&args.S_vh[idx], x);
    }
    for (int i = tid; i < nV; i += blockDim.x) if (sv_s[i] != 0.f) atomicAdd(&args.S_v[i], sv_s[i]);
    for (int j = tid; j < nH; j += blockDim.x) if (sh_s[j] != 0.f) atomicAdd(&args.S_h[j], sh_s[j]);
  }
}

size_t smem_bytes_needed(int nV, int nH, int groups, bool fuse_stats) {
  const int nVp = round_up(nV, 16), nHp = round_up(nH, 16);
  size_t n = (size_t)nVp * (nHp + 8) * 2 + 2 * kMaxUnits * sizeof(float) + (size_t)2 * groups * kSlots * 32 * sizeof(uint32_t);
  if (fuse_stats) n += (64 * 64 + 128) * sizeof(float) + (size_t)groups * 2 * 16 * kStatLd * 2;
  return n;
}

}  // namespace

bool smem_variant_supported(int nV, int nH) { return nV >= 1 && nH >= 1 && nV <= kMaxUnits && nH <= kMaxUnits; }

size_t smem_workspace_bytes(int, int, int64_t) { return 0; }

// Workspace of a call WITH statistics: bf16 copies of the final states for the tensor-core statistics of priors with
// 65 - 256 units per side (up to 64 units the statistics are accumulated inside the sampling kernel: nothing needed)
size_t smem_stats_workspace_bytes(int nV, int nH, int64_t B) {
  if (smem_can_fuse_stats(nV, nH) || B <= 0) return 0;
  return (size_t)B * (size_t)(round_up(nV, 64) + round_up(nH, 64)) * 2 + 256;
}

bool smem_can_fuse_stats(int nV, int nH) { return nV <= 64 && nH <= 64; }

int smem_block_gibbs(const rbm_b200_params* p, float* v, float* h, int64_t B, int k, int init_chains, uint64_t seed,
                     uint64_t chain_offset, uint64_t step_offset, void* ws, size_t ws_bytes, cudaStream_t st, float* S_vh,
                     float* S_v, float* S_h) {
  RBM_REQUIRE(smem_variant_supported(p->n_visible, p->n_hidden), RBM_B200_ERR_UNSUPPORTED,
              "SMEM variant supports priors up to %d x %d units, got %d x %d", kMaxUnits, kMaxUnits, p->n_visible,
              p->n_hidden);
  DeviceInfo di;
  int rc_di = get_device_info(&di);
  if (rc_di) return rc_di;
  const int sm_count = di.sm_count;
  SmemArgs a{};
  a.W = p->W_bf16; a.W32 = p->W; a.vbias = p->vbias; a.hbias = p->hbias; a.v = v; a.h = h; a.B = B;
  a.nV = p->n_visible; a.nH = p->n_hidden; a.k = k; a.init_chains = init_chains;
  a.seed = seed; a.chain_offset = chain_offset; a.step_offset = step_offset;
  const bool fuse = S_vh != nullptr && smem_can_fuse_stats(p->n_visible, p->n_hidden);
  a.S_vh = fuse ? S_vh : nullptr; a.S_v = fuse ? S_v : nullptr; a.S_h = fuse ? S_h : nullptr;
  // 65 - 256 units: the final sweep also leaves bf16 state tiles in the workspace; <vh>, <v>, <h> are then a tcgen05
  // GEMM + column sums over those tiles (umma_stats_tiles) instead of an fp32 SIMT pass over the fp32 outputs
  const bool tiles = S_vh != nullptr && !fuse && k > 0;
  if (tiles) {
    RBM_REQUIRE(ws != nullptr && ws_bytes >= smem_stats_workspace_bytes(p->n_visible, p->n_hidden, B), RBM_B200_ERR_WORKSPACE,
                "SMEM variant with statistics needs %zu bytes of workspace, got %zu",
                smem_stats_workspace_bytes(p->n_visible, p->n_hidden, B), ws_bytes);
    a.ld16v = round_up(a.nV, 64); a.ld16h = round_up(a.nH, 64);
    a.v16 = reinterpret_cast<uint16_t*>(((uintptr_t)ws + 127) & ~(uintptr_t)127);
    a.h16 = a.v16 + (size_t)B * a.ld16v;
  }
  const int vwords = ceil_div(a.nV, 64), hwords = ceil_div(a.nH, 64);
  const int maxw = std::max(vwords, hwords);
  const int64_t n_groups = ceil_div64(B, 16);
  // few chains -> split each group's units over more warps (latency); many chains -> 1 warp per group
  int wn = 1;
  const int wn_cap = std::min(vwords, hwords);
  while (wn * 2 <= wn_cap && wn * 2 <= 4 && n_groups * wn < (int64_t)sm_count * 8) wn *= 2;
  // whole-word square-ish priors (64, 128, 192, 256 units on both sides with equal word counts) take the
  // guard-free instantiation
  const bool full = (a.nV % 64 == 0) && (a.nH % 64 == 0) && vwords == hwords && (vwords == 1 || vwords == 2 || vwords == 4);
  // very few chains: split words in halves over twice the warps (one Philox group per warp)
  bool half = false;
  if (full && wn == vwords && n_groups * wn * 2 <= (int64_t)sm_count * 4) { half = true; wn *= 2; }
  int groups = 8 / wn;
  while (groups > 1 && ceil_div64(n_groups, groups) < sm_count) groups >>= 1;
  a.wn = wn; a.groups = groups;
  a.n_batches = ceil_div64(n_groups, groups);
  const size_t smem = smem_bytes_needed(a.nV, a.nH, groups, fuse);
  const int grid = (int)std::min<int64_t>(a.n_batches, sm_count);
  const int threads = 32 * wn * groups;
  auto launch = [&](auto kern) -> int {
    int rc = ensure_dynamic_smem(kern, di.device, (int)smem);
    if (rc) return rc;
    kern<<<grid, threads, smem, st>>>(a);
    if ((rc = check_launch("smem_gibbs_kernel"))) return rc;
    if (tiles)
      return umma_stats_tiles(reinterpret_cast<const void*>(a.v16), reinterpret_cast<const void*>(a.h16), B, a.nV, a.nH, S_vh, S_v, S_h, st);
    return 0;
  };
  if (half) {
    if (maxw == 1) return launch(smem_gibbs_kernel<1, true, true>);
    if (maxw == 2) return launch(smem_gibbs_kernel<2, true, true>);
    return launch(smem_gibbs_kernel<4, true, true>);
  }
  if (full) {
    if (maxw == 1) return launch(smem_gibbs_kernel<1, true, false>);
    if (maxw == 2) return launch(smem_gibbs_kernel<2, true, false>);
    return launch(smem_gibbs_kernel<4, true, false>);
  }
  if (maxw <= 1) return launch(smem_gibbs_kernel<1, false, false>);
  if (maxw <= 2) return launch(smem_gibbs_kernel<2, false, false>);
  return launch(smem_gibbs_kernel<4, false, false>);
}

}  // namespace rbm
