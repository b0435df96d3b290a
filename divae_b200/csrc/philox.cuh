// Philox4x32-10 counter-based generator and the draw contract shared by every kernel.
//
// Contract (restated by the CPU oracle in oracle/philox.py, pinned by the Random123 KATs):
//   key     = (seed & 0xffffffff, seed >> 32)
//   counter = (unit_block, chain_id, half_step_lo, tag | half_step_hi16 << 16)
//   unit c -> unit_block = (c / 32) * 4 + (c % 8) / 2,  slot = 2 * ((c % 32) / 8) + (c % 2)
//   the 128 output bits are eight 16-bit fields; slot s = word s/2, low half if s even.
//   u = (u16 * 65536 + e16) / 2^32 with e16 the same field of the call tagged | TAG_EXTRA,
//   a unit fires iff p >= u   (reference: `probs >= torch.rand`, models/samplers/pcd.py:35,49)
//
// The slot mapping makes one Philox call serve exactly the eight accumulator elements a
// thread owns in four consecutive m16n8 MMA tiles (SMEM variant) and eight of the 32
// consecutive columns a thread reads from TMEM (UMMA variant).
#pragma once
#include <cstdint>

namespace rbm {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

enum : uint32_t { TAG_GIBBS = 0, TAG_INIT = 1, TAG_AIS = 2, TAG_AIS_INIT = 3, TAG_EXTRA = 0x100 };

struct PhiloxKey { uint32_t k0, k1; };

// Per-call draw context: everything of the counter except unit_block and chain.
struct DrawCtx {
  uint32_t k0, k1;  // key
  uint32_t c2, c3;  // half-step low word, tag | half-step high 16 bits
};

__host__ __device__ inline DrawCtx make_draw_ctx(uint64_t seed, uint64_t half_step, uint32_t tag) {
  DrawCtx d;
  d.k0 = (uint32_t)seed; d.k1 = (uint32_t)(seed >> 32);
  d.c2 = (uint32_t)half_step;
  d.c3 = (tag & 0xFFFFu) | ((uint32_t)((half_step >> 32) & 0xFFFFu) << 16);
  return d;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1;
    c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += kPhiloxW0; k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 draw_block(const DrawCtx& d, uint32_t unit_block, uint32_t chain) {
  return philox4x32_10(unit_block, chain, d.c2, d.c3, d.k0, d.k1);
}
__device__ __forceinline__ uint4 draw_block_extra(const DrawCtx& d, uint32_t unit_block, uint32_t chain) {
  return philox4x32_10(unit_block, chain, d.c2, d.c3 | TAG_EXTRA, d.k0, d.k1);
}

__host__ __device__ __forceinline__ uint32_t unit_block_of(uint32_t c) { return (c >> 5) * 4u + ((c & 7u) >> 1); }
__host__ __device__ __forceinline__ uint32_t unit_slot_of(uint32_t c) { return 2u * ((c & 31u) >> 3) + (c & 1u); }

__device__ __forceinline__ uint32_t field16(const uint4& w, uint32_t slot) {
  uint32_t word = (slot >> 1) == 0 ? w.x : (slot >> 1) == 1 ? w.y : (slot >> 1) == 2 ? w.z : w.w;
  return (slot & 1u) ? (word >> 16) : (word & 0xFFFFu);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------
// Bernoulli decision  "sigmoid(x) >= u"  from a pre-activation x and the 16-bit field k16.
// Fast path decides from k16 alone using s = 1 + e^-x = 1/p:
//     p*65536 >= k16+1  <=>  (k16+1)*s <= 65536   -> 1      (u < (k16+1)/65536 <= p)
//     p*65536 <  k16    <=>   k16*s    >  65536   -> 0      (u >= k16/65536 > p)
// otherwise (probability 2^-16 per draw) the 16 refinement bits decide exactly.
// One MUFU (ex2) on the fast path; no reciprocal.
template <class ExtraFn>
__device__ __forceinline__ bool bernoulli_from_preact(float x, uint32_t k16, ExtraFn extra16) {
  const float e = ex2_approx(-1.4426950408889634f * x);  // e^-x; +inf for x << 0
  const float s = 1.0f + e;
  const float kf = __uint_as_float(0x4B000000u | k16) - 8388608.0f;  // exact float(k16)
  const float t = kf * s;
  if (t + s <= 65536.0f) return true;
  if (t > 65536.0f) return false;
  // ambiguous (or NaN from 0*inf): resolve with the refinement bits against p itself
  const float p = __frcp_rn(s);
  const float y = p * 65536.0f;
  const float fl = floorf(y);
  if (fl > kf) return true;
  if (fl < kf) return false;
  const float frac = (y - fl) * 65536.0f;  // exact in fp32
  return frac >= (float)extra16();
}

// Reference-form decision against an injected uniform: sigmoid(x) >= u, both fp32.
__device__ __forceinline__ float sigmoid_precise(float x) {
  return __frcp_rn(1.0f + expf(-x));
}

}  // namespace rbm
