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
#include <initializer_list>
#include <vector>

#include "umma_common.cuh"

namespace rbm {
namespace {

constexpr uint32_t TAG_HITS = 4;   // Philox tag of the hit draws (oracle/philox.py)

// x [B, n] fp32 -> [rows_pad, 2 Kp] bf16 = [x_hi | x_lo], zero padded; eight units per thread, 16-byte stores
__global__ void split_input_kernel(const float* __restrict__ x, int64_t B, int n, uint16_t* __restrict__ dst,
                                   int64_t rows_pad, int Kp) {
  const int groups = Kp >> 3;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_pad * groups) return;
  const int64_t r = idx / groups;
  const int c0 = (int)(idx % groups) * 8;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float val = (r < B && c < n) ? __ldg(x + r * n + c) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(val);
    const __nv_bfloat16 l = __float2bfloat16_rn(val - __bfloat162float(h));
    if (j & 1) { hi[j >> 1] |= (uint32_t)__bfloat16_as_ushort(h) << 16; lo[j >> 1] |= (uint32_t)__bfloat16_as_ushort(l) << 16; }
    else { hi[j >> 1] = __bfloat16_as_ushort(h); lo[j >> 1] = __bfloat16_as_ushort(l); }
  }
  uint16_t* row = dst + r * (2 * (int64_t)Kp);
  *reinterpret_cast<uint4*>(row + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(row + Kp + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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

// Gumbel smoothing / gate of posterior logits (hierarchical encoder side):
//   l = clamp(logit, -88, 88)                                   models/networks/hierarchicalEncoder.py:167
//   training: z = sigmoid((l + log(rho) - log(1 - rho)) * beta) utils/dists/gumbelmod.py:18-21
//   else:     z = heaviside(l + log(rho) - log(1 - rho))        gumbelmod.py:22-23  = [sigmoid(l) >= u], u = 1 - rho
// rho comes from the library's draw contract (unit = unit_offset + column, chain = chain_offset + row, tag 5):
// u32 = u16 * 65536 + e16 (main and refinement fields), rho = 1 - (u32 + 0.5) / 2^32 in (0, 1).
constexpr uint32_t TAG_SMOOTH = 5;

__global__ void gumbel_smooth_kernel(float* __restrict__ logits, float* __restrict__ samples, int64_t B, int n, float beta,
                                     int is_training, int clamp, DrawCtx ctx, uint64_t chain_offset, uint32_t unit_offset) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * n) return;
  const int64_t r = idx / n;
  const int cidx = (int)(idx % n);
  float l = logits[idx];
  if (clamp) {
    l = fminf(fmaxf(l, -88.f), 88.f);
    logits[idx] = l;
  }
  const uint32_t unit = unit_offset + (uint32_t)cidx;
  const uint32_t chain = (uint32_t)(chain_offset + (uint64_t)r);
  const uint32_t blk = unit_block_of(unit), slot = unit_slot_of(unit);
  const uint32_t k16 = field16(draw_block(ctx, blk, chain), slot);
  if (is_training) {
    const uint32_t e16 = field16(draw_block_extra(ctx, blk, chain), slot);
    const float u = ((float)(k16 * 65536u + e16) + 0.5f) * 2.3283064365386963e-10f;   // (u32 + 0.5) / 2^32, rounded to fp32
    const float rho = fminf(fmaxf(1.0f - u, 5.9604645e-8f), 1.0f - 5.9604645e-8f);
    samples[idx] = sigmoid_precise((l + logf(rho) - logf(1.0f - rho)) * beta);
  } else {
    const bool one = bernoulli_from_preact(l, k16, [&]() { return field16(draw_block_extra(ctx, blk, chain), slot); });
    samples[idx] = one ? 1.f : 0.f;
  }
}

// B [K, N] fp32 (row-major, N contiguous) -> [3 Kp, Np] bf16 = [B_hi ; B_lo ; B_hi] stacked along K, zero padded: the
// MN-major B operand of C = A B with split operands (the K blocks of [A_hi | A_hi | A_lo] meet them in this order)
__global__ void split_weight_rows_kernel(const float* __restrict__ Bm, int K, int N, uint16_t* __restrict__ dst, int Kp, int Np) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Kp * Np) return;
  const int r = (int)(idx / Np), c = (int)(idx % Np);
  const float val = (r < K && c < N) ? Bm[(int64_t)r * N + c] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(val);
  const __nv_bfloat16 lo = __float2bfloat16_rn(val - __bfloat162float(hi));
  dst[(int64_t)r * Np + c] = __bfloat16_as_ushort(hi);
  dst[((int64_t)Kp + r) * Np + c] = __bfloat16_as_ushort(lo);
  dst[((int64_t)2 * Kp + r) * Np + c] = __bfloat16_as_ushort(hi);
}

struct GemmPlan {
  int64_t rows_pad;
  int Kp, Np, bn;
  size_t off_A, off_B, off_bias, total;
};
GemmPlan make_gemm_plan(int64_t M, int N, int K) {
  GemmPlan g;
  g.rows_pad = ceil_div64(std::max<int64_t>(M, 1), 2 * BLOCK_M) * (2 * BLOCK_M);
  g.Kp = round_up(K, 64); g.Np = round_up(N, 64);
  g.bn = pick_block_n(g.Np, g.rows_pad);
  auto align = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  size_t o = 0;
  g.off_A = o; o = align(o + (size_t)g.rows_pad * 2 * g.Kp * 2);
  g.off_B = o; o = align(o + (size_t)round_up(g.Np, 256) * 3 * g.Kp * 2);
  g.off_bias = o; o = align(o + (size_t)round_up(g.Np, 256) * 4);
  g.total = o + 1024;
  return g;
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

// B operand MN-major ([K, N] row-major storage): C = A B
int launch_linear_nn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const HalfStepArgs& a,
                     int sm_count, cudaStream_t st) {
  return bn == 256 ? launch_half_step_impl<256, true, EPI_LINEAR_F32, 1, true>(ta, tb, to, a, sm_count, st)
                   : launch_half_step_impl<128, true, EPI_LINEAR_F32, 1, true>(ta, tb, to, a, sm_count, st);
}

template <int EPI>
int launch_linear(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const HalfStepArgs& a,
                  int sm_count, cudaStream_t st) {
  return bn == 256 ? launch_half_step_impl<256, false, EPI, 1, true>(ta, tb, to, a, sm_count, st)
                   : launch_half_step_impl<128, false, EPI, 1, true>(ta, tb, to, a, sm_count, st);
}

}  // namespace

int gumbel_smooth(float* logits, float* samples, int64_t B, int n, float beta, int is_training, int clamp, uint64_t seed,
                  uint64_t chain_offset, uint64_t step, uint32_t unit_offset, cudaStream_t st) {
  if (B == 0 || n == 0) return 0;
  gumbel_smooth_kernel<<<(unsigned)ceil_div64(B * n, 256), 256, 0, st>>>(logits, samples, B, n, beta, is_training, clamp,
                                                                        make_draw_ctx(seed, step, TAG_SMOOTH), chain_offset,
                                                                        unit_offset);
  return check_launch("gumbel_smooth_kernel");
}

// C [M, N] = act(A op(B) + bias) for fp32 A [M, K]: op(B) = B^T with B [N, K] (trans_b = 1: y = x W^T, an nn.Linear) or
// op(B) = B with B [K, N] (trans_b = 0: dx = dy W).  One launch of the tcgen05 kernel of a Gibbs half-step over the
// tripled K axis of the split operands; bias may be NULL.
size_t umma_gemm_workspace_bytes(int64_t M, int N, int K) { return make_gemm_plan(M, N, K).total; }

int umma_gemm(int trans_b, int64_t M, int N, int K, const float* A, const float* Bm, const float* bias, int relu, float* C,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  const GemmPlan g = make_gemm_plan(M, N, K);
  RBM_REQUIRE(ws != nullptr && ws_bytes >= g.total, RBM_B200_ERR_WORKSPACE, "gemm needs %zu bytes of workspace, got %zu", g.total, ws_bytes);
  if (M == 0) return 0;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  RBM_REQUIRE(di.cc_major == 10, RBM_B200_ERR_NO_DEVICE, "the tensor-core GEMM needs an sm_100 device (got cc %d.x)", di.cc_major);
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  uint16_t* As = reinterpret_cast<uint16_t*>(base + g.off_A);
  uint16_t* Bs = reinterpret_cast<uint16_t*>(base + g.off_B);
  float* bias_p = reinterpret_cast<float*>(base + g.off_bias);
  split_input_kernel<<<(unsigned)ceil_div64(g.rows_pad * (g.Kp / 8), 256), 256, 0, st>>>(A, M, K, As, g.rows_pad, g.Kp);
  const int64_t nb = (int64_t)g.Np * g.Kp;
  if (trans_b) split_weight_kernel<<<(unsigned)ceil_div64(nb, 256), 256, 0, st>>>(Bm, N, K, Bs, g.Np, g.Kp);
  else split_weight_rows_kernel<<<(unsigned)ceil_div64(nb, 256), 256, 0, st>>>(Bm, K, N, Bs, g.Kp, g.Np);
  const int np256 = round_up(g.Np, 256);
  if (bias != nullptr) pad_bias_kernel<<<ceil_div(np256, 256), 256, 0, st>>>(bias, bias_p, N, np256, 1.f);
  else RBM_CUDA_TRY(cudaMemsetAsync(bias_p, 0, sizeof(float) * (size_t)np256, st));
  if ((rc = check_launch("gemm staging kernels"))) return rc;
  CUtensorMap ta, tb, to;
  if ((rc = make_tmap(&ta, As, (uint64_t)g.rows_pad, (uint64_t)(2 * g.Kp), (uint64_t)(2 * g.Kp), BLOCK_M))) return rc;
  if (trans_b) rc = make_tmap(&tb, Bs, (uint64_t)g.Np, (uint64_t)(3 * g.Kp), (uint64_t)(3 * g.Kp), (uint32_t)g.bn);
  else rc = make_tmap(&tb, Bs, (uint64_t)(3 * g.Kp), (uint64_t)g.Np, (uint64_t)g.Np, BLOCK_K);
  if (rc) return rc;
  HalfStepArgs a{};
  a.nbias = bias_p;
  a.ldo = g.Np; a.m_tiles = (int)(g.rows_pad / BLOCK_M); a.n_tiles = ceil_div(g.Np, g.bn);
  a.k_blocks = 3 * g.Kp / BLOCK_K; a.n_valid = N; a.rows_valid = M;
  a.out_f32 = C; a.ld_out = N; a.relu = relu ? 1 : 0;
  a.f32_tma = (N % 4 == 0 && ((uintptr_t)C & 15) == 0) ? 1 : 0;
  if (a.f32_tma) {
    if ((rc = make_tmap_f32(&to, C, (uint64_t)M, (uint64_t)N, (uint64_t)N, 32, 16, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  } else {
    to = ta;
  }
  return trans_b ? launch_linear<EPI_LINEAR_F32>(g.bn, ta, tb, to, a, di.sm_count, st)
                 : launch_linear_nn(g.bn, ta, tb, to, a, di.sm_count, st);
}

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
    const int64_t n = pl.rows_pad * (K0p / 8);
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
    // fp32 output through staged TMA stores when the row pitch allows it (multiple of 16 bytes, aligned base);
    // the tensor map carries the REAL extent [B, out_features], so padding rows / columns are clipped
    a.f32_tma = (lp.lin->out_features % 4 == 0 && ((uintptr_t)dst_f32 & 15) == 0) ? 1 : 0;
    if (a.f32_tma) {
      if ((r = make_tmap_f32(&to, dst_f32, (uint64_t)B, (uint64_t)lp.lin->out_features, (uint64_t)lp.lin->out_features, 32, 16,
                             CU_TENSOR_MAP_SWIZZLE_64B))) return r;
    } else {
      to = ta;   // no TMA store: any valid map
    }
    return launch_linear<EPI_LINEAR_F32>(lp.bn, ta, tb, to, a, sm_count, st);
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

}  // namespace rbm
