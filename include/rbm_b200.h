/*
 * rbm_b200.h — C ABI of the B200-native RBM block-Gibbs library (librbm_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of ezeeEric/DiVAE: block Gibbs
 * sampling of the RBM prior.  The reference implements that path in Python/PyTorch
 * (no FFI exists in the reference); each entry point below names the reference
 * function it replaces (paths relative to the reference root).  The Python host
 * side (divae_b200/) binds these with ctypes and mirrors the reference classes.
 *
 * Conventions
 *  - plain C: pointers, sizes, integers.  No C++ or torch types.
 *  - every pointer is a DEVICE pointer valid on the current CUDA device unless
 *    stated otherwise; the caller owns every buffer; the library allocates
 *    nothing persistent (workspaces are caller-provided).
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    never synchronises, never touches the default stream, and is CUDA-graph
 *    capturable.
 *  - return value: 0 = ok, < 0 = rbm_b200 error (RBM_B200_ERR_*), > 0 = cudaError_t.
 *    rbm_b200_last_error() returns a thread-local description.
 *  - matrices are row-major.  States are float32 holding exactly 0.0f / 1.0f at
 *    the API boundary (the reference's dtype, models/samplers/pcd.py:35).
 *  - there is no CPU fallback: without a CUDA device every compute entry fails.
 *
 * Random draws follow the counter-based Philox4x32-10 contract documented in
 * DESIGN.md ("Draw contract") so results do not depend on tiling, kernel variant
 * or GPU count:  counter = (unit_block, chain_id, half_step, tag), key = seed.  chain_id is the low 32 bits
 * of (chain_offset + row): draws repeat after 2^32 chains under one seed; half_step is a 48-bit counter.
 */
#ifndef RBM_B200_H
#define RBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBM_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define RBM_B200_API __attribute__((visibility("default")))
#else
#define RBM_B200_API
#endif

enum {
  RBM_B200_OK = 0,
  RBM_B200_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, ...) */
  RBM_B200_ERR_UNSUPPORTED = -2, /* shape / variant not supported by the requested kernel */
  RBM_B200_ERR_WORKSPACE = -3,   /* workspace missing or too small */
  RBM_B200_ERR_NO_DEVICE = -4    /* no CUDA device / wrong architecture (needs sm_100) */
};

/* Kernel variant selector for the k-step sampler. */
enum {
  RBM_B200_VARIANT_AUTO = 0,
  RBM_B200_VARIANT_GENERAL = 1, /* fp32 SIMT, one launch per half-step, any shape, fp32 W */
  RBM_B200_VARIANT_SMEM = 2,    /* persistent: W (bf16) resident in shared memory, states on chip */
  RBM_B200_VARIANT_UMMA = 3,    /* tcgen05/TMEM GEMM per half-step fed by TMA, bf16 {0,1} state tiles */
  RBM_B200_VARIANT_UMMA_SPLIT = 4 /* UMMA with W carried as bf16 high + bf16 low parts (needs the fp32 W): the GEMM
                                     runs over a doubled K axis and the pre-activations are fp32-grade (~2^-17
                                     relative) instead of bf16-rounded-W grade (2^-9); about twice the tensor time */
};

/* Epilogue mode of a single half-step. */
enum {
  RBM_B200_MODE_SAMPLE = 0,   /* Bernoulli draw with the Philox contract              */
  RBM_B200_MODE_INJECTED = 1, /* Bernoulli draw against caller-supplied uniforms (test) */
  RBM_B200_MODE_PROBS = 2     /* return probabilities, no draw (mean-field)           */
};

/* RBM parameters (reference: RBM.__init__, models/rbm/rbm.py:28-34).
 * W is [n_visible, n_hidden] fp32.  W_bf16 is OPTIONAL: the same matrix rounded to bf16
 * (round-to-nearest-even).  With W_bf16 == NULL the SMEM and UMMA variants round the fp32 W
 * themselves while staging it (shared memory / workspace), so a caller that only keeps the
 * fp32 parameters can never sample from a stale low-precision copy. */
typedef struct {
  int32_t n_visible;
  int32_t n_hidden;
  const float *W;
  const uint16_t *W_bf16;
  const float *vbias; /* [n_visible]  "a" */
  const float *hbias; /* [n_hidden]   "c" */
} rbm_b200_params;

/* Library / device queries (host only). */
RBM_B200_API int rbm_b200_abi_version(void);
RBM_B200_API const char *rbm_b200_last_error(void);
/* Writes SM count, max opt-in shared memory per block (bytes) and compute
 * capability (major*10+minor) of the current device. */
RBM_B200_API int rbm_b200_device_info(int32_t *sm_count, int32_t *smem_optin, int32_t *cc);

/* One conditional half-step on B rows.
 *   direction 0: out[B,nH] = f(in[B,nV] W   + hbias)   PCD.hidden_samples   pcd.py:23-35, gibbsSampler.py:20-23
 *   direction 1: out[B,nV] = f(in[B,nH] W^T + vbias)   PCD.visible_samples  pcd.py:37-49, gibbsSampler.py:25-28
 * f = sigmoid then, by `mode`, a Philox draw (`p >= u`), a draw against
 * `uniforms` ([B, n_out] fp32), or nothing (probabilities).  `in` may be
 * real-valued.  fp32 W, fp32 accumulate.  `half_step` is the draw counter. */
RBM_B200_API int rbm_b200_half_step(const rbm_b200_params *p, int direction, const float *in, float *out,
                       int64_t B, int mode, const float *uniforms, uint64_t seed,
                       uint64_t chain_offset, uint64_t half_step, void *stream);

/* Bernoulli(1/2) chain (re-)initialisation on the device, replacing the CPU
 * `rand >= rand` + H2D copy of PCD.__init__ / block_gibbs_sampling
 * (pcd.py:16-17,59,66-67).  v[B,nV] <- {0,1}. */
RBM_B200_API int rbm_b200_init_chains(float *v, int64_t B, int32_t n_visible, uint64_t seed,
                         uint64_t chain_offset, uint64_t step_offset, void *stream);

/* Bytes of caller-provided workspace rbm_b200_block_gibbs needs for this shape
 * and variant (0 for the GENERAL variant). Host only. */
RBM_B200_API size_t rbm_b200_block_gibbs_workspace(int32_t n_visible, int32_t n_hidden, int64_t B, int variant);
/* Name of the kernel that runs the half-steps of a block-Gibbs / AIS call of this shape on the current device
 * (shape-based choice inside the library: "smem_gibbs_kernel", "umma_resident_kernel", "umma_chain_kernel",
 * "umma_half_step_kernel", "general_half_step_kernel").  For reports and tests; static string, host only. */
RBM_B200_API const char *rbm_b200_kernel_choice(int32_t n_visible, int32_t n_hidden, int64_t B, int variant);
/* Same for rbm_b200_block_gibbs_stats: the SMEM variant needs room for bf16 copies of the final
 * states when the prior has 65 - 256 units per side (its statistics are a tensor-core GEMM over those
 * tiles); without it the statistics are computed from the fp32 outputs by the GENERAL kernels. */
RBM_B200_API size_t rbm_b200_block_gibbs_stats_workspace(int32_t n_visible, int32_t n_hidden, int64_t B, int variant);
/* Number of CUDA kernels one rbm_b200_block_gibbs call with these arguments launches on the current
 * device (the UMMA variant runs all 2k half-steps in ONE dataflow kernel when a half-step has few
 * tiles per SM, else one kernel per half-step).  Host only; -1 without a device. */
RBM_B200_API int64_t rbm_b200_block_gibbs_launch_count(int32_t n_visible, int32_t n_hidden, int64_t B,
                                                       int32_t k, int variant);

/* k steps of block Gibbs sampling: v <- v0; k x { h ~ p(h|v); v ~ p(v|h) }.
 * Replaces the loop of PCD.block_gibbs_sampling (pcd.py:51-69).  Returns the
 * same pair as the reference: h was drawn from the previous v, v from that h.
 *   v      [B,nV] in/out.  If init_chains != 0 the initial state is drawn on
 *          the device (Bernoulli(1/2), as rbm_b200_init_chains) and the input
 *          contents are ignored.
 *   h      [B,nH] out.
 *   step_offset  first half-step counter; half-step t of this call draws with
 *          counter step_offset + t  (t = 2*step for h, 2*step+1 for v).
 *   chain_offset global id of row 0 (rows are chains chain_offset .. +B-1), so
 *          a shard of a larger job draws exactly what the unsharded job would.
 *   variant      RBM_B200_VARIANT_*; AUTO picks SMEM for priors whose bf16 W
 *          fits in shared memory, UMMA for large ones, GENERAL otherwise. */
RBM_B200_API int rbm_b200_block_gibbs(const rbm_b200_params *p, float *v, float *h, int64_t B, int32_t k,
                         int init_chains, uint64_t seed, uint64_t chain_offset,
                         uint64_t step_offset, int variant, void *workspace,
                         size_t workspace_bytes, void *stream);

/* Test hook: as rbm_b200_block_gibbs (GENERAL variant) but every draw compares
 * against caller-supplied uniforms, laid out as the reference's torch.rand call
 * order: for step s, uniforms_h[s] is [B,nH], uniforms_v[s] is [B,nV]
 * (two dense arrays [k,B,nH] and [k,B,nV]). */
RBM_B200_API int rbm_b200_block_gibbs_injected(const rbm_b200_params *p, float *v, float *h, int64_t B,
                                  int32_t k, const float *uniforms_h, const float *uniforms_v,
                                  void *stream);

/* k mean-field sweeps on probabilities, no draws: GibbsSampler.gibbs_sampling
 * (gibbsSampler.py:31-37).  left[B,nV] in/out, right[B,nH] out. */
RBM_B200_API int rbm_b200_meanfield(const rbm_b200_params *p, float *left, float *right, int64_t B,
                       int32_t k, void *stream);

/* Per-sample energy E[b] = -v.a - h.c - sum_j (vW)_j h_j : RBM.energy
 * (models/rbm/rbm.py:56-72).  v[B,nV], h[B,nH] may be real-valued.
 * workspace: rbm_b200_energy_workspace(...) bytes. */
RBM_B200_API size_t rbm_b200_energy_workspace(int32_t n_visible, int32_t n_hidden, int64_t B);
RBM_B200_API int rbm_b200_energy(const rbm_b200_params *p, const float *v, const float *h, float *E,
                    int64_t B, void *workspace, size_t workspace_bytes, void *stream);

/* Negative-phase statistics of a batch of chain states, the quantities autograd
 * extracts from GumBolt.energy_exp / DiVAEPP.kl_divergence
 * (models/autoencoders/gumbolt.py:102-134, dvaepp.py:141-159):
 *   S_vh[nV,nH] = scale * sum_b v_b h_b^T,  S_v[nV] = scale * sum_b v_b,
 *   S_h[nH] = scale * sum_b h_b.  Pass scale = 1/B_global so that a sum-allreduce
 *   over chain shards yields the global means.  Outputs are overwritten. */
RBM_B200_API int rbm_b200_negative_phase(const float *v, const float *h, int64_t B, int32_t n_visible,
                            int32_t n_hidden, float scale, float *S_vh, float *S_v, float *S_h,
                            void *stream);

/* The PCD negative phase in ONE call: rbm_b200_block_gibbs followed, on the same stream, by the
 * statistics of the returned (v, h) pair (rbm_b200_negative_phase semantics, outputs overwritten).
 * For priors of at most 64 x 64 units (every prior the reference's model configs build except
 * divae.yaml and the *_deep variant) the statistics are computed INSIDE the final sweep of the
 * persistent shared-memory kernel from the on-chip states (v^T h by tensor-core MMA, raw counts
 * added atomically, scaled once).  On the UMMA variant v^T h is a tcgen05 GEMM over the sampler's bf16
 * state tiles (chains split over CTAs, integer counts added with red.global.add.f32: exact for at most
 * 2^24 chains per call, more is refused); other shapes run the fp32 statistics kernels after the chain.
 * energy_exp (device scalar, may be NULL; needs the fp32 W): -(<W,S_vh> + a.S_v + c.S_h), i.e. the mean
 * energy of the batch when scale = 1/B — the value GumBolt.energy_exp returns for these samples.
 * Replaces gumbolt.py:102-104 / dvaepp.py:141-159 (sampler call + energy_exp on its samples). */
RBM_B200_API int rbm_b200_block_gibbs_stats(const rbm_b200_params *p, float *v, float *h, int64_t B,
                                            int32_t k, int init_chains, uint64_t seed, uint64_t chain_offset,
                                            uint64_t step_offset, int variant, void *workspace,
                                            size_t workspace_bytes, float scale, float *S_vh, float *S_v,
                                            float *S_h, float *energy_exp, void *stream);

/* One CD-k parameter update of a vanilla RBM, in place: Contrastive_Divergence.contrastive_divergence
 * (sandbox/rbm_standalone.py:66-121; the in-tree models/samplers/contrastiveDivergence.py does not parse).
 *   ph0 = sigma(v0 W + c); h0 = (ph0 >= u_0); k x { pv = sigma(h W^T + a); ph = sigma(pv W + c); h = (ph >= u_s) }
 *   dW <- momentum dW + (v0^T h0 - pv^T ph) - weight_cost W;  W += lr dW;  biases likewise with
 *   sum_b (v0 - pv) and sum_b (ph0 - ph);  *loss = sum (v0 - pv)^2.   All fp32 (pv, ph are probabilities).
 *   uniforms  NULL: independent Philox draws per unit and chain (draw counters draw_offset .. +k-1);
 *             else [k+1, n_hidden] fp32 rows, each broadcast over the batch like the reference's
 *             torch.rand(n_hidden) (rbm_standalone.py:72,84).
 *   W, vbias, hbias, dW, dvb, dhb  updated in place;  loss: device scalar;  k >= 1.
 * Batches of at most 1024 rows run as ONE cooperative kernel launch (phases separated by grid barriers),
 * larger ones as 3k + 4 launches; same arithmetic either way.
 * workspace: rbm_b200_cd_workspace(...) bytes. */
RBM_B200_API size_t rbm_b200_cd_workspace(int32_t n_visible, int32_t n_hidden, int64_t B);
RBM_B200_API int rbm_b200_cd_step(float *W, float *vbias, float *hbias, float *dW, float *dvb, float *dhb,
                                  int32_t n_visible, int32_t n_hidden, const float *v0, int64_t B, int32_t k,
                                  float learning_rate, float momentum, float weight_cost, const float *uniforms,
                                  uint64_t seed, uint64_t draw_offset, float *loss, void *workspace,
                                  size_t workspace_bytes, void *stream);

/* Annealed importance sampling of log Z (Salakhutdinov & Murray 2008, RBM form).  The
 * reference has NO estimator: RBM.get_logZ_value returns the literal 33.65
 * (models/rbm/rbm.py:51-54); this entry supplies the missing computation.
 * Base-rate model A = (W = 0, c = 0, visible bias a).  For chain b of this shard:
 *   v ~ p_A;  for k = 1..K:  logw[b] += sum_j softplus(beta_k x_j) - softplus(beta_{k-1} x_j),
 *   x = v W + c;  h ~ sigma(beta_k x);  v ~ sigma(a + beta_k h W^T)
 * and  log Z = n_hidden*log 2 + sum_i softplus(a_i) + logsumexp_b(logw) - log(total chains),
 * the logsumexp running over all shards (NCCL on the caller's side).
 *   betas   HOST pointer, n_betas doubles, betas[0] = 0 < ... < betas[n_betas-1] = 1.
 *   logw    DEVICE pointer, n_chains doubles (nats), overwritten.
 *   chain_offset  global id of this shard's first chain (draws are keyed by global id).
 * Needs rbm_b200_ais_workspace(...) bytes of workspace and either W_bf16 (W rounded to bf16) or, with
 * W_bf16 == NULL, the fp32 W, which is then carried as bf16 high + low parts (fp32-grade
 * pre-activations at about twice the tensor time; see RBM_B200_VARIANT_UMMA_SPLIT).
 * The whole schedule runs as ONE persistent launch (up to 2^18 betas; longer schedules fall back to
 * one launch per half-step, W_bf16 only). */
RBM_B200_API size_t rbm_b200_ais_workspace(int32_t n_visible, int32_t n_hidden, int64_t n_chains);
RBM_B200_API int rbm_b200_ais_run(const rbm_b200_params *p, int64_t n_chains, const double *betas,
                                  int32_t n_betas, uint64_t seed, uint64_t chain_offset, double *logw,
                                  void *workspace, size_t workspace_bytes, void *stream);
/* Same, with the weight representation chosen explicitly: variant = RBM_B200_VARIANT_UMMA (or AUTO)
 * rounds W to bf16 (from W_bf16 if given, else from the fp32 W inside the library);
 * RBM_B200_VARIANT_UMMA_SPLIT carries the fp32 W as bf16 high + low parts. */
RBM_B200_API int rbm_b200_ais_run_ex(const rbm_b200_params *p, int64_t n_chains, const double *betas,
                                     int32_t n_betas, uint64_t seed, uint64_t chain_offset, double *logw,
                                     int variant, void *workspace, size_t workspace_bytes, void *stream);

/* ---- next row: the decoder MLP applied to generated prior samples ------------
 * One fully connected layer y = x W^T + b in torch.nn.Linear layout
 * (models/networks/networks.py:33-35: nn.Linear(node[0], node[1])). */
typedef struct rbm_b200_linear {
  int32_t in_features, out_features;
  const float *weight; /* [out_features, in_features] row-major fp32, device */
  const float *bias;   /* [out_features] fp32, device */
} rbm_b200_linear;

/* A decoder: `n_trunk` shared layers, then zero or two heads of `n_head` layers each.  Hidden
 * activation ReLU (activation_fct: relu in the configs/model yaml files), last layer of a chain linear.
 *   BasicDecoder   (models/networks/basicCoders.py:28-42):  n_head = 0, trunk = _layers
 *   BasicDecoderV3 (basicCoders.py:63-84): trunk = _layers (all ReLU), head_hits = _layers2,
 *                  head_act = _layers3
 * The arrays of rbm_b200_linear are HOST memory; the parameters they point to are device memory. */
typedef struct rbm_b200_decoder {
  int32_t n_trunk, n_head;
  const rbm_b200_linear *trunk;
  const rbm_b200_linear *head_hits; /* NULL when there is no hits head */
  const rbm_b200_linear *head_act;  /* NULL when n_head = 0 */
} rbm_b200_decoder;

/* Replaces `self.decoder(prior_samples)` and the shower tail of generate_samples
 * (models/autoencoders/gumboltCaloV5.py:88-93, gumboltCalo.py:69-90, dvaepp.py:221-223):
 *   x        [B, trunk[0].in_features] fp32: cat([rbm_vis, rbm_hid, true_e], 1) or any latent input
 *   out_hits [B, n_out] head_hits output (logits), or NULL
 *   out_act  [B, n_out] head_act output - or the single chain's output when n_head = 0 -, or NULL
 *   samples  [B, n_out] relu(out_act) * [sigmoid(out_hits) >= u], or NULL: the reference's
 *            energy_activation(output_activations) * GumbelMod(output_hits, beta, is_training=False)
 *            (utils/dists/gumbelmod.py:18-24: heaviside(logits + log rho - log(1-rho)) with rho ~ U(0,1)
 *            fires with probability sigmoid(logits); u = 1 - rho).  u comes from the library's Philox
 *            contract: key seed, counter (unit block of the output unit, chain_offset + row, step, tag 4).
 * Every layer is one tcgen05 GEMM (bf16 operands split in high + low parts, three products, fp32
 * accumulation: relative error ~2^-16 against the fp32 reference).  Needs an sm_100 device.
 * workspace: rbm_b200_decoder_workspace(...) bytes (0 for an invalid description).  The staged (split,
 * padded) parameters sit at the start of the workspace at offsets that do not depend on B:
 *   params_staged = 0  stage them from the fp32 parameters (always correct);
 *   params_staged = 1  the caller guarantees that this workspace was last used by a params_staged = 0
 *                      call with the same description and that the parameters have not changed since
 *                      (generation loops: saves 2 small launches per layer). */
RBM_B200_API size_t rbm_b200_decoder_workspace(const rbm_b200_decoder *d, int64_t B);
RBM_B200_API int rbm_b200_decoder_forward(const rbm_b200_decoder *d, const float *x, int64_t B, float *out_hits,
                                          float *out_act, float *samples, uint64_t seed, uint64_t chain_offset,
                                          uint64_t step, int params_staged, void *workspace,
                                          size_t workspace_bytes, void *stream);

/* ---- next row: the posterior side, inference forward ----------------------------
 * Clamp + Gumbel smoothing (training) or gate (evaluation) of one hierarchy level's logits, replacing
 *   logits = torch.clamp(current_net(current_input), min=-88., max=88.)        hierarchicalEncoder.py:167
 *   samples = self.smoothing_dist_mod(logits, beta, is_training)               hierarchicalEncoder.py:176
 * with GumbelMod (utils/dists/gumbelmod.py:14-24): rho ~ U(0,1);
 *   is_training: samples = sigmoid((logits + log(rho) - log(1-rho)) * beta)
 *   else:        samples = heaviside(logits + log(rho) - log(1-rho), 0) = [sigmoid(logits) >= 1 - rho]
 * The level's MLP itself runs through rbm_b200_decoder_forward (a chain with n_head = 0).
 *   logits  [B, n] fp32, clamped IN PLACE when clamp != 0;   samples [B, n] fp32, overwritten.
 *   rho = 1 - (u32 + 0.5) / 2^32 with u32 from the Philox contract: key seed, counter (unit block of
 *   unit_offset + column, chain_offset + row, step, tag 5) - pass the level's first latent index as unit_offset. */
RBM_B200_API int rbm_b200_gumbel_smooth(float *logits, float *samples, int64_t B, int32_t n, float beta,
                                        int is_training, int clamp, uint64_t seed, uint64_t chain_offset,
                                        uint64_t step, uint32_t unit_offset, void *stream);

/* ---- backward passes of the rows next to the path --------------
 * The reference gets these gradients from torch autograd, i.e. from ATen matmuls:
 *   energy_exp on posterior samples          models/autoencoders/gumbolt.py:97-99,118-132
 *   mean-field half-steps                    models/samplers/gibbsSampler.py:20-28
 *   nn.Linear layers of the encoder/decoder  models/networks/hierarchicalEncoder.py:139-183, basicCoders.py:28-84
 * Here every product is a tcgen05 GEMM with fp32-grade accuracy: both operands are split into bf16 high + low
 * parts and three partial products (hi hi + hi lo + lo hi, ~2^-16 relative) accumulate in fp32 in TMEM.
 *
 * rbm_b200_gemm:    C [M,N] = act(A [M,K] . op(B) + bias)
 *   trans_b = 1: op(B) = B^T, B is [N,K] row-major (y = x W^T + b: an nn.Linear forward; dv = dE h W^T)
 *   trans_b = 0: op(B) = B,   B is [K,N] row-major (dx = dy W; dh = dE v W)
 *   bias [N] or NULL; relu != 0 applies max(0, .).  A, B, C fp32 row-major without padding.
 * rbm_b200_gemm_tn: C [M,N] = A^T B with A [kdim,M], B [kdim,N] row-major (dW = dy^T x; dW = -(v g)^T h);
 *   the K splits add their partial tiles with fp32 atomics: the summation order is not fixed. */
RBM_B200_API size_t rbm_b200_gemm_workspace(int64_t M, int32_t N, int32_t K);
RBM_B200_API int rbm_b200_gemm(int trans_b, int64_t M, int32_t N, int32_t K, const float *A, const float *B,
                  const float *bias, int relu, float *C, void *workspace, size_t workspace_bytes, void *stream);
RBM_B200_API size_t rbm_b200_gemm_tn_workspace(int64_t kdim, int32_t M, int32_t N);
RBM_B200_API int rbm_b200_gemm_tn(const float *A, const float *B, int64_t kdim, int32_t M, int32_t N, float *C,
                     void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RBM_B200_H */
