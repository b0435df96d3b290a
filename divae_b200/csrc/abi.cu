// extern "C" entry points of librbm_b200.so (include/rbm_b200.h) and variant dispatch.
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "philox.cuh"
#include "variants.h"

namespace rbm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

int get_device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  static bool valid[64] = {};
  static std::mutex mu;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  std::lock_guard<std::mutex> lock(mu);
  if (dev >= 0 && dev < 64 && valid[dev]) { *out = cache[dev]; return 0; }
  DeviceInfo d{};
  d.device = dev;
  RBM_CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
  RBM_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  RBM_CUDA_TRY(cudaDeviceGetAttribute(&d.l2_bytes, cudaDevAttrL2CacheSize, dev));
  if (dev >= 0 && dev < 64) { cache[dev] = d; valid[dev] = true; }
  *out = d;
  return 0;
}

int ensure_dynamic_smem_impl(const void* kern, int device, int bytes) {
  struct Entry { const void* kern; int device; int bytes; };
  static Entry table[256];
  static int n = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  Entry* hit = nullptr;
  for (int i = 0; i < n; ++i)
    if (table[i].kern == kern && table[i].device == device) { hit = &table[i]; break; }
  if (hit && hit->bytes >= bytes) return 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  if (hit) hit->bytes = bytes;
  else if (n < 256) table[n++] = Entry{kern, device, bytes};
  return 0;
}

static int check_params(const rbm_b200_params* p, bool need_f32) {
  RBM_REQUIRE(p != nullptr, RBM_B200_ERR_INVALID, "params is NULL");
  RBM_REQUIRE(p->n_visible > 0 && p->n_hidden > 0, RBM_B200_ERR_INVALID, "bad RBM shape %d x %d",
              p->n_visible, p->n_hidden);
  RBM_REQUIRE(p->vbias && p->hbias, RBM_B200_ERR_INVALID, "bias pointer is NULL");
  if (need_f32) RBM_REQUIRE(p->W != nullptr, RBM_B200_ERR_INVALID, "fp32 W is required by this entry");
  return 0;
}

static int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available (%s): librbm_b200 has no CPU fallback", cudaGetErrorString(e));
    return RBM_B200_ERR_NO_DEVICE;
  }
  return 0;
}

// Shape-based kernel selection (one CUDA library, no backend dispatch).  Measured on B200
// (profiles/r01_all_configs_1gpu.jsonl): the persistent shared-memory kernel wins for priors up to
// 128 units per side at any chain count (64x64: 9.0e11 vs 3.0e11 unit-updates/s) and for up to
// 256 units when chains are few (256x256 x 1024 chains: 0.11 ms vs 0.54 ms per call); with many
// chains the tcgen05 GEMM wins at 256 units (7.0e11 vs 5.6e11).
int pick_variant(const rbm_b200_params* p, int64_t B, int variant) {
  if (variant == RBM_B200_VARIANT_UMMA_SPLIT) return RBM_B200_VARIANT_UMMA;
  if (variant != RBM_B200_VARIANT_AUTO) return variant;
  const int nmax = p->n_visible > p->n_hidden ? p->n_visible : p->n_hidden;
  // (the bf16 kernels round the fp32 W themselves when no bf16 copy is handed in)
  if (smem_variant_supported(p->n_visible, p->n_hidden) && (nmax <= 128 || B < 8192)) return RBM_B200_VARIANT_SMEM;
  // Any chain count: a shard of a few chains must sample with the same bf16-rounded W as the job it belongs to
  // (the GENERAL path keeps W in fp32), see divae_b200/dist.py.
  if (umma_variant_supported(p->n_visible, p->n_hidden) && B >= 1) return RBM_B200_VARIANT_UMMA;
  if (smem_variant_supported(p->n_visible, p->n_hidden)) return RBM_B200_VARIANT_SMEM;
  return RBM_B200_VARIANT_GENERAL;
}

}  // namespace rbm

using namespace rbm;

extern "C" {

int rbm_b200_abi_version(void) { return RBM_B200_ABI_VERSION; }

const char* rbm_b200_last_error(void) { return g_err; }

int rbm_b200_device_info(int32_t* sm_count, int32_t* smem_optin, int32_t* cc) {
  int rc = require_device();
  if (rc) return rc;
  int dev = 0, v = 0, maj = 0, min = 0;
  RBM_CUDA_TRY(cudaGetDevice(&dev));
  RBM_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = v;
  RBM_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem_optin) *smem_optin = v;
  RBM_CUDA_TRY(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  RBM_CUDA_TRY(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (cc) *cc = maj * 10 + min;
  return 0;
}

int rbm_b200_half_step(const rbm_b200_params* p, int direction, const float* in, float* out, int64_t B,
                       int mode, const float* uniforms, uint64_t seed, uint64_t chain_offset,
                       uint64_t half_step, void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  RBM_REQUIRE(direction == 0 || direction == 1, RBM_B200_ERR_INVALID, "direction must be 0 or 1");
  RBM_REQUIRE(B >= 0, RBM_B200_ERR_INVALID, "negative batch");
  RBM_REQUIRE(B == 0 || (in && out), RBM_B200_ERR_INVALID, "in/out is NULL");
  RBM_REQUIRE(mode != RBM_B200_MODE_INJECTED || uniforms || B == 0, RBM_B200_ERR_INVALID,
              "injected mode needs uniforms");
  if ((rc = require_device())) return rc;
  return general_half_step(p, direction, in, out, B, mode, uniforms, seed, chain_offset, half_step,
                           TAG_GIBBS, (cudaStream_t)stream);
}

int rbm_b200_init_chains(float* v, int64_t B, int32_t n_visible, uint64_t seed, uint64_t chain_offset,
                         uint64_t step_offset, void* stream) {
  RBM_REQUIRE(B >= 0 && n_visible > 0, RBM_B200_ERR_INVALID, "bad shape");
  RBM_REQUIRE(B == 0 || v, RBM_B200_ERR_INVALID, "v is NULL");
  int rc = require_device();
  if (rc) return rc;
  return general_init_chains(v, B, n_visible, seed, chain_offset, step_offset, (cudaStream_t)stream);
}

size_t rbm_b200_block_gibbs_workspace(int32_t n_visible, int32_t n_hidden, int64_t B, int variant) {
  if (B <= 0) return 0;
  rbm_b200_params q{};
  q.n_visible = n_visible; q.n_hidden = n_hidden;
  switch (pick_variant(&q, B, variant)) {
    case RBM_B200_VARIANT_UMMA: return umma_workspace_bytes(n_visible, n_hidden, B);
    case RBM_B200_VARIANT_SMEM: return smem_workspace_bytes(n_visible, n_hidden, B);
    default: return 0;
  }
}

const char* rbm_b200_kernel_choice(int32_t n_visible, int32_t n_hidden, int64_t B, int variant) {
  if (B <= 0 || n_visible <= 0 || n_hidden <= 0) return "none";
  if (require_device()) return "unknown (no device)";
  rbm_b200_params q{};
  q.n_visible = n_visible; q.n_hidden = n_hidden;
  switch (pick_variant(&q, B, variant)) {
    case RBM_B200_VARIANT_UMMA: return umma_kernel_choice(n_visible, n_hidden, B, variant == RBM_B200_VARIANT_UMMA_SPLIT);
    case RBM_B200_VARIANT_SMEM: return "smem_gibbs_kernel";
    default: return "general_half_step_kernel";
  }
}

size_t rbm_b200_block_gibbs_stats_workspace(int32_t n_visible, int32_t n_hidden, int64_t B, int variant) {
  if (B <= 0 || n_visible <= 0 || n_hidden <= 0) return 0;
  rbm_b200_params q{};
  q.n_visible = n_visible; q.n_hidden = n_hidden;
  switch (pick_variant(&q, B, variant)) {
    case RBM_B200_VARIANT_UMMA: return umma_workspace_bytes(n_visible, n_hidden, B);
    case RBM_B200_VARIANT_SMEM: return smem_stats_workspace_bytes(n_visible, n_hidden, B);
    default: return 0;
  }
}

int64_t rbm_b200_block_gibbs_launch_count(int32_t n_visible, int32_t n_hidden, int64_t B, int32_t k, int variant) {
  if (B <= 0 || n_visible <= 0 || n_hidden <= 0 || k < 0) return 0;
  if (require_device()) return -1;
  rbm_b200_params q{};
  q.n_visible = n_visible; q.n_hidden = n_hidden;
  switch (pick_variant(&q, B, variant)) {
    case RBM_B200_VARIANT_UMMA: return umma_launch_count(n_visible, n_hidden, B, k, variant == RBM_B200_VARIANT_UMMA_SPLIT);
    case RBM_B200_VARIANT_SMEM: return 1;              // one persistent kernel: staging, 2k half-steps, outputs
    default: return 1 + 2 * (int64_t)k;                // GENERAL: initial state + one kernel per half-step
  }
}

static int block_gibbs_impl(const rbm_b200_params* p, float* v, float* h, int64_t B, int32_t k,
                            int init_chains, uint64_t seed, uint64_t chain_offset, uint64_t step_offset,
                            int variant, void* workspace, size_t workspace_bytes, void* stream, float* S_vh,
                            float* S_v, float* S_h, bool* stats_fused) {
  if (stats_fused) *stats_fused = false;
  int rc = check_params(p, false);
  if (rc) return rc;
  RBM_REQUIRE(B >= 0 && k >= 0, RBM_B200_ERR_INVALID, "negative batch or step count");
  RBM_REQUIRE(B == 0 || (v && h), RBM_B200_ERR_INVALID, "state pointer is NULL");
  RBM_REQUIRE(variant >= RBM_B200_VARIANT_AUTO && variant <= RBM_B200_VARIANT_UMMA_SPLIT, RBM_B200_ERR_INVALID,
              "unknown variant %d", variant);
  if ((rc = require_device())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return 0;
  const int chosen = pick_variant(p, B, variant);
  if (chosen == RBM_B200_VARIANT_SMEM) {
    RBM_REQUIRE(p->W_bf16 || p->W, RBM_B200_ERR_INVALID, "SMEM variant needs W (fp32) or W_bf16");
    // up to 64 units: counted inside the sampling kernel; 65 - 256 units: tensor-core GEMM over the bf16 state tiles the
    // final sweep leaves in the workspace (without a workspace: fp32 SIMT statistics of the returned states, below)
    const bool fuse = S_vh != nullptr && k > 0 &&
                      (smem_can_fuse_stats(p->n_visible, p->n_hidden) ||
                       (workspace != nullptr && workspace_bytes >= smem_stats_workspace_bytes(p->n_visible, p->n_hidden, B)));
    if (stats_fused) *stats_fused = fuse;
    return smem_block_gibbs(p, v, h, B, k, init_chains, seed, chain_offset, step_offset, workspace,
                            workspace_bytes, st, fuse ? S_vh : nullptr, fuse ? S_v : nullptr, fuse ? S_h : nullptr);
  }
  if (chosen == RBM_B200_VARIANT_UMMA) {
    const bool split = variant == RBM_B200_VARIANT_UMMA_SPLIT;
    if (split) RBM_REQUIRE(p->W, RBM_B200_ERR_INVALID, "UMMA_SPLIT variant needs the fp32 W");
    else RBM_REQUIRE(p->W_bf16 || p->W, RBM_B200_ERR_INVALID, "UMMA variant needs W (fp32) or W_bf16");
    const bool fuse = S_vh != nullptr && k > 0;
    if (stats_fused) *stats_fused = fuse;
    return umma_block_gibbs(p, v, h, B, k, init_chains, seed, chain_offset, step_offset, workspace,
                            workspace_bytes, st, fuse ? S_vh : nullptr, fuse ? S_v : nullptr, fuse ? S_h : nullptr, split);
  }
  RBM_REQUIRE(p->W, RBM_B200_ERR_INVALID, "GENERAL variant needs fp32 W");
  if (init_chains) {
    if ((rc = general_init_chains(v, B, p->n_visible, seed, chain_offset, step_offset, st))) return rc;
  }
  for (int s = 0; s < k; ++s) {
    if ((rc = general_half_step(p, 0, v, h, B, RBM_B200_MODE_SAMPLE, nullptr, seed, chain_offset,
                                step_offset + 2ull * s, TAG_GIBBS, st))) return rc;
    if ((rc = general_half_step(p, 1, h, v, B, RBM_B200_MODE_SAMPLE, nullptr, seed, chain_offset,
                                step_offset + 2ull * s + 1, TAG_GIBBS, st))) return rc;
  }
  return 0;
}

int rbm_b200_block_gibbs(const rbm_b200_params* p, float* v, float* h, int64_t B, int32_t k,
                         int init_chains, uint64_t seed, uint64_t chain_offset, uint64_t step_offset,
                         int variant, void* workspace, size_t workspace_bytes, void* stream) {
  return block_gibbs_impl(p, v, h, B, k, init_chains, seed, chain_offset, step_offset, variant, workspace,
                          workspace_bytes, stream, nullptr, nullptr, nullptr, nullptr);
}

int rbm_b200_block_gibbs_stats(const rbm_b200_params* p, float* v, float* h, int64_t B, int32_t k,
                               int init_chains, uint64_t seed, uint64_t chain_offset, uint64_t step_offset,
                               int variant, void* workspace, size_t workspace_bytes, float scale, float* S_vh,
                               float* S_v, float* S_h, float* energy_exp, void* stream) {
  RBM_REQUIRE(p != nullptr && S_vh && S_v && S_h, RBM_B200_ERR_INVALID, "NULL statistics pointer");
  RBM_REQUIRE(k >= 1, RBM_B200_ERR_INVALID, "the negative phase needs at least one Gibbs step (h is undefined for k = 0)");
  int rc = require_device();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // zero first: the fused path ADDS raw counts inside the sampling kernel
  if ((rc = stats_zero(S_vh, S_v, S_h, p->n_visible, p->n_hidden, st))) return rc;
  bool fused = false;
  if ((rc = block_gibbs_impl(p, v, h, B, k, init_chains, seed, chain_offset, step_offset, variant, workspace,
                             workspace_bytes, stream, S_vh, S_v, S_h, &fused))) return rc;
  if (fused) rc = stats_scale(S_vh, S_v, S_h, p->n_visible, p->n_hidden, scale, st);
  else rc = general_negative_phase(v, h, B, p->n_visible, p->n_hidden, scale, S_vh, S_v, S_h, st);
  if (rc || !energy_exp) return rc;
  RBM_REQUIRE(p->W != nullptr, RBM_B200_ERR_INVALID, "energy_exp needs the fp32 W");
  return stats_energy(p, S_vh, S_v, S_h, energy_exp, st);
}

int rbm_b200_block_gibbs_injected(const rbm_b200_params* p, float* v, float* h, int64_t B, int32_t k,
                                  const float* uniforms_h, const float* uniforms_v, void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  RBM_REQUIRE(B >= 0 && k >= 0, RBM_B200_ERR_INVALID, "negative batch or step count");
  RBM_REQUIRE(B == 0 || k == 0 || (v && h && uniforms_h && uniforms_v), RBM_B200_ERR_INVALID, "NULL pointer");
  if ((rc = require_device())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (int s = 0; s < k; ++s) {
    if ((rc = general_half_step(p, 0, v, h, B, RBM_B200_MODE_INJECTED,
                                uniforms_h + (size_t)s * B * p->n_hidden, 0, 0, 0, TAG_GIBBS, st))) return rc;
    if ((rc = general_half_step(p, 1, h, v, B, RBM_B200_MODE_INJECTED,
                                uniforms_v + (size_t)s * B * p->n_visible, 0, 0, 0, TAG_GIBBS, st))) return rc;
  }
  return 0;
}

int rbm_b200_meanfield(const rbm_b200_params* p, float* left, float* right, int64_t B, int32_t k,
                       void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  RBM_REQUIRE(B >= 0 && k >= 0, RBM_B200_ERR_INVALID, "negative batch or step count");
  RBM_REQUIRE(B == 0 || (left && right), RBM_B200_ERR_INVALID, "NULL pointer");
  if ((rc = require_device())) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (int s = 0; s < k; ++s) {
    if ((rc = general_half_step(p, 0, left, right, B, RBM_B200_MODE_PROBS, nullptr, 0, 0, 0, TAG_GIBBS, st))) return rc;
    if ((rc = general_half_step(p, 1, right, left, B, RBM_B200_MODE_PROBS, nullptr, 0, 0, 0, TAG_GIBBS, st))) return rc;
  }
  return 0;
}

size_t rbm_b200_energy_workspace(int32_t n_visible, int32_t n_hidden, int64_t B) {
  (void)n_visible;
  if (B <= 0 || n_hidden <= 0) return 0;
  return sizeof(float) * (size_t)B * (size_t)ceil_div(n_hidden, 64);
}

int rbm_b200_energy(const rbm_b200_params* p, const float* v, const float* h, float* E, int64_t B,
                    void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_params(p, true);
  if (rc) return rc;
  RBM_REQUIRE(B >= 0, RBM_B200_ERR_INVALID, "negative batch");
  if (B == 0) return 0;
  RBM_REQUIRE(v && h && E, RBM_B200_ERR_INVALID, "NULL pointer");
  RBM_REQUIRE(workspace && workspace_bytes >= rbm_b200_energy_workspace(p->n_visible, p->n_hidden, B),
              RBM_B200_ERR_WORKSPACE, "energy workspace too small");
  if ((rc = require_device())) return rc;
  return general_energy(p, v, h, E, B, (float*)workspace, (cudaStream_t)stream);
}

int rbm_b200_negative_phase(const float* v, const float* h, int64_t B, int32_t n_visible, int32_t n_hidden,
                            float scale, float* S_vh, float* S_v, float* S_h, void* stream) {
  RBM_REQUIRE(B >= 0 && n_visible > 0 && n_hidden > 0, RBM_B200_ERR_INVALID, "bad shape");
  RBM_REQUIRE(S_vh && S_v && S_h && (B == 0 || (v && h)), RBM_B200_ERR_INVALID, "NULL pointer");
  int rc = require_device();
  if (rc) return rc;
  return general_negative_phase(v, h, B, n_visible, n_hidden, scale, S_vh, S_v, S_h, (cudaStream_t)stream);
}

size_t rbm_b200_cd_workspace(int32_t n_visible, int32_t n_hidden, int64_t B) {
  if (B <= 0 || n_visible <= 0 || n_hidden <= 0) return 0;
  return general_cd_workspace(n_visible, n_hidden, B);
}

int rbm_b200_cd_step(float* W, float* vbias, float* hbias, float* dW, float* dvb, float* dhb, int32_t n_visible,
                     int32_t n_hidden, const float* v0, int64_t B, int32_t k, float learning_rate, float momentum,
                     float weight_cost, const float* uniforms, uint64_t seed, uint64_t draw_offset, float* loss,
                     void* workspace, size_t workspace_bytes, void* stream) {
  RBM_REQUIRE(W && vbias && hbias && dW && dvb && dhb && loss, RBM_B200_ERR_INVALID, "NULL parameter pointer");
  RBM_REQUIRE(n_visible > 0 && n_hidden > 0 && B > 0 && v0, RBM_B200_ERR_INVALID, "bad shape or NULL data");
  RBM_REQUIRE(k >= 1, RBM_B200_ERR_INVALID, "CD-k needs k >= 1");
  RBM_REQUIRE(workspace && workspace_bytes >= general_cd_workspace(n_visible, n_hidden, B), RBM_B200_ERR_WORKSPACE,
              "CD workspace too small");
  int rc = require_device();
  if (rc) return rc;
  return general_cd_step(W, vbias, hbias, dW, dvb, dhb, n_visible, n_hidden, v0, B, k, learning_rate, momentum,
                         weight_cost, uniforms, seed, draw_offset, loss, (float*)workspace, (cudaStream_t)stream);
}

size_t rbm_b200_ais_workspace(int32_t n_visible, int32_t n_hidden, int64_t n_chains) {
  if (n_chains <= 0 || n_visible <= 0 || n_hidden <= 0) return 0;
  return umma_ais_workspace_bytes(n_visible, n_hidden, n_chains);
}

int rbm_b200_ais_run(const rbm_b200_params* p, int64_t n_chains, const double* betas, int32_t n_betas,
                     uint64_t seed, uint64_t chain_offset, double* logw, void* workspace, size_t workspace_bytes,
                     void* stream) {
  int rc = check_params(p, false);
  if (rc) return rc;
  RBM_REQUIRE(p->W_bf16 != nullptr || p->W != nullptr, RBM_B200_ERR_INVALID, "AIS needs W_bf16, or the fp32 W for split weights");
  RBM_REQUIRE(n_chains >= 0, RBM_B200_ERR_INVALID, "negative chain count");
  RBM_REQUIRE(betas != nullptr && n_betas >= 2, RBM_B200_ERR_INVALID, "need at least two betas");
  RBM_REQUIRE(betas[0] == 0.0 && betas[n_betas - 1] == 1.0, RBM_B200_ERR_INVALID, "betas must run from 0 to 1");
  for (int i = 1; i < n_betas; ++i)
    RBM_REQUIRE(betas[i] > betas[i - 1], RBM_B200_ERR_INVALID, "betas must be strictly increasing");
  if (n_chains == 0) return 0;
  RBM_REQUIRE(logw != nullptr, RBM_B200_ERR_INVALID, "logw is NULL");
  if ((rc = require_device())) return rc;
  // W_bf16 == NULL: the fp32 W is carried as bf16 high + low parts (fp32-grade pre-activations, K axis doubled)
  return umma_ais_run(p, n_chains, betas, n_betas, seed, chain_offset, logw, workspace, workspace_bytes,
                      (cudaStream_t)stream, p->W_bf16 == nullptr);
}

int rbm_b200_ais_run_ex(const rbm_b200_params* p, int64_t n_chains, const double* betas, int32_t n_betas,
                        uint64_t seed, uint64_t chain_offset, double* logw, int variant, void* workspace,
                        size_t workspace_bytes, void* stream) {
  int rc = check_params(p, false);
  if (rc) return rc;
  RBM_REQUIRE(variant == RBM_B200_VARIANT_AUTO || variant == RBM_B200_VARIANT_UMMA || variant == RBM_B200_VARIANT_UMMA_SPLIT,
              RBM_B200_ERR_INVALID, "AIS runs on the UMMA kernels: variant must be AUTO, UMMA or UMMA_SPLIT (got %d)", variant);
  const bool split = variant == RBM_B200_VARIANT_UMMA_SPLIT;
  if (split) RBM_REQUIRE(p->W != nullptr, RBM_B200_ERR_INVALID, "split weights need the fp32 W");
  else RBM_REQUIRE(p->W_bf16 != nullptr || p->W != nullptr, RBM_B200_ERR_INVALID, "AIS needs W (fp32) or W_bf16");
  RBM_REQUIRE(n_chains >= 0, RBM_B200_ERR_INVALID, "negative chain count");
  RBM_REQUIRE(betas != nullptr && n_betas >= 2, RBM_B200_ERR_INVALID, "need at least two betas");
  RBM_REQUIRE(betas[0] == 0.0 && betas[n_betas - 1] == 1.0, RBM_B200_ERR_INVALID, "betas must run from 0 to 1");
  for (int i = 1; i < n_betas; ++i)
    RBM_REQUIRE(betas[i] > betas[i - 1], RBM_B200_ERR_INVALID, "betas must be strictly increasing");
  if (n_chains == 0) return 0;
  RBM_REQUIRE(logw != nullptr, RBM_B200_ERR_INVALID, "logw is NULL");
  if ((rc = require_device())) return rc;
  return umma_ais_run(p, n_chains, betas, n_betas, seed, chain_offset, logw, workspace, workspace_bytes,
                      (cudaStream_t)stream, split);
}

size_t rbm_b200_decoder_workspace(const rbm_b200_decoder* d, int64_t B) {
  if (d == nullptr || B < 0) return 0;
  return umma_decoder_workspace_bytes(d, B);
}

int rbm_b200_decoder_forward(const rbm_b200_decoder* d, const float* x, int64_t B, float* out_hits, float* out_act,
                             float* samples, uint64_t seed, uint64_t chain_offset, uint64_t step, int params_staged,
                             void* workspace, size_t workspace_bytes, void* stream) {
  RBM_REQUIRE(d != nullptr, RBM_B200_ERR_INVALID, "decoder description is NULL");
  RBM_REQUIRE(B >= 0, RBM_B200_ERR_INVALID, "negative batch");
  int rc = require_device();
  if (rc) return rc;
  return umma_decoder_forward(d, x, B, out_hits, out_act, samples, seed, chain_offset, step, params_staged, workspace,
                              workspace_bytes, (cudaStream_t)stream);
}

int rbm_b200_gumbel_smooth(float* logits, float* samples, int64_t B, int32_t n, float beta, int is_training, int clamp,
                           uint64_t seed, uint64_t chain_offset, uint64_t step, uint32_t unit_offset, void* stream) {
  RBM_REQUIRE(B >= 0 && n >= 0, RBM_B200_ERR_INVALID, "negative shape");
  RBM_REQUIRE((B == 0 || n == 0) || (logits && samples), RBM_B200_ERR_INVALID, "NULL pointer");
  int rc = require_device();
  if (rc) return rc;
  return gumbel_smooth(logits, samples, B, n, beta, is_training, clamp, seed, chain_offset, step, unit_offset,
                       (cudaStream_t)stream);
}

size_t rbm_b200_gemm_workspace(int64_t M, int32_t N, int32_t K) {
  if (M < 0 || N <= 0 || K <= 0) return 0;
  return umma_gemm_workspace_bytes(M, N, K);
}

int rbm_b200_gemm(int trans_b, int64_t M, int32_t N, int32_t K, const float* A, const float* B, const float* bias, int relu,
                  float* C, void* workspace, size_t workspace_bytes, void* stream) {
  RBM_REQUIRE(M >= 0 && N >= 1 && K >= 1 && N <= 16384 && K <= 16384, RBM_B200_ERR_INVALID,
              "gemm shape out of range: M %lld, N %d, K %d", (long long)M, N, K);
  RBM_REQUIRE(M == 0 || (A && B && C), RBM_B200_ERR_INVALID, "NULL pointer");
  int rc = require_device();
  if (rc) return rc;
  return umma_gemm(trans_b ? 1 : 0, M, N, K, A, B, bias, relu, C, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t rbm_b200_gemm_tn_workspace(int64_t kdim, int32_t M, int32_t N) {
  if (kdim < 0 || M <= 0 || N <= 0) return 0;
  return umma_gemm_tn_workspace_bytes(kdim, M, N);
}

int rbm_b200_gemm_tn(const float* A, const float* B, int64_t kdim, int32_t M, int32_t N, float* C, void* workspace,
                     size_t workspace_bytes, void* stream) {
  RBM_REQUIRE(kdim >= 0 && M >= 1 && N >= 1 && M <= 16384 && N <= 16384, RBM_B200_ERR_INVALID,
              "gemm shape out of range: kdim %lld, M %d, N %d", (long long)kdim, M, N);
  RBM_REQUIRE(C != nullptr && (kdim == 0 || (A && B)), RBM_B200_ERR_INVALID, "NULL pointer");
  int rc = require_device();
  if (rc) return rc;
  return umma_gemm_tn(A, B, kdim, M, N, C, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
