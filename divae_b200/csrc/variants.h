// Internal interface between the C ABI (abi.cu) and the kernel variants.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "../../include/rbm_b200.h"

namespace rbm {

// gibbs_general.cu
int general_half_step(const rbm_b200_params* p, int direction, const float* in, float* out, int64_t B,
                      int mode, const float* U, uint64_t seed, uint64_t chain_offset,
                      uint64_t half_step, uint32_t tag, cudaStream_t st, bool broadcast_u = false,
                      float* probs_out = nullptr);
size_t general_cd_workspace(int nV, int nH, int64_t B);
int general_cd_step(float* W, float* a, float* c, float* dW, float* da, float* dc, int nV, int nH, const float* v0,
                    int64_t B, int k, float lr, float momentum, float weight_cost, const float* uniforms,
                    uint64_t seed, uint64_t draw_offset, float* loss, float* ws, cudaStream_t st);
int general_init_chains(float* v, int64_t B, int nV, uint64_t seed, uint64_t chain_offset,
                        uint64_t step_offset, cudaStream_t st);
int general_energy(const rbm_b200_params* p, const float* v, const float* h, float* E, int64_t B,
                   float* part, cudaStream_t st);
int general_negative_phase(const float* v, const float* h, int64_t B, int nV, int nH, float scale,
                           float* S_vh, float* S_v, float* S_h, cudaStream_t st);
int stats_zero(float* S_vh, float* S_v, float* S_h, int nV, int nH, cudaStream_t st);
int stats_energy(const rbm_b200_params* p, const float* S_vh, const float* S_v, const float* S_h, float* E, cudaStream_t st);
int stats_scale(float* S_vh, float* S_v, float* S_h, int nV, int nH, float scale, cudaStream_t st);

// gibbs_smem.cu — persistent kernel, W (bf16) resident in shared memory
bool smem_variant_supported(int nV, int nH);
size_t smem_workspace_bytes(int nV, int nH, int64_t B);
size_t smem_stats_workspace_bytes(int nV, int nH, int64_t B);   // calls that also return <vh>, <v>, <h>
// S_vh/S_v/S_h non-null (and smem_can_fuse_stats): raw counts of the final (v,h) are atomically
// ADDED to them inside the same kernel (caller zeroes before, scales after).
bool smem_can_fuse_stats(int nV, int nH);
int smem_block_gibbs(const rbm_b200_params* p, float* v, float* h, int64_t B, int k, int init_chains,
                     uint64_t seed, uint64_t chain_offset, uint64_t step_offset, void* ws,
                     size_t ws_bytes, cudaStream_t st, float* S_vh = nullptr, float* S_v = nullptr,
                     float* S_h = nullptr);

// gibbs_umma.cu — tcgen05/TMEM GEMM per half-step fed by TMA
bool umma_variant_supported(int nV, int nH);
size_t umma_workspace_bytes(int nV, int nH, int64_t B);
int64_t umma_launch_count(int nV, int nH, int64_t B, int k, bool split_w);
const char* umma_kernel_choice(int nV, int nH, int64_t B, bool split_w);
// S_vh/S_v/S_h non-null: raw counts of the final (v, h) are ADDED to them (tensor-core V^T H over the
// bf16 state tiles, split over the chains); the caller zeroes before and scales after, as for the SMEM variant.
int umma_block_gibbs(const rbm_b200_params* p, float* v, float* h, int64_t B, int k, int init_chains,
                     uint64_t seed, uint64_t chain_offset, uint64_t step_offset, void* ws,
                     size_t ws_bytes, cudaStream_t st, float* S_vh = nullptr, float* S_v = nullptr,
                     float* S_h = nullptr, bool split_w = false);

// <vh>, <v>, <h> raw counts of bf16 {0,1} state tiles [B, round_up(nV, 64)] / [B, round_up(nH, 64)], ADDED to the outputs:
// tcgen05 GEMM V^T H + column sums (the SMEM variant hands its final states over this way for 65 - 256 units)
int umma_stats_tiles(const void* v_tiles, const void* h_tiles, int64_t B, int nV, int nH, float* S_vh, float* S_v,
                     float* S_h, cudaStream_t st);

// decoder MLP on generated samples: one tcgen05 GEMM per layer, split-bf16 operands
size_t umma_decoder_workspace_bytes(const rbm_b200_decoder* d, int64_t B);
int umma_decoder_forward(const rbm_b200_decoder* d, const float* x, int64_t B, float* out_hits, float* out_act,
                         float* samples, uint64_t seed, uint64_t chain_offset, uint64_t step, int params_staged,
                         void* ws, size_t ws_bytes, cudaStream_t st);

// fp32-grade GEMMs on the tensor cores (split bf16 operands) for the backward passes of the rows next to the path:
//   umma_gemm     C [M,N] = act(A [M,K] op(B) + bias), op(B) = B^T for B [N,K] (trans_b = 1) or B for B [K,N] (trans_b = 0)
//   umma_gemm_tn  C [M,N] = A^T B for A [kdim,M], B [kdim,N]
size_t umma_gemm_workspace_bytes(int64_t M, int N, int K);
int umma_gemm(int trans_b, int64_t M, int N, int K, const float* A, const float* B, const float* bias, int relu, float* C,
              void* ws, size_t ws_bytes, cudaStream_t st);
size_t umma_gemm_tn_workspace_bytes(int64_t kdim, int M, int N);
int umma_gemm_tn(const float* A, const float* B, int64_t kdim, int M, int N, float* C, void* ws, size_t ws_bytes, cudaStream_t st);

// posterior side: clamp + Gumbel smoothing / gate of encoder logits, Philox draws
int gumbel_smooth(float* logits, float* samples, int64_t B, int n, float beta, int is_training, int clamp, uint64_t seed,
                  uint64_t chain_offset, uint64_t step, uint32_t unit_offset, cudaStream_t st);

size_t umma_ais_workspace_bytes(int nV, int nH, int64_t M);
int umma_ais_run(const rbm_b200_params* p, int64_t M, const double* betas, int n_betas, uint64_t seed,
                 uint64_t chain_offset, double* logw, void* ws, size_t ws_bytes, cudaStream_t st, bool split_w = false);

}  // namespace rbm
