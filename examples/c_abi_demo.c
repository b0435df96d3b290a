/* Plain-C use of the drop-in boundary (include/rbm_b200.h): no Python, no torch, no C++.
 *
 *   gcc -std=c99 -O2 examples/c_abi_demo.c -Iinclude -I/usr/local/cuda/include \
 *       -L/usr/local/cuda/lib64 -lcudart -ldl -lm -o /tmp/c_abi_demo
 *   /tmp/c_abi_demo divae_b200/librbm_b200.so
 *
 * Samples 4096 chains x 20 block-Gibbs steps of a 96 x 80 RBM with W = 0 through rbm_b200_block_gibbs (AUTO variant)
 * and checks the exact answer for that prior: units are independent Bernoulli(sigmoid(bias)); then checks that the
 * library reports a workspace error (not a crash) for a short workspace.  The library is loaded with dlopen, the way
 * a host application would bind it. */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rbm_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

typedef int (*abi_version_fn)(void);
typedef const char *(*last_error_fn)(void);
typedef size_t (*workspace_fn)(int32_t, int32_t, int64_t, int);
typedef int (*block_gibbs_fn)(const rbm_b200_params *, float *, float *, int64_t, int32_t, int, uint64_t, uint64_t,
                              uint64_t, int, void *, size_t, void *);

static uint16_t bf16_of(float f) {      /* round to nearest even */
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

int main(int argc, char **argv) {
  const char *path = argc > 1 ? argv[1] : "divae_b200/librbm_b200.so";
  void *lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  abi_version_fn abi_version = (abi_version_fn)dlsym(lib, "rbm_b200_abi_version");
  last_error_fn last_error = (last_error_fn)dlsym(lib, "rbm_b200_last_error");
  workspace_fn workspace = (workspace_fn)dlsym(lib, "rbm_b200_block_gibbs_workspace");
  block_gibbs_fn block_gibbs = (block_gibbs_fn)dlsym(lib, "rbm_b200_block_gibbs");
  if (!abi_version || !last_error || !workspace || !block_gibbs) { fprintf(stderr, "missing symbol\n"); return 2; }
  if (abi_version() != 1) { fprintf(stderr, "ABI version %d\n", abi_version()); return 2; }

  enum { NV = 96, NH = 80, B = 4096, K = 20 };
  static float W[NV * NH], a[NV], c[NH], v[B * NV], h[B * NH];
  static uint16_t Wb[NV * NH];
  for (int i = 0; i < NV * NH; ++i) { W[i] = 0.f; Wb[i] = bf16_of(0.f); }
  for (int i = 0; i < NV; ++i) a[i] = -2.f + 4.f * (float)i / (NV - 1);
  for (int j = 0; j < NH; ++j) c[j] = 1.5f - 3.f * (float)j / (NH - 1);

  float *dW, *da, *dc, *dv, *dh;
  uint16_t *dWb;
  CHECK_CUDA(cudaMalloc((void **)&dW, sizeof W));   CHECK_CUDA(cudaMalloc((void **)&dWb, sizeof Wb));
  CHECK_CUDA(cudaMalloc((void **)&da, sizeof a));   CHECK_CUDA(cudaMalloc((void **)&dc, sizeof c));
  CHECK_CUDA(cudaMalloc((void **)&dv, sizeof v));   CHECK_CUDA(cudaMalloc((void **)&dh, sizeof h));
  CHECK_CUDA(cudaMemcpy(dW, W, sizeof W, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dWb, Wb, sizeof Wb, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(da, a, sizeof a, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dc, c, sizeof c, cudaMemcpyHostToDevice));

  rbm_b200_params p;
  p.n_visible = NV; p.n_hidden = NH; p.W = dW; p.W_bf16 = dWb; p.vbias = da; p.hbias = dc;
  int failures = 0;
  for (int variant = 0; variant <= 3; ++variant) {       /* AUTO, GENERAL, SMEM, UMMA */
    size_t ws_bytes = workspace(NV, NH, B, variant);
    void *ws = NULL;
    if (ws_bytes) CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
    cudaStream_t st;
    CHECK_CUDA(cudaStreamCreate(&st));
    int rc = block_gibbs(&p, dv, dh, B, K, /*init_chains=*/1, /*seed=*/0x5EEDull, /*chain_offset=*/0, /*step_offset=*/0,
                         variant, ws, ws_bytes, (void *)st);
    if (rc) { fprintf(stderr, "variant %d: rc %d: %s\n", variant, rc, last_error()); return 1; }
    CHECK_CUDA(cudaStreamSynchronize(st));
    CHECK_CUDA(cudaMemcpy(v, dv, sizeof v, cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(h, dh, sizeof h, cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (int i = 0; i < NV; ++i) {
      double m = 0.0;
      for (int b = 0; b < B; ++b) m += v[b * NV + i];
      const double pr = 1.0 / (1.0 + exp(-(double)a[i]));
      const double z = fabs(m / B - pr) / sqrt(pr * (1 - pr) / B);
      if (z > worst) worst = z;
    }
    for (int j = 0; j < NH; ++j) {
      double m = 0.0;
      for (int b = 0; b < B; ++b) m += h[b * NH + j];
      const double pr = 1.0 / (1.0 + exp(-(double)c[j]));
      const double z = fabs(m / B - pr) / sqrt(pr * (1 - pr) / B);
      if (z > worst) worst = z;
    }
    printf("variant %d: workspace %zu bytes, worst |z| of %d unit means = %.2f\n", variant, ws_bytes, NV + NH, worst);
    if (worst > 4.5) ++failures;
    if (ws_bytes > 1024) {     /* a short workspace is an error code, not a crash */
      rc = block_gibbs(&p, dv, dh, B, K, 1, 1, 0, 0, variant, ws, 1024, (void *)st);
      if (rc == 0) { fprintf(stderr, "variant %d accepted a short workspace\n", variant); ++failures; }
      else printf("variant %d: short workspace -> rc %d (%s)\n", variant, rc, last_error());
    }
    CHECK_CUDA(cudaStreamDestroy(st));
    if (ws) CHECK_CUDA(cudaFree(ws));
  }
  printf(failures ? "FAILED\n" : "c_abi_demo ok\n");
  return failures ? 1 : 0;
}
