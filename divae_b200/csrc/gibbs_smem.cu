#include "common.cuh"
#include "variants.h"
namespace rbm {
bool smem_variant_supported(int, int) { return false; }
size_t smem_workspace_bytes(int, int, int64_t) { return 0; }
int smem_block_gibbs(const rbm_b200_params*, float*, float*, int64_t, int, int, uint64_t, uint64_t, uint64_t, void*, size_t, cudaStream_t) {
  set_error("SMEM variant not built"); return RBM_B200_ERR_UNSUPPORTED;
}
}
