#include "common.cuh"
#include "variants.h"
namespace rbm {
bool umma_variant_supported(int, int) { return false; }
size_t umma_workspace_bytes(int, int, int64_t) { return 0; }
int umma_block_gibbs(const rbm_b200_params*, float*, float*, int64_t, int, int, uint64_t, uint64_t, uint64_t, void*, size_t, cudaStream_t) {
  set_error("UMMA variant not built"); return RBM_B200_ERR_UNSUPPORTED;
}
}
