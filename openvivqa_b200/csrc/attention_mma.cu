// placeholder until the tensor-core core lands (next commit)
#include "attention_simt.cuh"
namespace ovqa {
bool sdpa_fwd_mma_supported(const SdpaArgs&) { return false; }
int sdpa_fwd_mma_bf16(const SdpaArgs&, cudaStream_t) { return OVQA_ERR_BAD_SHAPE; }
}
