// internal: every kernel TU sees the public C ABI it implements
#pragma once
#include <cuda_runtime.h>
#include "../../include/trtllm_b200.h"

namespace tb {
// gemv_mma.cu: tensor-core decode projection (M <= 8) used by tb_gemv_fused whenever the shape is eligible
bool gemv_mma_eligible(int kind, int M, int K);
int gemv_mma_launch(int kind, void* y, float* y_f32, const void* x, const void* w, const void* w_scale, const float* sc,
                    const float* sr, int sc_per_channel, int sr_per_token, const void* residual, int M, int N, int K,
                    int swiglu, int prologue, const void* gamma, float eps, const void* const* pf, const unsigned* pf_lines,
                    cudaStream_t stream);
// context_attn_tc.cu: tcgen05 / TMEM / TMA causal prefill attention (no scratch)
int launch_flash_ctx_tc(void* out, const void* qkv, void* workspace, const int* input_lengths, int batch, int seq_len,
                        int num_heads, float qk_scale, cudaStream_t stream);
size_t flash_ctx_tc_workspace_bytes(int batch, int seq_len, int num_heads);
}  // namespace tb
