// oracle/_ref/libref_cutlass.so: the REFERENCE's own CUTLASS 2.10 GEMMs (compiled from /root/reference where they lie,
// never copied) behind a C ABI of raw device pointers, as the same-box baseline and GPU oracle for the SmoothQuant and
// weight-only projection GEMMs.  TEST / MEASUREMENT INFRASTRUCTURE ONLY — nothing in the product path loads this.
//
// Built as forward-compatible compute_90 PTX: CE/gemm/kernel/fpA_intB_gemm.h:475-487 static-asserts on
// __CUDA_ARCH__ > 900 and the runners reject SM > 90 at run time (int8_gemm_template.h:343-352,
// fpA_intB_gemm_template.h:345-356), so the templated launchers are called directly with arch::Sm80 — the Ampere
// mma.sync kernels the reference would run, JIT-compiled for this GPU.  A few (tile, stages) tactics of the reference's
// own candidate list are instantiated; the caller times them all and keeps the best, as the plugin's tactic profiler does.
//   K/cutlass_kernels/int8_gemm/int8_gemm_template.h:56-172   genericInt8GemmKernelLauncher
//   K/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:49-175 generic_mixed_gemm_kernelLauncher
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tensorrt_llm/kernels/cutlass_kernels/int8_gemm/int8_gemm_template.h"
#include "tensorrt_llm/kernels/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h"

using namespace tensorrt_llm::kernels::cutlass_kernels;
namespace tkc = tensorrt_llm::cutlass_extensions;
namespace tk = tensorrt_llm::common;
using cutlass::gemm::GemmShape;

extern "C" {

int ref_int8_gemm_num_tactics() { return 4; }

// C[m,n] half = (A[m,k] int8 . B[n,k]^T int8) * alphaCol[n] * alphaRow[m]   (per-token + per-channel)
int ref_int8_gemm_half(const int8_t* A, const int8_t* B, const float* alpha_col, const float* alpha_row, void* C, int m,
                       int n, int k, int per_channel, int per_token, int tactic, char* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  tkc::CutlassGemmConfig cfg;
  cfg.split_k_style = tkc::SplitKStyle::NO_SPLIT_K;
  cfg.split_k_factor = 1;
  const tk::QuantOption q = tk::QuantOption::make(per_channel != 0, per_token != 0);
  half* c = reinterpret_cast<half*>(C);
  try {
    switch (tactic) {
      case 0: cfg.stages = 3; genericInt8GemmKernelLauncher<half, cutlass::arch::Sm80, GemmShape<256, 128, 64>, GemmShape<64, 64, 64>, 3>(A, B, q, alpha_col, alpha_row, c, m, n, k, cfg, workspace, workspace_bytes, stream); break;
      case 1: cfg.stages = 3; genericInt8GemmKernelLauncher<half, cutlass::arch::Sm80, GemmShape<128, 256, 64>, GemmShape<64, 64, 64>, 3>(A, B, q, alpha_col, alpha_row, c, m, n, k, cfg, workspace, workspace_bytes, stream); break;
      case 2: cfg.stages = 4; genericInt8GemmKernelLauncher<half, cutlass::arch::Sm80, GemmShape<128, 128, 64>, GemmShape<64, 32, 64>, 4>(A, B, q, alpha_col, alpha_row, c, m, n, k, cfg, workspace, workspace_bytes, stream); break;
      case 3: cfg.stages = 4; genericInt8GemmKernelLauncher<half, cutlass::arch::Sm80, GemmShape<256, 128, 64>, GemmShape<64, 64, 64>, 4>(A, B, q, alpha_col, alpha_row, c, m, n, k, cfg, workspace, workspace_bytes, stream); break;
      default: return -1;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "[ref_int8_gemm_half] %s\n", e.what());
    return -2;
  }
  return (int) cudaGetLastError();
}

int ref_fpA_intB_gemm_num_tactics() { return 6; }

// C[m,n] half = A[m,k] half . dequant(B)   with B in the reference's pre-processed interleaved layout
// (libref_host.so: ref_symmetric_quantize), scales half [n].  bits = 8 | 4.  tactic: (tile, stages, split_k).
int ref_fpA_intB_gemm_half(const void* A, const void* B, const void* scales, void* C, int m, int n, int k, int bits,
                           int tactic, char* workspace, size_t workspace_bytes, cudaStream_t stream) {
  tkc::CutlassGemmConfig cfg;
  const int split_k[6] = {1, 1, 2, 4, 1, 2};
  cfg.split_k_factor = split_k[tactic % 6];
  cfg.split_k_style = cfg.split_k_factor > 1 ? tkc::SplitKStyle::SPLIT_K_SERIAL : tkc::SplitKStyle::NO_SPLIT_K;
  const half* a = reinterpret_cast<const half*>(A);
  const half* s = reinterpret_cast<const half*>(scales);
  half* c = reinterpret_cast<half*>(C);
  using NoBias = tkc::EpilogueOpNoBias;
  try {
#define TB_MIXED(WT, TM, WM, ST)                                                                                       \
  cfg.stages = ST;                                                                                                     \
  generic_mixed_gemm_kernelLauncher<half, WT, cutlass::arch::Sm80, NoBias, GemmShape<TM, 128, 64>,                    \
                                    GemmShape<WM, 32, 64>, ST>(a, reinterpret_cast<const WT*>(B), s, nullptr, c, m, n, \
                                                               k, cfg, workspace, workspace_bytes, stream)
    if (bits == 8) {
      switch (tactic) {
        case 0: TB_MIXED(uint8_t, 32, 32, 3); break;
        case 1: TB_MIXED(uint8_t, 32, 32, 4); break;
        case 2: TB_MIXED(uint8_t, 32, 32, 4); break;
        case 3: TB_MIXED(uint8_t, 32, 32, 4); break;
        case 4: TB_MIXED(uint8_t, 64, 64, 3); break;
        case 5: TB_MIXED(uint8_t, 64, 64, 3); break;
        default: return -1;
      }
    } else if (bits == 4) {
      switch (tactic) {
        case 0: TB_MIXED(cutlass::uint4b_t, 32, 32, 3); break;
        case 1: TB_MIXED(cutlass::uint4b_t, 32, 32, 4); break;
        case 2: TB_MIXED(cutlass::uint4b_t, 32, 32, 4); break;
        case 3: TB_MIXED(cutlass::uint4b_t, 32, 32, 4); break;
        case 4: TB_MIXED(cutlass::uint4b_t, 64, 64, 3); break;
        case 5: TB_MIXED(cutlass::uint4b_t, 64, 64, 3); break;
        default: return -1;
      }
    } else {
      return -1;
    }
#undef TB_MIXED
  } catch (const std::exception& e) {
    fprintf(stderr, "[ref_fpA_intB_gemm_half] %s\n", e.what());
    return -2;
  }
  return (int) cudaGetLastError();
}
}
