// Internal interface of the tensor-core engine (kfb_gemm.cu).
#pragma once
#include "kfb_common.cuh"

namespace kfb {

// D[b] = A[b] * B[b]^T with the epilogue described by `epi` (see include/kfb.h).
// k_splits > 1 cuts the contraction across CTAs (STORE epilogue, fp32 output only; partial
// products are combined with atomic adds, so the output must be accumulated into or zeroed).
// k_splits == 0 lets the engine choose.
int gemm_nt(const kfb_split& A, const kfb_split& B, const kfb_epilogue& epi, int precision,
            int k_splits, cudaStream_t stream);

void count_launch(int n = 1);

}  // namespace kfb
