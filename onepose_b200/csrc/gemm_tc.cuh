// tcgen05 / TMA GEMM core (sm_100a) -- declarations.  See gemm_tc.cu.
#pragma once
#include "gemm_common.cuh"

namespace opb {
// C (+ fused epilogue, GemmProblem::epi) = A . B^T on the 5th-gen tensor cores with the 3-pass fp16-split
// scheme.  Returns 0, -1 (bad shape / unsupported combination) or -2 (CUDA error).
// `timeline` (optional, debug): device buffer [n_ctas][64] of clock64 stamps per CTA.
int launch_gemm_tc(const GemmProblem& p, cudaStream_t stream, long long* timeline = nullptr);
}  // namespace opb
