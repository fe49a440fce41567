// tcgen05 / TMA GEMM core (sm_100a) -- declarations.  See gemm_tc.cu.
#pragma once
#include "gemm_common.cuh"

namespace opb {
// Plain mode: fp32 C (+bias) exactly like launch_gemm_simt, computed on the 5th-gen tensor
// cores with the 3-pass fp16-split scheme.  Returns 0, -1 (bad shape) or -2 (CUDA error).
// `timeline` (optional, debug): device buffer [n_ctas][64] of clock64 stamps per CTA.
int launch_gemm_tc_plain(const GemmProblem& p, cudaStream_t stream, long long* timeline = nullptr, int dbg = 0);
}  // namespace opb
