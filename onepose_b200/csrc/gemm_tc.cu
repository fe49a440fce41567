#include "gemm_tc.cuh"
namespace opb {
int launch_gemm_tc_plain(const GemmProblem&, cudaStream_t) { return -2; }
}
