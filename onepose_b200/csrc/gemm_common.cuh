// GEMM problem description shared by the tcgen05 core (product) and the SIMT fp32 core
// (bring-up / cross-check; tests only).  Both consume the SAME fp16-split operands, so the
// SIMT result is an fp32 reference for the tensor-core result on identical inputs.
#pragma once
#include "common.cuh"

namespace opb {

// C[batch][rows, n_out] = A[rows, K1+K2] . B[n_out, K1+K2]^T  (+ bias[n_out])
//   reduction columns [0,K1)      : a1[:, 0:K1]  with b1[n, 0:K1]
//   reduction columns [K1,K1+K2)  : a2[:, 0:K2]  with b2[(seg*n_out + n), 0:K2]   (b2_per_seg)
//                                                 or b2[n, 0:K2]
// where seg = L.seg_of_row(row) -- per-frame/side dynamic weights (the G fold).
// Batched form (score GEMM): batch z uses A rows offset z*a_batch_rows, B rows offset
// z*b_batch_rows (B operand = activations of the same frame), C offset z*c_batch_elems.
struct GemmProblem {
  CPlanes a1, a2, b1, b2;
  int K1, K2;
  int b2_per_seg;
  int rows, n_out;       // per batch; rows % 128 == 0, n_out % 128 == 0
  int batch;
  long long a_batch_rows, b_batch_rows, c_batch_elems;
  Layout L;
  const float* bias;     // [n_out] or nullptr
  int elu_cols;          // output columns [0, elu_cols) get elu(x)+1 after the bias
  float* c;              // fp32 [rows, ldc]
  int ldc;
};

int launch_gemm_simt(const GemmProblem& p, cudaStream_t stream);

}  // namespace opb
