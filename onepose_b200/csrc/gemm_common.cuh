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
// Epilogue of the tcgen05 core (what happens to the accumulator tile before it leaves the SM):
enum GemmEpilogue {
  EPI_F32 = 0,        // c[rows, ldc] fp32 = acc (+bias) (elu+1 on columns < elu_cols)
  EPI_F32_STATS = 1,  // EPI_F32 + InstanceNorm partial sums (sum, sum^2 per 32-row quarter and column) -> statpart
  EPI_QSCALE = 2,     // Q' = elu1(acc+bias) / (elu1(.) . Kmean_src + 1e-6/m_src) per head -> out planes   (n_out = 256)
  EPI_RESID = 3,      // out planes = resid planes + acc + bias  (x += delta, in place)                     (n_out = 256)
  EPI_L2NORM = 5,     // F.normalize(acc + bias) over the 256 columns -> out planes                          (n_out = 256)
  // dual-softmax tail on the batched score GEMM (one batch = one frame; rows = queries, columns = 3D points):
  EPI_SCORE_SUMS = 6, // e = exp((cos-1)/scale): per-tile row sums and per-32-row column sums -> rowsum_part / colsum_part
  EPI_SCORE_CONF = 7, // conf = e^2 / (rowsum*colsum) -> conf [B,N,M] (optional) + packed row / column arg-max (atomicMax)
  EPI_KV = 8,         // [K | V] projection -> row-major planes out[rows, 512]; elu+1 on K; pad rows zeroed; per-32-row column
                      // sums of K -> statpart [rows/32][256] (K mean of the linear attention)
  EPI_BIAS_PLANES = 10,  // out planes = acc + bias.  The residual GEMM of a layer uses it with the residual INSIDE the reduction:
                      // x_new = [hn | x] . [W1 | I]^T + b  (K2 = 256 identity block; x_hi*64 + x_lo*64 is exact in the fp32 accumulator),
                      // so the epilogue has no global loads -- every chunk of an epilogue ends in a proxy fence that waits for the
                      // thread's outstanding loads, which made the load-x-in-the-epilogue form (EPI_RESID) latency-bound
  EPI_QKV = 9,        // q,k,v projection (n_out = 768, q_tiles = 1): columns [0,256) = Q -> fp32 c[rows, ldc]; columns [256,768) = [K | V]
                      // -> out.hi[rows, 512] = fp16(64 * [elu1(K) | V]), pad rows zero -- the single-pass operands of kv_state_h.
                      // q_tiles = 0: k,v projection only (n_out = 512)
};

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
  // ---- tcgen05 core only ----
  int epi;               // GemmEpilogue
  Planes out;            // planes output [rows, out.ld] (EPI_QSCALE / EPI_RESID / EPI_L2NORM / EPI_KV)
  CPlanes resid;         // EPI_RESID input planes [rows, 256]
  const float* kmean;    // EPI_QSCALE: [S][256]
  int cross;             // EPI_QSCALE: source segment selection
  float* statpart;       // EPI_F32_STATS: [rows/32][n_out][2]
  long long a_batch_k, b_batch_k;   // batched along the reduction dimension: batch z starts at column z*a_batch_k (KV state)
  int mn_major;                     // operands are [K, rows]-shaped row-major tensors (reduction index = tensor ROW): the KV-state
                                    // GEMM  C[256,256] = K_piece^T . V_piece  reads the row-major K/V planes directly (UMMA MN-major)
  // A-operand conversion inside the tcgen05 core (gemm_tc.cu ACV_*): 1 = A1 = ReLU((a_raw - mu_seg) * rstd_seg)  (K1 columns),
  // 2 = A2 = elu1(a_raw) / (elu1(a_raw) . kmean_src + eps/m) per head (K2 = 256 columns); a_raw is fp32 [rows, a_raw_ld]
  int q_tiles;                      // EPI_QKV: leading 256-column n-tiles that are fp32 Q output (1: q,k,v projection; 0: k,v only)
  int a_conv;
  const float* a_raw;
  int a_raw_ld;
  const float* mu;                  // [S][512]
  const float* rstd;
  // EPI_SCORE_*: L gives N, M, n_pad, m_pad
  float inv_scale;
  float* rowsum_part;               // [B][n_out/256][n_pad]
  float* colsum_part;               // [B][rows/32][m_pad]
  const float* inv_rowsum;          // [B][n_pad]
  const float* inv_colsum;          // [B][m_pad]
  float* conf;                      // [B][N][M] or nullptr
  unsigned long long* rowbest;      // [B][N]
  unsigned long long* colbest;      // [B][M]
};

int launch_gemm_simt(const GemmProblem& p, cudaStream_t stream);

}  // namespace opb
