// tcgen05 linear-attention state kernel -- declaration.  See kv_state_tc.cu.
#pragma once
#include "common.cuh"

namespace opb {
// kv: fp32 [rows, ld] with K at column k_off and V at v_off (head-contiguous, bias added; k_activated: K already holds elu+1).
// partial: [rows/256][4 heads][64*64 + 64] per-slab sums (KV then Ksum), reduced per segment by kv_state_reduce.
// Returns 0 or -2 (CUDA error).
int launch_kv_state_tc(const float* kv, int ld, int k_off, int v_off, int k_activated, const Layout& L, float* partial, cudaStream_t stream);
// fp16 variant: kvh = fp16 [rows, 512] = 64 * [elu1(K) | V] with pad rows zero (written by the QKV GEMM, EPI_QKV); same partial layout.
int launch_kv_state_h(const __half* kvh, const Layout& L, float* partial, cudaStream_t stream);
}  // namespace opb
