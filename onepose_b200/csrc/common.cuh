// Shared device/host helpers for the GATsSPG matcher kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace opb {

constexpr int kD = 256;        // descriptor_dim
constexpr int kHeads = 4;
constexpr int kDh = 64;        // per-head dim
constexpr int kTileRows = 128; // row tile of every GEMM; segments are padded to it
constexpr float kLoScale = 2048.0f;          // lo plane carries (x - hi) * 2^11
constexpr float kLoInv = 1.0f / 2048.0f;
constexpr float kHalfMax = 65504.0f;

// Activation layout of one forward chunk: frame b owns rows
//   [b*R, b*R + n_pad)           query segment   (N valid)
//   [b*R + n_pad, (b+1)*R)       3D-point segment (M valid)
// with R = n_pad + m_pad, both multiples of kTileRows.  A "segment" is the unit of
// InstanceNorm statistics and of the linear-attention KV state
// (reference GATs_SuperGlue.py:126 and :77-78 reduce over the point dimension).
struct Layout {
  int B, N, M, n_pad, m_pad, R;
  __host__ __device__ int rows() const { return B * R; }
  __host__ __device__ int segs() const { return 2 * B; }
  __host__ __device__ int seg_of_row(int row) const {
    int b = row / R;
    return 2 * b + ((row - b * R) >= n_pad ? 1 : 0);
  }
  __host__ __device__ int seg_start(int seg) const { return (seg >> 1) * R + ((seg & 1) ? n_pad : 0); }
  __host__ __device__ int seg_valid(int seg) const { return (seg & 1) ? M : N; }
  __host__ __device__ int seg_padded(int seg) const { return (seg & 1) ? m_pad : n_pad; }
  // attention source segment: 'self' -> itself, 'cross' -> the other side of the frame
  // (reference GATs_SuperGlue.py:57-58, :62-63)
  __host__ __device__ int src_seg(int seg, int cross) const { return cross ? (seg ^ 1) : seg; }
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// fp32 -> (hi, lo) fp16 planes; x ~= hi + lo * 2^-11 to ~2^-22 relative.
__device__ __forceinline__ void split_f32(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}
__device__ __forceinline__ float join_f32(__half hi, __half lo) {
  return fmaf(__half2float(lo), kLoInv, __half2float(hi));
}
// elu(x) + 1  (reference GATs_SuperGlue.py:71-72)
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x + 1.f : expf(x); }

struct Planes {      // one fp16-split tensor [rows, ld]
  __half* hi;
  __half* lo;
  int ld;
};
struct CPlanes {
  const __half* hi;
  const __half* lo;
  int ld;
};

}  // namespace opb

#define OPB_CUDA_CHECK(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) return opb::fail_cuda(_e, #expr, __FILE__, __LINE__);    \
  } while (0)
