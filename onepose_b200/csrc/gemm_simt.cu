// SIMT fp32 GEMM over fp16-split operands: the arithmetic cross-check of the tcgen05 core.
// Reconstructs x = hi + lo*2^-11 in the loader and runs a classic 128x128x16 register-tiled
// FFMA kernel.  Not on the product path: reachable through opb_debug_gemm only (arithmetic cross-check of the tcgen05 core).
#include "gemm_common.cuh"

namespace opb {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, LDS = 132;

__device__ __forceinline__ void load_rowchunk(const __half* hi, const __half* lo, long long off, float (&v)[8]) {
  uint4 uh = *reinterpret_cast<const uint4*>(hi + off);
  uint4 ul = *reinterpret_cast<const uint4*>(lo + off);
  const __half* h = reinterpret_cast<const __half*>(&uh);
  const __half* l = reinterpret_cast<const __half*>(&ul);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = join_f32(h[j], l[j]);
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmProblem p) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];
  const int tid = threadIdx.x;
  const int z = blockIdx.z;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  const long long a_row0 = (long long)z * p.a_batch_rows + row0;
  const long long b_rowz = (long long)z * p.b_batch_rows;
  const int seg = p.b2_per_seg ? p.L.seg_of_row(row0) : 0;
  const int lr = tid >> 1, lk = (tid & 1) * 8;  // loader: row, k-offset
  const int ty = tid >> 4, tx = tid & 15;
  const int K = p.K1 + p.K2;

  float acc[8][8] = {};
  float ra[8], rb[8];

  auto fetch = [&](int k0) {
    if (k0 < p.K1) {
      load_rowchunk(p.a1.hi, p.a1.lo, (a_row0 + lr) * p.a1.ld + k0 + lk, ra);
      load_rowchunk(p.b1.hi, p.b1.lo, (b_rowz + col0 + lr) * p.b1.ld + k0 + lk, rb);
    } else {
      const int kk = k0 - p.K1;
      load_rowchunk(p.a2.hi, p.a2.lo, (a_row0 + lr) * p.a2.ld + kk + lk, ra);
      const long long brow = (p.b2_per_seg ? (long long)seg * p.n_out : 0ll) + col0 + lr;
      load_rowchunk(p.b2.hi, p.b2.lo, brow * p.b2.ld + kk + lk, rb);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      As[buf][lk + j][lr] = ra[j];
      Bs[buf][lk + j][lr] = rb[j];
    }
  };

  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  float* cz = p.c + (long long)z * p.c_batch_elems;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = row0 + (i >> 2) * 64 + ty * 4 + (i & 3);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c = col0 + jh * 64 + tx * 4;
      float4 o = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (p.bias) {
        o.x += p.bias[c]; o.y += p.bias[c + 1]; o.z += p.bias[c + 2]; o.w += p.bias[c + 3];
      }
      *reinterpret_cast<float4*>(cz + (long long)r * p.ldc + c) = o;
    }
  }
}
}  // namespace

int launch_gemm_simt(const GemmProblem& p, cudaStream_t stream) {
  if (p.rows % BM || p.n_out % BN || p.K1 % BK || p.K2 % BK || p.ldc % 4) return -1;
  dim3 grid(p.n_out / BN, p.rows / BM, p.batch);
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace opb
