// Bandwidth-bound helper kernels of the GATsSPG matcher (everything that is not a GEMM):
// layout changes, GATs aggregation, linear-attention state, InstanceNorm statistics,
// fp16-split re-packing, dual-softmax/arg-max tail.  sm_100a; plain coalesced
// warp-per-row / block-per-tile kernels sized in multiples of the SM count by the host.
#pragma once
#include "common.cuh"

namespace opb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// Channel-first fp32 [batch][C=256][n] -> point-major rows.  Row r of batch b lands at
// out row  b*out_batch_stride + row_offset + r.   (reference tensors are [B, D, n]:
// GATs_SuperGlue.py:184-186; the device works point-major so that a point is one 1 KB row.)
// MODE 0: fp16-split planes, MODE 1: fp32 rows.
// grid (ceil(n/32), batch), block (32, 8)
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void transpose_cf_to_rows(const float* __restrict__ in, int n, long long in_batch_stride,
                                     __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                                     float* __restrict__ out_f32, long long out_batch_stride_rows,
                                     int row_offset) {
  __shared__ float tile[kD][33];
  const int p0 = blockIdx.x * 32;
  const float* src = in + (long long)blockIdx.y * in_batch_stride;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int c = ty; c < kD; c += 8) {
    int p = p0 + tx;
    tile[c][tx] = (p < n) ? src[(long long)c * n + p] : 0.f;
  }
  __syncthreads();
  const int tid = ty * 32 + tx;  // 256 threads = 256 channels
  const long long row_base = (long long)blockIdx.y * out_batch_stride_rows + row_offset + p0;
  for (int p = 0; p < 32 && p0 + p < n; ++p) {
    float v = tile[tid][p];
    long long o = (row_base + p) * kD + tid;
    if (MODE == 0) {
      __half h, l;
      split_f32(v, h, l);
      out_hi[o] = h;
      out_lo[o] = l;
    } else {
      out_f32[o] = v;
    }
  }
}

// Copy the per-object 3D-point planes into the d-segment of every frame of the chunk.
// grid (ceil(m_pad*256/ (256*8)), B)
__global__ void broadcast_object_rows(const __half* __restrict__ src_hi, const __half* __restrict__ src_lo,
                                      __half* __restrict__ x_hi, __half* __restrict__ x_lo, Layout L) {
  const long long n_vec = (long long)L.m_pad * kD / 8;  // uint4 = 8 halves
  const long long dst0 = ((long long)blockIdx.y * L.R + L.n_pad) * kD / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    reinterpret_cast<uint4*>(x_hi)[dst0 + i] = reinterpret_cast<const uint4*>(src_hi)[i];
    reinterpret_cast<uint4*>(x_lo)[dst0 + i] = reinterpret_cast<const uint4*>(src_lo)[i];
  }
}

// ---------------------------------------------------------------------------------------
// GATs: frame-invariant leaf logits  s2[layer][r] = leaf_row[r] . (W a[:256])_layer
// (reference GATs.py:40,78: wh_2d @ a[:out] == h_2d @ (W a[:out]) in exact arithmetic).
// warp per leaf row, all 4 GATs layers from one read.   wa2: [4][256]
// ---------------------------------------------------------------------------------------
__global__ void gats_leaf_logits(const float* __restrict__ leaves, long long n_rows,
                                 const float* __restrict__ wa2, float* __restrict__ s2) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float w[4][8];
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int j = 0; j < 8; ++j) w[l][j] = wa2[l * kD + (j >> 2) * 128 + lane * 4 + (j & 3)];
  for (long long r = warp; r < n_rows; r += n_warps) {
    const float4* row = reinterpret_cast<const float4*>(leaves + r * kD);
    float4 a = row[lane], b = row[32 + lane];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(v[j], w[l][j], acc);
      acc = warp_sum(acc);
      if (lane == 0) s2[(long long)l * n_rows + r] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------
// GATs aggregation (reference GATs.py:35-88 with with_linear_transform=False):
//   s3 = h3 . (W a[256:]);  e_self = LeakyReLU(2 s3);  e_j = LeakyReLU(s3 + s2_j)
//   att = softmax(e);  h' = att_self h3 + sum_j att_j leaf_j ;  out = ELU(h')
// include_self=0:  h' = sum_j att_j leaf_j / 2 + h3 (GATs.py:64-67); additional: h' += h3 (:61).
// Warp per (point, frame); frames of one point are adjacent warps so the 8 leaf rows are
// served from L1/L2 after the first frame.  Updates the d-segment rows of X in place.
// ---------------------------------------------------------------------------------------
__global__ void gats_aggregate(__half* __restrict__ x_hi, __half* __restrict__ x_lo, Layout L,
                               const float* __restrict__ leaves, int n_leaf,
                               const float* __restrict__ s2 /*[M*n_leaf] this layer*/,
                               const float* __restrict__ wa3 /*[256]*/, int include_self, int additional,
                               float alpha) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total = (long long)L.M * L.B;
  if (warp >= total) return;
  const int b = (int)(warp % L.B);
  const int i = (int)(warp / L.B);
  const long long row = (long long)b * L.R + L.n_pad + i;
  // lane owns channels [lane*4, lane*4+4) and [128+lane*4, 128+lane*4+4)
  float h3[8];
  {
    const uint2* ph = reinterpret_cast<const uint2*>(x_hi + row * kD);
    const uint2* pl = reinterpret_cast<const uint2*>(x_lo + row * kD);
#pragma unroll
    for (int half_i = 0; half_i < 2; ++half_i) {
      uint2 uh = ph[half_i * 32 + lane], ul = pl[half_i * 32 + lane];
      const __half* hh = reinterpret_cast<const __half*>(&uh);
      const __half* hl = reinterpret_cast<const __half*>(&ul);
#pragma unroll
      for (int j = 0; j < 4; ++j) h3[half_i * 4 + j] = join_f32(hh[j], hl[j]);
    }
  }
  float s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s3 = fmaf(h3[j], wa3[(j >> 2) * 128 + lane * 4 + (j & 3)], s3);
  s3 = warp_sum(s3);
  auto lrelu = [alpha](float v) { return v > 0.f ? v : alpha * v; };
  // logits of the leaves: lane j (< n_leaf) holds e_j
  float e = -INFINITY;
  if (lane < n_leaf) e = lrelu(s3 + s2[(long long)i * n_leaf + lane]);
  float e_self = include_self ? lrelu(2.f * s3) : -INFINITY;
  float mx = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  mx = fmaxf(mx, e_self);
  float p = (lane < n_leaf) ? exp_fast(e - mx) : 0.f;
  float p_self = include_self ? exp_fast(e_self - mx) : 0.f;
  float denom = warp_sum(p) + p_self;
  float inv = 1.f / denom;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = include_self ? (p_self * inv) * h3[j] : 0.f;
  // leaves in batches of 4 rows: 8 independent 16-byte loads per lane in flight before the first use
  for (int c0 = 0; c0 < n_leaf; c0 += 4) {
    float4 u[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = min(c0 + k, n_leaf - 1);
      const float4* lr = reinterpret_cast<const float4*>(leaves + ((long long)i * n_leaf + c) * kD);
      u[k] = lr[lane];
      v[k] = lr[32 + lane];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = (c0 + k < n_leaf) ? __shfl_sync(0xffffffffu, p, (c0 + k) & 31) * inv : 0.f;
      acc[0] = fmaf(a, u[k].x, acc[0]); acc[1] = fmaf(a, u[k].y, acc[1]);
      acc[2] = fmaf(a, u[k].z, acc[2]); acc[3] = fmaf(a, u[k].w, acc[3]);
      acc[4] = fmaf(a, v[k].x, acc[4]); acc[5] = fmaf(a, v[k].y, acc[5]);
      acc[6] = fmaf(a, v[k].z, acc[6]); acc[7] = fmaf(a, v[k].w, acc[7]);
    }
  }
  uint2 oh[2], ol[2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = acc[j];
    if (!include_self) v = v * 0.5f + h3[j];
    else if (additional) v += h3[j];
    v = v > 0.f ? v : exp_fast(v) - 1.f;  // ELU (GATs.py:69-70)
    __half h, l;
    split_f32(v, h, l);
    reinterpret_cast<__half*>(&oh[j >> 2])[j & 3] = h;
    reinterpret_cast<__half*>(&ol[j >> 2])[j & 3] = l;
  }
  uint2* qh = reinterpret_cast<uint2*>(x_hi + row * kD);
  uint2* ql = reinterpret_cast<uint2*>(x_lo + row * kD);
  qh[lane] = oh[0]; qh[32 + lane] = oh[1];
  ql[lane] = ol[0]; ql[32 + lane] = ol[1];
}

// Same layer, warp per POINT looping over the frames of the chunk: the point's 8 leaf rows and leaf logits are read once
// into registers and reused for every frame (the leaves are per-object constants; reference GATs.py:46 reshapes the same
// tensor for every batch element).  Fast path for num_leaf == 8 (the released configuration, test_GATsSPG.yaml:5).
constexpr int kGatsFramesPerWarp = 8;
__global__ void __launch_bounds__(256) gats_aggregate_frames8(__half* __restrict__ x_hi, __half* __restrict__ x_lo, Layout L,
                                                              const float* __restrict__ leaves, const float* __restrict__ s2,
                                                              const float* __restrict__ wa3, int include_self, int additional, float alpha) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_groups = (L.B + kGatsFramesPerWarp - 1) / kGatsFramesPerWarp;   // frame groups: more warps in flight than points alone
  const int i = (int)(wid / n_groups), fg = (int)(wid % n_groups);
  if (i >= L.M) return;
  const int b_begin = fg * kGatsFramesPerWarp, b_end = min(L.B, b_begin + kGatsFramesPerWarp);
  float4 lu[8], lv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4* lr = reinterpret_cast<const float4*>(leaves + ((long long)i * 8 + c) * kD);
    lu[c] = lr[lane];
    lv[c] = lr[32 + lane];
  }
  const float s2l = lane < 8 ? s2[(long long)i * 8 + lane] : 0.f;
  float w3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w3[j] = wa3[(j >> 2) * 128 + lane * 4 + (j & 3)];
  auto lrelu = [alpha](float v) { return v > 0.f ? v : alpha * v; };
#pragma unroll 2
  for (int b = b_begin; b < b_end; ++b) {
    const long long row = (long long)b * L.R + L.n_pad + i;
    float h3[8];
    {
      const uint2* ph = reinterpret_cast<const uint2*>(x_hi + row * kD);
      const uint2* pl = reinterpret_cast<const uint2*>(x_lo + row * kD);
#pragma unroll
      for (int half_i = 0; half_i < 2; ++half_i) {
        uint2 uh = ph[half_i * 32 + lane], ul = pl[half_i * 32 + lane];
        const __half* hh = reinterpret_cast<const __half*>(&uh);
        const __half* hl = reinterpret_cast<const __half*>(&ul);
#pragma unroll
        for (int j = 0; j < 4; ++j) h3[half_i * 4 + j] = join_f32(hh[j], hl[j]);
      }
    }
    float s3 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s3 = fmaf(h3[j], w3[j], s3);
    s3 = warp_sum(s3);
    float e = lane < 8 ? lrelu(s3 + s2l) : -INFINITY;
    const float e_self = include_self ? lrelu(2.f * s3) : -INFINITY;
    float mx = e;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));   // lanes 0..7 hold the leaf logits
    mx = fmaxf(__shfl_sync(0xffffffffu, mx, 0), e_self);
    const float p = lane < 8 ? exp_fast(e - mx) : 0.f;
    const float p_self = include_self ? exp_fast(e_self - mx) : 0.f;
    float ps = p;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
    const float inv = 1.f / (__shfl_sync(0xffffffffu, ps, 0) + p_self);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = include_self ? (p_self * inv) * h3[j] : 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float a = __shfl_sync(0xffffffffu, p, c) * inv;
      acc[0] = fmaf(a, lu[c].x, acc[0]); acc[1] = fmaf(a, lu[c].y, acc[1]);
      acc[2] = fmaf(a, lu[c].z, acc[2]); acc[3] = fmaf(a, lu[c].w, acc[3]);
      acc[4] = fmaf(a, lv[c].x, acc[4]); acc[5] = fmaf(a, lv[c].y, acc[5]);
      acc[6] = fmaf(a, lv[c].z, acc[6]); acc[7] = fmaf(a, lv[c].w, acc[7]);
    }
    uint2 oh[2], ol[2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[j];
      if (!include_self) v = v * 0.5f + h3[j];
      else if (additional) v += h3[j];
      v = v > 0.f ? v : exp_fast(v) - 1.f;  // ELU (GATs.py:69-70); absolute error ~1e-7, the order of the fp16 split below
      __half hh, ll;
      split_f32(v, hh, ll);
      reinterpret_cast<__half*>(&oh[j >> 2])[j & 3] = hh;
      reinterpret_cast<__half*>(&ol[j >> 2])[j & 3] = ll;
    }
    uint2* qh = reinterpret_cast<uint2*>(x_hi + row * kD);
    uint2* ql = reinterpret_cast<uint2*>(x_lo + row * kD);
    qh[lane] = oh[0]; qh[32 + lane] = oh[1];
    ql[lane] = ol[0]; ql[32 + lane] = ol[1];
  }
}

// ---------------------------------------------------------------------------------------
// Linear-attention state (reference GATs_SuperGlue.py:71-78), per segment s and head h:
//   Kmean[s][h][d]     = (1/m) sum_rows elu1(K[r,h,d])
//   KVmean[s][h][d][q] = (1/m) sum_rows elu1(K[r,h,d]) * V[r,h,q]
// Input kv: fp32 [rows, ld] with K at column k_off and V at v_off (head-contiguous, bias added).
// Stage 1: block = (slab of kSlabRows rows, head, segment) -> partial sums (deterministic);
// Stage 2: fixed-order reduction over slabs and the 1/m scale.
// ---------------------------------------------------------------------------------------
constexpr int kKVPartial = kDh * kDh + kDh;  // 64x64 KV + 64 Ksum

// grid (row tiles), block 256 = 4 heads x (8 x 8 threads, 8(d) x 8(q) outputs each): one 128-row tile
// -> partial[tile][h][64*64 + 64].  `k_activated`: K already holds elu(k)+1 (fused GEMM epilogue).
__global__ void __launch_bounds__(256) kv_state_partial(const float* __restrict__ kv, int ld, int k_off, int v_off, int k_activated,
                                                        Layout L, float* __restrict__ partial) {
  const int tile = blockIdx.x;
  const int row0 = tile * kTileRows;
  const int seg = L.seg_of_row(row0);
  const int n_valid = min(kTileRows, L.seg_valid(seg) - (row0 - L.seg_start(seg)));   // <= 0 for all-pad tiles
  const int tid = threadIdx.x;
  const int h = tid >> 6, ty = (tid >> 3) & 7, tx = tid & 7;
  __shared__ __align__(16) float sK[16][kD];
  __shared__ __align__(16) float sV[16][kD];
  float acc[8][8] = {};
  float ks[8] = {};
  for (int r0 = 0; r0 < n_valid; r0 += 16) {
    // 16 rows x 256 K and V values: thread loads 4 float4 of each
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;          // 0..1023 float4 slots
      const int rr = idx >> 6, c4 = (idx & 63) * 4;
      float4 kq = make_float4(0.f, 0.f, 0.f, 0.f), vq = kq;
      if (r0 + rr < n_valid) {
        const float* rowp = kv + (long long)(row0 + r0 + rr) * ld;
        kq = *reinterpret_cast<const float4*>(rowp + k_off + c4);
        vq = *reinterpret_cast<const float4*>(rowp + v_off + c4);
        if (!k_activated) { kq.x = elu1(kq.x); kq.y = elu1(kq.y); kq.z = elu1(kq.z); kq.w = elu1(kq.w); }
      }
      *reinterpret_cast<float4*>(&sK[rr][c4]) = kq;
      *reinterpret_cast<float4*>(&sV[rr][c4]) = vq;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < 16; ++rr) {
      const float4 k0 = *reinterpret_cast<const float4*>(&sK[rr][h * kDh + ty * 8]);
      const float4 k1 = *reinterpret_cast<const float4*>(&sK[rr][h * kDh + ty * 8 + 4]);
      const float4 v0 = *reinterpret_cast<const float4*>(&sV[rr][h * kDh + tx * 8]);
      const float4 v1 = *reinterpret_cast<const float4*>(&sV[rr][h * kDh + tx * 8 + 4]);
      const float kk[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
      const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        ks[a] += kk[a];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[a][c] = fmaf(kk[a], vv[c], acc[a][c]);
      }
    }
    __syncthreads();
  }
  float* out = partial + ((long long)tile * kHeads + h) * kKVPartial;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    float* o = out + (ty * 8 + a) * kDh + tx * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
  }
  if (tx == 0) {
#pragma unroll
    for (int a = 0; a < 8; ++a) out[kDh * kDh + ty * 8 + a] = ks[a];
  }
}

// Same partial sums on the warp-level tensor-core path (mma.sync m16n8k16, fp16 operands, fp32 accumulate) with
// the fp16 hi/lo split of common.cuh (3 passes), reading the fp32 [K | V] rows the QKV GEMM wrote.  One block = one
// 128-row tile, all 4 heads; warp w owns head w/2 and 32 of its 64 K-channels (2 m16 tiles) x all 64 V-channels.
// K^T and V fragments come straight from row-major smem tiles through ldmatrix.trans.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(smem_row)));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int kKvLd = 264;   // smem row stride in halves (528 B): 16-byte aligned rows, conflict-free ldmatrix
constexpr int kKvRawStages = 2;   // 2 x 32 KB raw + 33 KB planes = 98 KB per block -> 2 blocks per SM
constexpr int kKvRawBytes = 16 * 512 * 4;                                  // one raw stage: 16 rows x [K 256 | V 256] fp32
constexpr int kKvSmemBytes = kKvRawStages * kKvRawBytes + 4 * 16 * kKvLd * 2;   // raw ring + 4 fp16 planes
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Stage flow (16 rows per stage): cp.async ring of raw fp32 rows (global latency hidden 2 stages ahead)
//   -> convert raw -> fp16 hi/lo planes in smem (elu+1 on K, pad rows already zero) -> ldmatrix.trans + 48 MMAs per warp.
__global__ void __launch_bounds__(256, 2) kv_state_partial_mma(const float* __restrict__ kv, int ld, int k_off, int v_off, int k_activated,
                                                               Layout L, float* __restrict__ partial) {
  extern __shared__ __align__(16) uint8_t kv_smem[];
  float* raw = reinterpret_cast<float*>(kv_smem);                                            // [stages][16][512]
  __half (*sKh)[kKvLd] = reinterpret_cast<__half (*)[kKvLd]>(kv_smem + kKvRawStages * kKvRawBytes);
  __half (*sKl)[kKvLd] = sKh + 16;
  __half (*sVh)[kKvLd] = sKh + 32;
  __half (*sVl)[kKvLd] = sKh + 48;
  const int tile = blockIdx.x;
  const int row0 = tile * kTileRows;
  const int seg = L.seg_of_row(row0);
  const int n_valid = min(kTileRows, L.seg_valid(seg) - (row0 - L.seg_start(seg)));   // <= 0 for all-pad tiles
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int h = warp >> 1, mh = (warp & 1) * 32;
  float acc[2][8][4] = {};
  float ks4[4] = {0.f, 0.f, 0.f, 0.f};            // K column sums of this thread's 4 columns (rows tid/128, +2, +4, ...)
  const int n_stages = n_valid > 0 ? (n_valid + 15) / 16 : 0;
  auto issue = [&](int s) {                       // raw rows of stage s -> ring slot s % kKvRawStages (rows past n_valid: zero-filled)
    float* dst = raw + (s % kKvRawStages) * (16 * 512);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;              // 0..2047 float4 slots: row = idx / 128, 128 float4 per row (K 64 | V 64)
      const int rr = idx >> 7, c4 = (idx & 127) * 4;
      const int r = s * 16 + rr;
      if (r < n_valid) {
        const float* rowp = kv + (long long)(row0 + r) * ld;
        cp_async16(dst + rr * 512 + c4, rowp + (c4 < 256 ? k_off + c4 : v_off + c4 - 256));
      } else {
        *reinterpret_cast<float4*>(dst + rr * 512 + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };
  if (n_stages > 0) issue(0);
  if (n_stages > 1) issue(1); else cp_async_commit();
  for (int s = 0; s < n_stages; ++s) {
    cp_async_wait<1>();                           // stage s has landed (one younger group may still be in flight)
    __syncthreads();                              // ... for every thread; also: previous stage's MMAs are done with the planes
    const float* src = raw + (s % kKvRawStages) * (16 * 512);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256;
      const int rr = idx >> 7, c4 = (idx & 127) * 4;
      float4 x = *reinterpret_cast<const float4*>(src + rr * 512 + c4);
      const bool is_k = c4 < 256;
      if (is_k) {
        if (!k_activated && s * 16 + rr < n_valid) { x.x = elu1(x.x); x.y = elu1(x.y); x.z = elu1(x.z); x.w = elu1(x.w); }
        ks4[0] += x.x; ks4[1] += x.y; ks4[2] += x.z; ks4[3] += x.w;
      }
      // packed split: hi = fp16x2(64 x), lo = fp16x2(64 x - hi)
      const float2 a = make_float2(x.x * kPre, x.y * kPre), b = make_float2(x.z * kPre, x.w * kPre);
      const __half2 ha = __float22half2_rn(a), hb = __float22half2_rn(b);
      const float2 fa = __half22float2(ha), fb = __half22float2(hb);
      const __half2 la = __float22half2_rn(make_float2(a.x - fa.x, a.y - fa.y)), lb = __float22half2_rn(make_float2(b.x - fb.x, b.y - fb.y));
      __half* dh = is_k ? &sKh[rr][c4] : &sVh[rr][c4 - 256];
      __half* dl = is_k ? &sKl[rr][c4] : &sVl[rr][c4 - 256];
      __half2 hv[2] = {ha, hb}, lv[2] = {la, lb};
      *reinterpret_cast<uint2*>(dh) = *reinterpret_cast<uint2*>(hv);
      *reinterpret_cast<uint2*>(dl) = *reinterpret_cast<uint2*>(lv);
    }
    __syncthreads();                              // planes complete; raw slot s is free again
    if (s + 2 < n_stages) issue(s + 2); else cp_async_commit();
    // A = K^T (m = K channel, k = row): 16x16 blocks of the row-major K tile, transposed on load
    uint32_t ah[2][4], al[2][4];
    const int a_row = (lane & 7) + 8 * (lane >> 4), a_col = 8 * ((lane >> 3) & 1);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int d0 = h * kDh + mh + mt * 16;
      ldmatrix_x4_trans(ah[mt], &sKh[a_row][d0 + a_col]);
      ldmatrix_x4_trans(al[mt], &sKl[a_row][d0 + a_col]);
    }
    // B = V (k = row, n = V channel): two n8 tiles per ldmatrix.x4.trans
    const int b_row = (lane & 7) + 8 * ((lane >> 3) & 1), b_col = 8 * (lane >> 4);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int q0 = h * kDh + np * 16;
      uint32_t bh[4], bl[4];
      ldmatrix_x4_trans(bh, &sVh[b_row][q0 + b_col]);
      ldmatrix_x4_trans(bl, &sVl[b_row][q0 + b_col]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          float (&d)[4] = acc[mt][np * 2 + s2];
          mma_f16(d, ah[mt], bh[2 * s2], bh[2 * s2 + 1]);
          mma_f16(d, ah[mt], bl[2 * s2], bl[2 * s2 + 1]);
          mma_f16(d, al[mt], bh[2 * s2], bh[2 * s2 + 1]);
        }
      }
    }
  }
  float* out = partial + ((long long)tile * kHeads + h) * kKVPartial;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int d = mh + mt * 16 + g, q = nt * 8 + 2 * t;
      *reinterpret_cast<float2*>(out + d * kDh + q) = make_float2(acc[mt][nt][0] * kProdInv, acc[mt][nt][1] * kProdInv);
      *reinterpret_cast<float2*>(out + (d + 8) * kDh + q) = make_float2(acc[mt][nt][2] * kProdInv, acc[mt][nt][3] * kProdInv);
    }
  // K column sums: threads t and t+128 hold the even / odd rows of the same 4 columns (K columns: (tid & 127) < 64)
  __syncthreads();
  float* red = raw;                               // raw ring is idle now
  if ((tid & 127) < 64) {
#pragma unroll
    for (int e = 0; e < 4; ++e) red[(tid >> 7) * 256 + (tid & 127) * 4 + e] = ks4[e];
  }
  __syncthreads();
  partial[((long long)tile * kHeads + (tid >> 6)) * kKVPartial + kDh * kDh + (tid & 63)] = red[tid] + red[256 + tid];
}

// grid (S*H, 17), block 256: fixed-order sum over the segment's tiles, scaled by 1/m
__global__ void kv_state_reduce(const float* __restrict__ partial, Layout L, int rows_per_partial,
                                float* __restrict__ kvmean /*[S][H][64][64]*/, float* __restrict__ kmean /*[S][H][64]*/) {
  const int sh = blockIdx.x;
  const int seg = sh / kHeads, h = sh % kHeads;
  const int i = blockIdx.y * 256 + threadIdx.x;
  if (i >= kKVPartial) return;
  const int t0 = L.seg_start(seg) / rows_per_partial;
  const int nt = (L.seg_valid(seg) + rows_per_partial - 1) / rows_per_partial;
  float s = 0.f;
  for (int t = 0; t < nt; ++t) s += partial[((long long)(t0 + t) * kHeads + h) * kKVPartial + i];
  s = L.seg_valid(seg) > 0 ? s * (1.f / (float)L.seg_valid(seg)) : 0.f;   // empty segment (object prologue / query-only pass)
  if (i < kDh * kDh) kvmean[(long long)sh * kDh * kDh + i] = s;
  else kmean[(long long)sh * kDh + (i - kDh * kDh)] = s;
}

// tcgen05 path of the linear-attention state (fuse level 2).  The [K | V] projection epilogue (EPI_KV) leaves elu1(K) and V
// as row-major fp16-split planes kv[rows, 512] with pad rows zeroed; one batched GEMM whose reduction index is the tensor
// ROW (MN-major UMMA operands) then yields, per 256-row piece, part[piece][256 (K channel)][256 (V channel)] = K_piece^T V_piece.
// kv_reduce_pieces: fixed-order sum over the pieces of a segment, diagonal head blocks only, 1/m scale; the last
// y-block of each (segment, head) combines the per-32-row K column sums (EPI_KV epilogue) into Kmean.
// grid (S*H, 17), block 256
__global__ void kv_reduce_pieces(const float* __restrict__ part, const float* __restrict__ ksum_part /*[rows/32][256]*/, Layout L,
                                 float* __restrict__ kvmean /*[S][H][64][64]*/, float* __restrict__ kmean /*[S][256]*/) {
  const int sh = blockIdx.x;
  const int seg = sh / kHeads, h = sh % kHeads;
  const int valid = L.seg_valid(seg);
  const float inv_m = valid > 0 ? 1.f / (float)valid : 0.f;
  if (blockIdx.y == 16) {
    if (threadIdx.x < kDh) {
      const int c = h * kDh + threadIdx.x;
      const int q0 = L.seg_start(seg) / 32, nq = (valid + 31) / 32;
      float s = 0.f;
      for (int t = 0; t < nq; ++t) s += ksum_part[(long long)(q0 + t) * 256 + c];
      kmean[seg * kD + c] = s * inv_m;
    }
    return;
  }
  const int i = blockIdx.y * 256 + threadIdx.x;      // d*64 + q
  const int d = i >> 6, q = i & 63;
  const int p0 = L.seg_start(seg) / 256;
  const int np = (valid + 255) / 256;
  float s = 0.f;
  for (int t = 0; t < np; ++t) s += part[((long long)(p0 + t) * 256 + h * kDh + d) * 256 + h * kDh + q];
  kvmean[(long long)sh * kDh * kDh + i] = s * inv_m;
}

// ---------------------------------------------------------------------------------------
// Q' = elu1(q) / (elu1(q) . Kmean_src + 1e-6/m_src)  per head  (GATs_SuperGlue.py:71,78-79
// with the /m, *m of :75,:79 folded into the means).  q: fp32 [rows, ldq] (cols 0..255,
// head-contiguous, bias added).  Output fp16-split planes [rows, 256].  Warp per row.
// ---------------------------------------------------------------------------------------
__global__ void q_scale_split(const float* __restrict__ q, int ldq, int activated, Layout L, int cross,
                              const float* __restrict__ kmean, __half* __restrict__ o_hi, __half* __restrict__ o_lo) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= L.rows()) return;
  const int seg = L.seg_of_row((int)row);
  const int src = L.src_seg(seg, cross);
  const float eps_m = 1e-6f / (float)L.seg_valid(src);
  const float4* qp = reinterpret_cast<const float4*>(q + row * ldq) + lane * 2;   // channels lane*8 .. +7
  const float4* kp = reinterpret_cast<const float4*>(kmean + (long long)src * kD) + lane * 2;
  float4 a = qp[0], b = qp[1], ka = kp[0], kb = kp[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (!activated) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = elu1(v[j]);
  }
  float kk[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) dot = fmaf(v[j], kk[j], dot);
  // head = 8 consecutive lanes
  dot += __shfl_xor_sync(0xffffffffu, dot, 1);
  dot += __shfl_xor_sync(0xffffffffu, dot, 2);
  dot += __shfl_xor_sync(0xffffffffu, dot, 4);
  const float z = 1.f / (dot + eps_m);
  uint4 oh, ol;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __half h, l;
    split_f32(v[j] * z, h, l);
    reinterpret_cast<__half*>(&oh)[j] = h;
    reinterpret_cast<__half*>(&ol)[j] = l;
  }
  reinterpret_cast<uint4*>(o_hi + row * kD)[lane] = oh;
  reinterpret_cast<uint4*>(o_lo + row * kD)[lane] = ol;
}

// ---------------------------------------------------------------------------------------
// Dynamic weight  G[s][c][h*64+d] = sum_q KVmean[src(s)][h][d][q] * W0m[c][h*64+q]
// where W0m = mlp.0.weight[:, 256:] @ merge.weight (folded on the host), so that
//   mlp.0([x ; merge(msg)]) = W0a x + G (Q')  + b   (GATs_SuperGlue.py:101,113,122).
// Output fp16-split planes [S][512][256].
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) g_fold(const float* __restrict__ kvmean, const float* __restrict__ w0m /*[512][256]*/,
                                              Layout L, int cross, __half* __restrict__ g_hi, __half* __restrict__ g_lo) {
  // grid (512/64 c-chunks, heads, S); block 256 = 16x16 threads, 4(c) x 4(d) outputs each
  __shared__ __align__(16) float sKVt[kDh][kDh + 4];   // [q][d]
  __shared__ __align__(16) float sWt[kDh][kDh + 4];    // [q][c]
  const int c0 = blockIdx.x * 64, h = blockIdx.y, seg = blockIdx.z;
  const int src = L.src_seg(seg, cross);
  const int tid = threadIdx.x;
  const float* kvs = kvmean + ((long long)src * kHeads + h) * kDh * kDh;   // [d][q]
  for (int i = tid; i < kDh * kDh; i += 256) {
    int a = i >> 6, b = i & 63;
    sKVt[b][a] = kvs[i];                                        // (d=a, q=b)
    sWt[b][a] = w0m[(long long)(c0 + a) * kD + h * kDh + b];    // (c=a, q=b)
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
#pragma unroll 8
  for (int q = 0; q < kDh; ++q) {
    const float4 w = *reinterpret_cast<const float4*>(&sWt[q][ty * 4]);
    const float4 k = *reinterpret_cast<const float4*>(&sKVt[q][tx * 4]);
    const float wv[4] = {w.x, w.y, w.z, w.w}, kv4[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(wv[a], kv4[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const long long o = ((long long)seg * 512 + c0 + ty * 4 + a) * kD + h * kDh + tx * 4;
    __half hh[4], ll[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) split_f32(acc[a][b], hh[b], ll[b]);
    *reinterpret_cast<uint2*>(g_hi + o) = *reinterpret_cast<uint2*>(hh);
    *reinterpret_cast<uint2*>(g_lo + o) = *reinterpret_cast<uint2*>(ll);
  }
}

// ---------------------------------------------------------------------------------------
// InstanceNorm1d(512) statistics over the valid rows of each segment
// (reference GATs_SuperGlue.py:126: no affine, biased variance, eps 1e-5).
// Stage 1: per 32-row quarter and channel: sum, sum of squares (fp32) -- written by the fused GEMM epilogue
// (EPI_F32_STATS) or by in_stats_partial (grid (tiles, 4), block 128) on the unfused path.
// Stage 2: per (segment, channel): fixed-order fp64 combine -> mean, rstd.
// ---------------------------------------------------------------------------------------
__global__ void in_stats_partial(const float* __restrict__ hid /*[rows,512]*/, Layout L, float* __restrict__ part /*[rows/32][512][2]*/) {
  const int tile = blockIdx.x;
  const int c = blockIdx.y * 128 + threadIdx.x;
  const int row0 = tile * kTileRows;
  const int seg = L.seg_of_row(row0);
  const int n_valid = L.seg_valid(seg) - (row0 - L.seg_start(seg));
  for (int qq = 0; qq < 4; ++qq) {
    float s = 0.f, s2 = 0.f;
    const int r_end = min(32, n_valid - qq * 32);
    for (int i = 0; i < r_end; ++i) {
      float v = hid[(long long)(row0 + qq * 32 + i) * 512 + c];
      s += v;
      s2 = fmaf(v, v, s2);
    }
    part[((long long)(tile * 4 + qq) * 512 + c) * 2 + 0] = s;
    part[((long long)(tile * 4 + qq) * 512 + c) * 2 + 1] = s2;
  }
}

// grid (S, 16), block (32 channels, 8 slices): fixed-order fp64 combine of the 32-row partials of a segment -> mean, rstd
__global__ void in_stats_final(const float* __restrict__ part, Layout L, float* __restrict__ mu, float* __restrict__ rstd) {
  __shared__ double sh[8][32][2];
  const int seg = blockIdx.x;
  const int c = blockIdx.y * 32 + threadIdx.x;
  const int slice = threadIdx.y;
  const int q0 = L.seg_start(seg) / 32;
  const int nq = (L.seg_valid(seg) + 31) / 32;
  double s = 0.0, s2 = 0.0;
  for (int t = slice; t < nq; t += 8) {
    const float2 v = reinterpret_cast<const float2*>(part)[(long long)(q0 + t) * 512 + c];
    s += (double)v.x;
    s2 += (double)v.y;
  }
  sh[slice][threadIdx.x][0] = s;
  sh[slice][threadIdx.x][1] = s2;
  __syncthreads();
  if (slice == 0) {
    s = 0.0; s2 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += sh[k][threadIdx.x][0]; s2 += sh[k][threadIdx.x][1]; }
    const double n = (double)(L.seg_valid(seg) > 0 ? L.seg_valid(seg) : 1);
    const double mean = s / n;
    double var = s2 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    mu[seg * 512 + c] = (float)mean;
    rstd[seg * 512 + c] = (float)(1.0 / sqrt(var + 1e-5));
  }
}

// HN = ReLU((hid - mu) * rstd) -> fp16-split planes [rows, 512].  One thread = 8 channels.
__global__ void norm_relu_split(const float* __restrict__ hid, Layout L, const float* __restrict__ mu,
                                const float* __restrict__ rstd, __half* __restrict__ o_hi, __half* __restrict__ o_lo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over rows*64
  if (idx >= (long long)L.rows() * 64) return;
  const int row = (int)(idx >> 6);
  const int c0 = (int)(idx & 63) * 8;
  const int seg = L.seg_of_row(row);
  const float4* hp = reinterpret_cast<const float4*>(hid + (long long)row * 512 + c0);
  const float4* mp = reinterpret_cast<const float4*>(mu + seg * 512 + c0);
  const float4* rp = reinterpret_cast<const float4*>(rstd + seg * 512 + c0);
  float4 a = hp[0], b = hp[1], ma = mp[0], mb = mp[1], ra = rp[0], rb = rp[1];
  float v[8] = {(a.x - ma.x) * ra.x, (a.y - ma.y) * ra.y, (a.z - ma.z) * ra.z, (a.w - ma.w) * ra.w,
                (b.x - mb.x) * rb.x, (b.y - mb.y) * rb.y, (b.z - mb.z) * rb.z, (b.w - mb.w) * rb.w};
  uint4 oh, ol;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __half h, l;
    split_f32(fmaxf(v[j], 0.f), h, l);
    reinterpret_cast<__half*>(&oh)[j] = h;
    reinterpret_cast<__half*>(&ol)[j] = l;
  }
  reinterpret_cast<uint4*>(o_hi + (long long)row * 512 + c0)[0] = oh;
  reinterpret_cast<uint4*>(o_lo + (long long)row * 512 + c0)[0] = ol;
}

// X <- X + delta  (reference GATs_SuperGlue.py:59,64), delta fp32 [rows,256] bias included.
__global__ void residual_update(__half* __restrict__ x_hi, __half* __restrict__ x_lo, const float* __restrict__ delta,
                                long long n_vec /* rows*256/8 */) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  uint4 uh = reinterpret_cast<uint4*>(x_hi)[i], ul = reinterpret_cast<uint4*>(x_lo)[i];
  const float4* dp = reinterpret_cast<const float4*>(delta) + i * 2;
  float4 a = dp[0], b = dp[1];
  float d[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = join_f32(reinterpret_cast<__half*>(&uh)[j], reinterpret_cast<__half*>(&ul)[j]) + d[j];
    __half h, l;
    split_f32(v, h, l);
    reinterpret_cast<__half*>(&uh)[j] = h;
    reinterpret_cast<__half*>(&ul)[j] = l;
  }
  reinterpret_cast<uint4*>(x_hi)[i] = uh;
  reinterpret_cast<uint4*>(x_lo)[i] = ul;
}

// Range guard of the fp16-split operand format (|x| < 1023): an overflow anywhere in the GNN turns into inf/NaN and
// stays in the residual stream, so scanning the final X hi-plane once per chunk detects it.  Sets *flag = 1.
__global__ void range_check(const __half* __restrict__ x_hi, long long n_vec /* elements / 8 */, int* __restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  const uint4 u = reinterpret_cast<const uint4*>(x_hi)[i];
  const unsigned w[4] = {u.x, u.y, u.z, u.w};
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) bad |= ((w[k] & 0x7C00u) == 0x7C00u) || ((w[k] & 0x7C000000u) == 0x7C000000u);   // exponent all ones: inf / NaN
  if (bad) atomicOr(flag, 1);
}

// planes -> fp32 (debug / tests)
__global__ void join_planes(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = join_f32(hi[i], lo[i]);
}
__global__ void split_planes(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) split_f32(x[i], hi[i], lo[i]);
}

// ---------------------------------------------------------------------------------------
// Tail.  P = F.normalize(final_proj(x)) (GATs_SuperGlue.py:209-213, eps 1e-12) -> split planes.
// Warp per row; proj fp32 [rows,256] bias included.
// ---------------------------------------------------------------------------------------
__global__ void l2_normalize_split(const float* __restrict__ proj, long long rows, __half* __restrict__ o_hi, __half* __restrict__ o_lo) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float4* p = reinterpret_cast<const float4*>(proj + row * kD) + lane * 2;
  float4 a = p[0], b = p[1];
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  uint4 oh, ol;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __half h, l;
    split_f32(v[j] * inv, h, l);
    reinterpret_cast<__half*>(&oh)[j] = h;
    reinterpret_cast<__half*>(&ol)[j] = l;
  }
  reinterpret_cast<uint4*>(o_hi + row * kD)[lane] = oh;
  reinterpret_cast<uint4*>(o_lo + row * kD)[lane] = ol;
}

// Dual softmax with a FIXED shift: the operands are unit vectors so score <= 1/scale; with
//   e[n,m] = exp((cos[n,m] - 1) / scale)   in (exp(-2/scale), 1]
// softmax(scores,1)*softmax(scores,2) (GATs_SuperGlue.py:218) = e^2 / (colsum[m] * rowsum[n]),
// no running max needed.  exp runs on the hardware ex2 unit (__expf, ~4e-7 relative here; the contract is 1e-4 absolute).  v0 (SIMT) path: cos matrix materialised in `s` [B][n_pad][m_pad].
// Row sums: warp per (b, n).
__global__ void score_row_sums(const float* __restrict__ s, Layout L, float inv_scale, float* __restrict__ rowsum /*[B][n_pad]*/) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)L.B * L.N) return;
  const int b = (int)(w / L.N), n = (int)(w % L.N);
  const float* row = s + ((long long)b * L.n_pad + n) * L.m_pad;
  float acc = 0.f;
  for (int m = lane; m < L.M; m += 32) acc += exp_fast((row[m] - 1.f) * inv_scale);
  acc = warp_sum(acc);
  if (lane == 0) rowsum[b * L.n_pad + n] = 1.f / acc;       // stored as the inverse
}
// Column sums: block = 32 columns x 8 row-groups (coalesced 128 B per warp-row), fixed-order combine.
// grid (ceil(M/32), B), block (32, 8)
__global__ void score_col_sums(const float* __restrict__ s, Layout L, float inv_scale, float* __restrict__ colsum /*[B][m_pad]*/) {
  __shared__ float part[8][33];
  const int b = blockIdx.y;
  const int m = blockIdx.x * 32 + threadIdx.x;
  const float* col = s + (long long)b * L.n_pad * L.m_pad + m;
  float acc = 0.f;
  if (m < L.M)
    for (int n = threadIdx.y; n < L.N; n += 8) acc += exp_fast((col[(long long)n * L.m_pad] - 1.f) * inv_scale);
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && m < L.M) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += part[g][threadIdx.x];
    colsum[b * L.m_pad + m] = 1.f / t;                        // stored as the inverse
  }
}

// Fused tail, stage between the two score-GEMM passes: fixed-order sums of the per-tile partials -> 1/rowsum, 1/colsum.
// grid (ceil((n_pad + m_pad)/256), B), block 256
__global__ void score_sums_finalize(const float* __restrict__ rowsum_part, const float* __restrict__ colsum_part, Layout L, int n_tiles,
                                    int q_groups, float* __restrict__ inv_rowsum, float* __restrict__ inv_colsum) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < L.n_pad) {
    float s = 0.f;
    for (int t = 0; t < n_tiles; ++t) s += rowsum_part[((long long)b * n_tiles + t) * L.n_pad + i];
    inv_rowsum[b * L.n_pad + i] = i < L.N ? 1.f / s : 0.f;
  } else if (i - L.n_pad < L.m_pad) {
    const int m = i - L.n_pad;
    const int qv = (L.N + 31) / 32;                     // quarters that hold valid query rows
    float s = 0.f;
    if (m < L.M)
      for (int t = 0; t < qv && t < q_groups; ++t) s += colsum_part[((long long)b * q_groups + t) * L.m_pad + m];
    inv_colsum[b * L.m_pad + m] = m < L.M ? 1.f / s : 0.f;
  }
}

// order-preserving pack of (positive float value, index) with lowest-index-wins ties
__device__ __forceinline__ unsigned long long pack_arg(float v, int idx) {
  return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}

// conf = e^2 / (rowsum*colsum); optional store to conf [B][N][M]; row/col arg-max via packed atomicMax.
// block = 32 rows x 128 cols tile; grid (ceil(M/128), ceil(N/32), B); block 128 threads (thread = column).
__global__ void conf_argmax_simt(const float* __restrict__ s, Layout L, float inv_scale, const float* __restrict__ rowsum,
                                 const float* __restrict__ colsum, float* __restrict__ conf,
                                 unsigned long long* __restrict__ rowbest /*[B][N]*/, unsigned long long* __restrict__ colbest /*[B][M]*/) {
  const int b = blockIdx.z;
  const int m = blockIdx.x * 128 + threadIdx.x;
  const int n0 = blockIdx.y * 32;
  const bool mv = m < L.M;
  const float inv_cs = mv ? colsum[b * L.m_pad + m] : 0.f;      // the sums kernels store inverses
  unsigned long long cbest = 0ull;
  const int lane = threadIdx.x & 31;
  for (int n = n0; n < min(n0 + 32, L.N); ++n) {
    float c = 0.f;
    if (mv) {
      float e = exp_fast((s[((long long)b * L.n_pad + n) * L.m_pad + m] - 1.f) * inv_scale);
      c = (e * rowsum[b * L.n_pad + n]) * (e * inv_cs);
      if (conf) conf[((long long)b * L.N + n) * L.M + m] = c;
      unsigned long long pk = pack_arg(c, n);
      cbest = pk > cbest ? pk : cbest;
    }
    // row arg-max over this warp's 32 columns: conf >= 0 so its bit pattern orders like an unsigned integer
    const unsigned bits = mv ? __float_as_uint(c) : 0u;
    const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
    const unsigned who = __ballot_sync(0xffffffffu, mv && bits == wmax);
    if (who && lane == __ffs(who) - 1) atomicMax(&rowbest[(long long)b * L.N + n], pack_arg(c, m));   // lowest column wins ties
  }
  if (mv && cbest) atomicMax(&colbest[(long long)b * L.M + m], cbest);
}

// Mutual nearest neighbour + threshold (reference GATs_SuperGlue.py:220-230).
// One block per frame is plenty (N+M <= ~20k); grid (B), block 256.
__global__ void mutual_match(const unsigned long long* __restrict__ rowbest, const unsigned long long* __restrict__ colbest,
                             int N, int M, float thr, long long* __restrict__ m0, long long* __restrict__ m1,
                             float* __restrict__ sc0, float* __restrict__ sc1) {
  const int b = blockIdx.x;
  const unsigned long long* rb = rowbest + (long long)b * N;
  const unsigned long long* cb = colbest + (long long)b * M;
  auto idx_of = [](unsigned long long p) { return (int)(0xFFFFFFFFu - (unsigned)(p & 0xFFFFFFFFull)); };
  auto val_of = [](unsigned long long p) { return __uint_as_float((unsigned)(p >> 32)); };
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    int i0 = idx_of(rb[n]);
    bool mutual = idx_of(cb[i0]) == n;
    float ms = mutual ? val_of(rb[n]) : 0.f;
    bool valid = mutual && ms > thr;
    m0[(long long)b * N + n] = valid ? (long long)i0 : -1ll;
    sc0[(long long)b * N + n] = ms;
  }
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    int i1 = idx_of(cb[m]);
    int i0 = idx_of(rb[i1]);
    bool mutual1 = i0 == m;
    // mscores0[i1], valid0[i1] recomputed (cheap) instead of read-after-write across threads
    bool mutual0_i1 = idx_of(cb[i0]) == i1;
    float ms0_i1 = mutual0_i1 ? val_of(rb[i1]) : 0.f;
    bool valid0_i1 = mutual0_i1 && ms0_i1 > thr;
    float ms1 = mutual1 ? ms0_i1 : 0.f;
    bool valid1 = mutual1 && valid0_i1;
    m1[(long long)b * M + m] = valid1 ? (long long)i1 : -1ll;
    sc1[(long long)b * M + m] = ms1;
  }
}

// ---------------------------------------------------------------------------------------
// Offline producer: segmented mean over variable-length tracks, fp64
// (reference feature_process.py:297-305).  offsets = exclusive prefix of seg_len.
// Warp per segment, lanes stride the D channels; sequential fp64 accumulation in row order
// then one division, which is what np.mean does for short axes (pairwise only kicks in >8 rows
// along a non-contiguous reduction? -- parity is checked to 1e-12 in the tests).
// ---------------------------------------------------------------------------------------
__global__ void segmented_mean_f64(const double* __restrict__ desc, const long long* __restrict__ seg_len,
                                   const long long* __restrict__ offsets, int M, int D, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long seg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seg >= M) return;
  const long long start = offsets[seg], len = seg_len[seg];
  for (int c = lane; c < D; c += 32) {
    double acc = 0.0;
    for (long long r = 0; r < len; ++r) acc += desc[(start + r) * D + c];
    out[seg * D + c] = acc / (double)len;
  }
}
__global__ void exclusive_scan_i64_single(const long long* __restrict__ in, long long* __restrict__ out, int n) {
  // tiny helper (M <= ~1e5): one thread; the producer is offline and HBM-bound elsewhere
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long acc = 0;
    for (int i = 0; i < n; ++i) { out[i] = acc; acc += in[i]; }
  }
}

}  // namespace opb
