// tcgen05 / TMA GEMM core for sm_100a.
//
//   C[128 x 256 tile] = sum_k  A_tile . B_tile^T      A, B: fp16 hi/lo(x2^11) planes, K-major
//
// executed as THREE tensor-core passes into TWO fp32 TMEM accumulators
//     D0 += A_hi . B_hi                      (main term)
//     D1 += A_hi . B_lo' + A_lo' . B_hi      (first-order correction, carries 2^11)
//     C   = D0 + 2^-11 * D1                  (epilogue)
// which reproduces fp32 products to ~2^-22 relative (tools/precision_ladder.py) -- the
// reference's arithmetic is fp32 (GATs_SuperGlue.py:191-193) and the contract is 1e-4 on conf.
//
// One CTA per 128x256 output tile, 256 threads, warp-specialised:
//   warp 0   : TMA producer  (cp.async.bulk.tensor 2D, SWIZZLE_128B boxes, mbarrier expect_tx)
//   warp 1   : MMA issuer    (one elected lane issues tcgen05.mma kind::f16, commits to mbarriers)
//   warp 2   : TMEM allocator (512 columns: D0 = cols [0,256), D1 = cols [256,512))
//   warps 4-7: epilogue      (tcgen05.ld 32x32b -> registers -> combine -> global)
// smem ring: kStages x { A_hi, A_lo (128x64 fp16 = 16 KB each), B_hi, B_lo (256x64 = 32 KB each) }.
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "gemm_tc.cuh"

namespace opb {
namespace {

constexpr int BM = 128, BN = 256, BK = 64;       // BK fp16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kStages = 2;
constexpr int kABytes = BM * BK * 2;             // 16 KB
constexpr int kBBytes = BN * BK * 2;             // 32 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // 96 KB
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kTmemCols = 512;
constexpr uint32_t kSpinLimit = 1u << 22;        // bounded waits: trap instead of hanging the GPU

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::f16, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   start address >> 4 | LBO (unused for swizzled K-major; 1) | SBO = 1024 B between 8-row groups
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=BN
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct TcParams {
  int K1, K2;            // reduction split (multiples of BK)
  int b2_per_seg;
  int n_out;
  long long a_batch_rows, b_batch_rows, c_batch_elems;
  Layout L;
  const float* bias;
  float* c;
  int ldc;
};

__global__ void __launch_bounds__(256, 1)
gemm_tc_plain_kernel(const __grid_constant__ CUtensorMap map_a1h, const __grid_constant__ CUtensorMap map_a1l,
                     const __grid_constant__ CUtensorMap map_a2h, const __grid_constant__ CUtensorMap map_a2l,
                     const __grid_constant__ CUtensorMap map_b1h, const __grid_constant__ CUtensorMap map_b1l,
                     const __grid_constant__ CUtensorMap map_b2h, const __grid_constant__ CUtensorMap map_b2l, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, z = blockIdx.z;
  const int row0 = m_tile * BM;                       // within batch z
  const int a_row = (int)(z * p.a_batch_rows) + row0; // coordinate in the A tensor maps
  const int b_row1 = (int)(z * p.b_batch_rows) + n_tile * BN;
  const int seg = p.b2_per_seg ? p.L.seg_of_row(row0) : 0;
  const int b_row2 = (p.b2_per_seg ? seg * p.n_out : 0) + n_tile * BN;
  const int nkb1 = p.K1 / BK, nkb = (p.K1 + p.K2) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a1h); prefetch_tmap(&map_a1l); prefetch_tmap(&map_b1h); prefetch_tmap(&map_b1l);
    if (p.K2) { prefetch_tmap(&map_a2h); prefetch_tmap(&map_a2l); prefetch_tmap(&map_b2h); prefetch_tmap(&map_b2l); }
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t it = kb / kStages;
        mbar_wait(&empty_bar[s], (it & 1) ^ 1);
        uint8_t* st = smem + s * kStageBytes;
        mbar_expect_tx(&full_bar[s], kStageBytes);
        if (kb < nkb1) {
          tma_load_2d(st, &map_a1h, &full_bar[s], kb * BK, a_row);
          tma_load_2d(st + kABytes, &map_a1l, &full_bar[s], kb * BK, a_row);
          tma_load_2d(st + 2 * kABytes, &map_b1h, &full_bar[s], kb * BK, b_row1);
          tma_load_2d(st + 2 * kABytes + kBBytes, &map_b1l, &full_bar[s], kb * BK, b_row1);
        } else {
          const int k2 = (kb - nkb1) * BK;
          tma_load_2d(st, &map_a2h, &full_bar[s], k2, a_row);
          tma_load_2d(st + kABytes, &map_a2l, &full_bar[s], k2, a_row);
          tma_load_2d(st + 2 * kABytes, &map_b2h, &full_bar[s], k2, b_row2);
          tma_load_2d(st + 2 * kABytes + kBBytes, &map_b2l, &full_bar[s], k2, b_row2);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t d0 = tmem_base, d1 = tmem_base + BN;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t it = kb / kStages;
        mbar_wait(&full_bar[s], it & 1);
        tc_fence_after();
        const uint32_t sa_h = smem_u32(smem + s * kStageBytes);
        const uint32_t sa_l = sa_h + kABytes, sb_h = sa_h + 2 * kABytes, sb_l = sb_h + kBBytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint32_t koff = k * UMMA_K * 2;   // bytes inside the 128 B swizzle row
          const uint64_t ah = make_desc_sw128(sa_h + koff), al = make_desc_sw128(sa_l + koff);
          const uint64_t bh = make_desc_sw128(sb_h + koff), bl = make_desc_sw128(sb_l + koff);
          const uint32_t acc = (kb | k) != 0;
          tc_mma_f16(d0, ah, bh, kIdesc, acc);
          tc_mma_f16(d1, ah, bl, kIdesc, acc);
          tc_mma_f16(d1, al, bh, kIdesc, 1u);
        }
        tc_commit(&empty_bar[s]);                 // frees the smem stage when these MMAs retire
      }
      tc_commit(tmem_full_bar);                   // accumulators complete
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp - 4;                       // TMEM lane quarter == warp_id % 4
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int r = row0 + q * 32 + lane;
    float* crow = p.c + (long long)z * p.c_batch_elems + (long long)r * p.ldc + (long long)n_tile * BN;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v0[32], v1[32];
      tmem_ld32(lane_base + c0, v0);
      tmem_ld32(lane_base + BN + c0, v1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o;
        o.x = fmaf(__uint_as_float(v1[j + 0]), kLoInv, __uint_as_float(v0[j + 0]));
        o.y = fmaf(__uint_as_float(v1[j + 1]), kLoInv, __uint_as_float(v0[j + 1]));
        o.z = fmaf(__uint_as_float(v1[j + 2]), kLoInv, __uint_as_float(v0[j + 2]));
        o.w = fmaf(__uint_as_float(v1[j + 3]), kLoInv, __uint_as_float(v0[j + 3]));
        if (p.bias) {
          const float4 bb = *reinterpret_cast<const float4*>(p.bias + n_tile * BN + c0 + j);
          o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        }
        *reinterpret_cast<float4*>(crow + c0 + j) = o;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------ host: tensor maps
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(f);
  });
  return fn;
}

// fp16 plane [rows, ld] (cols used: `cols`), box = BK x box_rows, SWIZZLE_128B.  Cached per (ptr, rows, cols, ld, box).
bool make_map(CUtensorMap* out, const __half* ptr, long long rows, int cols, int ld, int box_rows) {
  static std::map<std::tuple<const void*, long long, int, int, int>, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_tuple((const void*)ptr, rows, cols, ld, box_rows);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return true;
}

}  // namespace

int launch_gemm_tc_plain(const GemmProblem& p, cudaStream_t stream) {
  if (p.rows % BM || p.n_out % BN || p.K1 % BK || p.K2 % BK || p.K1 <= 0 || p.ldc % 4) return -1;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(gemm_tc_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return -2;
    attr_done = true;
  }
  const long long a_rows = (long long)(p.batch - 1) * p.a_batch_rows + p.rows;
  const long long b1_rows = (long long)(p.batch - 1) * p.b_batch_rows + p.n_out;
  const long long b2_rows = p.b2_per_seg ? (long long)p.L.segs() * p.n_out : p.n_out;
  CUtensorMap ma1h, ma1l, ma2h, ma2l, mb1h, mb1l, mb2h, mb2l;
  bool ok = make_map(&ma1h, p.a1.hi, a_rows, p.K1, p.a1.ld, BM) && make_map(&ma1l, p.a1.lo, a_rows, p.K1, p.a1.ld, BM) &&
            make_map(&mb1h, p.b1.hi, b1_rows, p.K1, p.b1.ld, BN) && make_map(&mb1l, p.b1.lo, b1_rows, p.K1, p.b1.ld, BN);
  if (ok && p.K2) {
    ok = make_map(&ma2h, p.a2.hi, a_rows, p.K2, p.a2.ld, BM) && make_map(&ma2l, p.a2.lo, a_rows, p.K2, p.a2.ld, BM) &&
         make_map(&mb2h, p.b2.hi, b2_rows, p.K2, p.b2.ld, BN) && make_map(&mb2l, p.b2.lo, b2_rows, p.K2, p.b2.ld, BN);
  } else if (ok) {
    ma2h = ma1h; ma2l = ma1l; mb2h = mb1h; mb2l = mb1l;
  }
  if (!ok) return -2;
  TcParams tp;
  tp.K1 = p.K1; tp.K2 = p.K2; tp.b2_per_seg = p.b2_per_seg; tp.n_out = p.n_out;
  tp.a_batch_rows = p.a_batch_rows; tp.b_batch_rows = p.b_batch_rows; tp.c_batch_elems = p.c_batch_elems;
  tp.L = p.L; tp.bias = p.bias; tp.c = p.c; tp.ldc = p.ldc;
  dim3 grid(p.n_out / BN, p.rows / BM, p.batch);
  gemm_tc_plain_kernel<<<grid, 256, kSmemBytes, stream>>>(ma1h, ma1l, ma2h, ma2l, mb1h, mb1l, mb2h, mb2l, tp);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace opb
