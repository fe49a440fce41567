// Host-side helpers shared by the translation units of the library (matcher api.cu, extractor superpoint.cu):
// device buffers that only ever grow, fp16-split plane pairs, the PDL-aware kernel launcher.
#pragma once
#include <utility>
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace opb {

// >= 0: byte value written into every fresh allocation that the algorithm does not require to be zero (opb_debug_set_ws_fill)
int ws_fill_byte();

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool grew = false;      // set by ensure() when it (re)allocated: the caller orders the fill against its stream
  cudaError_t ensure(size_t need, bool zero = false) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) return e;
    bytes = need;
    grew = true;
    if (zero) e = cudaMemset(p, 0, need);                                 // buffers whose untouched parts are read as zero (bd, flags)
    else if (ws_fill_byte() >= 0) e = cudaMemset(p, ws_fill_byte(), need);   // test hook: poison everything else
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// A window of rows of a [rows, 256] plane pair (the residual stream of a chunk, the query-only buffer, ...).
struct XView {
  __half* hi;
  __half* lo;
  CPlanes c(int ld) const { return CPlanes{hi, lo, ld}; }
  Planes m(int ld) const { return Planes{hi, lo, ld}; }
};

struct PlaneBuf {
  DevBuf hi, lo;
  cudaError_t ensure(size_t elems, bool zero = false) {
    cudaError_t e = hi.ensure(elems * sizeof(__half), zero);
    if (e != cudaSuccess) return e;
    return lo.ensure(elems * sizeof(__half), zero);
  }
  void release() { hi.release(); lo.release(); }
  CPlanes c(int ld, size_t off_elems = 0) const { return CPlanes{hi.as<__half>() + off_elems, lo.as<__half>() + off_elems, ld}; }
  Planes m(int ld, size_t off_elems = 0) const { return Planes{hi.as<__half>() + off_elems, lo.as<__half>() + off_elems, ld}; }
  XView view(size_t row_off = 0) const { return XView{hi.as<__half>() + row_off * kD, lo.as<__half>() + row_off * kD}; }
};

inline void split_host(const std::vector<double>& w, std::vector<__half>& hi, std::vector<__half>& lo) {
  hi.resize(w.size());
  lo.resize(w.size());
  for (size_t i = 0; i < w.size(); ++i) split_f32((float)w[i], hi[i], lo[i]);
}

// Every kernel of the library begins with griddep_sync() and is launched with programmatic stream serialization (PDL).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace opb
