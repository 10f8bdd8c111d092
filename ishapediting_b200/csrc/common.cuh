// common.cuh — shared host/device helpers for libishape_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/ishape_b200.h"

namespace isb {

// ---- host-side error plumbing -------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define ISB_CHECK_ARG(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      isb::set_error(__VA_ARGS__);                    \
      return ISB_ERR_ARG;                             \
    }                                                 \
  } while (0)

#define ISB_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) {                                                  \
      isb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                     cudaGetErrorString(_e));                                 \
      return ISB_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

#define ISB_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    cudaError_t _e = cudaPeekAtLastError();                                   \
    if (_e != cudaSuccess) {                                                  \
      isb::set_error("%s:%d launch -> %s", __FILE__, __LINE__,                \
                     cudaGetErrorString(_e));                                 \
      return ISB_ERR_CUDA;                                                    \
    }                                                                         \
    isb::count_launch();                                                      \
  } while (0)

inline cudaStream_t as_stream(isb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

bool is_initialised();
int num_sms();
// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint (no -lcuda).
typedef CUresult (*tensormap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tensormap_encode_fn get_tensormap_encode();

// ---- device helpers -------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch (PDL).  Every kernel of this library is launched with
// programmaticStreamSerialization: it may be scheduled while its predecessor in the stream is
// still draining, runs its input-independent prologue, and blocks in pdl_wait() until the
// predecessor has completed and flushed.  pdl_trigger() lets the successor start that early: the conv kernels
// issue it as their first statement (their successors have a real input-independent prologue); every other kernel
// issues pdl_wait() FIRST and pdl_trigger() right after it, so that waiting grids never cascade.  Rule: no global read of produced data and NO global write before
// pdl_wait().  ISB_PDL=0 in the environment turns the attribute off (plain stream order).
bool pdl_enabled();
bool pdl_enabled_conv();
// per-family opt-in (ISB_PDL_MASK bit f): 0 GroupNorm statistics / reduce, 1 GroupNorm apply, 2 fused GroupNorm,
// 3 fused attention, 4 everything else.  The launching function tags its launches with PdlFamily.
bool pdl_enabled_family(int family);
extern thread_local int t_pdl_family;
struct PdlFamily {
  int prev;
  explicit PdlFamily(int f) : prev(t_pdl_family) { t_pdl_family = f; }
  ~PdlFamily() { t_pdl_family = prev; }
};
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() || pdl_enabled_family(t_pdl_family)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float silu_f(float z) { return z / (1.0f + __expf(-z)); }
// d silu / dz = s * (1 + z * (1 - s)),  s = sigmoid(z)
__device__ __forceinline__ float silu_grad_f(float z) {
  float s = 1.0f / (1.0f + __expf(-z));
  return s * (1.0f + z * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 8 floats -> global, as fp32 (32 B) or bf16 (16 B); p is element pointer of the first value
__device__ __forceinline__ void store8(void* base, size_t elem_off, int dtype, const float* v) {
  if (dtype == ISB_BF16) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = u;
  } else {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}
__device__ __forceinline__ void load8(const float* p, float* v) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// plain (coherent) variant for buffers written earlier in the same kernel chain is
// not needed: kernels never read what they wrote in the same launch.

#endif  // __CUDACC__

}  // namespace isb
