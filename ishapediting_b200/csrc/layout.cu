// layout.cu — NCHW <-> NHWC transposes and the fp32 -> bf16 cast.
// The reference keeps everything NCHW (torch default); the kernels here are channels-last so
// that one pixel's channels are one contiguous TMA row.
#include "common.cuh"

namespace isb {

// tile: 32 pixels x 32 channels through shared memory; both sides coalesced.
template <typename TOut>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, TOut* __restrict__ dst, int C, int HW, int c_pad) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) v = __ldg(src + (static_cast<size_t>(n) * C + c) * HW + p);
    tile[ty + i * 8][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + i * 8, c = c0 + tx;
    if (p < HW && c < c_pad) {
      const float v = tile[tx][ty + i * 8];
      if constexpr (sizeof(TOut) == 2) dst[(static_cast<size_t>(n) * HW + p) * c_pad + c] = __float2bfloat16_rn(v);
      else dst[(static_cast<size_t>(n) * HW + p) * c_pad + c] = v;
    }
  }
}

template <typename TIn>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const TIn* __restrict__ src, float* __restrict__ dst, int C, int HW, int c_stride) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + i * 8, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) {
      if constexpr (sizeof(TIn) == 2) v = __bfloat162float(src[(static_cast<size_t>(n) * HW + p) * c_stride + c]);
      else v = src[(static_cast<size_t>(n) * HW + p) * c_stride + c];
    }
    tile[ty + i * 8][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, p = p0 + tx;
    if (c < C && p < HW) dst[(static_cast<size_t>(n) * C + c) * HW + p] = tile[tx][ty + i * 8];
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n8) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  load8(src + i * 8, v);
  store8(dst, i * 8, ISB_BF16, v);
}

// L2 prefetch of a byte range (weight panels of layers a few launches ahead): one bulk prefetch per 16 KiB chunk.
// The kernel only ISSUES the prefetches and exits; the data streams into L2 in the background.  No PDL wait: it reads
// nothing that a predecessor produces (weights are constant for the life of a plan).
constexpr size_t PF_CHUNK = 16384;
__global__ void __launch_bounds__(128)
prefetch_l2_kernel(const char* __restrict__ ptr, size_t bytes) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t off = i * PF_CHUNK;
  if (off >= bytes) return;
  size_t len = bytes - off;
  if (len > PF_CHUNK) len = PF_CHUNK;
  len &= ~static_cast<size_t>(15);
  if (len == 0) return;
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr + off), "r"(static_cast<uint32_t>(len)) : "memory");
}

}  // namespace isb

extern "C" {

int isb_prefetch_l2(const void* ptr, size_t bytes, isb_stream_t stream) {
  ISB_CHECK_ARG(ptr != nullptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "isb_prefetch_l2: pointer must be 16-byte aligned");
  if (bytes < 16) return ISB_OK;
  const size_t chunks = (bytes + isb::PF_CHUNK - 1) / isb::PF_CHUNK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>((chunks + 127) / 128));
  cfg.blockDim = dim3(128);
  cfg.stream = isb::as_stream(stream);
  ISB_CUDA(cudaLaunchKernelEx(&cfg, isb::prefetch_l2_kernel, static_cast<const char*>(ptr), bytes));
  isb::g_launches.fetch_add(1);
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

int isb_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C, int H, int W, int c_pad,
                     isb_stream_t stream) {
  ISB_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && c_pad >= C, "isb_nchw_to_nhwc: bad args");
  const int HW = H * W;
  dim3 grid(isb::cdiv(HW, 32), isb::cdiv(c_pad, 32), N);
  if (dst_dtype == ISB_BF16)
    ISB_CUDA(isb::launch(isb::nchw_to_nhwc_kernel<__nv_bfloat16>, grid, 256, 0, isb::as_stream(stream), src, static_cast<__nv_bfloat16*>(dst), C, HW, c_pad));
  else if (dst_dtype == ISB_F32)
    ISB_CUDA(isb::launch(isb::nchw_to_nhwc_kernel<float>, grid, 256, 0, isb::as_stream(stream), src, static_cast<float*>(dst), C, HW, c_pad));
  else ISB_CHECK_ARG(false, "isb_nchw_to_nhwc: bad dtype");
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

int isb_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int N, int C, int H, int W, int c_stride,
                     isb_stream_t stream) {
  ISB_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && c_stride >= C, "isb_nhwc_to_nchw: bad args");
  const int HW = H * W;
  dim3 grid(isb::cdiv(HW, 32), isb::cdiv(C, 32), N);
  if (src_dtype == ISB_BF16)
    ISB_CUDA(isb::launch(isb::nhwc_to_nchw_kernel<__nv_bfloat16>, grid, 256, 0, isb::as_stream(stream), static_cast<const __nv_bfloat16*>(src), dst, C, HW, c_stride));
  else if (src_dtype == ISB_F32)
    ISB_CUDA(isb::launch(isb::nhwc_to_nchw_kernel<float>, grid, 256, 0, isb::as_stream(stream), static_cast<const float*>(src), dst, C, HW, c_stride));
  else ISB_CHECK_ARG(false, "isb_nhwc_to_nchw: bad dtype");
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

int isb_cast_f32_bf16(const float* src, void* dst, size_t n, isb_stream_t stream) {
  ISB_CHECK_ARG(src && dst && n % 8 == 0, "isb_cast_f32_bf16: n must be a multiple of 8");
  if (n == 0) return ISB_OK;
  const size_t n8 = n / 8;
  ISB_CUDA(isb::launch(isb::cast_f32_bf16_kernel, isb::cdiv(n8, 256), 256, 0, isb::as_stream(stream), src, static_cast<__nv_bfloat16*>(dst), n8));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // extern "C"
