// decode_tc.cu — the triplane occupancy decoder (axisnetworks.py:517-562, visualize.py:79-98) with its two
// 128x128 layers (94 % of the 69 888 FLOP per point) on the 5th-generation tensor cores: tcgen05.mma, operands in
// shared memory, fp32 accumulators in TMEM.
//
// fp32-GRADE accuracy on half-precision tensor cores.  The decision boundary needs it (|2*pi*f@B| reaches hundreds
// of radians, the IoU gate is 0.999), and the reference computes in fp32.  Every operand is split into two fp16
// terms, x = x_hi + x_lo with x_hi = fp16(x), x_lo = fp16(x - x_hi): 22 mantissa bits.  A product keeps the three
// leading terms
//        a*w  ~=  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo                (dropped: a_lo*w_lo ~ 2^-22 |a w|)
// = three kind::f16 MMAs per layer with fp32 accumulation.  w_lo ~ 2^-12 |w| would be an fp16 SUBNORMAL for typical
// weights (|w| ~ 0.1), so it is stored scaled by 2^11 and its product is accumulated in a SECOND TMEM accumulator D'
// that the epilogue folds back as D + 2^-11 D' (exact scaling, no extra operand copy).  Measured against the fp32
// oracle: see tests/test_gpu_baseline_configs.py::test_decoder_iou_128.
// Range: layer-1 inputs are sines / cosines; layer-2 inputs are relu(h1) with |h1| <= max_r(sum_k |W1[r,k]| + |b1[r]|).
// The host passes that bound (isb_triplane_mlp.h1_bound); this kernel is used only when it is fp16-safe, otherwise the
// 3xTF32 mma.sync kernel of decode.cu runs (same results, slower).
//
// One persistent CTA per SM (220 KB of shared memory: both layers' weights as hi / lo' fp16 K-major 128B-swizzled
// UMMA tiles = 128 KB, one A operand pair = 64 KB, staging).  544 threads: warps 0-15 compute (sampling, Fourier
// features, operand splitting, epilogues; 4 warps per TMEM lane quarter), warp 16 issues the MMAs.  Per 128-point
// tile i, software-pipelined so that the tensor cores never wait for the sampling of their own tile:
//   S0  [sin | cos](2 pi u_i) -> A_hi / A_lo in UMMA layout                        -> mbarrier a_ready
//   M1  D1 = A_hi W1hi^T + A_lo W1hi^T,  D1' = A_hi W1lo'^T   (24 MMAs 128x128x16) -> tcgen05.commit d_ready
//   S1  (under M1) bilinear plane samples of tile i+1 -> F[32][128]
//   S2  h1 = relu(D1 + 2^-11 D1' + b1) -> split -> A_hi / A_lo (layer-2 operand)   -> a_ready
//   M2  D2, D2' likewise with W2                                                  -> d_ready
//   S3  (under M2) u_{i+1} = F @ B  (fp32 FFMA, 4x4 register tiles; stays in registers)
//   S4  logit = w3 . relu(D2 + 2^-11 D2' + b2) + b3  (fp32 registers)             -> global
#include "common.cuh"

namespace isb {
namespace dtc {

constexpr int TP = 128;                 // points per tile = UMMA M
constexpr int H = 128, F = 32, MF = 64; // hidden width, plane features, Fourier frequencies
constexpr int COMPUTE_THREADS = 512;    // 16 compute warps: 4 per TMEM lane quarter, 4 per scheduler (latency hiding)
constexpr int THREADS = COMPUTE_THREADS + 32;
constexpr int ATOM = TP * 128;          // one K-atom: 128 rows x 64 fp16 = 16 KiB
constexpr int OPER = 2 * ATOM;          // one 128x128 fp16 operand = 32 KiB
// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_W1H = 0, OFF_W1L = OPER, OFF_W2H = 2 * OPER, OFF_W2L = 3 * OPER;
constexpr int OFF_AH = 4 * OPER, OFF_AL = 5 * OPER;
constexpr int OFF_FOP = 6 * OPER;                    // plane features as a UMMA A operand: 128 rows x 64 fp16,
                                                     //   k 0..31 = f_hi, k 32..63 = f_lo' (one 16 KiB atom)
constexpr int OFF_BOP = OFF_FOP + ATOM;              // Fourier matrix as a B operand: 64 rows (frequencies) x 64 fp16,
                                                     //   k 0..31 = B_hi^T, k 32..63 = B_lo'^T (8 KiB)
constexpr int OFF_VEC = OFF_BOP + MF * 128;          // float b1[128], b2[128], w3[128]
constexpr int OFF_X = OFF_VEC + 3 * H * 4;           // float xch[3][128]
constexpr int SMEM_BYTES = OFF_X + 3 * TP * 4 + 1024;  // + alignment slack

struct Args {
  const float* planes; int R;
  const float* fourier_B; const float* w1; const float* b1; const float* w2; const float* b2;
  const float* w3; const float* b3;
  const float* lin; int res; long long idx0;
  const float* coords;
  long long npts;
  float* out;
  int fast_rows;   // grid mode, res % 128 == 0, <= 31 plane rows per 32 consecutive z: cooperative row sampling
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// bounded wait: a pipeline bug must trap (CUDA error), never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 6000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// K-major, 128B-swizzled operand tile (rows of 64 fp16 = 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// sin and cos of x to ~1 ulp (measured 7e-8 absolute against float64 for |x| <= 4e4, tools note in DESIGN.md):
// three-term Cody-Waite reduction by pi/2 and the two degree-7/8 kernels on |r| <= pi/4, i.e. the fast path of
// sincosf() without its slow-path branch, denormal handling and local-memory frame; arguments beyond the reduction's
// range (never reached by 2*pi*f@B in practice) take the library call.
__device__ __forceinline__ void sincos_cw(float x, float& sn, float& cs) {
  if (fabsf(x) > 40000.0f) {
    sincosf(x, &sn, &cs);
    return;
  }
  const float k = rintf(x * 0.636619772f);
  const int q = static_cast<int>(k);
  float r = fmaf(k, -1.57079601e+00f, x);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  const float r2 = r * r;
  float p = 2.86567956e-6f;
  p = fmaf(p, r2, -1.98559923e-4f);
  p = fmaf(p, r2, 8.33338592e-3f);
  p = fmaf(p, r2, -1.66666672e-1f);
  const float s = fmaf(p, r * r2, r);
  float c = 2.44677067e-5f;
  c = fmaf(c, r2, -1.38877297e-3f);
  c = fmaf(c, r2, 4.16666567e-2f);
  c = fmaf(c, r2, -0.5f);
  c = fmaf(c, r2, 1.0f);
  const float a = (q & 1) ? c : s, b = (q & 1) ? s : c;
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}

// byte offset of the 16-byte chunk holding elements [k8*8, k8*8+8) of row r inside a 128x128 fp16 operand
__device__ __forceinline__ uint32_t chunk_off(int r, int k8) {
  return static_cast<uint32_t>((k8 >> 3) * ATOM + (r >> 3) * 1024 + (r & 7) * 128 + (((k8 & 7) ^ (r & 7)) << 4));
}
// x = hi + lo (two fp16 terms); returns the packed pair of two consecutive elements
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
  const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
  hi = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
  lo = static_cast<uint32_t>(__half_as_ushort(l0)) | (static_cast<uint32_t>(__half_as_ushort(l1)) << 16);
}
// 8 consecutive k of row r -> A_hi / A_lo
__device__ __forceinline__ void store_chunk(uint32_t a_hi, uint32_t a_lo, int r, int k8, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
  const uint32_t off = chunk_off(r, k8);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

__device__ __forceinline__ void sample_plane8(const float* __restrict__ plane, int R, float gx, float gy, int c0, float* acc) {
  const float ix = ((gx + 1.0f) / 2.0f) * static_cast<float>(R - 1);       // grid_sample, align_corners=True
  const float iy = ((gy + 1.0f) / 2.0f) * static_cast<float>(R - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
  const float fx = ix - fx0, fy = iy - fy0;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int x = x0 + dx, y = y0 + dy;
      if (x < 0 || x >= R || y < 0 || y >= R) continue;                     // zeros padding
      const float w = (dx ? fx : 1.0f - fx) * (dy ? fy : 1.0f - fy);
      const float4* p = reinterpret_cast<const float4*>(plane + (static_cast<size_t>(y) * R + x) * F + c0);
      const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
      acc[0] = fmaf(w, v0.x, acc[0]); acc[1] = fmaf(w, v0.y, acc[1]); acc[2] = fmaf(w, v0.z, acc[2]); acc[3] = fmaf(w, v0.w, acc[3]);
      acc[4] = fmaf(w, v1.x, acc[4]); acc[5] = fmaf(w, v1.y, acc[5]); acc[6] = fmaf(w, v1.z, acc[6]); acc[7] = fmaf(w, v1.w, acc[7]);
    }
}

// plane samples of point (tile, pt), channels c0..c0+7 (zeros for points past the end)
template <bool GRID>
__device__ __forceinline__ void sample_point(const Args& a, long long tile, int pt, int c0, float* f) {
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  const long long i = tile * TP + pt;
  if (i >= a.npts) return;
  float cx, cy, cz;
  if (GRID) {
    const long long gi = a.idx0 + i;
    const int z = static_cast<int>(gi % a.res);
    const long long t = gi / a.res;
    const int y = static_cast<int>(t % a.res);
    const int x = static_cast<int>(t / a.res);
    cx = __ldg(a.lin + x); cy = __ldg(a.lin + y); cz = __ldg(a.lin + z);
  } else {
    cx = __ldg(a.coords + i * 3); cy = __ldg(a.coords + i * 3 + 1); cz = __ldg(a.coords + i * 3 + 2);
  }
  const size_t plane_sz = static_cast<size_t>(a.R) * a.R * F;
  sample_plane8(a.planes, a.R, cx, cy, c0, f);                 // xy plane: x->W, y->H
  sample_plane8(a.planes + plane_sz, a.R, cy, cz, c0, f);      // yz plane
  sample_plane8(a.planes + 2 * plane_sz, a.R, cx, cz, c0, f);  // xz plane
}

// Grid mode with res % 128 == 0: the 32 points of a warp are consecutive in z at one (x, y).  Then
//     f(z) = xy(x, y) + (1 - fy(z)) G[y0(z)] + fy(z) G[y0(z) + 1],    G[row] = lerp_x(yz plane)[row] + lerp_x(xz plane)[row]
// (bilinear interpolation separates, and both z-planes are sampled at the same row coordinate).  Lane j loads the
// combined row y0(lane 0) + j ONCE (4 pixels) and every lane fetches its two rows with warp shuffles: 16 loads per lane
// instead of 24, and 6x fewer distinct cache lines per tile than per-point sampling (the L1 is only ~28 KB next to 219 KB
// of shared memory: per-point sampling was 40 % of the kernel's stall samples, profiles/r02_decode_tc.md).
__device__ __forceinline__ void sample_rows_fast(const Args& a, long long tile, int wq, int lane, int c0, float* f) {
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  const long long gi0 = a.idx0 + tile * TP + wq * 32;
  const int z0 = static_cast<int>(gi0 % a.res);
  const long long t = gi0 / a.res;
  const int y = static_cast<int>(t % a.res);
  const int x = static_cast<int>(t / a.res);
  const float cx = __ldg(a.lin + x), cy = __ldg(a.lin + y), cz = __ldg(a.lin + z0 + lane);
  const int R = a.R;
  const size_t plane_sz = static_cast<size_t>(R) * R * F;
  sample_plane8(a.planes, R, cx, cy, c0, f);                    // xy plane: one broadcast sample for the warp
  const float iy = ((cz + 1.0f) / 2.0f) * static_cast<float>(R - 1);
  const float fy0 = floorf(iy);
  const int y0 = static_cast<int>(fy0);
  const float fy = iy - fy0;
  const int rbase = __shfl_sync(0xffffffffu, y0, 0);
  const int rlast = __shfl_sync(0xffffffffu, y0, 31) + 1;         // last row any lane of this warp reads
  const int row = rbase + lane;
  float g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  if (row >= 0 && row < R && row <= rlast) {
#pragma unroll
    for (int pl = 1; pl <= 2; ++pl) {
      const float gx = pl == 1 ? cy : cx;                         // yz plane: y -> W;  xz plane: x -> W
      const float ix = ((gx + 1.0f) / 2.0f) * static_cast<float>(R - 1);
      const float fx0 = floorf(ix);
      const int x0 = static_cast<int>(fx0);
      const float fx = ix - fx0;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int xx = x0 + dx;
        if (xx < 0 || xx >= R) continue;
        const float w = dx ? fx : 1.0f - fx;
        const float4* p = reinterpret_cast<const float4*>(a.planes + pl * plane_sz + (static_cast<size_t>(row) * R + xx) * F + c0);
        const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
        g[0] = fmaf(w, v0.x, g[0]); g[1] = fmaf(w, v0.y, g[1]); g[2] = fmaf(w, v0.z, g[2]); g[3] = fmaf(w, v0.w, g[3]);
        g[4] = fmaf(w, v1.x, g[4]); g[5] = fmaf(w, v1.y, g[5]); g[6] = fmaf(w, v1.z, g[6]); g[7] = fmaf(w, v1.w, g[7]);
      }
    }
  }
  const int j0 = y0 - rbase;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v0 = __shfl_sync(0xffffffffu, g[j], j0);
    const float v1 = __shfl_sync(0xffffffffu, g[j], j0 + 1);
    f[j] = fmaf(fy, v1, fmaf(1.0f - fy, v0, f[j]));
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}

// 8 plane features of one point -> f_hi (k = c0..c0+7) and f_lo' = (f - f_hi) * 2^11 (k = 32 + c0..) of the F operand
__device__ __forceinline__ void store_features(uint32_t fop, int r, int c0, const float* f) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half h0 = __float2half_rn(f[2 * j]), h1 = __float2half_rn(f[2 * j + 1]);
    const __half l0 = __float2half_rn((f[2 * j] - __half2float(h0)) * 2048.0f);
    const __half l1 = __float2half_rn((f[2 * j + 1] - __half2float(h1)) * 2048.0f);
    h[j] = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
    l[j] = static_cast<uint32_t>(__half_as_ushort(l0)) | (static_cast<uint32_t>(__half_as_ushort(l1)) << 16);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(fop + chunk_off(r, c0 >> 3)), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(fop + chunk_off(r, 4 + (c0 >> 3))), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

template <bool GRID>
__global__ void __launch_bounds__(THREADS, 1)
triplane_decode_tc_kernel(const Args a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_d, bar_u;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* sb1 = reinterpret_cast<float*>(gbase + OFF_VEC);
  float* sb2 = sb1 + H;
  float* sw3 = sb2 + H;
  float* sx = reinterpret_cast<float*>(gbase + OFF_X);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: weights -> hi / lo' fp16 UMMA tiles, small vectors, barriers, TMEM ----
  for (int i = tid; i < H * (H / 8); i += THREADS) {            // (row n, chunk k8) of W[n][k]
    const int n = i / (H / 8), k8 = i % (H / 8);
#pragma unroll
    for (int layer = 0; layer < 2; ++layer) {
      const float* W = layer ? a.w2 : a.w1;
      float v[8];
      load8(W + static_cast<size_t>(n) * H + k8 * 8, v);
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half h0 = __float2half_rn(v[2 * j]), h1 = __float2half_rn(v[2 * j + 1]);
        const __half l0 = __float2half_rn((v[2 * j] - __half2float(h0)) * 2048.0f);      // lo' = lo * 2^11
        const __half l1 = __float2half_rn((v[2 * j + 1] - __half2float(h1)) * 2048.0f);
        h[j] = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
        l[j] = static_cast<uint32_t>(__half_as_ushort(l0)) | (static_cast<uint32_t>(__half_as_ushort(l1)) << 16);
      }
      const uint32_t off = chunk_off(n, k8);
      const uint32_t dh = base + (layer ? OFF_W2H : OFF_W1H) + off, dl = base + (layer ? OFF_W2L : OFF_W1L) + off;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dh), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dl), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
    }
  }
  for (int i = tid; i < MF * (F / 8); i += THREADS) {           // Fourier matrix B[c][m] -> operand rows m, k = c
    const int m = i / (F / 8), k8 = i % (F / 8);
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v0 = __ldg(a.fourier_B + (k8 * 8 + 2 * j) * MF + m), v1 = __ldg(a.fourier_B + (k8 * 8 + 2 * j + 1) * MF + m);
      const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
      const __half l0 = __float2half_rn((v0 - __half2float(h0)) * 2048.0f), l1 = __float2half_rn((v1 - __half2float(h1)) * 2048.0f);
      h[j] = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
      l[j] = static_cast<uint32_t>(__half_as_ushort(l0)) | (static_cast<uint32_t>(__half_as_ushort(l1)) << 16);
    }
    const uint32_t bop = base + OFF_BOP;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bop + chunk_off(m, k8)), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(bop + chunk_off(m, 4 + k8)), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
  }
  if (tid < H) { sb1[tid] = __ldg(a.b1 + tid); sb2[tid] = __ldg(a.b2 + tid); sw3[tid] = __ldg(a.w3 + tid); }
  const float b3 = __ldg(a.b3);
  if (tid == 0) {
    mbar_init(smem_u32(&bar_a), COMPUTE_THREADS);
    mbar_init(smem_u32(&bar_d), 1);
    mbar_init(smem_u32(&bar_u), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();          // the weight tiles were written through the generic proxy; tcgen05.mma reads them
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_hi = base + OFF_AH, a_lo = base + OFF_AL;
  const long long ntiles = (a.npts + TP - 1) / TP;
  // contiguous tile range per CTA: consecutive tiles walk z, then y — the xz-plane rows of a z range are reused from L1
  // for every y, only the yz-plane rows are new per tile
  const long long tile_begin = ntiles * blockIdx.x / gridDim.x, tile_end = ntiles * (blockIdx.x + 1) / gridDim.x;

  if (warp == 16) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = f16, both K-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(H >> 3) << 17) | (static_cast<uint32_t>(TP >> 4) << 24);
      const uint64_t d_ah = make_desc_sw128(a_hi), d_al = make_desc_sw128(a_lo);
      // Fourier projection u = f @ B on the tensor cores too: M = 128, N = 64, K = 32 per term
      const uint32_t idesc_u = (1u << 4) | (static_cast<uint32_t>(MF >> 3) << 17) | (static_cast<uint32_t>(TP >> 4) << 24);
      const uint64_t d_f = make_desc_sw128(base + OFF_FOP), d_b = make_desc_sw128(base + OFF_BOP);
      const uint32_t ba = smem_u32(&bar_a), bd = smem_u32(&bar_d), bu = smem_u32(&bar_u);
      uint32_t ph = 0;
      auto fourier_mma = [&]() {
        // u (D1 columns 0..63) = f_hi B_hi^T;   u' (D1' columns 0..63) = f_hi B_lo'^T + f_lo' B_hi^T   (u = u + 2^-11 u')
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t hi = static_cast<uint64_t>(ks * 2), lo = static_cast<uint64_t>(4 + ks * 2);
          umma_f16(tmem + 128u, d_f + hi, d_b + lo, idesc_u, ks > 0 ? 1u : 0u);
          umma_f16(tmem + 128u, d_f + lo, d_b + hi, idesc_u, 1u);
          umma_f16(tmem, d_f + hi, d_b + hi, idesc_u, ks > 0 ? 1u : 0u);
        }
        umma_commit(bu);
      };
      if (tile_begin < tile_end) {                 // prologue: the first tile's features are in place
        mbar_wait(ba, ph);
        ph ^= 1u;
        fence_after();
        fourier_mma();
      }
      for (long long tile = tile_begin; tile < tile_end; ++tile) {
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          mbar_wait(ba, ph);
          ph ^= 1u;
          fence_after();
          const uint64_t d_wh = make_desc_sw128(base + (layer ? OFF_W2H : OFF_W1H));
          const uint64_t d_wl = make_desc_sw128(base + (layer ? OFF_W2L : OFF_W1L));
          const uint32_t dmain = tmem + static_cast<uint32_t>(layer * 256), dlow = dmain + 128u;
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks) {
            // 16 fp16 = 32 B along K inside a swizzle atom: +2 in (addr >> 4) units; second atom 16 KiB further
            const uint64_t ko = static_cast<uint64_t>((ks >> 2) * (ATOM >> 4) + (ks & 3) * 2);
            umma_f16(dlow, d_ah + ko, d_wl + ko, idesc, ks > 0 ? 1u : 0u);       // D' = A_hi W_lo'^T
            umma_f16(dmain, d_al + ko, d_wh + ko, idesc, ks > 0 ? 1u : 0u);      // small term first
            umma_f16(dmain, d_ah + ko, d_wh + ko, idesc, 1u);
          }
          umma_commit(bd);
          // D1 / D1' were drained by the layer-1 epilogue and the next tile's features were staged before it: project
          // them while the compute warps run the layer-2 epilogue
          if (layer == 1 && tile + 1 < tile_end) fourier_mma();
        }
      }
    }
  } else {
    // ===== compute warps =====
    // Software pipeline: while the tensor cores chew layer 1 of tile i the compute warps SAMPLE tile i+1, while they
    // chew layer 2 the warps PROJECT tile i+1 onto the Fourier frequencies (u stays in registers across the epilogue).
    const uint32_t ba = smem_u32(&bar_a), bd = smem_u32(&bar_d), bu = smem_u32(&bar_u);
    const uint32_t fop = base + OFF_FOP;
    uint32_t phd = 0, phu = 0;
    const int s_pt = (warp & 3) * 32 + lane, s_c0 = (warp >> 2) * 8;   // sampling: point, first of 8 channels
    const int q = warp & 3, cg = warp >> 2;                            // TMEM lane quarter / column group
    const int row = q * 32 + lane, col0 = cg * 32;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const float two_pi = 6.283185307179586f;
    if (tile_begin < tile_end) {           // prologue: features of the first tile -> F operand -> first projection
      float f[8];
      if (GRID && a.fast_rows) sample_rows_fast(a, tile_begin, warp & 3, lane, s_c0, f);
      else sample_point<GRID>(a, tile_begin, s_pt, s_c0, f);
      store_features(fop, s_pt, s_c0, f);
      fence_async_smem();
      mbar_arrive(ba);
    }
    for (long long tile = tile_begin; tile < tile_end; ++tile) {
      const long long next = tile + 1;
      const bool has_next = next < tile_end;               // CTA-uniform
      // ---- S0: u = D + 2^-11 D' (Fourier projection, 16 frequencies of this thread's point) -> [sin | cos](2 pi u)
      //          -> layer-1 operand ----
      mbar_wait(bu, phu);
      phu ^= 1u;
      fence_after();
      {
        float d[16], dl[16];
        tmem_ld16(trow + static_cast<uint32_t>(cg * 16), d);
        tmem_ld16(trow + static_cast<uint32_t>(128 + cg * 16), dl);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          float sv[8], cv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) sincos_cw(two_pi * fmaf(dl[g8 * 8 + j], 1.0f / 2048.0f, d[g8 * 8 + j]), sv[j], cv[j]);
          store_chunk(a_hi, a_lo, row, cg * 2 + g8, sv);            // k = m          (sin), m = 16 cg + 8 g8 + j
          store_chunk(a_hi, a_lo, row, 8 + cg * 2 + g8, cv);        // k = 64 + m     (cos)
        }
      }
      fence_async_smem();
      fence_before();
      mbar_arrive(ba);
      // ---- S1 (under MMA layer 1): sample the next tile into the F operand (its projection was consumed above) ----
      if (has_next) {
        float f[8];
        if (GRID && a.fast_rows) sample_rows_fast(a, next, warp & 3, lane, s_c0, f);
        else sample_point<GRID>(a, next, s_pt, s_c0, f);
        store_features(fop, s_pt, s_c0, f);
      }
      // ---- S2: h1 = relu(D1 + 2^-11 D1' + b1) -> layer-2 operand ----
      mbar_wait(bd, phd);
      phd ^= 1u;
      fence_after();
      {
        float d[32], dl[32];
        tmem_ld32(trow + static_cast<uint32_t>(col0), d);
        tmem_ld32(trow + static_cast<uint32_t>(128 + col0), dl);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const float4 ba4 = *reinterpret_cast<const float4*>(sb1 + col0 + g8 * 8);
          const float4 bb4 = *reinterpret_cast<const float4*>(sb1 + col0 + g8 * 8 + 4);
          const float bv[8] = {ba4.x, ba4.y, ba4.z, ba4.w, bb4.x, bb4.y, bb4.z, bb4.w};
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = g8 * 8 + j;
            v[j] = fmaxf(fmaf(dl[k], 1.0f / 2048.0f, d[k]) + bv[j], 0.f);
          }
          store_chunk(a_hi, a_lo, row, (col0 >> 3) + g8, v);
        }
      }
      fence_async_smem();          // covers the layer-2 operand AND the next tile's F operand (S1)
      fence_before();
      mbar_arrive(ba);
      // ---- S4: logit = w3 . relu(D2 + 2^-11 D2' + b2) + b3 ----
      mbar_wait(bd, phd);
      phd ^= 1u;
      fence_after();
      float acc = 0.f;
      {
        float d[32], dl[32];
        tmem_ld32(trow + static_cast<uint32_t>(256 + col0), d);
        tmem_ld32(trow + static_cast<uint32_t>(384 + col0), dl);
        tmem_ld_wait();
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 w4 = *reinterpret_cast<const float4*>(sw3 + col0 + k4 * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(sb2 + col0 + k4 * 4);
          const int k = k4 * 4;
          acc = fmaf(w4.x, fmaxf(fmaf(dl[k], 1.0f / 2048.0f, d[k]) + b4.x, 0.f), acc);
          acc = fmaf(w4.y, fmaxf(fmaf(dl[k + 1], 1.0f / 2048.0f, d[k + 1]) + b4.y, 0.f), acc);
          acc = fmaf(w4.z, fmaxf(fmaf(dl[k + 2], 1.0f / 2048.0f, d[k + 2]) + b4.z, 0.f), acc);
          acc = fmaf(w4.w, fmaxf(fmaf(dl[k + 3], 1.0f / 2048.0f, d[k + 3]) + b4.w, 0.f), acc);
        }
      }
      fence_before();
      if (cg > 0) sx[(cg - 1) * TP + row] = acc;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (cg == 0) {
        const long long i = tile * TP + row;
        if (i < a.npts) a.out[i] = (((acc + sx[row]) + sx[TP + row]) + sx[2 * TP + row]) + b3;
      }
      // (sx is rewritten only after the next tile's barriers: every reader is long past this point)
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace dtc

int decode_tc_init() {
  ISB_CUDA(cudaFuncSetAttribute(dtc::triplane_decode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dtc::SMEM_BYTES));
  ISB_CUDA(cudaFuncSetAttribute(dtc::triplane_decode_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dtc::SMEM_BYTES));
  return ISB_OK;
}

// h1_bound: caller-supplied upper bound of |layer-1 pre-activation| (0 = unknown)
bool decode_tc_usable(float h1_bound) { return h1_bound > 0.f && h1_bound <= 30000.0f; }

int decode_tc_launch(const float* planes, int R, const isb_triplane_mlp* w, const float* lin, int res, long long idx0,
                     const float* coords, long long npts, float* out, bool grid_mode, cudaStream_t st) {
  dtc::Args a{planes, R, w->fourier_B, w->w1, w->b1, w->w2, w->b2, w->w3, w->b3, lin, res, idx0, coords, npts, out, 0};
  // cooperative row sampling: whole tiles share (x, y), and the 32 z of a warp touch at most 31 plane rows
  if (grid_mode && res % dtc::TP == 0 && idx0 % dtc::TP == 0 && 31.0 * (R - 1) / (res - 1) + 2.0 <= 31.0) a.fast_rows = 1;
  const long long ntiles = (npts + dtc::TP - 1) / dtc::TP;
  const long long blocks = ntiles < num_sms() ? ntiles : num_sms();
  if (blocks < 1) return ISB_OK;
  if (grid_mode) ISB_CUDA(isb::launch(dtc::triplane_decode_tc_kernel<true>, static_cast<int>(blocks), dtc::THREADS, dtc::SMEM_BYTES, st, a));
  else ISB_CUDA(isb::launch(dtc::triplane_decode_tc_kernel<false>, static_cast<int>(blocks), dtc::THREADS, dtc::SMEM_BYTES, st, a));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // namespace isb
