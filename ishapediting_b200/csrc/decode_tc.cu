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
// One persistent CTA per SM (219 KB of shared memory: both layers' weights as hi / lo' fp16 K-major 128B-swizzled
// UMMA tiles = 128 KB, one A operand pair = 64 KB, staging).  288 threads: warps 0-7 compute (sampling, Fourier
// features, operand splitting, epilogues), warp 8 issues the MMAs.  Per 128-point tile:
//   P1  bilinear plane samples -> f[32]            (2 threads per point, 16 channels each)
//   P2  u = f @ B, [sin | cos](2 pi u) -> A_hi / A_lo in UMMA layout              -> mbarrier a_ready
//   M1  D1 = A_hi W1hi^T + A_lo W1hi^T,  D1' = A_hi W1lo'^T   (24 MMAs 128x128x16) -> tcgen05.commit d_ready
//   P3  h1 = relu(D1 + 2^-11 D1' + b1) -> split -> A_hi / A_lo (layer-2 operand)  -> a_ready
//   M2  D2, D2' likewise with W2                                                 -> d_ready
//   P4  logit = w3 . relu(D2 + 2^-11 D2' + b2) + b3  (fp32 registers)            -> global
#include "common.cuh"

namespace isb {
namespace dtc {

constexpr int TP = 128;                 // points per tile = UMMA M
constexpr int H = 128, F = 32, MF = 64; // hidden width, plane features, Fourier frequencies
constexpr int COMPUTE_THREADS = 256;
constexpr int THREADS = COMPUTE_THREADS + 32;
constexpr int ATOM = TP * 128;          // one K-atom: 128 rows x 64 fp16 = 16 KiB
constexpr int OPER = 2 * ATOM;          // one 128x128 fp16 operand = 32 KiB
// shared-memory map (bytes from the 1024-aligned base)
constexpr int OFF_W1H = 0, OFF_W1L = OPER, OFF_W2H = 2 * OPER, OFF_W2L = 3 * OPER;
constexpr int OFF_AH = 4 * OPER, OFF_AL = 5 * OPER;
constexpr int OFF_F = 6 * OPER;                      // float F[128][33]
constexpr int OFF_BM = OFF_F + TP * 33 * 4;          // float Bm[32][64]
constexpr int OFF_VEC = OFF_BM + F * MF * 4;         // float b1[128], b2[128], w3[128]
constexpr int OFF_X = OFF_VEC + 3 * H * 4;           // float xch[128]
constexpr int SMEM_BYTES = OFF_X + TP * 4 + 1024;    // + alignment slack

struct Args {
  const float* planes; int R;
  const float* fourier_B; const float* w1; const float* b1; const float* w2; const float* b2;
  const float* w3; const float* b3;
  const float* lin; int res; long long idx0;
  const float* coords;
  long long npts;
  float* out;
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

__device__ __forceinline__ void sample_plane16(const float* __restrict__ plane, int R, float gx, float gy, int c0, float* acc) {
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
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(p + q);
        acc[4 * q] = fmaf(w, v.x, acc[4 * q]); acc[4 * q + 1] = fmaf(w, v.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(w, v.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w, v.w, acc[4 * q + 3]);
      }
    }
}

template <bool GRID>
__global__ void __launch_bounds__(THREADS, 1)
triplane_decode_tc_kernel(const Args a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_d;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* sF = reinterpret_cast<float*>(gbase + OFF_F);
  float* sBm = reinterpret_cast<float*>(gbase + OFF_BM);
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
  for (int i = tid; i < F * MF; i += THREADS) sBm[i] = __ldg(a.fourier_B + i);
  if (tid < H) { sb1[tid] = __ldg(a.b1 + tid); sb2[tid] = __ldg(a.b2 + tid); sw3[tid] = __ldg(a.w3 + tid); }
  const float b3 = __ldg(a.b3);
  if (tid == 0) {
    mbar_init(smem_u32(&bar_a), COMPUTE_THREADS);
    mbar_init(smem_u32(&bar_d), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
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
  const size_t plane_sz = static_cast<size_t>(a.R) * a.R * F;

  if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D = f32, A = B = f16, both K-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(H >> 3) << 17) | (static_cast<uint32_t>(TP >> 4) << 24);
      const uint64_t d_ah = make_desc_sw128(a_hi), d_al = make_desc_sw128(a_lo);
      const uint32_t ba = smem_u32(&bar_a), bd = smem_u32(&bar_d);
      uint32_t ph = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
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
        }
      }
    }
  } else {
    // ===== compute warps =====
    const uint32_t ba = smem_u32(&bar_a), bd = smem_u32(&bar_d);
    uint32_t phd = 0;
    const int pt = tid >> 1, half = tid & 1;
    const int q = warp & 3, colh = warp >> 2;            // TMEM lane quarter / column half of this warp
    const int row = q * 32 + lane;
    const float two_pi = 6.283185307179586f;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // ---- P1: plane samples, 16 channels per thread ----
      {
        const long long i = tile * TP + pt;
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = 0.f;
        if (i < a.npts) {
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
          const int c0 = half * 16;
          sample_plane16(a.planes, a.R, cx, cy, c0, f);                 // xy plane: x->W, y->H
          sample_plane16(a.planes + plane_sz, a.R, cy, cz, c0, f);      // yz plane
          sample_plane16(a.planes + 2 * plane_sz, a.R, cx, cz, c0, f);  // xz plane
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) sF[pt * 33 + half * 16 + j] = f[j];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- P2: Fourier features of 32 frequencies per thread -> layer-1 operand ----
      {
        float u[32];
#pragma unroll
        for (int m = 0; m < 32; ++m) u[m] = 0.f;
        const float* bm = sBm + half * 32;
#pragma unroll 4
        for (int c = 0; c < F; ++c) {
          const float fc = sF[pt * 33 + c];
#pragma unroll
          for (int m4 = 0; m4 < 8; ++m4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bm + c * MF + m4 * 4);
            u[4 * m4] = fmaf(fc, b4.x, u[4 * m4]); u[4 * m4 + 1] = fmaf(fc, b4.y, u[4 * m4 + 1]);
            u[4 * m4 + 2] = fmaf(fc, b4.z, u[4 * m4 + 2]); u[4 * m4 + 3] = fmaf(fc, b4.w, u[4 * m4 + 3]);
          }
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float sv[8], cv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) sincosf(two_pi * u[g8 * 8 + j], &sv[j], &cv[j]);
          store_chunk(a_hi, a_lo, pt, half * 4 + g8, sv);            // k = m            (sin)
          store_chunk(a_hi, a_lo, pt, 8 + half * 4 + g8, cv);        // k = 64 + m       (cos)
        }
      }
      fence_async_smem();
      mbar_arrive(ba);
      // ---- P3: h1 = relu(D1 + 2^-11 D1' + b1) -> layer-2 operand ----
      mbar_wait(bd, phd);
      phd ^= 1u;
      fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = colh * 64 + c * 32;
        float d[32], dl[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(col0), d);
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(128 + col0), dl);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = g8 * 8 + j;
            v[j] = fmaxf(fmaf(dl[k], 1.0f / 2048.0f, d[k]) + sb1[col0 + k], 0.f);
          }
          store_chunk(a_hi, a_lo, row, (col0 >> 3) + g8, v);
        }
      }
      fence_async_smem();
      fence_before();
      mbar_arrive(ba);
      // ---- P4: logit = w3 . relu(D2 + 2^-11 D2' + b2) + b3 ----
      mbar_wait(bd, phd);
      phd ^= 1u;
      fence_after();
      float acc = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = colh * 64 + c * 32;
        float d[32], dl[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(256 + col0), d);
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(384 + col0), dl);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k)
          acc = fmaf(sw3[col0 + k], fmaxf(fmaf(dl[k], 1.0f / 2048.0f, d[k]) + sb2[col0 + k], 0.f), acc);
      }
      fence_before();
      if (colh == 1) sx[row] = acc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (colh == 0) {
        const long long i = tile * TP + row;
        if (i < a.npts) a.out[i] = (acc + sx[row]) + b3;
      }
      // (the next tile's P1 writes sF only; sx is rewritten after the next bar.sync 1 pair)
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 8) {
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
  dtc::Args a{planes, R, w->fourier_B, w->w1, w->b1, w->w2, w->b2, w->w3, w->b3, lin, res, idx0, coords, npts, out};
  const long long ntiles = (npts + dtc::TP - 1) / dtc::TP;
  const long long blocks = ntiles < num_sms() ? ntiles : num_sms();
  if (blocks < 1) return ISB_OK;
  if (grid_mode) ISB_CUDA(isb::launch(dtc::triplane_decode_tc_kernel<true>, static_cast<int>(blocks), dtc::THREADS, dtc::SMEM_BYTES, st, a));
  else ISB_CUDA(isb::launch(dtc::triplane_decode_tc_kernel<false>, static_cast<int>(blocks), dtc::THREADS, dtc::SMEM_BYTES, st, a));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // namespace isb
