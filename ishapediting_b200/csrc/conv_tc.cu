// conv_tc.cu — bf16 implicit-GEMM convolution (3x3 / 1x1 / linear) on the 5th-gen
// tensor cores: TMA (cp.async.bulk.tensor, 128B swizzle) -> smem ring -> tcgen05.mma
// with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
// GEMM view:  D[M = 128 output pixels, N = block_n output channels] +=
//             A[M, 64 input channels of one filter tap] * B[N, 64]^T
// per k-iteration.  The A tile of tap (kh,kw) is the NHWC box
// {64 ch, tw, th, nb} at (c0, w0+kw-1, h0+kh-1, n0): the TMA zero-fills the halo,
// so there is no im2col buffer and no materialised padding.  An optional second
// source (the 1x1 skip convolution of a ResBlock, unet.py:222,256) is appended to
// the K loop so `skip(x) + h` costs no extra pass.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = epilogue (TMEM lanes 32*(warp%4) ..).  Split-K partials go to a
// caller workspace; the last CTA to arrive for an output tile reduces them in fixed order
// (deterministic) and runs the epilogue — no separate reduction launch.
#include <stdlib.h>
#include "common.cuh"

namespace isb {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;
constexpr int TC_UMMA_K = 16;
constexpr int TC_A_STAGE = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KiB
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM_LIMIT = 227 * 1024 - 2048;  // dynamic part; the kernel also has ~150 B of static smem (barriers)

struct TcParams {
  int tw, th, nb;
  int tw_log2, pi_log2;   // tw and tw*th are powers of two
  int tiles_w, tiles_h, tiles_n;
  int N, H, W;
  int Cout, block_n;
  int Cin, chunks0, ntaps, chunks1;
  int k_iters, splits, stages;
  int w_tiled;     // 1: weights packed as [Cout/64][K/64][64][64] panels (8 KiB contiguous per TMA box row-group)
  int debug;       // profiling only: bit0 = producer skips the TMA loads, bit1 = MMA thread skips the MMAs
  int a_bytes;     // bytes of the A tile actually loaded per stage (= 16 KiB, or less when the image has < 128 pixels)
  int two_cta;     // 1: CTA pair (cta_group::2): 256-row MMA, each CTA stages its A half and half of the B tile
  int cluster;     // 1: the `splits` CTAs of an output tile form a thread-block cluster (DSMEM reduce)
  int tmem_cols;
  const float* bias;
  const float* residual;
  void* out;
  int out_dtype;
  int accumulate;
  float* partial;  // != nullptr when splits > 1: [tile][split][128][block_n] fp32
  int* counters;   // [tiles] arrival counters (zero between launches)
  unsigned long long* trace;  // profiling only (debug bit2): [cta][16] %globaltimer stamps of the pipeline phases
  // GroupNorm statistics of the output, fused into the epilogue: every CTA writes the (sum, sum of squares) of its
  // part of each group of 2^gn_cg_log2 consecutive output channels to gn_part[n][group][slot] (no atomics).
  float* gn_part;
  int gn_cg_log2, gn_slots, gn_groups;
  // gn_mode 2: the output is dy of a GroupNorm(+SiLU) whose input gb_x has the same shape; the partials are the two
  // sums of its backward pass, sum(dz*gamma') and sum(dz*gamma'*xhat)  (dz = dy * silu'(z), z = xhat*gamma' + beta')
  int gn_mode;
  const float* gb_x; const float* gb_gamma; const float* gb_beta; const float* gb_film; const float* gb_stats;
  int gb_film_stride, gb_silu;
};

// ---- PTX wrappers ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// profiling only: one phase stamp of this CTA (slot 15 holds the SM id)
__device__ __forceinline__ void tc_stamp(const unsigned long long* trace_c, int slot) {
  unsigned long long* trace = const_cast<unsigned long long*>(trace_c);
  if (trace) {
    const size_t cta = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    trace[cta * 16 + slot] = global_timer_ns();
  }
}
// profiling only: per-k-iteration SM-clock stamps of CTA (0,0,0), rows of 4 behind the 3000 CTA records
__device__ __forceinline__ void tc_stamp_iter(const unsigned long long* trace_c, int iter, int slot) {
  unsigned long long* trace = const_cast<unsigned long long*>(trace_c);
  if (trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && iter < 160)
    trace[3000 * 16 + iter * 4 + slot] = static_cast<unsigned long long>(clock64());
}
// Bounded wait: a pipeline bug must trap (CUDA error to the caller), never hang the GPU.  The fast path is
// a bare try_wait loop (the instruction itself suspends the thread for a HW time slice); the SM cycle
// counter is consulted only every 4096 failed probes — no %globaltimer reads on the critical path.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 6000000000ll) __trap();   // ~3 s at 1.9 GHz
    }
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);  // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                   // [16,30) LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;           // [32,46) SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;                   // [46,48) descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                   // [61,64) SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes from the shared memory of CTA `rank` of this cluster, at the same offset as local address `laddr`
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t laddr, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(raddr)
               : "memory");
  return v;
}

// ---- CTA-pair (cta_group::2) helpers ---------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads issued by either CTA of the pair; completion bytes are credited to the LEADER's mbarrier
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (when all MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// bias + residual (+ previous out) and store 8 consecutive output channels
__device__ __forceinline__ void epilogue_store8(const TcParams& p, float* f, size_t off, int col) {
  if (p.bias != nullptr) {
    float b[8];
    load8(p.bias + col, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += b[j];
  }
  if (p.residual != nullptr) {
    float b[8];
    load8(p.residual + off, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += b[j];
  }
  if (p.out_dtype == ISB_F32 && p.accumulate) {
    const float4* o = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + off);
    const float4 a = o[0], b = o[1];
    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
    f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
  }
  store8(p.out, off, p.out_dtype, f);
}

// Direct epilogue of one warp (32 accumulator rows = TMEM lanes 32q..32q+31).  tcgen05.ld hands every lane 32
// consecutive channels of ITS pixel row, and rows are Cout*4 bytes apart in memory: storing from that layout
// touches 32 different 128-byte lines per instruction (measured: 7-9 us per 128x128..256 tile, 4x the HBM time).
// So each 32x32 chunk is transposed through a padded per-warp staging area (stride 36 floats: conflict-free for
// both the row-wise 16-byte writes and the 4-rows-by-128-bytes reads) and leaves as full 128-byte row segments;
// bias / residual / accumulate are applied on the coalesced side.
constexpr int TC_STG_STRIDE = 36 * 4;                  // bytes per staged row
constexpr int TC_STG_WARP = 32 * TC_STG_STRIDE;        // 4608 B per warp
constexpr uint32_t TC_GN_STAGE_OFF = 4 * TC_STG_WARP;  // 2 KB of statistic partials behind the four transpose buffers
struct EpiDst {
  void* out;              // row-major [rows][ld]
  int dtype, ld;
  const float* bias;      // indexed by destination column, or nullptr
  const float* residual;  // same layout as out (fp32), or nullptr
  int accumulate;
  int col_limit;          // columns >= col_limit do not exist (multiple of 32)
  uint32_t gn_sm;         // != 0: shared-memory staging [4 warps][8 chunks][4 groups] float2 for the fused statistics
  int cg_log2;
  const TcParams* gb;     // != nullptr: accumulate the GroupNorm-backward sums instead of (sum, sum of squares)
};
// per-lane constants of the GroupNorm-backward terms for 4 consecutive channels of image n
struct GbConst {
  float ga[4], be[4], mean, rstd;
};
__device__ __forceinline__ void gb_load(const TcParams& p, int n, int c, GbConst& k) {
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gb_gamma + c));
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.gb_beta + c));
  k.ga[0] = g4.x; k.ga[1] = g4.y; k.ga[2] = g4.z; k.ga[3] = g4.w;
  k.be[0] = b4.x; k.be[1] = b4.y; k.be[2] = b4.z; k.be[3] = b4.w;
  if (p.gb_film != nullptr) {
    const float* f = p.gb_film + static_cast<size_t>(n) * p.gb_film_stride;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(f + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(f + p.Cout + c));
    k.ga[0] *= 1.0f + sc.x; k.ga[1] *= 1.0f + sc.y; k.ga[2] *= 1.0f + sc.z; k.ga[3] *= 1.0f + sc.w;
    k.be[0] = k.be[0] * (1.0f + sc.x) + sh.x; k.be[1] = k.be[1] * (1.0f + sc.y) + sh.y;
    k.be[2] = k.be[2] * (1.0f + sc.z) + sh.z; k.be[3] = k.be[3] * (1.0f + sc.w) + sh.w;
  }
  const int g = c >> p.gn_cg_log2;
  k.mean = __ldg(p.gb_stats + (n * p.gn_groups + g) * 2);
  k.rstd = __ldg(p.gb_stats + (n * p.gn_groups + g) * 2 + 1);
}
// adds this element vector's contribution: s1 += dz*gamma', s2 += dz*gamma'*xhat
__device__ __forceinline__ void gb_accumulate(const TcParams& p, const GbConst& k, float4 dy, size_t off, float& s1, float& s2) {
  const float4 x4 = __ldg(reinterpret_cast<const float4*>(p.gb_x + off));
  const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ds[4] = {dy.x, dy.y, dy.z, dy.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float xhat = (xs[j] - k.mean) * k.rstd;
    float d = ds[j];
    if (p.gb_silu) d *= silu_grad_f(fmaf(xhat, k.ga[j], k.be[j]));
    const float dzg = d * k.ga[j];
    s1 += dzg;
    s2 = fmaf(dzg, xhat, s2);
  }
}
// residual (+ previous out) and store 4 consecutive channels at element offset `off` (bias already added);
// returns the stored fp32 value (the statistics are taken over what the consumer will read)
__device__ __forceinline__ float4 epilogue_store4_nb(const EpiDst& e, float4 f, size_t off) {
  if (e.residual != nullptr) {
    const float4 r4 = __ldg(reinterpret_cast<const float4*>(e.residual + off));
    f.x += r4.x; f.y += r4.y; f.z += r4.z; f.w += r4.w;
  }
  if (e.dtype == ISB_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + off);
    if (e.accumulate) {
      const float4 a4 = *o;
      f.x += a4.x; f.y += a4.y; f.z += a4.z; f.w += a4.w;
    }
    *o = f;
  } else {
    uint2 u;
    u.x = pack_bf16x2(f.x, f.y);
    u.y = pack_bf16x2(f.z, f.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(e.out) + off) = u;
  }
  return f;
}
__device__ __forceinline__ float4 epilogue_store4(const EpiDst& e, float4 f, size_t off, int col) {
  if (e.bias != nullptr) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
    f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
  }
  return epilogue_store4_nb(e, f, off);
}
// `row` = destination row of this lane's accumulator row, `col_base` = destination column of accumulator column 0
__device__ __forceinline__ void epilogue_direct_warp(const EpiDst& e, int nchunks, uint32_t tmem_base, uint32_t stg,
                                                     int q, int lane, uint32_t row, bool valid, int col_base,
                                                     const unsigned long long* trace = nullptr, int gb_n = 0) {
  const int sub = lane >> 3;
  const int cv = (lane & 7) * 4;
  const uint32_t my_row = stg + static_cast<uint32_t>(lane) * TC_STG_STRIDE;
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  for (int c = 0; c < nchunks; ++c) {
    const int col0 = col_base + c * 32;
    if (col0 >= e.col_limit) break;  // warp-uniform
    uint32_t v[32];
    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), v);
    tmem_ld_wait();
    if (c == 0 && q == 2 && lane == 0) tc_stamp(trace, 11);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + static_cast<uint32_t>(j * 16)),
                   "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                   : "memory");
    __syncwarp();
    if (c == 0 && q == 2 && lane == 0) tc_stamp(trace, 12);
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col0 + cv));
    // all eight transposed reads first (back to back), then the dependent shuffle / add / store chains: the
    // volatile asm keeps program order, so one fused loop serialised load -> store eight times (~110 cycles each)
    float4 f[8];
#pragma unroll
    for (int g = 0; g < 8; ++g)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(f[g].x), "=f"(f[g].y), "=f"(f[g].z), "=f"(f[g].w)
                   : "r"(stg + static_cast<uint32_t>((g * 4 + sub) * TC_STG_STRIDE + cv * 4)));
    float gs = 0.f, gq = 0.f;
    GbConst gk;
    if (e.gb != nullptr) gb_load(*e.gb, gb_n, col0 + cv, gk);
    // The residual (and the previous output when accumulating) of all eight row groups is requested BEFORE the first
    // store: inside the store loop every load sat behind the previous iteration's store (the two pointers may alias as
    // far as the compiler knows), i.e. eight DRAM / L2 round trips in a row per 32-column chunk — a 256->256 conv at
    // 8 x 128 x 128 took 290 us with a residual and 171 us without.
    size_t offs[8];
    float4 r4[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int rsel = g * 4 + sub;
      const uint32_t m_row = __shfl_sync(0xffffffffu, row, rsel);
      offs[g] = static_cast<size_t>(m_row) * e.ld + col0 + cv;
      r4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (e.residual != nullptr) {
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if ((vmask >> (g * 4 + sub)) & 1u) r4[g] = __ldg(reinterpret_cast<const float4*>(e.residual + offs[g]));
    }
    EpiDst e2 = e;              // residual handled here; epilogue_store4_nb keeps the accumulate / dtype logic
    e2.residual = nullptr;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int rsel = g * 4 + sub;
      if (!((vmask >> rsel) & 1u)) continue;
      f[g].x += b4.x; f[g].y += b4.y; f[g].z += b4.z; f[g].w += b4.w;
      if (e.residual != nullptr) { f[g].x += r4[g].x; f[g].y += r4[g].y; f[g].z += r4[g].z; f[g].w += r4[g].w; }
      const size_t off = offs[g];
      const float4 o4 = epilogue_store4_nb(e2, f[g], off);
      if (e.gb != nullptr) {
        gb_accumulate(*e.gb, gk, o4, off, gs, gq);
      } else {
        gs += (o4.x + o4.y) + (o4.z + o4.w);
        gq = fmaf(o4.x, o4.x, fmaf(o4.y, o4.y, fmaf(o4.z, o4.z, fmaf(o4.w, o4.w, gq))));
      }
    }
    if (e.gn_sm != 0) {
      // the 32 rows of this warp: fold the four row-subsets, then the cg/4 lanes that share a group
      gs += __shfl_xor_sync(0xffffffffu, gs, 8);  gq += __shfl_xor_sync(0xffffffffu, gq, 8);
      gs += __shfl_xor_sync(0xffffffffu, gs, 16); gq += __shfl_xor_sync(0xffffffffu, gq, 16);
      const int lanes_per_group = 1 << (e.cg_log2 - 2);
      for (int w = 1; w < lanes_per_group; w <<= 1) {
        gs += __shfl_xor_sync(0xffffffffu, gs, w);
        gq += __shfl_xor_sync(0xffffffffu, gq, w);
      }
      if (sub == 0 && (lane & (lanes_per_group - 1)) == 0)
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(e.gn_sm + static_cast<uint32_t>(((q * 8 + c) * 4 + (lane >> (e.cg_log2 - 2))) * 8)),
                     "f"(gs), "f"(gq)
                     : "memory");
    }
    __syncwarp();
    if (c == 0 && q == 2 && lane == 0) tc_stamp(trace, 13);
  }
}

// ---- GroupNorm statistics of the conv output (fused; see TcParams::gn_part) ----------------------------------
// slot of this CTA's contribution within its image(s)
__device__ __forceinline__ int gn_slot(const TcParams& p, int tile_w, int tile_h, int split) {
  const int per_tile = p.cluster ? p.splits : 1;
  return (p.nb == 1 ? (tile_h * p.tiles_w + tile_w) * per_tile : 0) + (p.cluster ? split : 0);
}
// direct epilogue: the four warps' staged sums -> gn_part (threads et = 0..127, all must call)
__device__ __forceinline__ void gn_flush_direct(const TcParams& p, uint32_t gn_sm, int et, int n0, int cout0, int slot) {
  asm volatile("bar.sync 1, 128;" ::: "memory");
  int cols = p.Cout - cout0;
  if (cols > p.block_n) cols = p.block_n;
  const int ngroups = cols >> p.gn_cg_log2;
  const int gpc = 32 >> p.gn_cg_log2;        // groups per 32-column chunk
  const int warps_per_img = 4 / p.nb;        // nb = 1 or 2 images per 128-row tile
  for (int i = et; i < ngroups * p.nb; i += 128) {
    const int img = i / ngroups, gi = i - img * ngroups;
    if (n0 + img >= p.N) continue;
    const int c = gi / gpc, k = gi - c * gpc;
    float s = 0.f, q = 0.f;
    for (int w = img * warps_per_img; w < (img + 1) * warps_per_img; ++w) {
      float a, b;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(gn_sm + static_cast<uint32_t>(((w * 8 + c) * 4 + k) * 8)) : "memory");
      s += a;
      q += b;
    }
    float2* dst = reinterpret_cast<float2*>(p.gn_part) +
                  (static_cast<size_t>(n0 + img) * p.gn_groups + (cout0 >> p.gn_cg_log2) + gi) * p.gn_slots + slot;
    *dst = make_float2(s, q);
  }
}
// split-K fold: every thread owns one 4-channel vector (fixed) over some rows of ONE image
__device__ __forceinline__ void gn_flush_fold(const TcParams& p, uint32_t gn_sm, int et, float gs, float gq, int n0,
                                              int my_img, bool has_rows, int cout0, int slot) {
  const int vec_per_row = p.block_n / 4;     // 16, 32 or 64
  const int lanes_per_group = 1 << (p.gn_cg_log2 - 2);
  for (int w = 1; w < lanes_per_group; w <<= 1) {
    gs += __shfl_xor_sync(0xffffffffu, gs, w);
    gq += __shfl_xor_sync(0xffffffffu, gq, w);
  }
  for (int w = vec_per_row; w < 32; w <<= 1) {   // several rows per warp instruction (block_n = 64)
    gs += __shfl_xor_sync(0xffffffffu, gs, w);
    gq += __shfl_xor_sync(0xffffffffu, gq, w);
  }
  const int warp = et >> 5, lane = et & 31;
  const int span = vec_per_row < 32 ? vec_per_row : 32;                  // vectors covered by one warp
  const int v0 = vec_per_row > 32 ? (32 * warp) % vec_per_row : 0;       // first vector of this warp
  if (lane < span && (lane & (lanes_per_group - 1)) == 0) {
    const int gi = (v0 + lane) >> (p.gn_cg_log2 - 2);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(gn_sm + static_cast<uint32_t>((warp * 64 + gi) * 8)), "f"(gs), "f"(gq) : "memory");
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  int cols = p.Cout - cout0;
  if (cols > p.block_n) cols = p.block_n;
  const int ngroups = cols >> p.gn_cg_log2;
  for (int gi = et; gi < ngroups; gi += 128) {
    const int v = gi << (p.gn_cg_log2 - 2);
    float s = 0.f, q = 0.f;
    for (int w = 0; w < 4; ++w) {
      const int wv0 = vec_per_row > 32 ? (32 * w) % vec_per_row : 0;
      if (v < wv0 || v >= wv0 + span) continue;
      float a, b;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(gn_sm + static_cast<uint32_t>((w * 64 + gi) * 8)) : "memory");
      s += a;
      q += b;
    }
    for (int img = 0; img < p.nb; ++img) {
      if (n0 + img >= p.N) continue;
      const bool mine = has_rows && img == my_img;
      float2* dst = reinterpret_cast<float2*>(p.gn_part) +
                    (static_cast<size_t>(n0 + img) * p.gn_groups + (cout0 >> p.gn_cg_log2) + gi) * p.gn_slots + slot;
      *dst = mine ? make_float2(s, q) : make_float2(0.f, 0.f);
    }
  }
}

// Cluster split-K fold.  Every CTA of the cluster has written its fp32 partial tile (row-major [128][block_n],
// coalesced) to the L2-resident workspace; CTA `split` owns rows [split*R, (split+1)*R) of the VALID rows of the
// output tile and sums them over the S partials in rank order (deterministic).  All loads of a batch (16 x 16 B per
// thread) are issued before the first add.  (The first version of this fold pulled the partials from the peers'
// shared memory through DSMEM: the phase trace showed ~6 B/cycle/SM, 3-8 us per launch; L2 sustains 10x that.)
template <int S>
__device__ __forceinline__ void cluster_fold(const TcParams& p, const EpiDst& e, const float* tile_base, int split,
                                             int et, int n0, int h0, int w0, int cout0, int gn_slot_fold) {
  constexpr int U = 16 / S;
  const int per_img = p.tw * p.th;
  int valid_rows = (p.N - n0) * per_img;
  if (valid_rows > TC_BLOCK_M) valid_rows = TC_BLOCK_M;
  const int R = (valid_rows + S - 1) / S;
  const int row0 = split * R;
  int rows = valid_rows - row0;
  if (rows > R) rows = R;
  const int vec_per_row = p.block_n / 4;
  const int nvec = rows * vec_per_row;          // <= 0 when this rank owns no valid row
  const size_t tile_elems = static_cast<size_t>(TC_BLOCK_M) * p.block_n;
  float gs = 0.f, gq = 0.f;
  GbConst gk;
  if (e.gb != nullptr) {     // this thread's 4-channel vector is the same in every batch; its rows lie in one image
    const int my_n = n0 + (rows > 0 ? row0 >> p.pi_log2 : 0);
    const int my_c = cout0 + (et % vec_per_row) * 4;
    gb_load(p, my_n < p.N ? my_n : p.N - 1, my_c < p.Cout ? my_c : 0, gk);
  }
  EpiDst e2 = e;                // the residual is loaded together with the partials (not behind the previous store)
  e2.residual = nullptr;
  for (int base = et; base < nvec; base += 128 * U) {
    float4 buf[U][S];
    float4 rv[U];
    size_t offv[U];
    int cc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * 128;
      const int rl = idx / vec_per_row;
      const int rr = row0 + rl;
      cc[u] = (idx - rl * vec_per_row) * 4;
      const int rn = rr >> p.pi_log2;
      const int rrem = rr & (per_img - 1);
      const int rh = rrem >> p.tw_log2;
      const int rw = rrem & (p.tw - 1);
      const size_t mm = (static_cast<size_t>(n0 + rn) * p.H + (h0 + rh)) * p.W + (w0 + rw);
      offv[u] = mm * p.Cout + cout0 + cc[u];
      if (idx < nvec) {
        const float* src = tile_base + static_cast<size_t>(rr) * p.block_n + cc[u];
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) buf[u][s2] = __ldcg(reinterpret_cast<const float4*>(src + s2 * tile_elems));
        if (e.residual != nullptr && cout0 + cc[u] < p.Cout) rv[u] = __ldg(reinterpret_cast<const float4*>(e.residual + offv[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * 128;
      const int col = cout0 + cc[u];
      if (idx >= nvec || col >= p.Cout) continue;
      float4 f = buf[u][0];
#pragma unroll
      for (int s2 = 1; s2 < S; ++s2) {
        f.x += buf[u][s2].x; f.y += buf[u][s2].y; f.z += buf[u][s2].z; f.w += buf[u][s2].w;
      }
      if (et == 0 && base == 0 && u == 0) tc_stamp(p.trace, 10);
      if (e.bias != nullptr) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
        f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
      }
      if (e.residual != nullptr) { f.x += rv[u].x; f.y += rv[u].y; f.z += rv[u].z; f.w += rv[u].w; }
      const float4 o4 = epilogue_store4_nb(e2, f, offv[u]);
      if (e.gb != nullptr) {
        gb_accumulate(p, gk, o4, offv[u], gs, gq);
      } else {
        gs += (o4.x + o4.y) + (o4.z + o4.w);
        gq = fmaf(o4.x, o4.x, fmaf(o4.y, o4.y, fmaf(o4.z, o4.z, fmaf(o4.w, o4.w, gq))));
      }
    }
  }
  if (p.gn_part != nullptr)
    gn_flush_fold(p, e.gn_sm, et, gs, gq, n0, rows > 0 ? row0 >> p.pi_log2 : 0, rows > 0, cout0, gn_slot_fold);
}

// ---- the kernel -------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
               const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  pdl_trigger();   // let the next kernel's prologue overlap this kernel (it blocks in its own pdl_wait)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int split_is_last;
  if (threadIdx.x == 0) {
    tc_stamp(p.trace, 0);
    if (p.trace) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const size_t cta = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      p.trace[cta * 16 + 15] = smid;
    }
  }

  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024 B alignment
  // Images with fewer than 128 pixels (8x8): only the valid rows of the A tile are loaded and a stage shrinks to them,
  // so the ring holds more WEIGHT bytes (what these layers stream).  The MMA still reads 128 rows — the rows past the
  // loaded ones alias the stage's weight tile: garbage that lands in accumulator rows the epilogue never stores.
  const uint32_t a_stage = static_cast<uint32_t>(p.a_bytes);
  const uint32_t stage_bytes = a_stage + static_cast<uint32_t>(p.block_n) * 128u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // output tile of this CTA
  int mt = blockIdx.x;
  const int tile_w = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int tile_h = mt % p.tiles_h;
  const int tile_n = mt / p.tiles_h;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = tile_n * p.nb;
  const int cout0 = blockIdx.y * p.block_n;
  const int split = blockIdx.z;
  const int k_begin = static_cast<int>(static_cast<long long>(p.k_iters) * split / p.splits);
  const int k_end = static_cast<int>(static_cast<long long>(p.k_iters) * (split + 1) / p.splits);
  EpiDst out_dst;
  out_dst.out = p.out;
  out_dst.dtype = p.out_dtype;
  out_dst.ld = p.Cout;
  out_dst.bias = p.bias;
  out_dst.residual = p.residual;
  out_dst.accumulate = p.accumulate;
  out_dst.col_limit = p.Cout;
  out_dst.gn_sm = p.gn_part != nullptr ? tiles_addr + TC_GN_STAGE_OFF : 0u;
  out_dst.cg_log2 = p.gn_cg_log2;
  out_dst.gb = (p.gn_part != nullptr && p.gn_mode == 2) ? &p : nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&tmem_slot)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) tc_stamp(p.trace, 1);

  if (warp == 0) {
    // ===== TMA producer =====
    // One thread, latency-bound: every dependent instruction costs its full latency, so the loop carries its
    // state (stage, phase, filter tap, channel chunk) incrementally — no integer divisions per k-iteration.
    if (lane == 0) {
      const int seg0_iters = p.ntaps * p.chunks0;
      const int niter = k_end - k_begin;
      const int npre = niter < p.stages ? niter : p.stages;
      const bool skip_tma = (p.debug & 1) != 0;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      // The weight panels never change during a step: start streaming them into the ring before the
      // producer of our activations has even finished (PDL), then wait and fetch the activations.
      // (weight K coordinate of k-iteration `it` is simply it * 64 in both K segments)
      for (int i = 0; i < npre; ++i) {
        mbar_expect_tx(full0 + 8u * i, skip_tma ? 0u : stage_bytes);
        if (!skip_tma) {
          const uint32_t b_dst = tiles_addr + static_cast<uint32_t>(i) * stage_bytes + a_stage;
          if (p.w_tiled) tma_load_4d(b_dst, &mapB, full0 + 8u * i, 0, 0, k_begin + i, cout0 / 64);
          else tma_load_2d(b_dst, &mapB, full0 + 8u * i, (k_begin + i) * TC_BLOCK_K, cout0);
        }
      }
      pdl_wait();
      tc_stamp(p.trace, 2);
      // activation-side iterator
      int tap, chunk, dh = 0, dw = 0;
      bool seg1 = k_begin >= seg0_iters;
      if (seg1) {
        tap = p.ntaps;
        chunk = k_begin - seg0_iters;
      } else {
        tap = k_begin / p.chunks0;
        chunk = k_begin - tap * p.chunks0;
        if (p.ntaps == 9) {
          dh = tap / 3 - 1;
          dw = tap % 3 - 1;
        }
      }
      auto load_a = [&](uint32_t a_dst, uint32_t fb) {
        if (!seg1) tma_load_4d(a_dst, &mapA, fb, chunk * TC_BLOCK_K, w0 + dw, h0 + dh, n0);
        else tma_load_4d(a_dst, &mapA2, fb, chunk * TC_BLOCK_K, w0, h0, n0);
        if (++chunk == p.chunks0 && !seg1) {
          chunk = 0;
          if (++tap == p.ntaps) {
            seg1 = true;
            dh = 0;
            dw = 0;
          } else if (p.ntaps == 9 && ++dw == 2) {
            dw = -1;
            ++dh;
          }
        }
      };
      if (!skip_tma)
        for (int i = 0; i < npre; ++i) load_a(tiles_addr + static_cast<uint32_t>(i) * stage_bytes, full0 + 8u * i);
      int s = npre == p.stages ? 0 : npre;
      uint32_t ph = 0;                     // parity of the "slot free" phase being waited for
      for (int i = npre; i < niter; ++i) {
        mbar_wait(empty0 + 8u * s, ph);
        tc_stamp_iter(p.trace, i, 0);
        const uint32_t fb = full0 + 8u * s;
        const uint32_t a_dst = tiles_addr + static_cast<uint32_t>(s) * stage_bytes;
        mbar_expect_tx(fb, skip_tma ? 0u : stage_bytes);
        if (!skip_tma) {
          load_a(a_dst, fb);
          if (p.w_tiled) tma_load_4d(a_dst + a_stage, &mapB, fb, 0, 0, k_begin + i, cout0 / 64);
          else tma_load_2d(a_dst + a_stage, &mapB, fb, (k_begin + i) * TC_BLOCK_K, cout0);
        }
        tc_stamp_iter(p.trace, i, 1);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, K-major both, N=block_n, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                             (static_cast<uint32_t>(p.block_n >> 3) << 17) |
                             (static_cast<uint32_t>(TC_BLOCK_M >> 4) << 24);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const int niter = k_end - k_begin;
      const bool skip_mma = (p.debug & 2) != 0;
      int s = 0;
      uint32_t ph = 0;
      uint64_t adesc = make_desc_sw128(tiles_addr);
      const uint64_t desc_step = static_cast<uint64_t>(stage_bytes >> 4);
      const uint64_t desc_b_off = static_cast<uint64_t>(a_stage >> 4);
      const uint64_t adesc0 = adesc;
      for (int i = 0; i < niter; ++i) {
        mbar_wait(full0 + 8u * s, ph);
        tc_fence_after();
        if (i == 0) tc_stamp(p.trace, 3);
        tc_stamp_iter(p.trace, i, 2);
        if (!skip_mma) {
          const uint64_t bdesc = adesc + desc_b_off;
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in (addr >> 4) units
            umma_bf16(tmem_base, adesc + static_cast<uint64_t>(k * 2),
                      bdesc + static_cast<uint64_t>(k * 2), idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty0 + 8u * s);  // frees the smem slot when these MMAs retire
        tc_stamp_iter(p.trace, i, 3);
        adesc += desc_step;
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
          adesc = adesc0;
        }
      }
      umma_commit(smem_u32(&tmem_full_bar));   // accumulator complete
      tc_stamp(p.trace, 4);
    }
  } else {
    // ===== epilogue warps 2..5 =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int per_img = p.tw * p.th;
    const int pn = r / per_img;
    const int rem = r - pn * per_img;
    const int ph_ = rem / p.tw;
    const int pw_ = rem - ph_ * p.tw;
    const int n = n0 + pn;
    const bool valid = n < p.N;
    const size_t m = (static_cast<size_t>(n) * p.H + (h0 + ph_)) * p.W + (w0 + pw_);
    const int nchunks = p.block_n / 32;

    mbar_wait(smem_u32(&tmem_full_bar), 0);
    tc_fence_after();
    if (threadIdx.x == 64) tc_stamp(p.trace, 5);
    pdl_wait();   // residual / accumulate reads and every global write come after the predecessor

    bool do_final = true;
    const size_t tile_id = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
    if (p.cluster) {
      // Cluster split-K: park the fp32 partial tile in the workspace (stays in L2); the fold happens after the
      // cluster barrier below.
      do_final = false;
      EpiDst pe;
      pe.out = p.partial + (tile_id * p.splits + split) * static_cast<size_t>(TC_BLOCK_M) * p.block_n;
      pe.dtype = ISB_F32;
      pe.ld = p.block_n;
      pe.bias = nullptr;
      pe.residual = nullptr;
      pe.accumulate = 0;
      pe.col_limit = p.Cout - cout0 < p.block_n ? p.Cout - cout0 : p.block_n;
      pe.gn_sm = 0;
      pe.cg_log2 = 0;
      pe.gb = nullptr;
      epilogue_direct_warp(pe, nchunks, tmem_base, tiles_addr + static_cast<uint32_t>(q) * TC_STG_WARP, q, lane,
                           static_cast<uint32_t>(r), valid, 0, p.trace);
    } else if (p.splits > 1) {
      // Split-K: every CTA parks its fp32 partial tile in the workspace; the LAST CTA to arrive for this
      // output tile folds all partials in split order (deterministic) and runs the real epilogue.
      float* mine = p.partial + ((tile_id * p.splits + split) * TC_BLOCK_M + r) * p.block_n;
      for (int c = 0; c < nchunks; ++c) {
        if (cout0 + c * 32 >= p.Cout) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), v);
        tmem_ld_wait();
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(mine + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const int prev = atomicAdd(p.counters + tile_id, 1);
        const int last = (prev == p.splits - 1);
        if (last) p.counters[tile_id] = 0;  // leave the workspace ready for the next launch
        split_is_last = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      do_final = split_is_last != 0;
      if (do_final) __threadfence();
    }

    if (do_final && p.splits == 1) {
      epilogue_direct_warp(out_dst, nchunks, tmem_base, tiles_addr + static_cast<uint32_t>(q) * TC_STG_WARP, q, lane,
                           static_cast<uint32_t>(m), valid, cout0, nullptr, valid ? n : p.N - 1);
      if (p.gn_part != nullptr)
        gn_flush_direct(p, out_dst.gn_sm, static_cast<int>(threadIdx.x) - 64, n0, cout0, gn_slot(p, tile_w, tile_h, 0));
    } else if (do_final) {
      // fold: the 128 epilogue threads sweep the tile as a flat array (coalesced 32 B per thread),
      // summing the partials in split order
      const int et = threadIdx.x - 64;
      const int vec_per_row = p.block_n / 8;
      const int nvec = TC_BLOCK_M * vec_per_row;
      const size_t tile_elems = static_cast<size_t>(TC_BLOCK_M) * p.block_n;
      const float* tile_base = p.partial + tile_id * p.splits * tile_elems;
      for (int idx = et; idx < nvec; idx += 128) {
        const int rr = idx / vec_per_row;
        const int cc = (idx - rr * vec_per_row) * 8;
        const int col = cout0 + cc;
        const int rn = rr / per_img;
        const int rrem = rr - rn * per_img;
        const int rh = rrem / p.tw;
        const int rw = rrem - rh * p.tw;
        if (n0 + rn >= p.N || col >= p.Cout) continue;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* src = tile_base + static_cast<size_t>(rr) * p.block_n + cc;
        for (int s2 = 0; s2 < p.splits; ++s2) {
          const float4 u0 = __ldcg(reinterpret_cast<const float4*>(src));
          const float4 u1 = __ldcg(reinterpret_cast<const float4*>(src) + 1);
          f[0] += u0.x; f[1] += u0.y; f[2] += u0.z; f[3] += u0.w;
          f[4] += u1.x; f[5] += u1.y; f[6] += u1.z; f[7] += u1.w;
          src += tile_elems;
        }
        const size_t mm = (static_cast<size_t>(n0 + rn) * p.H + (h0 + rh)) * p.W + (w0 + rw);
        epilogue_store8(p, f, mm * p.Cout + col, col);
      }
    }
  }

  if (threadIdx.x == 64) tc_stamp(p.trace, 6);
  if (p.cluster) {
    cluster_sync_all();   // every CTA of the cluster has parked its partial tile
    if (threadIdx.x == 64) tc_stamp(p.trace, 7);
    if (warp >= 2) {
      const int et = threadIdx.x - 64;
      const size_t tile_id = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
      const float* tile_base = p.partial + tile_id * p.splits * static_cast<size_t>(TC_BLOCK_M) * p.block_n;
      if (p.splits == 2) cluster_fold<2>(p, out_dst, tile_base, split, et, n0, h0, w0, cout0, gn_slot(p, tile_w, tile_h, split));
      else if (p.splits == 4) cluster_fold<4>(p, out_dst, tile_base, split, et, n0, h0, w0, cout0, gn_slot(p, tile_w, tile_h, split));
      else if (p.splits == 8) cluster_fold<8>(p, out_dst, tile_base, split, et, n0, h0, w0, cout0, gn_slot(p, tile_w, tile_h, split));
      else cluster_fold<16>(p, out_dst, tile_base, split, et, n0, h0, w0, cout0, gn_slot(p, tile_w, tile_h, split));
    }
    if (threadIdx.x == 64) tc_stamp(p.trace, 8);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
  if (threadIdx.x == 64) tc_stamp(p.trace, 9);
}

// ---- CTA-pair kernel: two SMs of one TPC cooperate on a 256 x block_n output tile ---------------
// (cta_group::2).  Each CTA stages its own 128 pixel rows of A and HALF of the B (weight) tile, so the
// shared-memory fill per FLOP — what bounds the single-CTA kernel on the 64^2 / 128^2 layers
// (profiles/r01_ncu_full_conv_tc.md) — drops by 1/3 (block_n = 128) to 1/2 (block_n = 256).  The leader CTA's
// MMA thread issues tcgen05.mma.cta_group::2 for the pair; TMA completions from both CTAs are credited to the
// leader's "full" barrier; tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both.
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
                const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  const int half_n = p.block_n / 2;
  const uint32_t stage_bytes = TC_A_STAGE + static_cast<uint32_t>(half_n) * 128u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  int mt = blockIdx.x;
  const int tile_w = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int tile_h = mt % p.tiles_h;
  const int tile_n = mt / p.tiles_h;
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = tile_n * p.nb;
  const int cout0 = blockIdx.y * p.block_n;
  EpiDst out_dst;
  out_dst.out = p.out;
  out_dst.dtype = p.out_dtype;
  out_dst.ld = p.Cout;
  out_dst.bias = p.bias;
  out_dst.residual = p.residual;
  out_dst.accumulate = p.accumulate;
  out_dst.col_limit = p.Cout;
  out_dst.gn_sm = p.gn_part != nullptr ? tiles_addr + TC_GN_STAGE_OFF : 0u;
  out_dst.cg_log2 = p.gn_cg_log2;
  out_dst.gb = (p.gn_part != nullptr && p.gn_mode == 2) ? &p : nullptr;
  const int niter = p.k_iters;
  if (threadIdx.x == 0) {
    tc_stamp(p.trace, 0);
    if (p.trace) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      const size_t cta = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      p.trace[cta * 16 + 15] = smid;
    }
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 2);    // leader's arrive.expect_tx + the peer's remote arrive
      mbar_init(smem_u32(&empty_bar[s]), 1);   // one multicast commit per phase
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();     // barriers of both CTAs initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) tc_stamp(p.trace, 1);

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====  (incremental loop state, see conv_tc_kernel)
    if (lane == 0) {
      const int npre = niter < p.stages ? niter : p.stages;
      const uint32_t full0 = mapa_u32(smem_u32(&full_bar[0]), 0);     // the LEADER's "full" barriers
      const uint32_t full0_local = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const int brow = cout0 + static_cast<int>(rank) * half_n;
      // weights first (they do not depend on the predecessor kernel), activations after the PDL wait
      for (int i = 0; i < npre; ++i) {
        if (leader) mbar_expect_tx(full0_local + 8u * i, 2u * stage_bytes);
        else mbar_arrive_cluster(full0 + 8u * i);
        tma2_load_2d(tiles_addr + static_cast<uint32_t>(i) * stage_bytes + TC_A_STAGE, &mapB, full0 + 8u * i,
                     i * TC_BLOCK_K, brow);
      }
      pdl_wait();
      tc_stamp(p.trace, 2);
      int tap = 0, chunk = 0, dh = p.ntaps == 9 ? -1 : 0, dw = dh;
      bool seg1 = p.ntaps * p.chunks0 == 0;
      auto load_a = [&](uint32_t a_dst, uint32_t fb) {
        if (!seg1) tma2_load_4d(a_dst, &mapA, fb, chunk * TC_BLOCK_K, w0 + dw, h0 + dh, n0);
        else tma2_load_4d(a_dst, &mapA2, fb, chunk * TC_BLOCK_K, w0, h0, n0);
        if (++chunk == p.chunks0 && !seg1) {
          chunk = 0;
          if (++tap == p.ntaps) {
            seg1 = true;
            dh = 0;
            dw = 0;
          } else if (p.ntaps == 9 && ++dw == 2) {
            dw = -1;
            ++dh;
          }
        }
      };
      for (int i = 0; i < npre; ++i) load_a(tiles_addr + static_cast<uint32_t>(i) * stage_bytes, full0 + 8u * i);
      int s = npre == p.stages ? 0 : npre;
      uint32_t ph = 0;
      for (int i = npre; i < niter; ++i) {
        mbar_wait(empty0 + 8u * s, ph);
        tc_stamp_iter(p.trace, i, 0);
        const uint32_t fb = full0 + 8u * s;
        if (leader) mbar_expect_tx(full0_local + 8u * s, 2u * stage_bytes);
        else mbar_arrive_cluster(fb);
        const uint32_t a_dst = tiles_addr + static_cast<uint32_t>(s) * stage_bytes;
        load_a(a_dst, fb);
        tma2_load_2d(a_dst + TC_A_STAGE, &mapB, fb, i * TC_BLOCK_K, brow);
        tc_stamp_iter(p.trace, i, 1);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: leader CTA only, one thread for the pair =====
    if (leader && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                             (static_cast<uint32_t>(p.block_n >> 3) << 17) |
                             (static_cast<uint32_t>(256 >> 4) << 24);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      int s = 0;
      uint32_t ph = 0;
      const uint64_t adesc0 = make_desc_sw128(tiles_addr);
      uint64_t adesc = adesc0;
      const uint64_t desc_step = static_cast<uint64_t>(stage_bytes >> 4);
      const uint64_t desc_b_off = static_cast<uint64_t>(TC_A_STAGE >> 4);
      for (int i = 0; i < niter; ++i) {
        mbar_wait(full0 + 8u * s, ph);
        tc_fence_after();
        if (i == 0) tc_stamp(p.trace, 3);
        tc_stamp_iter(p.trace, i, 2);
        const uint64_t bdesc = adesc + desc_b_off;
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
          umma2_bf16(tmem_base, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                     (i > 0 || k > 0) ? 1u : 0u);
        umma2_commit_pair(empty0 + 8u * s);
        tc_stamp_iter(p.trace, i, 3);
        adesc += desc_step;
        if (++s == p.stages) {
          s = 0;
          ph ^= 1u;
          adesc = adesc0;
        }
      }
      umma2_commit_pair(smem_u32(&tmem_full_bar));
      tc_stamp(p.trace, 4);
    }
  } else {
    // ===== epilogue warps 2..5 (both CTAs, each its own 128 rows) =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int per_img = p.tw * p.th;
    const int pn = r / per_img;
    const int rem = r - pn * per_img;
    const int ph_ = rem / p.tw;
    const int pw_ = rem - ph_ * p.tw;
    const int n = n0 + pn;
    const bool valid = n < p.N;
    const size_t m = (static_cast<size_t>(n) * p.H + (h0 + ph_)) * p.W + (w0 + pw_);
    mbar_wait(smem_u32(&tmem_full_bar), 0);
    tc_fence_after();
    if (threadIdx.x == 64) tc_stamp(p.trace, 5);
    pdl_wait();
    epilogue_direct_warp(out_dst, p.block_n / 32, tmem_base, tiles_addr + static_cast<uint32_t>(q) * TC_STG_WARP, q,
                         lane, static_cast<uint32_t>(m), valid, cout0, nullptr, valid ? n : p.N - 1);
    if (p.gn_part != nullptr)
      gn_flush_direct(p, out_dst.gn_sm, static_cast<int>(threadIdx.x) - 64, n0, cout0, gn_slot(p, tile_w, tile_h, 0));
  }

  if (threadIdx.x == 64) tc_stamp(p.trace, 6);
  tc_fence_before();
  cluster_sync_all();     // the pair's TMEM is released together; nobody exits while the peer still needs its smem
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
  if (threadIdx.x == 64) tc_stamp(p.trace, 9);
}

// ---- host side --------------------------------------------------------------
static void* g_trace = nullptr;   // profiling only: phase stamps of launches made with debug bit 2
void conv_tc_set_trace(void* ptr) { g_trace = ptr; }

struct TcPlan {
  TcParams p;
  int a_nb;         // images per A-tile TMA box (p.nb, or fewer when only valid rows are loaded)
  int gn_slots;     // contributions per (image, group) this launch writes when gn statistics are fused (0: unsupported)
  int smem_bytes;
  size_t ws_bytes;
  size_t counter_bytes;
  dim3 grid;
};

static int plan_tc(const isb_conv_desc* d, TcPlan* plan) {
  ISB_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv_tc: ksize %d unsupported", d->ksize);
  ISB_CHECK_ARG(d->Cin > 0 && d->Cin % 64 == 0, "conv_tc: Cin=%d must be a multiple of 64", d->Cin);
  ISB_CHECK_ARG(d->a2 == nullptr || (d->Cin2 > 0 && d->Cin2 % 64 == 0),
                "conv_tc: Cin2=%d must be a multiple of 64", d->Cin2);
  ISB_CHECK_ARG(d->Cout > 0 && d->Cout % 32 == 0, "conv_tc: Cout=%d must be a multiple of 32", d->Cout);
  ISB_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "conv_tc: bad N/H/W");
  TcParams& p = plan->p;
  p.N = d->N; p.H = d->H; p.W = d->W;
  // spatial tile of 128 output pixels
  if (d->W >= 128) {
    ISB_CHECK_ARG(d->W % 128 == 0, "conv_tc: W=%d must be a power of two <= 128 or a multiple of 128", d->W);
    p.tw = 128;
  } else {
    ISB_CHECK_ARG((d->W & (d->W - 1)) == 0 && d->W >= 8, "conv_tc: W=%d must be a power of two >= 8", d->W);
    p.tw = d->W;
  }
  p.th = 128 / p.tw;
  if (p.th > d->H) {
    ISB_CHECK_ARG((d->H & (d->H - 1)) == 0, "conv_tc: H=%d must be a power of two when H*W < 128", d->H);
    p.th = d->H;
  }
  ISB_CHECK_ARG(d->H % p.th == 0, "conv_tc: H=%d not divisible by tile height %d", d->H, p.th);
  p.nb = 128 / (p.tw * p.th);
  p.tw_log2 = 31 - __builtin_clz(p.tw);
  p.pi_log2 = 31 - __builtin_clz(p.tw * p.th);
  p.tiles_w = d->W / p.tw;
  p.tiles_h = d->H / p.th;
  p.tiles_n = cdiv(d->N, p.nb);
  p.Cout = d->Cout;
  p.Cin = d->Cin;
  p.chunks0 = d->Cin / 64;
  p.ntaps = d->ksize * d->ksize;
  p.chunks1 = d->a2 ? d->Cin2 / 64 : 0;
  p.k_iters = p.ntaps * p.chunks0 + p.chunks1;
  const int mtiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // N tile / split-K / pipeline depth, from the cold-weight sweeps (profiles/r01_conv_tune_sweep_v4.txt):
  //  * a launch has ~2 us of start latency and 3-6 us of epilogue, so the aim is enough CTAs to overlap those
  //    phases — up to two per SM — not "one wave";
  //  * split-K (cluster of 2/4/8 CTAs per output tile, partials folded through L2).  Layers with few output tiles
  //    split as long as every CTA keeps >= 2 k-iterations; layers that already have >= 48 tiles only while the
  //    CTAs keep >= 8 (the fold costs more than the iterations it saves);
  //  * shared memory per CTA stays <= 96 KB, so that the CTAs of the NEXT launch can become resident while this
  //    launch drains (programmatic dependent launch): their prologue and weight prefetch overlap our epilogue.
  //    Deeper pipelines (6 stages, 144-192 KB) measured up to 2x slower in a chain of launches for that reason.
  const int sms = num_sms();
  // policy knobs (experiments only; the defaults are the tuned values)
  auto knob = [](const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
  };
  const int cta_factor = knob("ISB_TC_CTA_FACTOR", 2), need_few = knob("ISB_TC_NEED_FEW", 2),
            need_many = knob("ISB_TC_NEED_MANY", 8), small_bn = knob("ISB_TC_SMALL_BN", 64),
            max_split = knob("ISB_TC_MAX_SPLIT", 8);
  auto splits_for = [&](int tiles) {
    int sp = 1;
    const int need = tiles >= 48 ? need_many : need_few;
    while (sp < max_split) {
      const int next = sp * 2;
      if (static_cast<long long>(tiles) * next > static_cast<long long>(cta_factor) * sms || p.k_iters / next < need) break;
      sp = next;
    }
    return sp;
  };
  int bn = d->block_n;
  if (bn == 0) {
    if (d->Cout < 128) bn = d->Cout;                                  // 32,64,96
    else if (d->Cout % 128 != 0 && d->Cout <= 256) bn = d->Cout;      // e.g. 192: one exact tile
    else bn = ((d->ksize == 1 || mtiles <= 2) && d->Cout % 64 == 0) ? small_bn : 128;   // 1x1 / tiny images: 64-wide
    if (bn == 64 && d->Cout % 128 == 0 && static_cast<long long>(mtiles) * (d->Cout / 64) > 2LL * sms)
      bn = 128;       // already more 64-wide tiles than two waves (batched launches): wider tiles re-read A less
  }
  ISB_CHECK_ARG(bn >= 32 && bn <= 256 && bn % 32 == 0, "conv_tc: block_n=%d must be a multiple of 32 in [32,256]", bn);
  p.block_n = bn;
  p.tmem_cols = bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256;
  int ntiles = cdiv(d->Cout, bn);
  int tiles = mtiles * ntiles;
  int splits = d->split_k ? d->split_k : splits_for(tiles);
  ISB_CHECK_ARG(splits >= 1 && splits <= p.k_iters, "conv_tc: split_k=%d out of range (k_iters=%d)", splits, p.k_iters);
  p.splits = splits;
  p.cluster = (splits == 2 || splits == 4 || splits == 8 || splits == 16) ? 1 : 0;   // other counts: workspace fold
  // (16 CTAs per cluster is beyond the portable limit: conv_tc_init() opts in, cudaFuncAttributeNonPortableClusterSizeAllowed)
  // CTA pairs for the big, un-split layers: 256-wide N tile, both m-tiles of a pair share the weight tile
  int two = d->two_cta;
  if (two == 0 && d->block_n == 0 && splits == 1 && mtiles % 2 == 0 && mtiles >= 64 && d->Cout % 128 == 0 && !d->w_tiled) {
    two = 1;          // sweep (profiles/r01_conv_tune_pair.txt): pairs with 128-wide tiles, 4 stages, two pairs' CTAs per SM
    bn = 128;
    p.block_n = bn;
    p.tmem_cols = 128;
  }
  if (two != 1) two = 0;
  if (two) {
    ISB_CHECK_ARG(splits == 1 && mtiles % 2 == 0 && (bn == 128 || bn == 256) && d->Cout % bn == 0,
                  "conv_tc: CTA-pair mode needs split_k=1, an even number of pixel tiles (%d) and block_n 128/256 dividing Cout", mtiles);
    ISB_CHECK_ARG(!d->w_tiled, "conv_tc: CTA-pair mode uses row-major weights");
  }
  p.two_cta = two;
  if (p.two_cta) {
    ntiles = cdiv(d->Cout, bn);
    tiles = mtiles * ntiles;
  }
  // A tile bytes per stage: the whole 128-row tile, or only the valid rows when ALL tiles are partial (N*H*W < 128)
  p.a_bytes = TC_A_STAGE;
  int a_nb = p.nb;
  static const bool compact_a = [] {
    const char* e = getenv("ISB_COMPACT_A");
    return e == nullptr || atoi(e) != 0;
  }();
  if (compact_a && !p.two_cta && p.tiles_n == 1 && d->N < p.nb && bn >= 64 && !d->w_tiled) {
    a_nb = d->N;
    p.a_bytes = a_nb * p.tw * p.th * 128;
  }
  plan->a_nb = a_nb;
  const int stage_bytes = p.a_bytes + (p.two_cta ? bn / 2 : bn) * 128;
  const int max_stages = (TC_SMEM_LIMIT - 1024) / stage_bytes;
  int stages = d->stages;
  if (stages == 0) {
    if (p.two_cta) {
      stages = bn <= 128 ? 4 : 5;
    } else if (p.a_bytes < TC_A_STAGE) {
      stages = (96 * 1024) / stage_bytes;            // same 96 KB budget (PDL co-residency), more slots
      if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
      const int per_cta = cdiv(p.k_iters, splits);
      if (stages > per_cta) stages = per_cta < 2 ? 2 : per_cta;
    } else {
      stages = bn <= 64 ? 4 : bn <= 128 ? 3 : 4;    // <= 96 KB (see above); 256-wide tiles cannot, they get 4
      const int per_cta = cdiv(p.k_iters, splits);
      if (stages > per_cta) stages = per_cta < 2 ? 2 : per_cta;
    }
  }
  if (stages > max_stages) stages = max_stages;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  ISB_CHECK_ARG(stages >= 2, "conv_tc: not enough shared memory for 2 stages");
  p.stages = stages;
  plan->smem_bytes = stages * stage_bytes;
  plan->smem_bytes += 1024;
  if (d->min_smem_bytes > plan->smem_bytes) plan->smem_bytes = d->min_smem_bytes < TC_SMEM_LIMIT ? d->min_smem_bytes : TC_SMEM_LIMIT;
  plan->counter_bytes = (static_cast<size_t>(mtiles) * ntiles * sizeof(int) + 255) & ~static_cast<size_t>(255);
  // split-K partial tiles [tile][split][128][bn] fp32 behind the arrival counters (the cluster path does not use
  // the counters); re-written by every launch, so they live in L2
  plan->ws_bytes = splits > 1 ? plan->counter_bytes + static_cast<size_t>(mtiles) * ntiles * splits * TC_BLOCK_M * bn * sizeof(float) : 0;
  plan->grid = dim3(mtiles, ntiles, splits);
  p.bias = d->bias;
  p.residual = d->residual;
  p.out = d->out;
  p.out_dtype = d->out_dtype;
  p.accumulate = d->accumulate;
  p.partial = nullptr;
  p.counters = nullptr;
  p.w_tiled = d->w_tiled;
  p.debug = d->debug_flags;
  p.trace = nullptr;
  // fused GroupNorm statistics: groups of 8/16/32 channels inside power-of-two N tiles, cluster or no split
  p.gn_part = nullptr;
  p.gn_cg_log2 = 0;
  p.gn_groups = 0;
  plan->gn_slots = 0;
  if (d->gn_cg == 8 || d->gn_cg == 16 || d->gn_cg == 32) {
    const bool ok = (bn == 64 || bn == 128 || bn == 256) && d->Cout % bn == 0 && (splits == 1 || p.cluster) &&
                    d->out_dtype == ISB_F32 && p.nb <= 2;
    if (ok) {
      plan->gn_slots = (p.nb == 1 ? p.tiles_w * p.tiles_h : 1) * (p.cluster ? splits : 1);
      p.gn_cg_log2 = d->gn_cg == 8 ? 3 : d->gn_cg == 16 ? 4 : 5;
      p.gn_groups = d->Cout / d->gn_cg;
      p.gn_slots = d->gn_slots;
      p.gn_part = d->gn_partials;
    }
  }
  p.gn_mode = d->gn_mode == 2 ? 2 : 1;
  p.gb_x = d->gb_x; p.gb_gamma = d->gb_gamma; p.gb_beta = d->gb_beta; p.gb_film = d->gb_film; p.gb_stats = d->gb_stats;
  p.gb_film_stride = d->gb_film_stride;
  p.gb_silu = d->gb_silu;
  return ISB_OK;
}

static int encode_act_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int tw, int th,
                          int nb) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_tensormap_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim,
                                      gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation %dx%dx%dx%d, box 64x%dx%dx%d) failed: %d", N, H, W, C, tw, th,
              nb, (int)r);
    return ISB_ERR_CUDA;
  }
  return ISB_OK;
}

static int encode_weight_map(CUtensorMap* m, const void* ptr, int Cout, int Ktot, int bn) {
  cuuint64_t gdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t gstr[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_tensormap_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim,
                                      gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weight %dx%d, box 64x%d) failed: %d", Cout, Ktot, bn, (int)r);
    return ISB_ERR_CUDA;
  }
  return ISB_OK;
}

// Panel-tiled weights: [Cout/64][K/64][64 rows][64 k]; a box {64,64,1,bn/64} lands in shared memory exactly
// like the 2-D box {64,bn} of the row-major layout, but every 64x64 panel is one contiguous 8 KiB read and
// consecutive k-chunks of a 64-row group are adjacent in memory -> the K loop streams HBM sequentially.
static int encode_weight_map_tiled(CUtensorMap* m, const void* ptr, int Cout, int Ktot, int bn) {
  const cuuint64_t kc = (cuuint64_t)Ktot / 64, groups = (cuuint64_t)Cout / 64;
  cuuint64_t gdim[4] = {64, 64, kc, groups};
  cuuint64_t gstr[3] = {128, 8192, kc * 8192};
  cuuint32_t box[4] = {64, 64, 1, (cuuint32_t)(bn / 64)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_tensormap_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim,
                                      gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(tiled weight %dx%d, bn %d) failed: %d", Cout, Ktot, bn, (int)r);
    return ISB_ERR_CUDA;
  }
  return ISB_OK;
}

int conv_tc_init() {
  ISB_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
  ISB_CUDA(cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
  ISB_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  return ISB_OK;
}

int conv_tc_gn_slots(const isb_conv_desc* d) {
  TcPlan plan;
  if (plan_tc(d, &plan) != ISB_OK) return 0;
  return plan.gn_slots;
}

size_t conv_tc_workspace(const isb_conv_desc* d) {
  TcPlan plan;
  if (plan_tc(d, &plan) != ISB_OK) return 0;
  return plan.ws_bytes;
}

int conv_tc_launch(const isb_conv_desc* d, void* ws, size_t ws_bytes, cudaStream_t stream) {
  TcPlan plan;
  int rc = plan_tc(d, &plan);
  if (rc) return rc;
  ISB_CHECK_ARG(d->out_dtype == ISB_F32 || d->out_dtype == ISB_BF16, "conv_tc: bad out dtype");
  ISB_CHECK_ARG(!(d->accumulate && d->out_dtype != ISB_F32), "conv_tc: accumulate needs fp32 out");
  ISB_CHECK_ARG((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
                "conv_tc: pointers must be 16-byte aligned");
  TcParams& p = plan.p;
  if (d->gn_partials != nullptr) {
    ISB_CHECK_ARG(plan.gn_slots > 0, "conv_tc: GroupNorm statistics cannot be fused for this launch (gn_cg=%d, block_n=%d, split_k=%d)", d->gn_cg, p.block_n, p.splits);
    ISB_CHECK_ARG(d->gn_slots == plan.gn_slots, "conv_tc: gn_slots=%d but this launch writes %d per (image, group)", d->gn_slots, plan.gn_slots);
    ISB_CHECK_ARG(d->gn_mode != 2 || (d->gb_x && d->gb_gamma && d->gb_beta && d->gb_stats && (d->gb_film == nullptr || d->gb_film_stride >= 2 * d->Cout)),
                  "conv_tc: gn_mode 2 needs gb_x, gb_gamma, gb_beta, gb_stats (and gb_film_stride >= 2*Cout with gb_film)");
  }
  if (p.splits > 1) {
    if (ws == nullptr || ws_bytes < plan.ws_bytes) {
      set_error("conv_tc: workspace %zu bytes < required %zu", ws_bytes, plan.ws_bytes);
      return ISB_ERR_WORKSPACE;
    }
    p.counters = static_cast<int*>(ws);
    p.partial = reinterpret_cast<float*>(static_cast<char*>(ws) + plan.counter_bytes);
  }
  if (p.debug & 4) p.trace = static_cast<unsigned long long*>(g_trace);   // profiling: isb_debug_set_trace()
  CUtensorMap mapA, mapA2, mapB;
  rc = encode_act_map(&mapA, d->a, d->N, d->H, d->W, d->Cin, p.tw, p.th, plan.a_nb);
  if (rc) return rc;
  if (d->a2) {
    rc = encode_act_map(&mapA2, d->a2, d->N, d->H, d->W, d->Cin2, p.tw, p.th, plan.a_nb);
    if (rc) return rc;
  } else {
    mapA2 = mapA;
  }
  const int Ktot = p.ntaps * d->Cin + (d->a2 ? d->Cin2 : 0);
  if (d->w_tiled) {
    ISB_CHECK_ARG(d->Cout % 64 == 0 && p.block_n % 64 == 0, "conv_tc: panel-tiled weights need Cout %% 64 == 0 and block_n %% 64 == 0 (Cout=%d, block_n=%d)", d->Cout, p.block_n);
    rc = encode_weight_map_tiled(&mapB, d->w, d->Cout, Ktot, p.block_n);
  } else {
    rc = encode_weight_map(&mapB, d->w, d->Cout, Ktot, p.two_cta ? p.block_n / 2 : p.block_n);
  }
  if (rc) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = plan.grid;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled_conv()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p.cluster || p.two_cta) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = p.two_cta ? 2 : 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = p.two_cta ? 1 : p.splits;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (p.two_cta) ISB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc2_kernel, mapA, mapA2, mapB, p));
  else ISB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel, mapA, mapA2, mapB, p));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // namespace isb
