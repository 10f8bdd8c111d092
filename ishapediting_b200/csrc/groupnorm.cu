// groupnorm.cu — GroupNorm32 (+FiLM scale/shift) (+SiLU) (+2x avg-pool / nearest-up) over a
// (possibly two-source, i.e. channel-concatenated) fp32 NHWC tensor, forward and input-gradient
// backward.  HBM-bound: every thread owns 8 consecutive channels of one pixel (32 B fp32 loads,
// 16 B bf16 stores); statistics are reduced in two deterministic stages (per-block partials in
// double, the last-arriving block of a group folds them in fixed order).
//
// Reference: nn.py:16-18 (GroupNorm32, fp32 math), unet.py:183-184,207-208,245-253 (ResBlock
// in/out layers with FiLM), unet.py:285 (attention norm), unet.py:613-614 (out), and
// unet.py:107,136 (nearest 2x / avg-pool inside up/down ResBlocks).
#include <stdlib.h>
#include "common.cuh"

namespace isb {

// The apply kernels are pure streaming passes: bytes in flight per SM = resident threads x bytes per thread.  With 76-80
// registers only 3 CTAs of 256 threads fit (37.5 % occupancy: ncu --set full, profiles/r02_ncu_full_gn.md) and the
// 512-CTA grid of a 128x128x256 layer ran as 1.15 waves; capping the registers at 64 (4 CTAs / SM) makes it one wave.
// (Measured: the cap costs more in spills than the occupancy brings — 4.94 vs 4.88 ms per step — so the default stays 3.
// Sizing gn_apply_part's pixel chunks so that its grid is ONE wave at 3 CTAs per SM (512 pixels per CTA at 128^2) was
// measured too: 4.77 / 4.87 ms with and without alike, inside the box-to-box spread; not kept.)
// Also measured and dropped (round 2, A/B of two builds on one box, interleaved): folding the conv-epilogue partials with
// one warp per group (one L2 round trip instead of four) + the wide-load part kernel for the two-source GroupNorms +
// an unrolled backward reduce: batch-1 step 4.94 vs 4.88 ms (slower), batch-8 step 17.84 vs 17.85 ms (no change) —
// although ncu shows these kernels at 2.5-2.8 TB/s with 33-37 % occupancy when run alone at batch 8
// (gpurun_out -> profiles/r02_ncu_gn_batch8.md), inside the graph they are not what the step waits for.
// What DID pay at batch 8 is the whole-row geometry further down (gn_apply_rows_kernel and its backward siblings, used
// for tensors >= 64 MB only): 4.1 / 5.5 / 3.4 TB/s instead of 2.5-2.75 / 4.2 / 2.5, batch-8 step 17.31 vs 17.90 ms.
#ifndef GN_MIN_BLOCKS
#define GN_MIN_BLOCKS 3
#endif
constexpr int GN_MAX_SPLITS = 64;
constexpr int GN_MAX_NG = 4096;   // arrival counters at the head of the scratch buffer (N * groups <= 4096)

struct GnArgs {
  const float* x1; const float* x2;
  int C1, C2, C, Cg, groups;
  int N, H, W, HW;
  float eps;
  const float* gamma; const float* beta;
  const float* film; int film_stride;
  int silu, resample;
  float* stats;       // [N,G,2] mean,rstd
  // scratch
  int* counters;      // [N*G]
  double2* partials;  // [N*G*GN_MAX_SPLITS]
  float* bstats;      // [N,G,2]  (backward: s1, s2)
  int splits;
};

__device__ __forceinline__ void gn_load_x8(const GnArgs& a, int n, int pix, int c0, float* v) {
  if (c0 < a.C1) load8(a.x1 + (static_cast<size_t>(n) * a.HW + pix) * a.C1 + c0, v);
  else load8(a.x2 + (static_cast<size_t>(n) * a.HW + pix) * a.C2 + (c0 - a.C1), v);
}
// effective per-(n,c) affine: z = xhat * ga + be
__device__ __forceinline__ void gn_affine8(const GnArgs& a, int n, int c0, float* ga, float* be) {
  load8(a.gamma + c0, ga);
  load8(a.beta + c0, be);
  if (a.film != nullptr) {
    float sc[8], sh[8];
    load8(a.film + static_cast<size_t>(n) * a.film_stride + c0, sc);
    load8(a.film + static_cast<size_t>(n) * a.film_stride + a.C + c0, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ga[j] = ga[j] * (1.0f + sc[j]);
      be[j] = be[j] * (1.0f + sc[j]) + sh[j];
    }
  }
}

// Fold per-block partials: the last block to arrive for group (n,g) sums them in index order.
// Returns true in thread 0 of that last block with the totals in (t0,t1).
__device__ __forceinline__ bool gn_block_reduce_and_fold(const GnArgs& a, int ng, int split, double v0, double v1,
                                                         double& t0, double& t1) {
  __shared__ double sh0[8], sh1[8];
  __shared__ int is_last;
  v0 = warp_sum_d(v0);
  v1 = warp_sum_d(v1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh0[warp] = v0; sh1[warp] = v1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0, s1 = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { s0 += sh0[w]; s1 += sh1[w]; }
    a.partials[static_cast<size_t>(ng) * GN_MAX_SPLITS + split] = make_double2(s0, s1);
    __threadfence();
    const int prev = atomicAdd(a.counters + ng, 1);
    is_last = (prev == a.splits - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // one partial per thread (independent loads: one L2 round trip in total, not one per partial), summed in order
  __shared__ double2 folded[GN_MAX_SPLITS];
  const double2* pp = a.partials + static_cast<size_t>(ng) * GN_MAX_SPLITS;
  if (static_cast<int>(threadIdx.x) < a.splits) folded[threadIdx.x] = __ldcg(pp + threadIdx.x);
  __syncthreads();
  if (threadIdx.x != 0) return false;
  double s0 = 0, s1 = 0;
  for (int s = 0; s < a.splits; ++s) { s0 += folded[s].x; s1 += folded[s].y; }
  a.counters[ng] = 0;  // leave the scratch zeroed for the next call
  t0 = s0; t1 = s1;
  return true;
}

// grid (splits, G, N)
__global__ void __launch_bounds__(256)
gn_stats_kernel(const GnArgs a) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const int split = blockIdx.x, g = blockIdx.y, n = blockIdx.z;
  const int V = a.Cg / 8;
  const long long total = static_cast<long long>(a.HW) * V;
  const long long beg = total * split / a.splits, end = total * (split + 1) / a.splits;
  float s = 0.f, ss = 0.f;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const int pix = static_cast<int>(i / V);
    const int c0 = g * a.Cg + static_cast<int>(i % V) * 8;
    float v[8];
    gn_load_x8(a, n, pix, c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += v[j]; ss = fmaf(v[j], v[j], ss); }
  }
  double t0, t1;
  if (gn_block_reduce_and_fold(a, n * a.groups + g, split, static_cast<double>(s), static_cast<double>(ss), t0, t1)) {
    const double m = static_cast<double>(a.HW) * a.Cg;
    const double mean = t0 / m;
    double var = t1 / m - mean * mean;
    if (var < 0) var = 0;
    a.stats[(n * a.groups + g) * 2 + 0] = static_cast<float>(mean);
    a.stats[(n * a.groups + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
}

struct GnFwdOut {
  void* y; int y_dtype;
  void* raw; int raw_dtype;
  float* xres;
};

// z/act for 8 channels of one input pixel
__device__ __forceinline__ void gn_act8(const GnArgs& a, const float* x, float mean, float rstd, const float* ga,
                                        const float* be, float* out) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float z = (x[j] - mean) * rstd * ga[j] + be[j];
    out[j] = a.silu ? silu_f(z) : z;
  }
}

// one (output-grid pixel, 8-channel vector) of the apply pass
__device__ __forceinline__ void gn_apply_one(const GnArgs& a, const GnFwdOut& o, int n, int h, int w, int c0,
                                             float mean, float rstd) {
  const int Ho = a.resample == 1 ? a.H / 2 : a.H, Wo = a.resample == 1 ? a.W / 2 : a.W;
  float ga[8], be[8];
  gn_affine8(a, n, c0, ga, be);
  if (a.resample == 0) {
    const int pix = h * a.W + w;
    float x[8], y[8];
    gn_load_x8(a, n, pix, c0, x);
    gn_act8(a, x, mean, rstd, ga, be, y);
    const size_t off = (static_cast<size_t>(n) * a.HW + pix) * a.C + c0;
    store8(o.y, off, o.y_dtype, y);
    if (o.raw) store8(o.raw, off, o.raw_dtype, x);
  } else if (a.resample == 1) {  // 2x2 average pool of the activation (and of x for the skip)
    float ysum[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int pix = (2 * h + dy) * a.W + (2 * w + dx);
        float x[8], y[8];
        gn_load_x8(a, n, pix, c0, x);
        gn_act8(a, x, mean, rstd, ga, be, y);
        if (o.raw) store8(o.raw, (static_cast<size_t>(n) * a.HW + pix) * a.C + c0, o.raw_dtype, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) { ysum[j] += y[j]; xsum[j] += x[j]; }
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) { ysum[j] *= 0.25f; xsum[j] *= 0.25f; }
    const size_t off = ((static_cast<size_t>(n) * Ho + h) * Wo + w) * a.C + c0;
    store8(o.y, off, o.y_dtype, ysum);
    if (o.xres) store8(o.xres, off, ISB_F32, xsum);
  } else {  // nearest 2x upsample
    const int pix = h * a.W + w;
    float x[8], y[8];
    gn_load_x8(a, n, pix, c0, x);
    gn_act8(a, x, mean, rstd, ga, be, y);
    if (o.raw) store8(o.raw, (static_cast<size_t>(n) * a.HW + pix) * a.C + c0, o.raw_dtype, x);
    const int H2 = a.H * 2, W2 = a.W * 2;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const size_t off = ((static_cast<size_t>(n) * H2 + (2 * h + dy)) * W2 + (2 * w + dx)) * a.C + c0;
        store8(o.y, off, o.y_dtype, y);
        if (o.xres) store8(o.xres, off, ISB_F32, x);
      }
  }
}

__global__ void __launch_bounds__(256, GN_MIN_BLOCKS)
gn_apply_kernel(const GnArgs a, const GnFwdOut o) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const int CV = a.C / 8;
  const int Ho = a.resample == 1 ? a.H / 2 : a.H, Wo = a.resample == 1 ? a.W / 2 : a.W;  // thread grid
  const long long total = static_cast<long long>(a.N) * Ho * Wo * CV;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = static_cast<int>(idx % CV);
  long long t = idx / CV;
  const int w = static_cast<int>(t % Wo); t /= Wo;
  const int h = static_cast<int>(t % Ho);
  const int n = static_cast<int>(t / Ho);
  const int c0 = cv * 8;
  const int g = c0 / a.Cg;
  gn_apply_one(a, o, n, h, w, c0, a.stats[(n * a.groups + g) * 2], a.stats[(n * a.groups + g) * 2 + 1]);
}

// Statistics accumulated by the producer conv's epilogue (isb_conv_desc.gn_partials): ONE launch normalises the
// tensor.  A CTA owns a block of 32 channels (1, 2 or 4 whole groups: Cg = 32, 16, 8) over a chunk of pixels, so it
// needs the statistics of at most four groups: it folds their `slots` partials itself, in fixed order (8 threads per
// group, then the 8 run sums in order), and goes on to apply.  No statistics pass, no finalize launch.
// grid (pixel chunks, C/32, N), 256 threads: thread t owns channel vector t & 3 of pixels (t >> 2) + 64 j.
constexpr int GN_PART_UNROLL = 4;
__global__ void __launch_bounds__(256, GN_MIN_BLOCKS)
gn_apply_part_kernel(const GnArgs a, const GnFwdOut o, const float2* __restrict__ partials, int slots, int ppc) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ double2 runs[32];
  __shared__ float s_mean[4], s_rstd[4];
  const int tid = threadIdx.x;
  const int cb = blockIdx.y, n = blockIdx.z;
  const int gpb = 32 / a.Cg;              // groups in this 32-channel block
  const int g0 = cb * gpb;
  const int v = tid & 3, pr = tid >> 2;
  const int c0 = cb * 32 + v * 8;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(a.HW, p0 + ppc);
  // everything that does not depend on the statistics is requested first: the first batch of pixels and the affine
  // parameters travel while the partials are being folded (one L2 round trip instead of three in a row)
  float x[GN_PART_UNROLL][8];
#pragma unroll
  for (int u = 0; u < GN_PART_UNROLL; ++u)
    if (p0 + pr + u * 64 < p1) load8(a.x1 + (static_cast<size_t>(n) * a.HW + (p0 + pr + u * 64)) * a.C + c0, x[u]);
  float ga[8], be[8];
  gn_affine8(a, n, c0, ga, be);
  if (tid < gpb * 8) {
    const int gl = tid >> 3, part = tid & 7;
    const float2* pp = partials + static_cast<size_t>(n * a.groups + g0 + gl) * slots;
    const int beg = slots * part / 8, end = slots * (part + 1) / 8;
    double s0 = 0, s1 = 0;
    int i = beg;
    for (; i + 4 <= end; i += 4) {
      float2 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = __ldcg(pp + i + j);
#pragma unroll
      for (int j = 0; j < 4; ++j) { s0 += w[j].x; s1 += w[j].y; }
    }
    for (; i < end; ++i) {
      const float2 w = __ldcg(pp + i);
      s0 += w.x;
      s1 += w.y;
    }
    runs[tid] = make_double2(s0, s1);
  }
  __syncthreads();
  if (tid < gpb) {
    double s0 = 0, s1 = 0;
    for (int k = 0; k < 8; ++k) { s0 += runs[tid * 8 + k].x; s1 += runs[tid * 8 + k].y; }
    const double m = static_cast<double>(a.HW) * a.Cg;
    const double mean = s0 / m;
    double var = s1 / m - mean * mean;
    if (var < 0) var = 0;
    s_mean[tid] = static_cast<float>(mean);
    s_rstd[tid] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
    if (blockIdx.x == 0) {     // kept for the backward pass
      a.stats[(n * a.groups + g0 + tid) * 2 + 0] = s_mean[tid];
      a.stats[(n * a.groups + g0 + tid) * 2 + 1] = s_rstd[tid];
    }
  }
  __syncthreads();
  const int gl = (v * 8) / a.Cg;
  const float mean = s_mean[gl], rstd = s_rstd[gl];
#pragma unroll
  for (int j = 0; j < 8; ++j) {   // z = x * ga' + be'
    ga[j] *= rstd;
    be[j] -= mean * ga[j];
  }
  for (int p = p0 + pr; p < p1; p += GN_PART_UNROLL * 64) {
    if (p != p0 + pr) {
#pragma unroll
      for (int u = 0; u < GN_PART_UNROLL; ++u)
        if (p + u * 64 < p1) load8(a.x1 + (static_cast<size_t>(n) * a.HW + (p + u * 64)) * a.C + c0, x[u]);
    }
#pragma unroll
    for (int u = 0; u < GN_PART_UNROLL; ++u)
      if (p + u * 64 < p1) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(x[u][j], ga[j], be[j]);
          y[j] = a.silu ? silu_f(z) : z;
        }
        const size_t off = (static_cast<size_t>(n) * a.HW + (p + u * 64)) * a.C + c0;
        store8(o.y, off, o.y_dtype, y);
        if (o.raw) store8(o.raw, off, o.raw_dtype, x[u]);
      }
  }
}

// ---- large tensors (beyond L2: the throughput mode, batch >= 4 at 64^2 / 128^2) ---------------------------------
// There the apply pass is an HBM stream and what counts is bytes in flight and DRAM locality.  ncu at batch 8
// (profiles/r02_ncu_gn_batch8.md): the 32-channel-column geometry above reaches 2.5 TB/s (a warp touches eight 128 B
// pieces 1 KB apart, and every CTA first waits for its partial fold), the one-vector-per-thread kernel 2.75 TB/s (32 B
// in flight per thread), while the kernels that read whole rows with several loads in flight reach 4.2 TB/s.  So:
// statistics first (one warp per (image, group) folds the conv-epilogue partials), then whole pixel rows — thread =
// (channel vector, pixel row), consecutive threads = consecutive 32 B of one pixel — with GN_ROWS_UNROLL pixels in flight.
__global__ void __launch_bounds__(256)
gn_fold_stats_kernel(const GnArgs a, const float2* __restrict__ partials, int slots) {
  pdl_wait();
  pdl_trigger();
  const int ng = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (ng >= a.N * a.groups) return;
  double s0 = 0, s1 = 0;
  for (int i = threadIdx.x & 31; i < slots; i += 32) {
    const float2 w = __ldcg(partials + static_cast<size_t>(ng) * slots + i);
    s0 += w.x;
    s1 += w.y;
  }
  s0 = warp_sum_d(s0);
  s1 = warp_sum_d(s1);
  if ((threadIdx.x & 31) == 0) {
    const double m = static_cast<double>(a.HW) * a.Cg;
    const double mean = s0 / m;
    double var = s1 / m - mean * mean;
    if (var < 0) var = 0;
    a.stats[ng * 2 + 0] = static_cast<float>(mean);
    a.stats[ng * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  }
}

constexpr int GN_ROWS_UNROLL = 4;
// grid (pixel chunks, N), block = (C/8) * rows threads
__global__ void __launch_bounds__(512)
gn_apply_rows_kernel(const GnArgs a, const GnFwdOut o, int rows, int ppc) {
  pdl_wait();
  pdl_trigger();
  const int CV = a.C / 8;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int c0 = cv * 8;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(a.HW, p0 + ppc);
  float x[GN_ROWS_UNROLL][8];
#pragma unroll
  for (int u = 0; u < GN_ROWS_UNROLL; ++u)
    if (p0 + r + u * rows < p1) gn_load_x8(a, n, p0 + r + u * rows, c0, x[u]);
  float ga[8], be[8];
  gn_affine8(a, n, c0, ga, be);
  const int sg = n * a.groups + c0 / a.Cg;
  const float mean = a.stats[sg * 2], rstd = a.stats[sg * 2 + 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) {   // z = x * ga' + be'
    ga[j] *= rstd;
    be[j] -= mean * ga[j];
  }
  for (int p = p0 + r; p < p1; p += GN_ROWS_UNROLL * rows) {
    if (p != p0 + r) {
#pragma unroll
      for (int u = 0; u < GN_ROWS_UNROLL; ++u)
        if (p + u * rows < p1) gn_load_x8(a, n, p + u * rows, c0, x[u]);
    }
#pragma unroll
    for (int u = 0; u < GN_ROWS_UNROLL; ++u)
      if (p + u * rows < p1) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(x[u][j], ga[j], be[j]);
          y[j] = a.silu ? silu_f(z) : z;
        }
        const size_t off = (static_cast<size_t>(n) * a.HW + (p + u * rows)) * a.C + c0;
        store8(o.y, off, o.y_dtype, y);
        if (o.raw) store8(o.raw, off, o.raw_dtype, x[u]);
      }
  }
}

// ---- thread-block-cluster helpers (the fused kernels below run CS CTAs per (image, group)) -------------
__device__ __forceinline__ uint32_t gn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t gn_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void gn_cluster_sync() {
  __syncwarp();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double gn_ld_dsmem_f64(const double* local, uint32_t rank) {
  const uint32_t laddr = static_cast<uint32_t>(__cvta_generic_to_shared(local));
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(raddr) : "memory");
  return v;
}
// block partials (thread 0 holds them) -> cluster totals in rank order (deterministic), returned to thread 0
__device__ __forceinline__ void gn_cluster_fold(double* s_part, double& t0, double& t1) {
  if (threadIdx.x == 0) { s_part[0] = t0; s_part[1] = t1; }
  gn_cluster_sync();
  if (threadIdx.x == 0) {
    const uint32_t cs = gn_cluster_size();
    double a0 = 0, a1 = 0;
    for (uint32_t r = 0; r < cs; ++r) { a0 += gn_ld_dsmem_f64(s_part, r); a1 += gn_ld_dsmem_f64(s_part + 1, r); }
    t0 = a0; t1 = a1;
  }
}

// Small tensors: ONE launch, one CTA per (image, group): statistics pass, block reduction, apply pass
// (the second read of the group's slice hits L1/L2).  Saves a launch and the global round trip.
__global__ void __launch_bounds__(512)
gn_fused_fwd_kernel(const GnArgs a, const GnFwdOut o) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ double sh0[16], sh1[16];
  __shared__ double s_part[2];
  __shared__ float s_mean, s_rstd;
  const int g = blockIdx.x, n = blockIdx.y;
  const int rank = static_cast<int>(gn_cluster_rank()), cs = static_cast<int>(gn_cluster_size());
  const int V = a.Cg / 8;
  const int total = a.HW * V;
  const int beg = static_cast<int>(static_cast<long long>(total) * rank / cs);
  const int end = static_cast<int>(static_cast<long long>(total) * (rank + 1) / cs);
  float s = 0.f, ss = 0.f;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    float v[8];
    gn_load_x8(a, n, i / V, g * a.Cg + (i % V) * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += v[j]; ss = fmaf(v[j], v[j], ss); }
  }
  double d0 = warp_sum_d(static_cast<double>(s)), d1 = warp_sum_d(static_cast<double>(ss));
  if ((threadIdx.x & 31) == 0) { sh0[threadIdx.x >> 5] = d0; sh1[threadIdx.x >> 5] = d1; }
  __syncthreads();
  double t0 = 0, t1 = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) { t0 += sh0[w]; t1 += sh1[w]; }
  gn_cluster_fold(s_part, t0, t1);
  if (threadIdx.x == 0) {
    const double m = static_cast<double>(a.HW) * a.Cg;
    const double mean = t0 / m;
    double var = t1 / m - mean * mean;
    if (var < 0) var = 0;
    s_mean = static_cast<float>(mean);
    s_rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
    if (rank == 0) {
      a.stats[(n * a.groups + g) * 2 + 0] = s_mean;
      a.stats[(n * a.groups + g) * 2 + 1] = s_rstd;
    }
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  const int Ho = a.resample == 1 ? a.H / 2 : a.H, Wo = a.resample == 1 ? a.W / 2 : a.W;
  const int total_o = Ho * Wo * V;
  const int obeg = static_cast<int>(static_cast<long long>(total_o) * rank / cs);
  const int oend = static_cast<int>(static_cast<long long>(total_o) * (rank + 1) / cs);
  for (int i = obeg + threadIdx.x; i < oend; i += blockDim.x) {
    const int pix = i / V;
    gn_apply_one(a, o, n, pix / Wo, pix % Wo, g * a.Cg + (i % V) * 8, mean, rstd);
  }
  gn_cluster_sync();   // peers may still be reading this CTA's partials
}

// ---- backward ---------------------------------------------------------------
struct GnBwdArgs {
  const float* dy;
  const float* gres; int gres_at_input;
  float* gx1; int acc1; void* gx1_lo;
  float* gx2; int acc2; void* gx2_lo;
  int lo_dtype;
};

// gradient arriving at input pixel (h,w) from a tensor living at OUTPUT resolution, pulled back
// through the resample (transpose of avg-pool / nearest-up).
__device__ __forceinline__ void gn_pull8(const GnArgs& a, const float* src, int n, int h, int w, int c0, float* v) {
  if (a.resample == 0) {
    load8(src + ((static_cast<size_t>(n) * a.H + h) * a.W + w) * a.C + c0, v);
  } else if (a.resample == 1) {
    const int Ho = a.H / 2, Wo = a.W / 2;
    load8(src + ((static_cast<size_t>(n) * Ho + (h >> 1)) * Wo + (w >> 1)) * a.C + c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
  } else {
    const int H2 = a.H * 2, W2 = a.W * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float t[8];
        load8(src + ((static_cast<size_t>(n) * H2 + (2 * h + dy)) * W2 + (2 * w + dx)) * a.C + c0, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
      }
  }
}

// dz * gamma'' and xhat for 8 channels
__device__ __forceinline__ void gn_bwd_terms8(const GnArgs& a, const GnBwdArgs& b, int n, int h, int w, int c0,
                                              float mean, float rstd, float* dzg, float* xhat) {
  float x[8], ga[8], be[8], dy[8];
  gn_load_x8(a, n, h * a.W + w, c0, x);
  gn_affine8(a, n, c0, ga, be);
  gn_pull8(a, b.dy, n, h, w, c0, dy);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    xhat[j] = (x[j] - mean) * rstd;
    float d = dy[j];
    if (a.silu) d *= silu_grad_f(xhat[j] * ga[j] + be[j]);
    dzg[j] = d * ga[j];
  }
}

__global__ void __launch_bounds__(256)
gn_bwd_reduce_kernel(const GnArgs a, const GnBwdArgs b) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const int split = blockIdx.x, g = blockIdx.y, n = blockIdx.z;
  const int V = a.Cg / 8;
  const long long total = static_cast<long long>(a.HW) * V;
  const long long beg = total * split / a.splits, end = total * (split + 1) / a.splits;
  const float mean = a.stats[(n * a.groups + g) * 2], rstd = a.stats[(n * a.groups + g) * 2 + 1];
  float s1 = 0.f, s2 = 0.f;
  for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const int pix = static_cast<int>(i / V);
    const int c0 = g * a.Cg + static_cast<int>(i % V) * 8;
    float dzg[8], xhat[8];
    gn_bwd_terms8(a, b, n, pix / a.W, pix % a.W, c0, mean, rstd, dzg, xhat);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += dzg[j]; s2 = fmaf(dzg[j], xhat[j], s2); }
  }
  double t0, t1;
  if (gn_block_reduce_and_fold(a, n * a.groups + g, split, static_cast<double>(s1), static_cast<double>(s2), t0, t1)) {
    const double m = static_cast<double>(a.HW) * a.Cg;
    a.bstats[(n * a.groups + g) * 2 + 0] = static_cast<float>(t0 / m);
    a.bstats[(n * a.groups + g) * 2 + 1] = static_cast<float>(t1 / m);
  }
}

__device__ __forceinline__ void gn_bwd_apply_one(const GnArgs& a, const GnBwdArgs& b, int n, int h, int w, int c0,
                                                 float mean, float rstd, float m1, float m2) {
  float dzg[8], xhat[8], dx[8];
  gn_bwd_terms8(a, b, n, h, w, c0, mean, rstd, dzg, xhat);
#pragma unroll
  for (int j = 0; j < 8; ++j) dx[j] = rstd * (dzg[j] - (m1 + xhat[j] * m2));
  if (b.gres != nullptr) {
    float r[8];
    if (b.gres_at_input) load8(b.gres + ((static_cast<size_t>(n) * a.H + h) * a.W + w) * a.C + c0, r);
    else gn_pull8(a, b.gres, n, h, w, c0, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) dx[j] += r[j];
  }
  const size_t pix = (static_cast<size_t>(n) * a.H + h) * a.W + w;
  float* gx; void* lo; int acc; size_t off;
  if (c0 < a.C1) { gx = b.gx1; lo = b.gx1_lo; acc = b.acc1; off = pix * a.C1 + c0; }
  else { gx = b.gx2; lo = b.gx2_lo; acc = b.acc2; off = pix * a.C2 + (c0 - a.C1); }
  if (gx == nullptr && lo == nullptr) return;
  if (acc && gx != nullptr) {
    const float4* p = reinterpret_cast<const float4*>(gx + off);
    const float4 u = p[0], v = p[1];
    dx[0] += u.x; dx[1] += u.y; dx[2] += u.z; dx[3] += u.w;
    dx[4] += v.x; dx[5] += v.y; dx[6] += v.z; dx[7] += v.w;
  }
  if (gx != nullptr) store8(gx, off, ISB_F32, dx);
  if (lo != nullptr) store8(lo, off, b.lo_dtype, dx);
}

__global__ void __launch_bounds__(256, GN_MIN_BLOCKS)
gn_bwd_apply_kernel(const GnArgs a, const GnBwdArgs b) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const int CV = a.C / 8;
  const long long total = static_cast<long long>(a.N) * a.HW * CV;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = static_cast<int>(idx % CV);
  long long t = idx / CV;
  const int w = static_cast<int>(t % a.W); t /= a.W;
  const int h = static_cast<int>(t % a.H);
  const int n = static_cast<int>(t / a.H);
  const int c0 = cv * 8;
  const int sg = n * a.groups + c0 / a.Cg;
  gn_bwd_apply_one(a, b, n, h, w, c0, a.stats[sg * 2], a.stats[sg * 2 + 1], a.bstats[sg * 2], a.bstats[sg * 2 + 1]);
}

// per-thread constants of the backward kernels: z = xhat * ga + be
struct GnRowBwdConst {
  float ga[8], be[8], mean, rstd;
};
__device__ __forceinline__ void gn_row_bwd_terms(const GnArgs& a, const GnRowBwdConst& k, const float* x, const float* dy,
                                                 float* dzg, float* xhat) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    xhat[j] = (x[j] - k.mean) * k.rstd;
    float d = dy[j];
    if (a.silu) d *= silu_grad_f(xhat[j] * k.ga[j] + k.be[j]);
    dzg[j] = d * k.ga[j];
  }
}

// Backward with the reduction terms already accumulated by the dgrad conv that produced dy (isb_conv_desc.gn_mode 2):
// same geometry as gn_apply_part_kernel — a CTA owns 32 channels (<= 4 groups), folds their partials itself and
// applies.  One launch instead of reduce + apply.
constexpr int GN_PART_UNROLL_BWD = 2;
__global__ void __launch_bounds__(256)
gn_bwd_apply_part_kernel(const GnArgs a, const GnBwdArgs b, const float2* __restrict__ partials, int slots, int ppc) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ double2 runs[32];
  __shared__ float s_m1[4], s_m2[4];
  const int tid = threadIdx.x;
  const int cb = blockIdx.y, n = blockIdx.z;
  const int gpb = 32 / a.Cg;
  const int g0 = cb * gpb;
  const int v = tid & 3, pr = tid >> 2;
  const int c0 = cb * 32 + v * 8;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(a.HW, p0 + ppc);
  const bool acc_gx = b.acc1 && b.gx1 != nullptr;
  float x[GN_PART_UNROLL_BWD][8], dy[GN_PART_UNROLL_BWD][8], rs[GN_PART_UNROLL_BWD][8], old[GN_PART_UNROLL_BWD][8];
  auto fetch = [&](int p) {
#pragma unroll
    for (int u = 0; u < GN_PART_UNROLL_BWD; ++u)
      if (p + u * 64 < p1) {
        const size_t off = (static_cast<size_t>(n) * a.HW + (p + u * 64)) * a.C + c0;
        load8(a.x1 + off, x[u]);
        load8(b.dy + off, dy[u]);
        if (b.gres != nullptr) load8(b.gres + off, rs[u]);
        if (acc_gx) {
          const float4* q = reinterpret_cast<const float4*>(b.gx1 + off);
          const float4 u0 = q[0], u1 = q[1];
          old[u][0] = u0.x; old[u][1] = u0.y; old[u][2] = u0.z; old[u][3] = u0.w;
          old[u][4] = u1.x; old[u][5] = u1.y; old[u][6] = u1.z; old[u][7] = u1.w;
        }
      }
  };
  fetch(p0 + pr);     // in flight while the partials are folded
  GnRowBwdConst k;
  gn_affine8(a, n, c0, k.ga, k.be);
  const int sg = n * a.groups + c0 / a.Cg;
  k.mean = a.stats[sg * 2];
  k.rstd = a.stats[sg * 2 + 1];
  if (tid < gpb * 8) {
    const int gl = tid >> 3, part = tid & 7;
    const float2* pp = partials + static_cast<size_t>(n * a.groups + g0 + gl) * slots;
    const int beg = slots * part / 8, end = slots * (part + 1) / 8;
    double s0 = 0, s1 = 0;
    int i = beg;
    for (; i + 4 <= end; i += 4) {
      float2 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = __ldcg(pp + i + j);
#pragma unroll
      for (int j = 0; j < 4; ++j) { s0 += w[j].x; s1 += w[j].y; }
    }
    for (; i < end; ++i) {
      const float2 w = __ldcg(pp + i);
      s0 += w.x;
      s1 += w.y;
    }
    runs[tid] = make_double2(s0, s1);
  }
  __syncthreads();
  if (tid < gpb) {
    double s0 = 0, s1 = 0;
    for (int q = 0; q < 8; ++q) { s0 += runs[tid * 8 + q].x; s1 += runs[tid * 8 + q].y; }
    const double m = static_cast<double>(a.HW) * a.Cg;
    s_m1[tid] = static_cast<float>(s0 / m);
    s_m2[tid] = static_cast<float>(s1 / m);
  }
  __syncthreads();
  const int gl = (v * 8) / a.Cg;
  const float m1 = s_m1[gl], m2 = s_m2[gl];
  for (int p = p0 + pr; p < p1; p += GN_PART_UNROLL_BWD * 64) {
    if (p != p0 + pr) fetch(p);
#pragma unroll
    for (int u = 0; u < GN_PART_UNROLL_BWD; ++u)
      if (p + u * 64 < p1) {
        float dzg[8], xhat[8], dx[8];
        gn_row_bwd_terms(a, k, x[u], dy[u], dzg, xhat);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dx[j] = k.rstd * (dzg[j] - (m1 + xhat[j] * m2));
          if (b.gres != nullptr) dx[j] += rs[u][j];
          if (acc_gx) dx[j] += old[u][j];
        }
        const size_t off = (static_cast<size_t>(n) * a.HW + (p + u * 64)) * a.C + c0;
        if (b.gx1 != nullptr) store8(b.gx1, off, ISB_F32, dx);
        if (b.gx1_lo != nullptr) store8(b.gx1_lo, off, b.lo_dtype, dx);
      }
  }
}

// Backward reduction for large tensors (see gn_apply_rows_kernel): whole pixel rows with GN_ROWS_UNROLL pixels in
// flight; a CTA covers ALL groups of a pixel chunk and writes one partial per group, the last CTA to arrive for a
// group folds that group's partials in chunk order (deterministic).  grid (chunks <= GN_MAX_SPLITS, N).
__global__ void __launch_bounds__(512)
gn_bwd_reduce_rows_kernel(const GnArgs a, const GnBwdArgs b, int rows, int ppc, int chunks) {
  pdl_wait();
  pdl_trigger();
  __shared__ float2 vals[512];
  __shared__ int s_last[64];
  const int CV = a.C / 8;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int c0 = cv * 8;
  const int p0 = chunk * ppc;
  const int p1 = min(a.HW, p0 + ppc);
  GnRowBwdConst k;
  gn_affine8(a, n, c0, k.ga, k.be);
  {
    const int sg = n * a.groups + c0 / a.Cg;
    k.mean = a.stats[sg * 2];
    k.rstd = a.stats[sg * 2 + 1];
  }
  float s1 = 0.f, s2 = 0.f;
  for (int p = p0 + r; p < p1; p += GN_ROWS_UNROLL * rows) {
    float x[GN_ROWS_UNROLL][8], dy[GN_ROWS_UNROLL][8];
#pragma unroll
    for (int u = 0; u < GN_ROWS_UNROLL; ++u)
      if (p + u * rows < p1) {
        gn_load_x8(a, n, p + u * rows, c0, x[u]);
        load8(b.dy + (static_cast<size_t>(n) * a.HW + (p + u * rows)) * a.C + c0, dy[u]);
      }
#pragma unroll
    for (int u = 0; u < GN_ROWS_UNROLL; ++u)
      if (p + u * rows < p1) {
        float dzg[8], xhat[8];
        gn_row_bwd_terms(a, k, x[u], dy[u], dzg, xhat);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1 += dzg[j]; s2 = fmaf(dzg[j], xhat[j], s2); }
      }
  }
  vals[threadIdx.x] = make_float2(s1, s2);
  __syncthreads();
  const int V = a.Cg / 8;
  if (static_cast<int>(threadIdx.x) < a.groups) {
    const int g = threadIdx.x, ng = n * a.groups + g;
    double d0 = 0, d1 = 0;
    for (int rr = 0; rr < rows; ++rr)
      for (int v = 0; v < V; ++v) {
        const float2 t = vals[rr * CV + g * V + v];
        d0 += t.x;
        d1 += t.y;
      }
    a.partials[static_cast<size_t>(ng) * GN_MAX_SPLITS + chunk] = make_double2(d0, d1);
    __threadfence();
    const int prev = atomicAdd(a.counters + ng, 1);
    s_last[g] = (prev == chunks - 1);
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < a.groups && s_last[threadIdx.x]) {
    __threadfence();
    const int ng = n * a.groups + threadIdx.x;
    const double2* pp = a.partials + static_cast<size_t>(ng) * GN_MAX_SPLITS;
    double t0 = 0, t1 = 0;
    for (int c = 0; c < chunks; ++c) {
      const double2 w = __ldcg(pp + c);
      t0 += w.x;
      t1 += w.y;
    }
    a.counters[ng] = 0;  // leave the scratch zeroed for the next call
    const double m = static_cast<double>(a.HW) * a.Cg;
    a.bstats[ng * 2 + 0] = static_cast<float>(t0 / m);
    a.bstats[ng * 2 + 1] = static_cast<float>(t1 / m);
  }
}

// Backward apply for large tensors, same row geometry; two pixels (4-8 independent 32 B loads) in flight per thread.
constexpr int GN_ROWS_UNROLL_BWD = 2;
__global__ void __launch_bounds__(512)
gn_bwd_apply_rows_kernel(const GnArgs a, const GnBwdArgs b, int rows, int ppc) {
  pdl_wait();
  pdl_trigger();
  const int CV = a.C / 8;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int c0 = cv * 8;
  const int p0 = blockIdx.x * ppc;
  const int p1 = min(a.HW, p0 + ppc);
  GnRowBwdConst k;
  gn_affine8(a, n, c0, k.ga, k.be);
  const int sg = n * a.groups + c0 / a.Cg;
  k.mean = a.stats[sg * 2];
  k.rstd = a.stats[sg * 2 + 1];
  const float m1 = a.bstats[sg * 2], m2 = a.bstats[sg * 2 + 1];
  // destination of this thread's channels: first or second source of the concatenation
  const bool first = c0 < a.C1;
  float* gx = first ? b.gx1 : b.gx2;
  void* lo = first ? b.gx1_lo : b.gx2_lo;
  const bool acc = (first ? b.acc1 : b.acc2) && gx != nullptr;
  const int Cd = first ? a.C1 : a.C2, cd = first ? c0 : c0 - a.C1;
  if (gx == nullptr && lo == nullptr) return;
  for (int p = p0 + r; p < p1; p += GN_ROWS_UNROLL_BWD * rows) {
    float x[GN_ROWS_UNROLL_BWD][8], dy[GN_ROWS_UNROLL_BWD][8], rs[GN_ROWS_UNROLL_BWD][8], old[GN_ROWS_UNROLL_BWD][8];
#pragma unroll
    for (int u = 0; u < GN_ROWS_UNROLL_BWD; ++u)
      if (p + u * rows < p1) {
        const size_t pix = static_cast<size_t>(n) * a.HW + (p + u * rows);
        gn_load_x8(a, n, p + u * rows, c0, x[u]);
        load8(b.dy + pix * a.C + c0, dy[u]);
        if (b.gres != nullptr) load8(b.gres + pix * a.C + c0, rs[u]);
        if (acc) {
          const float4* q = reinterpret_cast<const float4*>(gx + pix * Cd + cd);
          const float4 u0 = q[0], u1 = q[1];
          old[u][0] = u0.x; old[u][1] = u0.y; old[u][2] = u0.z; old[u][3] = u0.w;
          old[u][4] = u1.x; old[u][5] = u1.y; old[u][6] = u1.z; old[u][7] = u1.w;
        }
      }
#pragma unroll
    for (int u = 0; u < GN_ROWS_UNROLL_BWD; ++u)
      if (p + u * rows < p1) {
        float dzg[8], xhat[8], dx[8];
        gn_row_bwd_terms(a, k, x[u], dy[u], dzg, xhat);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dx[j] = k.rstd * (dzg[j] - (m1 + xhat[j] * m2));
          if (b.gres != nullptr) dx[j] += rs[u][j];
          if (acc) dx[j] += old[u][j];
        }
        const size_t off = (static_cast<size_t>(n) * a.HW + (p + u * rows)) * Cd + cd;
        if (gx != nullptr) store8(gx, off, ISB_F32, dx);
        if (lo != nullptr) store8(lo, off, b.lo_dtype, dx);
      }
  }
}

// single-launch backward for small tensors (see gn_fused_fwd_kernel)
__global__ void __launch_bounds__(512)
gn_fused_bwd_kernel(const GnArgs a, const GnBwdArgs b) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ double sh0[16], sh1[16];
  __shared__ double s_part[2];
  __shared__ float s_m1, s_m2;
  const int g = blockIdx.x, n = blockIdx.y;
  const int rank = static_cast<int>(gn_cluster_rank()), cs = static_cast<int>(gn_cluster_size());
  const int V = a.Cg / 8;
  const int total = a.HW * V;
  const int beg = static_cast<int>(static_cast<long long>(total) * rank / cs);
  const int end = static_cast<int>(static_cast<long long>(total) * (rank + 1) / cs);
  const float mean = a.stats[(n * a.groups + g) * 2], rstd = a.stats[(n * a.groups + g) * 2 + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const int pix = i / V;
    float dzg[8], xhat[8];
    gn_bwd_terms8(a, b, n, pix / a.W, pix % a.W, g * a.Cg + (i % V) * 8, mean, rstd, dzg, xhat);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1 += dzg[j]; s2 = fmaf(dzg[j], xhat[j], s2); }
  }
  double d0 = warp_sum_d(static_cast<double>(s1)), d1 = warp_sum_d(static_cast<double>(s2));
  if ((threadIdx.x & 31) == 0) { sh0[threadIdx.x >> 5] = d0; sh1[threadIdx.x >> 5] = d1; }
  __syncthreads();
  double t0 = 0, t1 = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x >> 5); ++w) { t0 += sh0[w]; t1 += sh1[w]; }
  gn_cluster_fold(s_part, t0, t1);
  if (threadIdx.x == 0) {
    const double m = static_cast<double>(a.HW) * a.Cg;
    s_m1 = static_cast<float>(t0 / m);
    s_m2 = static_cast<float>(t1 / m);
  }
  __syncthreads();
  const float m1 = s_m1, m2 = s_m2;
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const int pix = i / V;
    gn_bwd_apply_one(a, b, n, pix / a.W, pix % a.W, g * a.Cg + (i % V) * 8, mean, rstd, m1, m2);
  }
  gn_cluster_sync();   // peers may still be reading this CTA's partials
}

// one CTA per (image, group) only pays off for tiny slices (8x8, 16x16): measured, larger slices are
// faster with the two wide kernels (32 CTAs cannot pull enough bandwidth)
static long long gn_fused_max() {
  static const long long v = [] {
    const char* e = getenv("ISB_GN_FUSED_MAX");
    return e ? atoll(e) : 8192LL;
  }();
  return v;
}
// tensors at least this large (fp32 bytes) take the whole-row apply kernel (ISB_GN_ROWS_MIN_MB, default 64 MB: beyond
// what stays in the 126 MB L2 together with the conv's operands; batch 1 never gets there)
static bool gn_use_rows(const GnArgs& a) {
  const char* e = getenv("ISB_GN_ROWS_MIN_MB");
  const double min_mb = e ? atof(e) : 64.0;
  if (min_mb < 0 || a.resample != 0) return false;
  const int CV = a.C / 8;
  if (CV % 32 != 0 || CV > 512) return false;
  return static_cast<double>(a.N) * a.HW * a.C * 4.0 >= min_mb * 1048576.0;
}
static cudaError_t gn_launch_rows(const GnArgs& a, const GnFwdOut& o, cudaStream_t st) {
  const int CV = a.C / 8;
  const int rows = CV >= 256 ? 1 : 256 / CV;
  const int ppc = rows * GN_ROWS_UNROLL * 4;          // four iterations per CTA
  return launch(gn_apply_rows_kernel, dim3(cdiv(a.HW, ppc), a.N), dim3(CV * rows), 0, st, a, o, rows, ppc);
}
static bool gn_use_fused(const GnArgs& a) { return static_cast<long long>(a.HW) * a.Cg <= gn_fused_max(); }
// CTAs per (image, group) for the fused kernels: enough CTAs to pull bandwidth, DSMEM fold of the partial sums
static int gn_cluster_size_for(const GnArgs& a) {
  const long long e = static_cast<long long>(a.HW) * a.Cg;
  return e <= 8192 ? 1 : e <= 32768 ? 2 : e <= 65536 ? 4 : 8;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster_z(void (*kernel)(KArgs...), dim3 grid, dim3 block, int cs, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cs > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = cs;
    ++na;
  }
  if (pdl_enabled() || pdl_enabled_family(t_pdl_family)) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int gn_fill_args(const isb_gn_desc* d, void* scratch, GnArgs* a) {
  ISB_CHECK_ARG(d->x1 != nullptr && d->C1 > 0, "groupnorm: x1 missing");
  ISB_CHECK_ARG(scratch != nullptr, "groupnorm: scratch missing");
  a->x1 = d->x1; a->x2 = d->x2;
  a->C1 = d->C1; a->C2 = d->x2 ? d->C2 : 0;
  a->C = a->C1 + a->C2;
  a->groups = d->groups;
  ISB_CHECK_ARG(d->groups > 0 && a->C % d->groups == 0, "groupnorm: C=%d not divisible by groups=%d", a->C, d->groups);
  a->Cg = a->C / d->groups;
  ISB_CHECK_ARG(a->Cg % 8 == 0 && a->C1 % 8 == 0, "groupnorm: channels per group (%d) and C1 (%d) must be multiples of 8", a->Cg, a->C1);
  a->N = d->N; a->H = d->H; a->W = d->W; a->HW = d->H * d->W;
  ISB_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "groupnorm: bad N/H/W");
  ISB_CHECK_ARG(d->resample >= 0 && d->resample <= 2, "groupnorm: bad resample");
  ISB_CHECK_ARG(d->resample != 1 || (d->H % 2 == 0 && d->W % 2 == 0), "groupnorm: avg-pool needs even H,W");
  a->eps = d->eps;
  a->gamma = d->gamma; a->beta = d->beta;
  ISB_CHECK_ARG(d->gamma && d->beta && d->stats, "groupnorm: gamma/beta/stats missing");
  a->film = d->film; a->film_stride = d->film_stride;
  ISB_CHECK_ARG(d->film == nullptr || d->film_stride >= 2 * a->C, "groupnorm: film_stride too small");
  a->silu = d->silu; a->resample = d->resample;
  a->stats = d->stats;
  const size_t ng = static_cast<size_t>(d->N) * d->groups;
  // The arrival counters live in a FIXED-size region at the head of the scratch buffer, whatever N is: callers
  // reuse one buffer for plans of different batch sizes, and the counters must be 0 between launches.  (With an
  // N-dependent layout a batch-1 launch wrote its partial sums where a batch-8 launch keeps the counters of images
  // 1..3 — the statistics of exactly those images then came out wrong.)
  ISB_CHECK_ARG(ng <= GN_MAX_NG, "groupnorm: N*groups=%zu exceeds the %d arrival counters of the scratch layout", ng, GN_MAX_NG);
  char* s = static_cast<char*>(scratch);
  a->counters = reinterpret_cast<int*>(s);
  size_t off = GN_MAX_NG * sizeof(int);
  a->partials = reinterpret_cast<double2*>(s + off);
  off += ng * GN_MAX_SPLITS * sizeof(double2);
  a->bstats = reinterpret_cast<float*>(s + off);
  const long long vec = static_cast<long long>(a->HW) * (a->Cg / 8);
  long long splits = (4LL * num_sms()) / static_cast<long long>(ng);
  const long long by_work = vec / 512;
  if (splits > by_work) splits = by_work;
  if (splits > GN_MAX_SPLITS) splits = GN_MAX_SPLITS;
  if (splits < 1) splits = 1;
  a->splits = static_cast<int>(splits);
  return ISB_OK;
}

}  // namespace isb

extern "C" {

size_t isb_gn_scratch_bytes(int N, int groups) {
  const size_t ng = static_cast<size_t>(N) * groups;
  size_t off = isb::GN_MAX_NG * sizeof(int);      // fixed counter region, see gn_fill_args
  off += ng * isb::GN_MAX_SPLITS * sizeof(double2);
  off += ng * 2 * sizeof(float);
  return off;
}

int isb_gn_forward(const isb_gn_desc* d, void* scratch, isb_stream_t stream) {
  ISB_CHECK_ARG(d != nullptr, "isb_gn_forward: null desc");
  isb::GnArgs a;
  int rc = isb::gn_fill_args(d, scratch, &a);
  if (rc) return rc;
  ISB_CHECK_ARG(d->y != nullptr, "isb_gn_forward: y missing");
  cudaStream_t st = isb::as_stream(stream);
  isb::GnFwdOut o{d->y, d->y_dtype, d->raw, d->raw_dtype, d->xres};
  if (d->partials != nullptr) {
    ISB_CHECK_ARG(d->x2 == nullptr && d->resample == 0 && d->partial_slots > 0 && d->groups == 32 &&
                      (a.Cg == 8 || a.Cg == 16 || a.Cg == 32),
                  "isb_gn_forward: fused statistics need a single source, no resample, 32 groups of 8/16/32 channels");
    if (isb::gn_use_rows(a)) {
      { isb::PdlFamily fam(0);
        ISB_CUDA(isb::launch(isb::gn_fold_stats_kernel, dim3(isb::cdiv(a.N * a.groups, 8)), dim3(256), 0, st, a,
                             reinterpret_cast<const float2*>(d->partials), d->partial_slots)); }
      ISB_LAUNCH_CHECK();
      { isb::PdlFamily fam(1); ISB_CUDA(isb::gn_launch_rows(a, o, st)); }
      ISB_LAUNCH_CHECK();
      return ISB_OK;
    }
    const int ppc = 64 * isb::GN_PART_UNROLL;
    isb::PdlFamily fam(1);
    ISB_CUDA(isb::launch(isb::gn_apply_part_kernel, dim3(isb::cdiv(a.HW, ppc), a.C / 32, a.N), dim3(256), 0, st, a, o,
                         reinterpret_cast<const float2*>(d->partials), d->partial_slots, ppc));
    ISB_LAUNCH_CHECK();
    return ISB_OK;
  }
  if (isb::gn_use_fused(a)) {
    const int cs = isb::gn_cluster_size_for(a);
    isb::PdlFamily fam(2);
    ISB_CUDA(isb::launch_cluster_z(isb::gn_fused_fwd_kernel, dim3(a.groups, a.N, cs), dim3(512), cs, st, a, o));
    ISB_LAUNCH_CHECK();
    return ISB_OK;
  }
  { isb::PdlFamily fam(0); ISB_CUDA(isb::launch(isb::gn_stats_kernel, dim3(a.splits, a.groups, a.N), 256, 0, st, a)); }
  ISB_LAUNCH_CHECK();
  if (isb::gn_use_rows(a)) {
    { isb::PdlFamily fam(1); ISB_CUDA(isb::gn_launch_rows(a, o, st)); }
    ISB_LAUNCH_CHECK();
    return ISB_OK;
  }
  const int Ho = a.resample == 1 ? a.H / 2 : a.H, Wo = a.resample == 1 ? a.W / 2 : a.W;
  const long long total = static_cast<long long>(a.N) * Ho * Wo * (a.C / 8);
  { isb::PdlFamily fam(1); ISB_CUDA(isb::launch(isb::gn_apply_kernel, isb::cdiv(total, 256), 256, 0, st, a, o)); }
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

int isb_gn_backward(const isb_gn_bwd_desc* d, void* scratch, isb_stream_t stream) {
  ISB_CHECK_ARG(d != nullptr && d->dy != nullptr, "isb_gn_backward: dy missing");
  isb::GnArgs a;
  int rc = isb::gn_fill_args(&d->f, scratch, &a);
  if (rc) return rc;
  ISB_CHECK_ARG(d->gx1 != nullptr || d->gx2 != nullptr || d->gx1_lo != nullptr || d->gx2_lo != nullptr,
                "isb_gn_backward: no gradient output");
  isb::GnBwdArgs b{d->dy, d->gres, d->gres_at_input, d->gx1, d->acc1, d->gx1_lo,
                   d->gx2, d->acc2, d->gx2_lo, d->lo_dtype};
  cudaStream_t st = isb::as_stream(stream);
  if (d->partials != nullptr) {
    ISB_CHECK_ARG(d->f.x2 == nullptr && d->f.resample == 0 && d->partial_slots > 0 && d->f.groups == 32 &&
                      (a.Cg == 8 || a.Cg == 16 || a.Cg == 32) && d->gx2 == nullptr && d->gx2_lo == nullptr,
                  "isb_gn_backward: fused reduction terms need a single source, no resample, 32 groups of 8/16/32 channels");
    const int ppc = 64 * isb::GN_PART_UNROLL_BWD * 2;
    isb::PdlFamily fam(1);
    ISB_CUDA(isb::launch(isb::gn_bwd_apply_part_kernel, dim3(isb::cdiv(a.HW, ppc), a.C / 32, a.N), dim3(256), 0, st, a, b,
                         reinterpret_cast<const float2*>(d->partials), d->partial_slots, ppc));
    ISB_LAUNCH_CHECK();
    return ISB_OK;
  }
  if (isb::gn_use_fused(a)) {
    const int cs = isb::gn_cluster_size_for(a);
    isb::PdlFamily fam(2);
    ISB_CUDA(isb::launch_cluster_z(isb::gn_fused_bwd_kernel, dim3(a.groups, a.N, cs), dim3(512), cs, st, a, b));
    ISB_LAUNCH_CHECK();
    return ISB_OK;
  }
  if (isb::gn_use_rows(a) && a.groups <= 64) {
    const int CV = a.C / 8;
    const int rows = CV >= 256 ? 1 : 256 / CV;
    const int step = rows * isb::GN_ROWS_UNROLL;
    int ppc = isb::cdiv(a.HW, isb::GN_MAX_SPLITS);
    ppc = isb::cdiv(ppc, step) * step;
    const int chunks = isb::cdiv(a.HW, ppc);
    isb::PdlFamily fam(0);
    ISB_CUDA(isb::launch(isb::gn_bwd_reduce_rows_kernel, dim3(chunks, a.N), dim3(CV * rows), 0, st, a, b, rows, ppc, chunks));
  } else {
    isb::PdlFamily fam(0);
    ISB_CUDA(isb::launch(isb::gn_bwd_reduce_kernel, dim3(a.splits, a.groups, a.N), 256, 0, st, a, b));
  }
  ISB_LAUNCH_CHECK();
  if (isb::gn_use_rows(a)) {
    const int CV = a.C / 8;
    const int rows = CV >= 256 ? 1 : 256 / CV;
    const int ppc = rows * isb::GN_ROWS_UNROLL_BWD * 4;
    isb::PdlFamily fam(1);
    ISB_CUDA(isb::launch(isb::gn_bwd_apply_rows_kernel, dim3(isb::cdiv(a.HW, ppc), a.N), dim3(CV * rows), 0, st, a, b, rows, ppc));
    ISB_LAUNCH_CHECK();
    return ISB_OK;
  }
  const long long total = static_cast<long long>(a.N) * a.HW * (a.C / 8);
  { isb::PdlFamily fam(1); ISB_CUDA(isb::launch(isb::gn_bwd_apply_kernel, isb::cdiv(total, 256), 256, 0, st, a, b)); }
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // extern "C"
