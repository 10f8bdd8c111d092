// attention_flash.cu — fused multi-head self-attention for the bf16 mode of the NFD UNet
// (reference: guided_diffusion/unet.py:337-354, QKVAttentionLegacy.forward, 64 channels per head).
//
// The unfused path (attention.cu) materialises the [heads, T, T] fp32 probability tensor: 33.5 MB per
// 32x32-resolution layer, written by QK^T, rewritten by softmax, read by PV and read three more times by the
// backward pass — ~1.3 ms of the 5.7 ms guided step.  Here the probabilities never leave the SM:
//   forward     one CTA per (64 queries, head): S = QK^T -> online softmax -> O += PV; writes O and the
//               per-row log-sum-exp (log2 domain)
//   backward    ONE launch, two kinds of CTA:
//     dq half   one CTA per (64 queries, head): recomputes P from the log-sum-exp, dP = dO V^T,
//               dS = P o (dP - delta) * scale, dQ += dS K;  also writes delta = rowsum(dO o O)
//     dkv half  one CTA per (64 keys, head): recomputes P^T (and delta), dV += P^T dO, dK += dS^T Q
// No atomics: every output element has exactly one writer and a fixed summation order (deterministic).
// Tensor cores through mma.sync m16n8k16 (bf16 in, fp32 accumulate), operands staged with cp.async into padded
// shared-memory tiles and fetched with ldmatrix.  (The GEMMs here are 64-deep and ~40 GFLOP per step in total:
// latency-bound, not worth a tcgen05/TMEM pipeline.)
#include "common.cuh"

namespace isb {

constexpr int FA_D = 64;                    // channels per head
constexpr int FA_B = 64;                    // rows per tile (queries or keys)
constexpr int FA_LDB = 144;                 // padded tile row in BYTES (72 bf16): conflict-free ldmatrix
constexpr int FA_TILE = FA_B * FA_LDB;      // 9216 B

struct FaParams {
  const __nv_bfloat16* qkv;     // [N][T][heads][3][64]
  __nv_bfloat16* out;           // [N][T][heads*64]     (forward: written; backward: read)
  const __nv_bfloat16* d_out;   // [N][T][heads*64]
  float* lse;                   // [N][heads][T]  row log-sum-exp of the scaled scores, log2 domain
  float* delta;                 // [N][heads][T]  rowsum(dO o O)
  __nv_bfloat16* d_qkv;         // [N][T][heads][3][64]
  int T, heads;
  float scale_log2, scale;      // scale = 1/sqrt(ch) (q and k each carry ch^-1/4, unet.py:348-351)
};

__device__ __forceinline__ uint32_t fa_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void fa_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// 64 rows x 64 bf16 (128 B) from global (row stride in elements) into a padded tile; 128 threads
__device__ __forceinline__ void fa_load_tile(uint32_t tile, const __nv_bfloat16* g, size_t row_stride, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 128;
    const int row = idx >> 3, ch = idx & 7;
    cp_async16(tile + static_cast<uint32_t>(row * FA_LDB + ch * 16), g + static_cast<size_t>(row) * row_stride + ch * 8);
  }
}
// A fragments (16 rows of this warp x 64 k) of a row-major tile: 4 k-steps
__device__ __forceinline__ void fa_load_a(uint32_t (&a)[4][4], uint32_t tile, int warp, int lane) {
  const uint32_t base = tile + static_cast<uint32_t>((warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * FA_LDB + (lane >> 4) * 16);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], base + static_cast<uint32_t>(ks * 32));
}
// C[16 x 64] (+)= A[16 x 64] * Bt^T, Bt stored [n = 64 rows][k = 64]  ("NT": scores, dP)
__device__ __forceinline__ void fa_gemm_nt(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
  const uint32_t base = tile + static_cast<uint32_t>(((lane & 7) + (lane >> 4) * 8) * FA_LDB + ((lane >> 3) & 1) * 16);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(b, base + static_cast<uint32_t>(np * 16 * FA_LDB + ks * 32));
      fa_mma(c[2 * np], a[ks], b[0], b[1]);
      fa_mma(c[2 * np + 1], a[ks], b[2], b[3]);
    }
}
// C[16 x 64] += P[16 x 64] * B, B stored [k = 64 rows][n = 64]  ("NN": PV, dS K, P^T dO, dS^T Q)
__device__ __forceinline__ void fa_gemm_nn(float (&c)[8][4], const uint32_t (&pa)[4][4], uint32_t tile, int lane) {
  const uint32_t base = tile + static_cast<uint32_t>(((lane & 7) + ((lane >> 3) & 1) * 8) * FA_LDB + (lane >> 4) * 16);
#pragma unroll
  for (int kt = 0; kt < 4; ++kt)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_trans(b, base + static_cast<uint32_t>(kt * 16 * FA_LDB + np * 32));
      fa_mma(c[2 * np], pa[kt], b[0], b[1]);
      fa_mma(c[2 * np + 1], pa[kt], b[2], b[3]);
    }
}
// accumulator tile (16 x 64 fp32) -> bf16 A fragments for the next GEMM (k = the 64 columns)
__device__ __forceinline__ void fa_pack_a(uint32_t (&pa)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    pa[kt][0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
    pa[kt][1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
    pa[kt][2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
    pa[kt][3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
  }
}
__device__ __forceinline__ void fa_zero(float (&c)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}
// 16 x 64 accumulator of this warp -> bf16 rows (g, g+8) of a [rows][row_stride] global matrix
__device__ __forceinline__ void fa_store_rows(__nv_bfloat16* dst, size_t row_stride, const float (&c)[8][4], int warp,
                                              int lane, float s0, float s1) {
  const int g = lane >> 2, t = lane & 3;
  __nv_bfloat16* r0 = dst + static_cast<size_t>(warp * 16 + g) * row_stride + 2 * t;
  __nv_bfloat16* r1 = r0 + 8 * row_stride;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    *reinterpret_cast<uint32_t*>(r0 + dt * 8) = pack_bf16x2(c[dt][0] * s0, c[dt][1] * s0);
    *reinterpret_cast<uint32_t*>(r1 + dt * 8) = pack_bf16x2(c[dt][2] * s1, c[dt][3] * s1);
  }
}

// ---- forward -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fa_fwd_kernel(const FaParams p) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  extern __shared__ __align__(16) uint8_t fa_smem[];
  const uint32_t sQ = fa_smem_u32(fa_smem);
  auto sK = [&](int j) { return sQ + static_cast<uint32_t>((1 + 2 * (j & 1)) * FA_TILE); };
  auto sV = [&](int j) { return sQ + static_cast<uint32_t>((2 + 2 * (j & 1)) * FA_TILE); };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int qb = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const size_t C3 = static_cast<size_t>(3) * p.heads * FA_D;
  const __nv_bfloat16* base = p.qkv + static_cast<size_t>(n) * p.T * C3 + static_cast<size_t>(h) * 3 * FA_D;
  fa_load_tile(sQ, base + static_cast<size_t>(qb) * FA_B * C3, C3, tid);
  fa_load_tile(sK(0), base + FA_D, C3, tid);
  fa_load_tile(sV(0), base + 2 * FA_D, C3, tid);
  cp_async_commit();
  const int nkv = p.T / FA_B;
  float o[8][4];
  fa_zero(o);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  uint32_t aq[4][4];
  for (int j = 0; j < nkv; ++j) {
    if (j + 1 < nkv) {
      const __nv_bfloat16* nb = base + static_cast<size_t>(j + 1) * FA_B * C3;
      fa_load_tile(sK((j + 1) & 1), nb + FA_D, C3, tid);
      fa_load_tile(sV((j + 1) & 1), nb + 2 * FA_D, C3, tid);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (j == 0) fa_load_a(aq, sQ, warp, lane);
    float s[8][4];
    fa_zero(s);
    fa_gemm_nt(s, aq, sK(j & 1), lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] *= p.scale_log2;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
    const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - mn0);
      s[nt][1] = exp2f(s[nt][1] - mn0);
      s[nt][2] = exp2f(s[nt][2] - mn1);
      s[nt][3] = exp2f(s[nt][3] - mn1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * c0 + rs0;
    l1 = l1 * c1 + rs1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= c0; o[dt][1] *= c0;
      o[dt][2] *= c1; o[dt][3] *= c1;
    }
    uint32_t pa[4][4];
    fa_pack_a(pa, s);
    fa_gemm_nn(o, pa, sV(j & 1), lane);
    __syncthreads();   // this stage is overwritten by the prefetch of iteration j+1
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  const size_t C = static_cast<size_t>(p.heads) * FA_D;
  const int row0 = qb * FA_B;
  fa_store_rows(p.out + (static_cast<size_t>(n) * p.T + row0) * C + static_cast<size_t>(h) * FA_D, C, o, warp, lane,
                1.0f / l0, 1.0f / l1);
  if (t == 0) {
    float* L = p.lse + (static_cast<size_t>(n) * p.heads + h) * p.T + row0 + warp * 16 + g;
    L[0] = m0 + log2f(l0);
    L[8] = m1 + log2f(l1);
  }
}

// delta[row] = sum_c dO[row][c] * O[row][c] for the 64 rows of a tile: two threads per row, 32 channels each
__device__ __forceinline__ void fa_delta_fetch(const __nv_bfloat16* dO, const __nv_bfloat16* O, size_t C, int tid,
                                               uint4 (&ua)[4], uint4 (&ub)[4]) {
  const int row = tid >> 1, half = tid & 1;
  const uint4* a = reinterpret_cast<const uint4*>(dO + static_cast<size_t>(row) * C + half * 32);
  const uint4* b = reinterpret_cast<const uint4*>(O + static_cast<size_t>(row) * C + half * 32);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ua[i] = a[i];
    ub[i] = b[i];
  }
}
__device__ __forceinline__ float fa_delta_dot(const uint4 (&ua)[4], const uint4 (&ub)[4]) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162* pa2 = reinterpret_cast<const __nv_bfloat162*>(&ua[i]);
    const __nv_bfloat162* pb2 = reinterpret_cast<const __nv_bfloat162*>(&ub[i]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __bfloat1622float2(pa2[e]), fb = __bfloat1622float2(pb2[e]);
      acc = fmaf(fa.x, fb.x, acc);
      acc = fmaf(fa.y, fb.y, acc);
    }
  }
  return acc + __shfl_xor_sync(0xffffffffu, acc, 1);
}

// ---- backward 1: dQ (and delta) ---------------------------------------------------------------------------
__device__ __forceinline__ void fa_bwd_dq_body(const FaParams& p, uint8_t* fa_smem, float* Dsm, int qb, int h, int n) {
  const uint32_t sQ = fa_smem_u32(fa_smem), sdO = sQ + FA_TILE;
  auto sK = [&](int j) { return sQ + static_cast<uint32_t>((2 + 2 * (j & 1)) * FA_TILE); };
  auto sV = [&](int j) { return sQ + static_cast<uint32_t>((3 + 2 * (j & 1)) * FA_TILE); };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2;
  const size_t C3 = static_cast<size_t>(3) * p.heads * FA_D, C = static_cast<size_t>(p.heads) * FA_D;
  const __nv_bfloat16* base = p.qkv + static_cast<size_t>(n) * p.T * C3 + static_cast<size_t>(h) * 3 * FA_D;
  const int row0 = qb * FA_B;
  const __nv_bfloat16* dO = p.d_out + (static_cast<size_t>(n) * p.T + row0) * C + static_cast<size_t>(h) * FA_D;
  const __nv_bfloat16* O = p.out + (static_cast<size_t>(n) * p.T + row0) * C + static_cast<size_t>(h) * FA_D;
  fa_load_tile(sQ, base + static_cast<size_t>(row0) * C3, C3, tid);
  fa_load_tile(sdO, dO, C, tid);
  fa_load_tile(sK(0), base + FA_D, C3, tid);
  fa_load_tile(sV(0), base + 2 * FA_D, C3, tid);
  cp_async_commit();
  const size_t stat0 = (static_cast<size_t>(n) * p.heads + h) * p.T + row0;
  {  // delta[row] = sum_c dO[row][c] * O[row][c]: two threads per row
    uint4 ua[4], ub[4];
    fa_delta_fetch(dO, O, C, tid, ua, ub);
    const float acc = fa_delta_dot(ua, ub);
    if ((tid & 1) == 0) {
      Dsm[tid >> 1] = acc;
      p.delta[stat0 + (tid >> 1)] = acc;
    }
  }
  const float L0 = p.lse[stat0 + warp * 16 + g], L1 = p.lse[stat0 + warp * 16 + g + 8];
  const int nkv = p.T / FA_B;
  float dq[8][4];
  fa_zero(dq);
  uint32_t aq[4][4], ado[4][4];
  float d0 = 0.f, d1 = 0.f;
  for (int j = 0; j < nkv; ++j) {
    if (j + 1 < nkv) {
      const __nv_bfloat16* nb = base + static_cast<size_t>(j + 1) * FA_B * C3;
      fa_load_tile(sK((j + 1) & 1), nb + FA_D, C3, tid);
      fa_load_tile(sV((j + 1) & 1), nb + 2 * FA_D, C3, tid);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (j == 0) {
      fa_load_a(aq, sQ, warp, lane);
      fa_load_a(ado, sdO, warp, lane);
      d0 = Dsm[warp * 16 + g];
      d1 = Dsm[warp * 16 + g + 8];
    }
    float s[8][4], dp[8][4];
    fa_zero(s);
    fa_zero(dp);
    fa_gemm_nt(s, aq, sK(j & 1), lane);
    fa_gemm_nt(dp, ado, sV(j & 1), lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * p.scale_log2 - L0) * (dp[nt][0] - d0) * p.scale;
      s[nt][1] = exp2f(s[nt][1] * p.scale_log2 - L0) * (dp[nt][1] - d0) * p.scale;
      s[nt][2] = exp2f(s[nt][2] * p.scale_log2 - L1) * (dp[nt][2] - d1) * p.scale;
      s[nt][3] = exp2f(s[nt][3] * p.scale_log2 - L1) * (dp[nt][3] - d1) * p.scale;
    }
    uint32_t pa[4][4];
    fa_pack_a(pa, s);
    fa_gemm_nn(dq, pa, sK(j & 1), lane);
    __syncthreads();
  }
  fa_store_rows(p.d_qkv + (static_cast<size_t>(n) * p.T + row0) * C3 + static_cast<size_t>(h) * 3 * FA_D, C3, dq, warp,
                lane, 1.0f, 1.0f);
}

// ---- backward 2: dK, dV ------------------------------------------------------------------------------------
__device__ __forceinline__ void fa_bwd_dkv_body(const FaParams& p, uint8_t* fa_smem, float (*Ls)[FA_B], float (*Ds)[FA_B],
                                                int kb, int h, int n) {
  const uint32_t sK = fa_smem_u32(fa_smem), sV = sK + FA_TILE;
  auto sQ = [&](int i) { return sK + static_cast<uint32_t>((2 + 2 * (i & 1)) * FA_TILE); };
  auto sdO = [&](int i) { return sK + static_cast<uint32_t>((3 + 2 * (i & 1)) * FA_TILE); };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = lane & 3;
  const size_t C3 = static_cast<size_t>(3) * p.heads * FA_D, C = static_cast<size_t>(p.heads) * FA_D;
  const __nv_bfloat16* base = p.qkv + static_cast<size_t>(n) * p.T * C3 + static_cast<size_t>(h) * 3 * FA_D;
  const __nv_bfloat16* dO = p.d_out + static_cast<size_t>(n) * p.T * C + static_cast<size_t>(h) * FA_D;
  const __nv_bfloat16* O = p.out + static_cast<size_t>(n) * p.T * C + static_cast<size_t>(h) * FA_D;
  const size_t stat0 = (static_cast<size_t>(n) * p.heads + h) * p.T;
  const int key0 = kb * FA_B;
  fa_load_tile(sK, base + static_cast<size_t>(key0) * C3 + FA_D, C3, tid);
  fa_load_tile(sV, base + static_cast<size_t>(key0) * C3 + 2 * FA_D, C3, tid);
  fa_load_tile(sQ(0), base, C3, tid);
  fa_load_tile(sdO(0), dO, C, tid);
  cp_async_commit();
  // delta = rowsum(dO o O) is recomputed here for every query block (the dq half of the launch runs concurrently
  // and cannot hand it over): the 64 rows are fetched into registers one iteration ahead
  uint4 ua[4], ub[4];
  fa_delta_fetch(dO, O, C, tid, ua, ub);
  if (tid < FA_B) Ls[0][tid] = p.lse[stat0 + tid];
  {
    const float d = fa_delta_dot(ua, ub);
    if ((tid & 1) == 0) Ds[0][tid >> 1] = d;
  }
  const int nq = p.T / FA_B;
  float dk[8][4], dv[8][4];
  fa_zero(dk);
  fa_zero(dv);
  uint32_t ak[4][4], av[4][4];
  for (int i = 0; i < nq; ++i) {
    if (i + 1 < nq) {
      const size_t r = static_cast<size_t>(i + 1) * FA_B;
      fa_load_tile(sQ((i + 1) & 1), base + r * C3, C3, tid);
      fa_load_tile(sdO((i + 1) & 1), dO + r * C, C, tid);
      if (tid < FA_B) Ls[(i + 1) & 1][tid] = p.lse[stat0 + r + tid];
      fa_delta_fetch(dO + r * C, O + r * C, C, tid, ua, ub);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (i == 0) {
      fa_load_a(ak, sK, warp, lane);
      fa_load_a(av, sV, warp, lane);
    }
    float st[8][4], dpt[8][4];     // S^T and dP^T: rows = keys of this warp, columns = the 64 queries of block i
    fa_zero(st);
    fa_zero(dpt);
    fa_gemm_nt(st, ak, sQ(i & 1), lane);
    fa_gemm_nt(dpt, av, sdO(i & 1), lane);
    const float* Lq = Ls[i & 1];
    const float* Dq = Ds[i & 1];
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int q0 = nt * 8 + 2 * t;
      const float La = Lq[q0], Lb = Lq[q0 + 1], Da = Dq[q0], Db = Dq[q0 + 1];
      const float p0 = exp2f(st[nt][0] * p.scale_log2 - La), p1 = exp2f(st[nt][1] * p.scale_log2 - Lb);
      const float p2 = exp2f(st[nt][2] * p.scale_log2 - La), p3 = exp2f(st[nt][3] * p.scale_log2 - Lb);
      st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
      dpt[nt][0] = p0 * (dpt[nt][0] - Da) * p.scale;
      dpt[nt][1] = p1 * (dpt[nt][1] - Db) * p.scale;
      dpt[nt][2] = p2 * (dpt[nt][2] - Da) * p.scale;
      dpt[nt][3] = p3 * (dpt[nt][3] - Db) * p.scale;
    }
    fa_pack_a(pa, st);
    fa_gemm_nn(dv, pa, sdO(i & 1), lane);
    fa_pack_a(pa, dpt);
    fa_gemm_nn(dk, pa, sQ(i & 1), lane);
    if (i + 1 < nq) {      // the rows fetched at the top of this iteration have arrived by now
      const float d = fa_delta_dot(ua, ub);
      if ((tid & 1) == 0) Ds[(i + 1) & 1][tid >> 1] = d;
    }
    __syncthreads();
  }
  __nv_bfloat16* dst = p.d_qkv + (static_cast<size_t>(n) * p.T + key0) * C3 + static_cast<size_t>(h) * 3 * FA_D;
  fa_store_rows(dst + FA_D, C3, dk, warp, lane, 1.0f, 1.0f);
  fa_store_rows(dst + 2 * FA_D, C3, dv, warp, lane, 1.0f, 1.0f);
}

// One launch for the whole backward: the first T/64 CTAs of a (head, image) take the key blocks (dK, dV: four
// GEMMs per tile pair, the longer job, scheduled first), the other T/64 the query blocks (dQ).  The two halves share
// nothing but inputs, so they fill twice as many SMs as either kernel alone (128 CTAs of 4 warps each at T = 1024
// left most of the machine idle) and one launch latency disappears.
__global__ void __launch_bounds__(128)
fa_bwd_kernel(const FaParams p) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  extern __shared__ __align__(16) uint8_t fa_smem[];
  __shared__ float stat_sm[4][FA_B];
  const int nblk = p.T / FA_B;
  const int h = blockIdx.y, n = blockIdx.z;
  if (static_cast<int>(blockIdx.x) < nblk) fa_bwd_dkv_body(p, fa_smem, &stat_sm[0], &stat_sm[2], blockIdx.x, h, n);
  else fa_bwd_dq_body(p, fa_smem, stat_sm[0], static_cast<int>(blockIdx.x) - nblk, h, n);
}

constexpr int FA_SMEM_FWD = 5 * FA_TILE;   // 46080
constexpr int FA_SMEM_BWD = 6 * FA_TILE;   // 55296

// attention_tc.cu: the forward on tcgen05 / TMEM / TMA for T % 128 == 0
int attention_tc_init();
bool attention_tc_usable(int T, int ch);
int attention_tc_forward(const void* qkv, int N, int T, int heads, void* out, float* lse, float scale_log2, cudaStream_t st);

int attention_flash_init() {
  ISB_CUDA(cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_FWD));
  ISB_CUDA(cudaFuncSetAttribute(fa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM_BWD));
  return attention_tc_init();
}

static int fa_check(const char* who, int N, int T, int heads, int ch) {
  ISB_CHECK_ARG(is_initialised(), "%s: isb_init() has not been called", who);
  ISB_CHECK_ARG(N > 0 && heads > 0 && T > 0 && T % FA_B == 0, "%s: T=%d must be a positive multiple of 64", who, T);
  ISB_CHECK_ARG(ch == FA_D, "%s: %d channels per head unsupported (fused path is built for 64)", who, ch);
  return ISB_OK;
}

}  // namespace isb

extern "C" {

int isb_attention_flash_forward(const void* qkv, int N, int T, int heads, int ch, void* out, float* lse,
                                isb_stream_t stream) {
  ISB_CHECK_ARG(qkv && out && lse, "isb_attention_flash_forward: null pointer");
  int rc = isb::fa_check("isb_attention_flash_forward", N, T, heads, ch);
  if (rc) return rc;
  isb::FaParams p{};
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.T = T;
  p.heads = heads;
  p.scale = 1.0f / sqrtf(static_cast<float>(ch));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  if (isb::attention_tc_usable(T, ch))      // T % 128 == 0: the tcgen05 / TMEM / TMA kernel (attention_tc.cu)
    return isb::attention_tc_forward(qkv, N, T, heads, out, lse, p.scale_log2, isb::as_stream(stream));
  isb::PdlFamily fam(3);
  ISB_CUDA(isb::launch(isb::fa_fwd_kernel, dim3(T / isb::FA_B, heads, N), dim3(128), isb::FA_SMEM_FWD,
                       isb::as_stream(stream), p));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

int isb_attention_flash_backward(const void* qkv, const void* out, const void* d_out, const float* lse, int N, int T,
                                 int heads, int ch, float* delta, void* d_qkv, isb_stream_t stream) {
  ISB_CHECK_ARG(qkv && out && d_out && lse && delta && d_qkv, "isb_attention_flash_backward: null pointer");
  int rc = isb::fa_check("isb_attention_flash_backward", N, T, heads, ch);
  if (rc) return rc;
  isb::FaParams p{};
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.out = static_cast<__nv_bfloat16*>(const_cast<void*>(out));
  p.d_out = static_cast<const __nv_bfloat16*>(d_out);
  p.lse = const_cast<float*>(lse);
  p.delta = delta;
  p.d_qkv = static_cast<__nv_bfloat16*>(d_qkv);
  p.T = T;
  p.heads = heads;
  p.scale = 1.0f / sqrtf(static_cast<float>(ch));
  p.scale_log2 = p.scale * 1.4426950408889634f;
  const dim3 grid(2 * (T / isb::FA_B), heads, N);
  isb::PdlFamily fam(3);
  ISB_CUDA(isb::launch(isb::fa_bwd_kernel, grid, dim3(128), isb::FA_SMEM_BWD, isb::as_stream(stream), p));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // extern "C"
