// attention.cu — QKVAttentionLegacy core (unet.py:337-354) and its input-gradient backward,
// fp32 on the FFMA pipes: strided batched GEMMs (64x64x16 tiles) + row softmax kernels.
// T <= 1024, head dim 64; attention is 1.9 % of the step's FLOPs, the score
// matrix [N,heads,T,T] is kept for the backward instead of being recomputed.
//
// qkv layout (legacy interleave, unet.py:346): channel block of head h is [q(ch) | k(ch) | v(ch)].
#include "common.cuh"

namespace isb {

constexpr int BG_BM = 64, BG_BN = 64, BG_BK = 16, BG_PAD = 4;

struct BgemmParams {
  const float* A; const float* B; void* C;
  int M, N, K;
  int lda, ldb, ldc;
  long long a_b0, a_b1, b_b0, b_b1, c_b0, c_b1;  // batch strides: outer (image), inner (head)
  int inner;                                      // heads
  float alpha;
  int c_dtype;
};

// C[m,n] = alpha * sum_k A(m,k) * B(n,k)
// A_K: A(m,k) = A[m*lda + k] (k contiguous) else A[k*lda + m] (m contiguous); same for B with n.
template <bool A_K, bool B_K>
__global__ void __launch_bounds__(256)
bgemm_kernel(const BgemmParams p) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ __align__(16) float As[BG_BK][BG_BM + BG_PAD];
  __shared__ __align__(16) float Bs[BG_BK][BG_BN + BG_PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int bz = blockIdx.z;
  const int bo = bz / p.inner, bi = bz % p.inner;
  const float* A = p.A + bo * p.a_b0 + bi * p.a_b1;
  const float* B = p.B + bo * p.b_b0 + bi * p.b_b1;
  const int m0 = blockIdx.x * BG_BM, n0 = blockIdx.y * BG_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BG_BK) {
    float4 av, bv;
    if (A_K) av = __ldg(reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + (tid >> 2)) * p.lda + k0 + (tid & 3) * 4));
    else     av = __ldg(reinterpret_cast<const float4*>(A + static_cast<size_t>(k0 + (tid >> 4)) * p.lda + m0 + (tid & 15) * 4));
    if (B_K) bv = __ldg(reinterpret_cast<const float4*>(B + static_cast<size_t>(n0 + (tid >> 2)) * p.ldb + k0 + (tid & 3) * 4));
    else     bv = __ldg(reinterpret_cast<const float4*>(B + static_cast<size_t>(k0 + (tid >> 4)) * p.ldb + n0 + (tid & 15) * 4));
    __syncthreads();
    if (A_K) {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      As[kk][r] = av.x; As[kk + 1][r] = av.y; As[kk + 2][r] = av.z; As[kk + 3][r] = av.w;
    } else {
      *reinterpret_cast<float4*>(&As[tid >> 4][(tid & 15) * 4]) = av;
    }
    if (B_K) {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      Bs[kk][r] = bv.x; Bs[kk + 1][r] = bv.y; Bs[kk + 2][r] = bv.z; Bs[kk + 3][r] = bv.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = bv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BG_BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w};
      const float br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  const long long coff = bo * p.c_b0 + bi * p.c_b1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t off = static_cast<size_t>(coff) + static_cast<size_t>(m0 + ty * 4 + i) * p.ldc + n0 + tx * 4;
    const float v0 = acc[i][0] * p.alpha, v1 = acc[i][1] * p.alpha, v2 = acc[i][2] * p.alpha, v3 = acc[i][3] * p.alpha;
    if (p.c_dtype == ISB_BF16) {
      uint2 u;
      u.x = pack_bf16x2(v0, v1);
      u.y = pack_bf16x2(v2, v3);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off) = u;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + off) = make_float4(v0, v1, v2, v3);
    }
  }
}

// ---- bf16 tensor-core variant (mma.sync m16n8k16, fp32 accumulate) -------------------------
// Same contract as bgemm_kernel; operands are read as fp32 and rounded to bf16 on their way into
// shared memory ("bf16 mode": q, k, v, P and dS are bf16 MMA operands, softmax stays fp32).
constexpr int TG_BM = 64, TG_BN = 64, TG_BK = 32, TG_LD = TG_BK + 8;  // 80-byte rows: conflict-free fragment loads

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// A [64 rows x 32 k] operand tile travels global(fp32) -> registers -> smem(bf16, rows x k, k contiguous).
// Two-phase so that the next tile's global loads are in flight while the tensor cores chew the current one.
//   KMAJOR  (row*ld + k): 4 x float4 along k per thread, 8-byte smem stores.
//   !KMAJOR (k*ld + row): thread owns one row and 16 consecutive k; 16 scalar loads, each coalesced across
//                         the warp (consecutive rows), two conflict-free 16-byte smem stores.
template <bool KMAJOR>
__device__ __forceinline__ void tg_fetch(const float* __restrict__ src, int ld, int row0, int k0, float* r) {
  const int tid = threadIdx.x;
  if (KMAJOR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128;
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(row0 + (idx >> 3)) * ld + k0 + (idx & 7) * 4));
      r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
  } else {
    const int row = tid & 63, kb = (tid >> 6) * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = __ldg(src + static_cast<size_t>(k0 + kb + j) * ld + row0 + row);
  }
}
template <bool KMAJOR>
__device__ __forceinline__ void tg_stash(const float* r, __nv_bfloat16 (*sm)[TG_LD]) {
  const int tid = threadIdx.x;
  if (KMAJOR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128;
      uint2 u;
      u.x = pack_bf16x2(r[4 * i], r[4 * i + 1]);
      u.y = pack_bf16x2(r[4 * i + 2], r[4 * i + 3]);
      *reinterpret_cast<uint2*>(&sm[idx >> 3][(idx & 7) * 4]) = u;
    }
  } else {
    const int row = tid & 63, kb = (tid >> 6) * 16;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 u;
      u.x = pack_bf16x2(r[8 * h], r[8 * h + 1]);
      u.y = pack_bf16x2(r[8 * h + 2], r[8 * h + 3]);
      u.z = pack_bf16x2(r[8 * h + 4], r[8 * h + 5]);
      u.w = pack_bf16x2(r[8 * h + 6], r[8 * h + 7]);
      *reinterpret_cast<uint4*>(&sm[row][kb + 8 * h]) = u;
    }
  }
}

template <bool A_K, bool B_K>
__global__ void __launch_bounds__(128)
bgemm_mma_kernel(const BgemmParams p) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ __align__(16) __nv_bfloat16 As[TG_BM][TG_LD];
  __shared__ __align__(16) __nv_bfloat16 Bs[TG_BN][TG_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = lane >> 2, t = lane & 3;
  const int bz = blockIdx.z;
  const int bo = bz / p.inner, bi = bz % p.inner;
  const float* A = p.A + bo * p.a_b0 + bi * p.a_b1;
  const float* B = p.B + bo * p.b_b0 + bi * p.b_b1;
  const int m0 = blockIdx.x * TG_BM, n0 = blockIdx.y * TG_BN;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

  float ra[16], rb[16];
  tg_fetch<A_K>(A, p.lda, m0, 0, ra);
  tg_fetch<B_K>(B, p.ldb, n0, 0, rb);
  for (int k0 = 0; k0 < p.K; k0 += TG_BK) {
    __syncthreads();   // previous tile's fragment loads are done
    tg_stash<A_K>(ra, As);
    tg_stash<B_K>(rb, Bs);
    __syncthreads();
    if (k0 + TG_BK < p.K) {
      tg_fetch<A_K>(A, p.lda, m0, k0 + TG_BK, ra);
      tg_fetch<B_K>(B, p.ldb, n0, k0 + TG_BK, rb);
    }
#pragma unroll
    for (int kk = 0; kk < TG_BK; kk += 16) {
      uint32_t af[2][4], bf[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int r = wm * 32 + mi * 16 + g;
        af[mi][0] = *reinterpret_cast<const uint32_t*>(&As[r][kk + 2 * t]);
        af[mi][1] = *reinterpret_cast<const uint32_t*>(&As[r + 8][kk + 2 * t]);
        af[mi][2] = *reinterpret_cast<const uint32_t*>(&As[r][kk + 2 * t + 8]);
        af[mi][3] = *reinterpret_cast<const uint32_t*>(&As[r + 8][kk + 2 * t + 8]);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int c = wn * 32 + ni * 8 + g;
        bf[ni][0] = *reinterpret_cast<const uint32_t*>(&Bs[c][kk + 2 * t]);
        bf[ni][1] = *reinterpret_cast<const uint32_t*>(&Bs[c][kk + 2 * t + 8]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) mma_bf16_16816(acc[mi][ni], af[mi], bf[ni]);
    }
  }
  const long long coff = bo * p.c_b0 + bi * p.c_b1;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = m0 + wm * 32 + mi * 16 + g + h * 8;
        const int col = n0 + wn * 32 + ni * 8 + 2 * t;
        const size_t off = static_cast<size_t>(coff) + static_cast<size_t>(row) * p.ldc + col;
        const float v0 = acc[mi][ni][2 * h] * p.alpha, v1 = acc[mi][ni][2 * h + 1] * p.alpha;
        if (p.c_dtype == ISB_BF16) *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off) = pack_bf16x2(v0, v1);
        else *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.C) + off) = make_float2(v0, v1);
      }
}

template <bool A_K, bool B_K>
static int bgemm_mma(const BgemmParams& p, int batches, cudaStream_t st) {
  ISB_CHECK_ARG(p.M % TG_BM == 0 && p.N % TG_BN == 0 && p.K % TG_BK == 0, "attention gemm (mma): M=%d N=%d K=%d must be multiples of 64/64/32", p.M, p.N, p.K);
  dim3 grid(p.M / TG_BM, p.N / TG_BN, batches);
  ISB_CUDA(isb::launch(bgemm_mma_kernel<A_K, B_K>, grid, 128, 0, st, p));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

template <bool A_K, bool B_K>
static int bgemm(const BgemmParams& p, int batches, cudaStream_t st) {
  ISB_CHECK_ARG(p.M % BG_BM == 0 && p.N % BG_BN == 0 && p.K % BG_BK == 0, "attention gemm: M=%d N=%d K=%d must be multiples of 64/64/16", p.M, p.N, p.K);
  dim3 grid(p.M / BG_BM, p.N / BG_BN, batches);
  ISB_CUDA(isb::launch(bgemm_kernel<A_K, B_K>, grid, 256, 0, st, p));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

template <bool A_K, bool B_K>
static int bgemm_any(const BgemmParams& p, int batches, bool tensor_cores, cudaStream_t st) {
  return tensor_cores ? bgemm_mma<A_K, B_K>(p, batches, st) : bgemm<A_K, B_K>(p, batches, st);
}

// one block (128 threads) per row of length T: in-place fp32 softmax (unet.py:352)
__global__ void __launch_bounds__(128)
softmax_rows_kernel(float* __restrict__ s, int T) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ float red[4];
  float* row = s + static_cast<size_t>(blockIdx.x) * T;
  const int tid = threadIdx.x;
  float v[8];  // T <= 1024
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = tid + i * 128;
    v[i] = c < T ? row[c] : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = tid + i * 128;
    v[i] = c < T ? expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  sum = (red[0] + red[1]) + (red[2] + red[3]);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = tid + i * 128;
    if (c < T) row[c] = v[i] * inv;
  }
}

// dS = alpha * P * (dP - sum_s dP*P), in place on dp
__global__ void __launch_bounds__(128)
softmax_bwd_rows_kernel(const float* __restrict__ probs, float* __restrict__ dp, int T, float alpha) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ float red[4];
  const float* prow = probs + static_cast<size_t>(blockIdx.x) * T;
  float* drow = dp + static_cast<size_t>(blockIdx.x) * T;
  const int tid = threadIdx.x;
  float pv[8], dv[8];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = tid + i * 128;
    pv[i] = c < T ? prow[c] : 0.f;
    dv[i] = c < T ? drow[c] : 0.f;
    dot = fmaf(pv[i], dv[i], dot);
  }
  dot = warp_sum(dot);
  if ((tid & 31) == 0) red[tid >> 5] = dot;
  __syncthreads();
  dot = (red[0] + red[1]) + (red[2] + red[3]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = tid + i * 128;
    if (c < T) drow[c] = alpha * pv[i] * (dv[i] - dot);
  }
}

}  // namespace isb

extern "C" {

int isb_attention_forward(const float* qkv, int N, int T, int heads, int ch, float* probs, void* out,
                          int out_dtype, isb_stream_t stream) {
  ISB_CHECK_ARG(qkv && probs && out, "isb_attention_forward: null pointer");
  ISB_CHECK_ARG(T % 64 == 0 && T <= 1024 && ch % 64 == 0, "isb_attention_forward: T=%d (<=1024, %%64) ch=%d (%%64) unsupported", T, ch);
  cudaStream_t st = isb::as_stream(stream);
  const int C = heads * ch;
  const long long TT = static_cast<long long>(T) * T;
  isb::BgemmParams p{};
  // S = Q K^T / sqrt(ch)
  p.A = qkv; p.B = qkv + ch; p.C = probs;
  p.M = T; p.N = T; p.K = ch;
  p.lda = 3 * C; p.ldb = 3 * C; p.ldc = T;
  p.a_b0 = static_cast<long long>(T) * 3 * C; p.a_b1 = 3 * ch;
  p.b_b0 = p.a_b0; p.b_b1 = p.a_b1;
  p.c_b0 = heads * TT; p.c_b1 = TT;
  p.inner = heads; p.alpha = 1.0f / sqrtf(static_cast<float>(ch)); p.c_dtype = ISB_F32;
  const bool tc = out_dtype == ISB_BF16;   // bf16 mode -> tensor cores; fp32 mode -> FFMA
  int rc = isb::bgemm_any<true, true>(p, N * heads, tc, st);
  if (rc) return rc;
  ISB_CUDA(isb::launch(isb::softmax_rows_kernel, N * heads * T, 128, 0, st, probs, T));
  ISB_LAUNCH_CHECK();
  // O = P V
  p.A = probs; p.B = qkv + 2 * ch; p.C = out;
  p.M = T; p.N = ch; p.K = T;
  p.lda = T; p.ldb = 3 * C; p.ldc = C;
  p.a_b0 = heads * TT; p.a_b1 = TT;
  p.b_b0 = static_cast<long long>(T) * 3 * C; p.b_b1 = 3 * ch;
  p.c_b0 = static_cast<long long>(T) * C; p.c_b1 = ch;
  p.alpha = 1.0f; p.c_dtype = out_dtype;
  return isb::bgemm_any<true, false>(p, N * heads, tc, st);
}

int isb_attention_backward(const float* qkv, const float* probs, const float* d_out, int N, int T, int heads,
                           int ch, float* tmp, void* d_qkv, int lo_dtype, isb_stream_t stream) {
  ISB_CHECK_ARG(qkv && probs && d_out && tmp && d_qkv, "isb_attention_backward: null pointer");
  ISB_CHECK_ARG(T % 64 == 0 && T <= 1024 && ch % 64 == 0, "isb_attention_backward: T=%d ch=%d unsupported", T, ch);
  cudaStream_t st = isb::as_stream(stream);
  const int C = heads * ch;
  const long long TT = static_cast<long long>(T) * T;
  const long long qkv_b0 = static_cast<long long>(T) * 3 * C;
  const size_t esz = lo_dtype == ISB_BF16 ? 2 : 4;
  char* dq = static_cast<char*>(d_qkv);
  const float alpha = 1.0f / sqrtf(static_cast<float>(ch));
  isb::BgemmParams p{};
  p.inner = heads;
  // dV[s,c] = sum_t P[t,s] dO[t,c]
  p.A = probs; p.B = d_out; p.C = dq + static_cast<size_t>(2 * ch) * esz;
  p.M = T; p.N = ch; p.K = T;
  p.lda = T; p.ldb = C; p.ldc = 3 * C;
  p.a_b0 = heads * TT; p.a_b1 = TT;
  p.b_b0 = static_cast<long long>(T) * C; p.b_b1 = ch;
  p.c_b0 = qkv_b0; p.c_b1 = 3 * ch;
  p.alpha = 1.0f; p.c_dtype = lo_dtype;
  const bool tc = lo_dtype == ISB_BF16;
  int rc = isb::bgemm_any<false, false>(p, N * heads, tc, st);
  if (rc) return rc;
  // dP[t,s] = sum_c dO[t,c] V[s,c]
  p.A = d_out; p.B = qkv + 2 * ch; p.C = tmp;
  p.M = T; p.N = T; p.K = ch;
  p.lda = C; p.ldb = 3 * C; p.ldc = T;
  p.a_b0 = static_cast<long long>(T) * C; p.a_b1 = ch;
  p.b_b0 = qkv_b0; p.b_b1 = 3 * ch;
  p.c_b0 = heads * TT; p.c_b1 = TT;
  p.alpha = 1.0f; p.c_dtype = ISB_F32;
  rc = isb::bgemm_any<true, true>(p, N * heads, tc, st);
  if (rc) return rc;
  ISB_CUDA(isb::launch(isb::softmax_bwd_rows_kernel, N * heads * T, 128, 0, st, probs, tmp, T, alpha));
  ISB_LAUNCH_CHECK();
  // dQ[t,c] = sum_s dS[t,s] K[s,c]
  p.A = tmp; p.B = qkv + ch; p.C = dq;
  p.M = T; p.N = ch; p.K = T;
  p.lda = T; p.ldb = 3 * C; p.ldc = 3 * C;
  p.a_b0 = heads * TT; p.a_b1 = TT;
  p.b_b0 = qkv_b0; p.b_b1 = 3 * ch;
  p.c_b0 = qkv_b0; p.c_b1 = 3 * ch;
  p.alpha = 1.0f; p.c_dtype = lo_dtype;
  rc = isb::bgemm_any<true, false>(p, N * heads, tc, st);
  if (rc) return rc;
  // dK[s,c] = sum_t dS[t,s] Q[t,c]
  p.A = tmp; p.B = qkv; p.C = dq + static_cast<size_t>(ch) * esz;
  p.lda = T; p.ldb = 3 * C; p.ldc = 3 * C;
  return isb::bgemm_any<false, false>(p, N * heads, tc, st);
}

}  // extern "C"
