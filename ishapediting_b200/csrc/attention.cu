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

template <bool A_K, bool B_K>
static int bgemm(const BgemmParams& p, int batches, cudaStream_t st) {
  ISB_CHECK_ARG(p.M % BG_BM == 0 && p.N % BG_BN == 0 && p.K % BG_BK == 0, "attention gemm: M=%d N=%d K=%d must be multiples of 64/64/16", p.M, p.N, p.K);
  dim3 grid(p.M / BG_BM, p.N / BG_BN, batches);
  bgemm_kernel<A_K, B_K><<<grid, 256, 0, st>>>(p);
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

// one block (128 threads) per row of length T: in-place fp32 softmax (unet.py:352)
__global__ void __launch_bounds__(128)
softmax_rows_kernel(float* __restrict__ s, int T) {
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
  int rc = isb::bgemm<true, true>(p, N * heads, st);
  if (rc) return rc;
  isb::softmax_rows_kernel<<<N * heads * T, 128, 0, st>>>(probs, T);
  ISB_LAUNCH_CHECK();
  // O = P V
  p.A = probs; p.B = qkv + 2 * ch; p.C = out;
  p.M = T; p.N = ch; p.K = T;
  p.lda = T; p.ldb = 3 * C; p.ldc = C;
  p.a_b0 = heads * TT; p.a_b1 = TT;
  p.b_b0 = static_cast<long long>(T) * 3 * C; p.b_b1 = 3 * ch;
  p.c_b0 = static_cast<long long>(T) * C; p.c_b1 = ch;
  p.alpha = 1.0f; p.c_dtype = out_dtype;
  return isb::bgemm<true, false>(p, N * heads, st);
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
  int rc = isb::bgemm<false, false>(p, N * heads, st);
  if (rc) return rc;
  // dP[t,s] = sum_c dO[t,c] V[s,c]
  p.A = d_out; p.B = qkv + 2 * ch; p.C = tmp;
  p.M = T; p.N = T; p.K = ch;
  p.lda = C; p.ldb = 3 * C; p.ldc = T;
  p.a_b0 = static_cast<long long>(T) * C; p.a_b1 = ch;
  p.b_b0 = qkv_b0; p.b_b1 = 3 * ch;
  p.c_b0 = heads * TT; p.c_b1 = TT;
  p.alpha = 1.0f; p.c_dtype = ISB_F32;
  rc = isb::bgemm<true, true>(p, N * heads, st);
  if (rc) return rc;
  isb::softmax_bwd_rows_kernel<<<N * heads * T, 128, 0, st>>>(probs, tmp, T, alpha);
  ISB_LAUNCH_CHECK();
  // dQ[t,c] = sum_s dS[t,s] K[s,c]
  p.A = tmp; p.B = qkv + ch; p.C = dq;
  p.M = T; p.N = ch; p.K = T;
  p.lda = T; p.ldb = 3 * C; p.ldc = 3 * C;
  p.a_b0 = heads * TT; p.a_b1 = TT;
  p.b_b0 = qkv_b0; p.b_b1 = 3 * ch;
  p.c_b0 = qkv_b0; p.c_b1 = 3 * ch;
  p.alpha = 1.0f; p.c_dtype = lo_dtype;
  rc = isb::bgemm<true, false>(p, N * heads, st);
  if (rc) return rc;
  // dK[s,c] = sum_t dS[t,s] Q[t,c]
  p.A = tmp; p.B = qkv; p.C = dq + static_cast<size_t>(ch) * esz;
  p.lda = T; p.ldb = 3 * C; p.ldc = 3 * C;
  return isb::bgemm<false, false>(p, N * heads, st);
}

}  // extern "C"
