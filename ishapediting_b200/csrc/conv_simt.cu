// conv_simt.cu — fp32 ("fp32 mode") implicit-GEMM convolution on the FFMA pipes, plus the
// isb_conv2d dispatcher.  Same contract as conv_tc.cu; used when operands are fp32 so that
// eps / guidance gradients can be checked against the reference at 1e-4 (BASELINE north_star).
#include "common.cuh"

namespace isb {

size_t conv_tc_workspace(const isb_conv_desc* d);
int conv_tc_gn_slots(const isb_conv_desc* d);
int conv_tc_launch(const isb_conv_desc* d, void* ws, size_t ws_bytes, cudaStream_t stream);

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 16, SM_PAD = 4;

struct SimtParams {
  const float* a; const float* a2; const float* w;
  const float* bias; const float* residual; float* out;
  int N, H, W, Cin, Cin2, Cout, ksize, accumulate;
  int Ktot;
  long long M;
};

__global__ void __launch_bounds__(256)
conv_simt_kernel(const SimtParams p) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ float As[SM_BK][SM_BM + SM_PAD];
  __shared__ float Bs[SM_BK][SM_BN + SM_PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = static_cast<long long>(blockIdx.x) * SM_BM;
  const int cout0 = blockIdx.y * SM_BN;

  // loader coordinates
  const int lrow = tid >> 2;   // 0..63
  const int lk = (tid & 3) * 4;  // 0,4,8,12
  const long long lm = m0 + lrow;
  const bool lm_ok = lm < p.M;
  int ln = 0, lh = 0, lw = 0;
  if (lm_ok) {
    lw = static_cast<int>(lm % p.W);
    long long t = lm / p.W;
    lh = static_cast<int>(t % p.H);
    ln = static_cast<int>(t / p.H);
  }
  const int lco = cout0 + lrow;
  const bool lco_ok = lco < p.Cout;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ntaps = p.ksize * p.ksize;
  const int chunks0 = p.Cin / SM_BK;
  const int chunks1 = p.a2 ? p.Cin2 / SM_BK : 0;
  const int iters = ntaps * chunks0 + chunks1;
  for (int it = 0; it < iters; ++it) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    int kglob;
    if (it < ntaps * chunks0) {
      const int tap = it / chunks0;
      const int chunk = it - tap * chunks0;
      int dh = 0, dw = 0;
      if (p.ksize == 3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
      const int hh = lh + dh, ww = lw + dw;
      if (lm_ok && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) {
        const size_t off = ((static_cast<size_t>(ln) * p.H + hh) * p.W + ww) * p.Cin + chunk * SM_BK + lk;
        av = __ldg(reinterpret_cast<const float4*>(p.a + off));
      }
      kglob = tap * p.Cin + chunk * SM_BK;
    } else {
      const int chunk = it - ntaps * chunks0;
      if (lm_ok) {
        const size_t off = static_cast<size_t>(lm) * p.Cin2 + chunk * SM_BK + lk;
        av = __ldg(reinterpret_cast<const float4*>(p.a2 + off));
      }
      kglob = ntaps * p.Cin + chunk * SM_BK;
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lco_ok) bv = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(lco) * p.Ktot + kglob + lk));
    __syncthreads();  // previous iteration's reads done
    As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
    Bs[lk + 0][lrow] = bv.x; Bs[lk + 1][lrow] = bv.y; Bs[lk + 2][lrow] = bv.z; Bs[lk + 3][lrow] = bv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SM_BK; ++k) {
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

  const int col = cout0 + tx * 4;
  if (col >= p.Cout) return;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const size_t off = static_cast<size_t>(m) * p.Cout + col;
    float4 v = make_float4(acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w);
    if (p.residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.residual + off));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    float4* dst = reinterpret_cast<float4*>(p.out + off);
    if (p.accumulate) {
      const float4 o = *dst;
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *dst = v;
  }
}

static int conv_simt_launch(const isb_conv_desc* d, cudaStream_t stream) {
  ISB_CHECK_ARG(d->ksize == 1 || d->ksize == 3, "conv_simt: ksize %d unsupported", d->ksize);
  ISB_CHECK_ARG(d->Cin > 0 && d->Cin % 16 == 0, "conv_simt: Cin=%d must be a multiple of 16", d->Cin);
  ISB_CHECK_ARG(d->a2 == nullptr || (d->Cin2 > 0 && d->Cin2 % 16 == 0), "conv_simt: Cin2=%d must be a multiple of 16", d->Cin2);
  ISB_CHECK_ARG(d->Cout > 0 && d->Cout % 4 == 0, "conv_simt: Cout=%d must be a multiple of 4", d->Cout);
  ISB_CHECK_ARG(d->out_dtype == ISB_F32, "conv_simt: fp32 mode writes fp32 only");
  SimtParams p;
  p.a = static_cast<const float*>(d->a);
  p.a2 = static_cast<const float*>(d->a2);
  p.w = static_cast<const float*>(d->w);
  p.bias = d->bias; p.residual = d->residual; p.out = static_cast<float*>(d->out);
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cin2 = d->a2 ? d->Cin2 : 0;
  p.Cout = d->Cout; p.ksize = d->ksize; p.accumulate = d->accumulate;
  p.Ktot = d->ksize * d->ksize * d->Cin + p.Cin2;
  p.M = static_cast<long long>(d->N) * d->H * d->W;
  dim3 grid(cdiv(p.M, SM_BM), cdiv(d->Cout, SM_BN));
  ISB_CUDA(isb::launch(conv_simt_kernel, grid, 256, 0, stream, p));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // namespace isb

extern "C" {

size_t isb_conv2d_workspace(const isb_conv_desc* d) {
  if (d == nullptr) return 0;
  if (d->a_dtype == ISB_BF16) return isb::conv_tc_workspace(d);
  return 0;
}

int isb_conv2d_gn_slots(const isb_conv_desc* d) {
  if (d == nullptr || d->a_dtype != ISB_BF16) return 0;    // the fp32 FFMA path computes no fused statistics
  return isb::conv_tc_gn_slots(d);
}

int isb_conv2d(const isb_conv_desc* d, void* workspace, size_t workspace_bytes, isb_stream_t stream) {
  if (!isb::is_initialised()) {
    isb::set_error("isb_conv2d: isb_init() has not been called");
    return ISB_ERR_NOT_INIT;
  }
  ISB_CHECK_ARG(d != nullptr && d->a != nullptr && d->w != nullptr && d->out != nullptr, "isb_conv2d: null pointer");
  if (d->a_dtype == ISB_BF16) return isb::conv_tc_launch(d, workspace, workspace_bytes, isb::as_stream(stream));
  ISB_CHECK_ARG(d->a_dtype == ISB_F32, "isb_conv2d: bad a_dtype %d", d->a_dtype);
  ISB_CHECK_ARG(d->gn_partials == nullptr, "isb_conv2d: fused GroupNorm statistics need the bf16 path");
  return isb::conv_simt_launch(d, isb::as_stream(stream));
}

}  // extern "C"
