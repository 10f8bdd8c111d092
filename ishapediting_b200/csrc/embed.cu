// embed.cu — timestep-embedding path: sinusoid (nn.py:102-120) -> time_embed MLP
// (unet.py:471-475) -> every ResBlock's emb_layers Linear at once (unet.py:199-205,245).
// The result depends only on t, so the host may cache it per step index; the GEMVs are
// weight-bandwidth bound (43 M fp32 weights): one warp per output row, float4 loads.
#include "common.cuh"

namespace isb {

// emb[n, i] = cos(t*f_i) for i < half, sin(t*f_{i-half}) otherwise; freqs supplied by the host
// (computed with the reference's own expression so the table is bit-identical).
__global__ void sinusoid_kernel(const float* __restrict__ t, const float* __restrict__ freqs, int N, int half,
                                float* __restrict__ emb) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * half) return;
  const int n = idx / half, i = idx % half;
  const float arg = t[n] * freqs[i];     // timesteps[:, None].float() * freqs[None]  (nn.py:116)
  emb[n * 2 * half + i] = cosf(arg);
  emb[n * 2 * half + half + i] = sinf(arg);
}

// out[n, r] = act(sum_k W[r,k] * in[n,k] + b[r]); one warp per row r, all n.
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ in,
                 float* __restrict__ out, int R, int K, int N, int silu_out) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float4* w4 = reinterpret_cast<const float4*>(W + static_cast<size_t>(r) * K);
  for (int n = 0; n < N; ++n) {
    const float4* x4 = reinterpret_cast<const float4*>(in + static_cast<size_t>(n) * K);
    float acc = 0.f;
    for (int k = lane; k < K / 4; k += 32) {
      const float4 w = __ldg(w4 + k);
      const float4 x = x4[k];
      acc = fmaf(w.x, x.x, acc); acc = fmaf(w.y, x.y, acc);
      acc = fmaf(w.z, x.z, acc); acc = fmaf(w.w, x.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + b[r];
      if (silu_out) v = v / (1.0f + expf(-v));
      out[static_cast<size_t>(n) * R + r] = v;
    }
  }
}

}  // namespace isb

extern "C" {

int isb_time_embed(const float* t, const float* freqs, int N, int model_ch, int hidden, const float* w1,
                   const float* b1, const float* w2, const float* b2, const float* w_all, const float* b_all,
                   int rows_all, float* scratch, float* film_all, isb_stream_t stream) {
  ISB_CHECK_ARG(t && freqs && w1 && b1 && w2 && b2 && w_all && b_all && scratch && film_all, "isb_time_embed: null pointer");
  ISB_CHECK_ARG(model_ch % 8 == 0 && hidden % 4 == 0 && N > 0, "isb_time_embed: model_ch %% 8, hidden %% 4 required");
  cudaStream_t st = isb::as_stream(stream);
  float* emb = scratch;                       // [N, model_ch]
  float* h1 = emb + static_cast<size_t>(N) * model_ch;   // [N, hidden]
  float* h2 = h1 + static_cast<size_t>(N) * hidden;      // [N, hidden] = silu(time_embed(emb))
  const int half = model_ch / 2;
  ISB_CUDA(isb::launch(isb::sinusoid_kernel, isb::cdiv(N * half, 128), 128, 0, st, t, freqs, N, half, emb));
  ISB_LAUNCH_CHECK();
  ISB_CUDA(isb::launch(isb::gemv_rows_kernel, isb::cdiv(hidden, 8), 256, 0, st, w1, b1, emb, h1, hidden, model_ch, N, 1));
  ISB_LAUNCH_CHECK();
  // every consumer applies SiLU first (unet.py:200), so store silu(emb) directly
  ISB_CUDA(isb::launch(isb::gemv_rows_kernel, isb::cdiv(hidden, 8), 256, 0, st, w2, b2, h1, h2, hidden, hidden, N, 1));
  ISB_LAUNCH_CHECK();
  ISB_CUDA(isb::launch(isb::gemv_rows_kernel, isb::cdiv(rows_all, 8), 256, 0, st, w_all, b_all, h2, film_all, rows_all, hidden, N, 0));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // extern "C"
