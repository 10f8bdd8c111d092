// drag.cu — drag guidance: motion-supervision loss + mask regulariser and their analytic
// gradient w.r.t. the intermediate UNet feature (drag_utils.py:141-159, 351-383).
//
// The reference materialises resize_feat_align (permute/interpolate/cat), two
// (3,170,B,15625) grid_sample outputs and lets autograd scatter-add the gradient back.  Here:
//   * resize_feat_align is an index map (chan_map / inv_map) applied on the fly;
//   * the (2r+1)^3 patch lattice projects onto (2r+1)^2 distinct 2-D points per plane and
//     handle, each with multiplicity (2r+1): the host passes the distinct points + weights;
//   * kernel 1 samples origin@patch and edit@shift bilinearly (align_corners=True, zeros
//     padding) and stores the per-point loss derivative g;
//   * kernel 2 is a GATHER over (plane, pixel) (no atomics -> bit-reproducible): each CTA compacts,
//     in point order, the samples whose footprint covers its pixel (bounding boxes prune whole
//     handles), then one thread per aligned channel accumulates them, adds the mask-regulariser
//     term and writes dLoss/dfeat straight into the NHWC feature gradient;
//   * kernel 3 folds the loss partials in fixed order.
#include "common.cuh"

namespace isb {

struct DragArgs {
  const float* feat; int S; int Cf;
  const float* origin; int Ca;
  const int32_t* chan_map; const int32_t* inv_map;
  const float* patch_xy; const float* shift_xy; const float* weight;
  int npts; int group_size; int ngroups;
  const int32_t* bbox;
  const uint8_t* mask; int mask_count;
  float inv_count; float cof; int loss_type;
  float* g; float* pt_info; double* partial; int n_gather_blocks;
  float* loss; float* d_feat;
  const float* dyn;   // optional device [2]: inv_count, 1/(Ca*mask_count) — overrides the by-value scalars
};

struct Bilin {
  int x0, y0;
  float fx, fy;
};
// torch grid_sampler_compute_source_index, align_corners=True
__device__ __forceinline__ Bilin bilin_setup(float gx, float gy, int S) {
  const float ix = ((gx + 1.0f) / 2.0f) * static_cast<float>(S - 1);
  const float iy = ((gy + 1.0f) / 2.0f) * static_cast<float>(S - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  Bilin b;
  b.x0 = static_cast<int>(fx0);
  b.y0 = static_cast<int>(fy0);
  b.fx = ix - fx0;
  b.fy = iy - fy0;
  return b;
}

// grid (npts, 3); block loops over the Ca aligned channels
__global__ void __launch_bounds__(192)
drag_sample_kernel(const DragArgs a) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const int j = blockIdx.x, pl = blockIdx.y;
  const int S = a.S;
  const size_t pj = static_cast<size_t>(pl) * a.npts + j;
  const Bilin bp = bilin_setup(a.patch_xy[pj * 2], a.patch_xy[pj * 2 + 1], S);
  const Bilin bs = bilin_setup(a.shift_xy[pj * 2], a.shift_xy[pj * 2 + 1], S);
  const float wt = a.weight[j];
  if (threadIdx.x == 0) {
    float4 info = make_float4(static_cast<float>(bs.x0), static_cast<float>(bs.y0), bs.fx, bs.fy);
    reinterpret_cast<float4*>(a.pt_info)[pj] = info;
  }
  float lsum = 0.f;
  for (int ch = threadIdx.x; ch < a.Ca; ch += blockDim.x) {
    float pv = 0.f, sv = 0.f;
    const int src_c = a.chan_map[pl * a.Ca + ch];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        {
          const int x = bp.x0 + dx, y = bp.y0 + dy;
          if (x >= 0 && x < S && y >= 0 && y < S) {
            const float w = (dx ? bp.fx : 1.0f - bp.fx) * (dy ? bp.fy : 1.0f - bp.fy);
            pv = fmaf(w, __ldg(a.origin + ((static_cast<size_t>(pl) * S + y) * S + x) * a.Ca + ch), pv);
          }
        }
        {
          const int x = bs.x0 + dx, y = bs.y0 + dy;
          if (x >= 0 && x < S && y >= 0 && y < S) {
            const float w = (dx ? bs.fx : 1.0f - bs.fx) * (dy ? bs.fy : 1.0f - bs.fy);
            sv = fmaf(w, __ldg(a.feat + (static_cast<size_t>(y) * S + x) * a.Cf + src_c), sv);
          }
        }
      }
    const float diff = sv - pv;
    float gv;
    if (a.loss_type == 0) { gv = 2.0f * diff; lsum = fmaf(diff, diff, lsum); }
    else { gv = static_cast<float>((diff > 0.f) - (diff < 0.f)); lsum += fabsf(diff); }
    a.g[pj * a.Ca + ch] = wt * gv;
  }
  __shared__ float red[6];
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    a.partial[pj] = s * wt;
  }
}

// one CTA per (plane, pixel), one thread per aligned channel.  The warps first build — in point order, so
// the summation order is fixed — the short list of sample points whose bilinear footprint covers this
// pixel (bounding boxes prune whole handles), then every channel thread walks that list.
constexpr int DG_THREADS = 192;
constexpr int DG_MAXLIST = 1024;

__global__ void __launch_bounds__(DG_THREADS)
drag_gather_kernel(const DragArgs a) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  __shared__ int s_j[DG_MAXLIST];
  __shared__ float s_w[DG_MAXLIST];
  __shared__ int s_cnt[DG_THREADS / 32 + 1];
  __shared__ float red[DG_THREADS / 32];
  const int S = a.S;
  const int pix = blockIdx.x, pl = blockIdx.y;
  const int x = pix % S, y = pix / S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = DG_THREADS / 32;

  // Four handles cover a few percent of the 3 x S x S pixels: a pixel outside every handle's bounding box has no
  // contributing sample point and skips both list passes (CTA-uniform test, so the barriers inside stay legal).
  bool any = false;
  for (int gi = 0; gi < a.ngroups; ++gi) {
    const int4 bb = __ldg(reinterpret_cast<const int4*>(a.bbox) + pl * a.ngroups + gi);
    any |= !(x < bb.x || x > bb.y + 1 || y < bb.z || y > bb.w + 1);
  }
  int base = 0, total = 0;
  if (any) {
    // pass 1: count per warp (each warp owns a contiguous segment of the point list)
    const int seg = (a.npts + nwarps - 1) / nwarps;
    const int jbeg = warp * seg, jend = min(a.npts, jbeg + seg);
    int my_count = 0;
    for (int j0 = jbeg; j0 < jend; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < jend) {
        const int4 bb = __ldg(reinterpret_cast<const int4*>(a.bbox) + pl * a.ngroups + j / a.group_size);
        if (!(x < bb.x || x > bb.y + 1 || y < bb.z || y > bb.w + 1)) {
          const float4 info = reinterpret_cast<const float4*>(a.pt_info)[static_cast<size_t>(pl) * a.npts + j];
          const int ddx = x - static_cast<int>(info.x), ddy = y - static_cast<int>(info.y);
          hit = ddx >= 0 && ddx <= 1 && ddy >= 0 && ddy <= 1;
        }
      }
      my_count += __popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0) s_cnt[warp] = my_count;
    __syncthreads();
    base = 0;
    for (int w = 0; w < nwarps; ++w) {
      if (w < warp) base += s_cnt[w];
      total += s_cnt[w];
    }
    // pass 2: ordered compaction into the shared list
    if (total > 0 && total <= DG_MAXLIST) {
      int pos = base;
      for (int j0 = jbeg; j0 < jend; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        float w = 0.f;
        if (j < jend) {
          const int4 bb = __ldg(reinterpret_cast<const int4*>(a.bbox) + pl * a.ngroups + j / a.group_size);
          if (!(x < bb.x || x > bb.y + 1 || y < bb.z || y > bb.w + 1)) {
            const float4 info = reinterpret_cast<const float4*>(a.pt_info)[static_cast<size_t>(pl) * a.npts + j];
            const int ddx = x - static_cast<int>(info.x), ddy = y - static_cast<int>(info.y);
            hit = ddx >= 0 && ddx <= 1 && ddy >= 0 && ddy <= 1;
            if (hit) w = (ddx ? info.z : 1.0f - info.z) * (ddy ? info.w : 1.0f - info.w);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const int k = pos + __popc(m & ((1u << lane) - 1u));
          s_j[k] = j;
          s_w[k] = w;
        }
        pos += __popc(m);
      }
    }
  }
  __syncthreads();

  float msq = 0.f;
  const bool masked = a.cof > 0.f && a.mask[(static_cast<size_t>(pl) * S + y) * S + x];
  const float inv_count = a.dyn ? a.dyn[0] : a.inv_count;
  const float norm = !masked ? 0.f : (a.dyn ? a.dyn[1] : 1.0f / (static_cast<float>(a.Ca) * static_cast<float>(a.mask_count)));
  for (int ch = threadIdx.x; ch < a.Ca; ch += DG_THREADS) {
    float acc = 0.f;
    if (total <= DG_MAXLIST) {
      for (int k = 0; k < total; ++k)
        acc = fmaf(s_w[k], a.g[(static_cast<size_t>(pl) * a.npts + s_j[k]) * a.Ca + ch], acc);
    } else {   // pathological overlap: walk the whole list (same order)
      for (int j = 0; j < a.npts; ++j) {
        const float4 info = reinterpret_cast<const float4*>(a.pt_info)[static_cast<size_t>(pl) * a.npts + j];
        const int ddx = x - static_cast<int>(info.x), ddy = y - static_cast<int>(info.y);
        if (ddx < 0 || ddx > 1 || ddy < 0 || ddy > 1) continue;
        const float w = (ddx ? info.z : 1.0f - info.z) * (ddy ? info.w : 1.0f - info.w);
        acc = fmaf(w, a.g[(static_cast<size_t>(pl) * a.npts + j) * a.Ca + ch], acc);
      }
    }
    float d = -inv_count * acc;
    const int src_c = a.chan_map[pl * a.Ca + ch];
    const size_t fidx = static_cast<size_t>(pix) * a.Cf + src_c;
    if (masked) {
      const float df = a.feat[fidx] - __ldg(a.origin + ((static_cast<size_t>(pl) * S + y) * S + x) * a.Ca + ch);
      if (a.loss_type == 0) { d -= a.cof * 2.0f * df * norm; msq += df * df; }
      else { d -= a.cof * static_cast<float>((df > 0.f) - (df < 0.f)) * norm; msq += fabsf(df); }
    }
    a.d_feat[fidx] = d;
  }
  if (pl == 0)   // channels resize_feat_align drops get a zero gradient
    for (int c = threadIdx.x; c < a.Cf; c += DG_THREADS)
      if (a.inv_map[c] < 0) a.d_feat[static_cast<size_t>(pix) * a.Cf + c] = 0.f;
  msq = warp_sum(msq);
  if (lane == 0) red[warp] = msq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0;
    for (int w = 0; w < nwarps; ++w) sum += red[w];
    a.partial[static_cast<size_t>(3) * a.npts + static_cast<size_t>(pl) * S * S + pix] = sum;
  }
}

__global__ void __launch_bounds__(1024)
drag_loss_kernel(const DragArgs a) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  // one CTA (the sum must have a fixed order) but 1024 threads with four loads in flight each: this kernel sits
  // between the gather and the backward pass on the critical path
  __shared__ double red[2][32];
  double motion = 0, mask = 0;
  const int n_motion = 3 * a.npts;
#pragma unroll 4
  for (int i = threadIdx.x; i < n_motion; i += blockDim.x) motion += a.partial[i];
#pragma unroll 4
  for (int i = threadIdx.x; i < a.n_gather_blocks; i += blockDim.x) mask += a.partial[n_motion + i];
  motion = warp_sum_d(motion);
  mask = warp_sum_d(mask);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = motion; red[1][threadIdx.x >> 5] = mask; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mo = 0, ma = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { mo += red[0][w]; ma += red[1][w]; }
    const double inv_count = a.dyn ? a.dyn[0] : a.inv_count;
    const double mnorm = a.dyn ? a.dyn[1] : (a.mask_count > 0 ? 1.0 / (static_cast<double>(a.Ca) * a.mask_count) : 0.0);
    double loss = -mo * inv_count;
    if (a.cof > 0.f) loss -= static_cast<double>(a.cof) * ma * mnorm;
    *a.loss = static_cast<float>(loss);
  }
}

__global__ void __launch_bounds__(256)
resize_feat_align_kernel(const float* __restrict__ feat, int S, int Cf, const int32_t* __restrict__ chan_map,
                         float* __restrict__ out, int Ca) {
  pdl_wait();      // predecessor complete and flushed
  pdl_trigger();   // only now may the successor become resident (no cascade of waiting grids)
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = 3LL * S * S * Ca;
  if (idx >= total) return;
  const int ch = static_cast<int>(idx % Ca);
  const long long t = idx / Ca;
  const int pix = static_cast<int>(t % (S * S));
  const int pl = static_cast<int>(t / (S * S));
  out[idx] = __ldg(feat + static_cast<size_t>(pix) * Cf + chan_map[pl * Ca + ch]);
}

}  // namespace isb

extern "C" {

size_t isb_drag_partial_len(int S, int Cf, int npts) {
  (void)Cf;
  return static_cast<size_t>(3) * npts + static_cast<size_t>(3) * S * S;
}

int isb_drag_loss_grad(const isb_drag_desc* d, isb_stream_t stream) {
  ISB_CHECK_ARG(d && d->feat && d->origin && d->chan_map && d->inv_map && d->patch_xy && d->shift_xy && d->weight &&
                    d->bbox && d->g && d->pt_info && d->partial && d->loss && d->d_feat,
                "isb_drag_loss_grad: null pointer");
  ISB_CHECK_ARG(d->S > 1 && d->Cf > 0 && d->Ca > 0 && d->npts > 0, "isb_drag_loss_grad: bad shape");
  ISB_CHECK_ARG(d->group_size > 0 && d->npts % d->group_size == 0, "isb_drag_loss_grad: npts must be a multiple of group_size");
  ISB_CHECK_ARG(d->cof <= 0.f || (d->mask != nullptr && (d->mask_count > 0 || d->dyn_scalars != nullptr)), "isb_drag_loss_grad: mask required when cof > 0");
  const int gblocks = 3 * d->S * d->S;
  ISB_CHECK_ARG(static_cast<size_t>(d->partial_len) >= isb_drag_partial_len(d->S, d->Cf, d->npts), "isb_drag_loss_grad: partial buffer too small");
  isb::DragArgs a{d->feat, d->S, d->Cf, d->origin, d->Ca, d->chan_map, d->inv_map,
                  d->patch_xy, d->shift_xy, d->weight, d->npts, d->group_size, d->npts / d->group_size,
                  d->bbox, d->mask, d->mask_count, d->inv_count, d->cof, d->loss_type,
                  d->g, d->pt_info, d->partial, gblocks, d->loss, d->d_feat, d->dyn_scalars};
  cudaStream_t st = isb::as_stream(stream);
  ISB_CUDA(isb::launch(isb::drag_sample_kernel, dim3(d->npts, 3), 192, 0, st, a));
  ISB_LAUNCH_CHECK();
  ISB_CUDA(isb::launch(isb::drag_gather_kernel, dim3(d->S * d->S, 3), isb::DG_THREADS, 0, st, a));
  ISB_LAUNCH_CHECK();
  ISB_CUDA(isb::launch(isb::drag_loss_kernel, 1, 1024, 0, st, a));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

int isb_resize_feat_align(const float* feat, int S, int Cf, const int32_t* chan_map, float* out, int Ca,
                          isb_stream_t stream) {
  ISB_CHECK_ARG(feat && chan_map && out && S > 0 && Cf > 0 && Ca > 0, "isb_resize_feat_align: bad args");
  const long long total = 3LL * S * S * Ca;
  ISB_CUDA(isb::launch(isb::resize_feat_align_kernel, isb::cdiv(total, 256), 256, 0, isb::as_stream(stream), feat, S, Cf, chan_map, out, Ca));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // extern "C"
