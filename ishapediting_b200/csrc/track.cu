// track.cu — OPT-IN point tracking: nearest-feature search around the current handle points
// (BASELINE.json north_star: "point tracking becomes a nearest-feature search with a warp-level argmin").
//
// The reference has NO tracking function (handles and targets are fixed for a whole edit,
// drag_utils.py:305-321), so this is an extension with PARITY UNPINNED; it is checked for self-consistency only
// (tests/test_gpu_track.py: the distance tables against a float64 torch restatement, the argmin index bit-exactly
// against a brute-force search over the same fp32 distances).
//
// Search space: the (2r+1)^3 voxel lattice around each handle point p (offsets in make_offsets order, index =
// ((i+r)*side + (j+r))*side + (k+r)).  The triplane feature of a 3-D point q is the concatenation of the bilinear
// samples of the three aligned planes at (qx,qy), (qy,qz), (qx,qz) (drag_utils.py:318-321,355-358), so the L1
// distance to the handle's stored feature f0 separates:
//     D(i,j,k) = T_xy[i][j] + T_yz[j][k] + T_xz[i][k]
// with three (2r+1)^2 tables per handle.  Kernel 1 fills the tables (one warp per table entry, lanes over channels,
// fixed shuffle-tree order); kernel 2 scans the lattice per handle and takes the argmin with a warp-level
// (value, index) shuffle reduction — ties go to the LOWEST index, like torch.argmin on the flattened lattice.
#include "common.cuh"

namespace isb {

struct TrackArgs {
  const float* feat; int S; int Cf; int Ca;
  const int32_t* chan_map;
  const float* f0;        // [B,3,Ca]
  const float* center;    // [B,3]
  int B, r, side;
  float voxel;
  float* table;           // [B,3,side*side]
  int32_t* out_idx;       // [B]
  float* out_dist;        // [B]
  float* out_pts;         // [B,3]
};

// one warp per (handle, plane, table entry)
__global__ void __launch_bounds__(256)
track_table_kernel(const TrackArgs a) {
  pdl_wait();
  pdl_trigger();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int per = a.side * a.side;
  if (warp >= a.B * 3 * per) return;
  const int e = warp % per, pl = (warp / per) % 3, h = warp / (3 * per);
  const int ia = e / a.side, ib = e - ia * a.side;              // offsets along the plane's (u, v) axes
  const int au = pl == 1 ? 1 : 0, av = pl == 0 ? 1 : 2;         // xy, yz, xz
  const float u = a.center[h * 3 + au] + a.voxel * static_cast<float>(ia - a.r);
  const float v = a.center[h * 3 + av] + a.voxel * static_cast<float>(ib - a.r);
  const int S = a.S;
  const float ix = ((u + 1.0f) / 2.0f) * static_cast<float>(S - 1);     // grid_sample, align_corners=True
  const float iy = ((v + 1.0f) / 2.0f) * static_cast<float>(S - 1);
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
  const float fx = ix - fx0, fy = iy - fy0;
  const float* f0 = a.f0 + (static_cast<size_t>(h) * 3 + pl) * a.Ca;
  float acc = 0.f;
  for (int ch = lane; ch < a.Ca; ch += 32) {
    const int src_c = a.chan_map[pl * a.Ca + ch];
    float sv = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int x = x0 + dx, y = y0 + dy;
        if (x >= 0 && x < S && y >= 0 && y < S) {                 // zeros padding
          const float w = (dx ? fx : 1.0f - fx) * (dy ? fy : 1.0f - fy);
          sv = fmaf(w, __ldg(a.feat + (static_cast<size_t>(y) * S + x) * a.Cf + src_c), sv);
        }
      }
    acc += fabsf(sv - f0[ch]);
  }
  acc = warp_sum(acc);
  if (lane == 0) a.table[warp] = acc;
}

__device__ __forceinline__ void argmin_combine(float& v, int& i, float ov, int oi) {
  if (ov < v || (ov == v && oi < i)) {
    v = ov;
    i = oi;
  }
}

// one CTA per handle
__global__ void __launch_bounds__(256)
track_argmin_kernel(const TrackArgs a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s_tab[];        // three tables of this handle
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const int h = blockIdx.x, side = a.side, per = side * side;
  for (int e = threadIdx.x; e < 3 * per; e += blockDim.x) s_tab[e] = a.table[static_cast<size_t>(h) * 3 * per + e];
  __syncthreads();
  const float* txy = s_tab;
  const float* tyz = s_tab + per;
  const float* txz = s_tab + 2 * per;
  float best = __int_as_float(0x7f800000);   // +inf
  int besti = 0x7fffffff;
  const int total = per * side;
  for (int q = threadIdx.x; q < total; q += blockDim.x) {
    const int k = q % side, j = (q / side) % side, i = q / per;
    const float d = (txy[i * side + j] + tyz[j * side + k]) + txz[i * side + k];
    argmin_combine(best, besti, d, q);
  }
  // warp-level argmin: (value, index) pairs through the shuffle tree
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
    argmin_combine(best, besti, ov, oi);
  }
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = besti;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? s_v[threadIdx.x] : __int_as_float(0x7f800000);
    besti = threadIdx.x < (blockDim.x >> 5) ? s_i[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
      argmin_combine(best, besti, ov, oi);
    }
    if (threadIdx.x == 0) {
      a.out_idx[h] = besti;
      a.out_dist[h] = best;
      const int k = besti % side, j = (besti / side) % side, i = besti / per;
      a.out_pts[h * 3 + 0] = a.center[h * 3 + 0] + a.voxel * static_cast<float>(i - a.r);
      a.out_pts[h * 3 + 1] = a.center[h * 3 + 1] + a.voxel * static_cast<float>(j - a.r);
      a.out_pts[h * 3 + 2] = a.center[h * 3 + 2] + a.voxel * static_cast<float>(k - a.r);
    }
  }
}

}  // namespace isb

extern "C" {

int isb_track_points(const isb_track_desc* d, isb_stream_t stream) {
  if (!isb::is_initialised()) {
    isb::set_error("isb_track_points: isb_init() has not been called");
    return ISB_ERR_NOT_INIT;
  }
  ISB_CHECK_ARG(d && d->feat && d->chan_map && d->f0 && d->center && d->table && d->out_idx && d->out_dist && d->out_pts,
                "isb_track_points: null pointer");
  ISB_CHECK_ARG(d->B > 0 && d->r >= 0 && d->r <= 31 && d->S > 1 && d->Ca > 0 && d->Cf > 0,
                "isb_track_points: bad sizes (B=%d, r=%d (<= 31), S=%d, Ca=%d)", d->B, d->r, d->S, d->Ca);
  isb::TrackArgs a;
  a.feat = d->feat; a.S = d->S; a.Cf = d->Cf; a.Ca = d->Ca;
  a.chan_map = d->chan_map; a.f0 = d->f0; a.center = d->center;
  a.B = d->B; a.r = d->r; a.side = 2 * d->r + 1; a.voxel = d->voxel;
  a.table = d->table; a.out_idx = d->out_idx; a.out_dist = d->out_dist; a.out_pts = d->out_pts;
  cudaStream_t st = isb::as_stream(stream);
  const int per = a.side * a.side;
  const long long warps = static_cast<long long>(a.B) * 3 * per;
  ISB_CUDA(isb::launch(isb::track_table_kernel, isb::cdiv(warps * 32, 256), 256, 0, st, a));
  ISB_LAUNCH_CHECK();
  ISB_CUDA(isb::launch(isb::track_argmin_kernel, a.B, 256, static_cast<size_t>(3 * per) * sizeof(float), st, a));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // extern "C"
