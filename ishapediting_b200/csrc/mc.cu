// mc.cu — the meshing tail of the decode path on the device:
//   * marching cubes over the logit volume, replacing `mcubes.marching_cubes(pred, 0)` + `vertices / res * 2 - 1`
//     (triplane_decoder/visualize.py:100-101): classify -> scan -> generate, shared vertices, deterministic order;
//   * uniform Laplacian smoothing, replacing Open3D's `filter_smooth_simple(number_of_iterations=10)`
//     (drag_utils.py:300): adjacency from the triangles, v_i <- (v_i + sum_{j in N(i)} v_j) / (1 + |N(i)|), float64.
// The case table is DERIVED (triplane_decoder/mc_table.py -> mc_table.inc); PyMCubes is not installed offline, so
// parity with that binary is unpinned — the CPU restatement oracle/mcubes_oracle.py is the checker (vertex set equal,
// element-wise equal triangles, closed oriented manifolds on analytic volumes).
//
// HBM-bound integer / byte work: one thread per grid point, x-major (the volume's own layout, z fastest -> coalesced),
// no tensor cores.  Per point: 4 B read (+ neighbours from L1/L2), 10 B of classification written; emission touches
// only the cells / edges that cross the surface.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "mc_table.inc"

namespace isb {

__constant__ signed char c_mc_tri[256 * 16];
__constant__ unsigned char c_mc_cnt[256];
// edge e = (owning corner, axis); corner i = (x,y,z) offsets, Bourke numbering (mc_table.py)
__constant__ int c_mc_edge_corner[12] = {0, 1, 3, 0, 4, 5, 7, 4, 0, 1, 2, 3};
__constant__ int c_mc_edge_axis[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};
__constant__ int c_mc_cx[8] = {0, 1, 1, 0, 0, 1, 1, 0};
__constant__ int c_mc_cy[8] = {0, 0, 1, 1, 0, 0, 1, 1};
__constant__ int c_mc_cz[8] = {0, 0, 0, 0, 1, 1, 1, 1};

int mc_init() {
  ISB_CUDA(cudaMemcpyToSymbol(c_mc_tri, MC_TRI_TABLE_H, sizeof(MC_TRI_TABLE_H)));
  ISB_CUDA(cudaMemcpyToSymbol(c_mc_cnt, MC_TRI_COUNT_H, sizeof(MC_TRI_COUNT_H)));
  return ISB_OK;
}

struct McWs {
  unsigned char* vflags;   // [n] bit a: the grid edge from this point along +axis a crosses the isovalue
  unsigned char* cube;     // [n] case index of the cell whose corner 0 is this point (0 for non-cells)
  int* vsum;               // [n] inclusive scan of popcount(vflags)
  int* tsum;               // [n] inclusive scan of triangle counts
  void* cub_tmp;
  size_t cub_bytes;
};

static size_t mc_cub_bytes(long long n) {
  size_t b = 0;
  cub::DeviceScan::InclusiveSum(nullptr, b, static_cast<int*>(nullptr), static_cast<int*>(nullptr), static_cast<int>(n));
  return b;
}
static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
static McWs mc_layout(void* ws, long long n) {
  char* p = static_cast<char*>(ws);
  McWs w;
  w.vflags = reinterpret_cast<unsigned char*>(p); p += align256(n);
  w.cube = reinterpret_cast<unsigned char*>(p); p += align256(n);
  w.vsum = reinterpret_cast<int*>(p); p += align256(n * 4);
  w.tsum = reinterpret_cast<int*>(p); p += align256(n * 4);
  w.cub_tmp = p;
  w.cub_bytes = mc_cub_bytes(n);
  return w;
}

__global__ void __launch_bounds__(256)
mc_classify_kernel(const float* __restrict__ vol, int res, float iso, McWs w) {
  pdl_wait();
  pdl_trigger();
  const long long n = static_cast<long long>(res) * res * res;
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int z = static_cast<int>(p % res);
  const long long t = p / res;
  const int y = static_cast<int>(t % res), x = static_cast<int>(t / res);
  const long long sx = static_cast<long long>(res) * res, sy = res;
  const bool b0 = vol[p] < iso;
  const bool hx = x + 1 < res, hy = y + 1 < res, hz = z + 1 < res;
  const bool bx = hx ? vol[p + sx] < iso : b0;
  const bool by = hy ? vol[p + sy] < iso : b0;
  const bool bz = hz ? vol[p + 1] < iso : b0;
  const unsigned flags = (b0 != bx ? 1u : 0u) | (b0 != by ? 2u : 0u) | (b0 != bz ? 4u : 0u);
  unsigned cube = 0;
  if (hx && hy && hz) {
    cube = (b0 ? 1u : 0u) | (bx ? 2u : 0u) | (vol[p + sx + sy] < iso ? 4u : 0u) | (by ? 8u : 0u) | (bz ? 16u : 0u) |
           (vol[p + sx + 1] < iso ? 32u : 0u) | (vol[p + sx + sy + 1] < iso ? 64u : 0u) | (vol[p + sy + 1] < iso ? 128u : 0u);
  }
  w.vflags[p] = static_cast<unsigned char>(flags);
  w.cube[p] = static_cast<unsigned char>(cube);
  w.vsum[p] = __popc(flags);
  w.tsum[p] = c_mc_cnt[cube];
}

__global__ void mc_totals_kernel(McWs w, long long n, long long* counts) {
  pdl_wait();
  pdl_trigger();
  counts[0] = w.vsum[n - 1];
  counts[1] = w.tsum[n - 1];
}

// vertex id of the crossing on the grid edge (q, axis)
__device__ __forceinline__ int mc_vertex_id(const McWs& w, long long q, int axis) {
  const unsigned f = w.vflags[q];
  return w.vsum[q] - __popc(f) + __popc(f & ((1u << axis) - 1u));
}

__global__ void __launch_bounds__(256)
mc_emit_kernel(const float* __restrict__ vol, int res, float iso, McWs w, float scale_div, float* __restrict__ verts,
               int* __restrict__ tris) {
  pdl_wait();
  pdl_trigger();
  const long long n = static_cast<long long>(res) * res * res;
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const unsigned flags = w.vflags[p];
  const unsigned cube = w.cube[p];
  if (flags == 0 && (cube == 0 || cube == 255)) return;
  const int z = static_cast<int>(p % res);
  const long long t = p / res;
  const int y = static_cast<int>(t % res), x = static_cast<int>(t / res);
  const long long stride[3] = {static_cast<long long>(res) * res, res, 1};
  if (flags) {
    const float v0 = vol[p];
    int vid = w.vsum[p] - __popc(flags);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!((flags >> a) & 1u)) continue;
      const float v1 = vol[p + stride[a]];
      const float tt = __fdiv_rn(__fsub_rn(iso, v0), __fsub_rn(v1, v0));      // linear interpolation, IEEE ops
      float c[3] = {static_cast<float>(x), static_cast<float>(y), static_cast<float>(z)};
      c[a] = __fadd_rn(c[a], tt);
      if (scale_div > 0.f) {                                                   // visualize.py:101: v / res * 2 - 1
#pragma unroll
        for (int k = 0; k < 3; ++k) c[k] = __fadd_rn(__fmul_rn(__fdiv_rn(c[k], scale_div), 2.0f), -1.0f);
      }
      verts[3LL * vid] = c[0]; verts[3LL * vid + 1] = c[1]; verts[3LL * vid + 2] = c[2];
      ++vid;
    }
  }
  const int nt = c_mc_cnt[cube];
  if (nt) {
    long long tbase = 3LL * (w.tsum[p] - nt);
    for (int k = 0; k < 3 * nt; ++k) {
      const int e = c_mc_tri[cube * 16 + k];
      const int cn = c_mc_edge_corner[e];
      const long long q = p + c_mc_cx[cn] * stride[0] + c_mc_cy[cn] * stride[1] + c_mc_cz[cn];
      tris[tbase + k] = mc_vertex_id(w, q, c_mc_edge_axis[e]);
    }
  }
}

// ---- uniform Laplacian smoothing ------------------------------------------------------------------------------
struct SmoothWs {
  int* deg;      // [nv] inclusive scan of the (duplicated) neighbour counts
  int* cursor;   // [nv]
  int* ucnt;     // [nv] unique neighbours
  int* nbr;      // [6 nt]
  double* a;     // [3 nv]
  double* b;     // [3 nv]
  void* cub_tmp;
  size_t cub_bytes;
};
static SmoothWs smooth_layout(void* ws, long long nv, long long nt) {
  char* p = static_cast<char*>(ws);
  SmoothWs s;
  s.deg = reinterpret_cast<int*>(p); p += align256(nv * 4);
  s.cursor = reinterpret_cast<int*>(p); p += align256(nv * 4);
  s.ucnt = reinterpret_cast<int*>(p); p += align256(nv * 4);
  s.nbr = reinterpret_cast<int*>(p); p += align256(nt * 24);
  s.a = reinterpret_cast<double*>(p); p += align256(nv * 24);
  s.b = reinterpret_cast<double*>(p); p += align256(nv * 24);
  s.cub_tmp = p;
  s.cub_bytes = mc_cub_bytes(nv);
  return s;
}

__global__ void __launch_bounds__(256) adj_count_kernel(const int* __restrict__ tris, long long nt, int* deg) {
  pdl_wait();
  pdl_trigger();
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nt) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) atomicAdd(deg + tris[3 * t + k], 2);
}
__global__ void __launch_bounds__(256) adj_fill_kernel(const int* __restrict__ tris, long long nt, SmoothWs s) {
  pdl_wait();
  pdl_trigger();
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const int v[3] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int me = v[k];
    const int begin = me ? s.deg[me - 1] : 0;
    const int pos = atomicAdd(s.cursor + me, 2);
    s.nbr[begin + pos] = v[(k + 1) % 3];
    s.nbr[begin + pos + 1] = v[(k + 2) % 3];
  }
}
// sort each vertex's short neighbour list and drop duplicates: the result does not depend on the atomics' order
__global__ void __launch_bounds__(256) adj_unique_kernel(long long nv, SmoothWs s) {
  pdl_wait();
  pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const int begin = i ? s.deg[i - 1] : 0, len = s.deg[i] - begin;
  int* l = s.nbr + begin;
  for (int a = 1; a < len; ++a) {
    const int key = l[a];
    int b = a - 1;
    while (b >= 0 && l[b] > key) { l[b + 1] = l[b]; --b; }
    l[b + 1] = key;
  }
  int u = 0;
  for (int a = 0; a < len; ++a)
    if (a == 0 || l[a] != l[a - 1]) l[u++] = l[a];
  s.ucnt[i] = u;
}
__global__ void __launch_bounds__(256) smooth_load_kernel(const float* __restrict__ v, long long n3, double* a) {
  pdl_wait();
  pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n3) a[i] = static_cast<double>(v[i]);
}
__global__ void __launch_bounds__(256) smooth_iter_kernel(long long nv, SmoothWs s, const double* __restrict__ src, double* __restrict__ dst) {
  pdl_wait();
  pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const int begin = i ? s.deg[i - 1] : 0, u = s.ucnt[i];
  double x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
  for (int k = 0; k < u; ++k) {
    const long long j = s.nbr[begin + k];
    x += src[3 * j]; y += src[3 * j + 1]; z += src[3 * j + 2];
  }
  const double d = 1.0 + static_cast<double>(u);
  dst[3 * i] = x / d; dst[3 * i + 1] = y / d; dst[3 * i + 2] = z / d;
}
__global__ void __launch_bounds__(256) smooth_store_kernel(const double* __restrict__ a, long long n3, float* v) {
  pdl_wait();
  pdl_trigger();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n3) v[i] = static_cast<float>(a[i]);
}

}  // namespace isb

extern "C" {

size_t isb_mc_workspace_bytes(int res) {
  if (res < 2) return 0;
  const long long n = static_cast<long long>(res) * res * res;
  return 2 * isb::align256(n) + 2 * isb::align256(n * 4) + isb::align256(isb::mc_cub_bytes(n)) + 256;
}

int isb_mc_count(const float* vol, int res, float iso, void* ws, size_t ws_bytes, int64_t* counts, isb_stream_t stream) {
  ISB_CHECK_ARG(vol && ws && counts && res >= 2 && res <= 1024, "isb_mc_count: bad arguments (res=%d)", res);
  ISB_CHECK_ARG(ws_bytes >= isb_mc_workspace_bytes(res), "isb_mc_count: workspace %zu < %zu bytes", ws_bytes, isb_mc_workspace_bytes(res));
  const long long n = static_cast<long long>(res) * res * res;
  isb::McWs w = isb::mc_layout(ws, n);
  cudaStream_t st = isb::as_stream(stream);
  ISB_CUDA(isb::launch(isb::mc_classify_kernel, isb::cdiv(n, 256), 256, 0, st, vol, res, iso, w));
  ISB_LAUNCH_CHECK();
  size_t tb = w.cub_bytes;
  ISB_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.vsum, w.vsum, static_cast<int>(n), st));
  tb = w.cub_bytes;
  ISB_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.tsum, w.tsum, static_cast<int>(n), st));
  ISB_CUDA(isb::launch(isb::mc_totals_kernel, 1, 1, 0, st, w, n, reinterpret_cast<long long*>(counts)));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

int isb_mc_emit(const float* vol, int res, float iso, void* ws, size_t ws_bytes, float scale_div, float* verts,
                int32_t* tris, isb_stream_t stream) {
  ISB_CHECK_ARG(vol && ws && res >= 2 && res <= 1024, "isb_mc_emit: bad arguments");
  ISB_CHECK_ARG(ws_bytes >= isb_mc_workspace_bytes(res), "isb_mc_emit: workspace too small");
  const long long n = static_cast<long long>(res) * res * res;
  isb::McWs w = isb::mc_layout(ws, n);
  ISB_CUDA(isb::launch(isb::mc_emit_kernel, isb::cdiv(n, 256), 256, 0, isb::as_stream(stream), vol, res, iso, w, scale_div,
                       verts, reinterpret_cast<int*>(tris)));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

size_t isb_mesh_smooth_workspace_bytes(int64_t nv, int64_t nt) {
  if (nv <= 0 || nt <= 0) return 256;
  return 3 * isb::align256(nv * 4) + isb::align256(nt * 24) + 2 * isb::align256(nv * 24) + isb::align256(isb::mc_cub_bytes(nv)) + 256;
}

int isb_mesh_smooth_simple(float* verts, int64_t nv, const int32_t* tris, int64_t nt, int iterations, void* ws,
                           size_t ws_bytes, isb_stream_t stream) {
  ISB_CHECK_ARG(iterations >= 0, "isb_mesh_smooth_simple: iterations < 0");
  if (nv <= 0 || nt <= 0 || iterations == 0) return ISB_OK;
  ISB_CHECK_ARG(verts && tris && ws, "isb_mesh_smooth_simple: null pointer");
  ISB_CHECK_ARG(nv < (1LL << 31) && 6 * nt < (1LL << 31), "isb_mesh_smooth_simple: mesh too large for 32-bit indices");
  ISB_CHECK_ARG(ws_bytes >= isb_mesh_smooth_workspace_bytes(nv, nt), "isb_mesh_smooth_simple: workspace too small");
  isb::SmoothWs s = isb::smooth_layout(ws, nv, nt);
  cudaStream_t st = isb::as_stream(stream);
  ISB_CUDA(cudaMemsetAsync(s.deg, 0, nv * 4, st));
  ISB_CUDA(cudaMemsetAsync(s.cursor, 0, nv * 4, st));
  const int tb = isb::cdiv(nt, 256), vb = isb::cdiv(nv, 256), cb = isb::cdiv(3 * nv, 256);
  ISB_CUDA(isb::launch(isb::adj_count_kernel, tb, 256, 0, st, reinterpret_cast<const int*>(tris), static_cast<long long>(nt), s.deg));
  ISB_LAUNCH_CHECK();
  size_t cbytes = s.cub_bytes;
  ISB_CUDA(cub::DeviceScan::InclusiveSum(s.cub_tmp, cbytes, s.deg, s.deg, static_cast<int>(nv), st));
  ISB_CUDA(isb::launch(isb::adj_fill_kernel, tb, 256, 0, st, reinterpret_cast<const int*>(tris), static_cast<long long>(nt), s));
  ISB_LAUNCH_CHECK();
  ISB_CUDA(isb::launch(isb::adj_unique_kernel, vb, 256, 0, st, static_cast<long long>(nv), s));
  ISB_LAUNCH_CHECK();
  ISB_CUDA(isb::launch(isb::smooth_load_kernel, cb, 256, 0, st, static_cast<const float*>(verts), static_cast<long long>(3 * nv), s.a));
  ISB_LAUNCH_CHECK();
  double* src = s.a;
  double* dst = s.b;
  for (int it = 0; it < iterations; ++it) {
    ISB_CUDA(isb::launch(isb::smooth_iter_kernel, vb, 256, 0, st, static_cast<long long>(nv), s, static_cast<const double*>(src), dst));
    ISB_LAUNCH_CHECK();
    double* tmp = src; src = dst; dst = tmp;
  }
  ISB_CUDA(isb::launch(isb::smooth_store_kernel, cb, 256, 0, st, static_cast<const double*>(src), static_cast<long long>(3 * nv), verts));
  ISB_LAUNCH_CHECK();
  return ISB_OK;
}

}  // extern "C"
