// unet.cu — the UNet execution plan behind an opaque handle (the handle-level C ABI of SURVEY.md §8b).
//
// isb_unet_create / isb_unet_load_weight / isb_unet_finalize / isb_unet_forward / isb_unet_backward_input replace
// `UNetModel.__init__` + `load_state_dict` + `forward` (neural_field_diffusion/guided_diffusion/unet.py:396-671) and
// the `loss.backward()` of drag_utils.py:383 down to `img.grad`, for a host that is not Python: the block structure,
// the weight packing (OIHW fp32 -> K-major bf16 panels for the forward AND the flipped/transposed backward-data
// panels), the activation layout in the caller's workspace and the launch schedule all live here.  The Python host
// (guided_diffusion/unet.py `_Plan`) drives the SAME per-operator entry points in the SAME order, so both produce
// bit-identical results (tests/test_gpu_native_unet.py).
//
// Ownership: the handle owns the packed weights (cudaMalloc); every activation, gradient, scratch and split-K
// workspace lives in ONE caller-provided buffer of isb_unet_workspace_bytes() bytes (zero-filled once through
// isb_unet_workspace_init).  All launches go to the caller's stream; nothing synchronises after isb_unet_finalize, so
// forward / backward are CUDA-graph capturable.  One handle serves one stream at a time (it keeps the "which tensors
// carry a gradient" state of the pass in flight).
#include <stdlib.h>
#include <string.h>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "common.cuh"

namespace isb {
namespace un {

#define ISB_TRY(expr)            \
  do {                           \
    const int _rc = (expr);      \
    if (_rc != 0) return _rc;    \
  } while (0)

constexpr size_t kAlign = 256;
static inline size_t esize(int dt) { return dt == ISB_BF16 ? 2 : 4; }

// ---- weight packing kernels ------------------------------------------------------------------------------------
// w [Co][Ci][ks][ks] fp32 (the reference's nn.Conv2d / Conv1d layout) ->
//   forward panel  out[co][col_off + tap*Ci_pad + ci]            (K-major rows per output channel)
//   dgrad panel    out[ci][tap*Co + co] = w[co][ci][taps flipped] (backward-data as a forward conv over dy)
// rows / columns of padded input channels (ci >= Ci) are zero.
__global__ void pack_conv_kernel(const float* __restrict__ w, void* __restrict__ out, int out_dtype, int Co, int Ci,
                                 int kk, int Ci_pad, long long ld, int col_off, int dgrad) {
  const long long total = static_cast<long long>(Co) * kk * Ci_pad;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int co, ci, tap;
  long long dst;
  if (!dgrad) {
    const long long per = static_cast<long long>(kk) * Ci_pad;
    co = static_cast<int>(idx / per);
    const long long rem = idx % per;
    tap = static_cast<int>(rem / Ci_pad);
    ci = static_cast<int>(rem % Ci_pad);
    dst = co * ld + col_off + rem;
  } else {
    const long long per = static_cast<long long>(kk) * Co;
    ci = static_cast<int>(idx / per);
    const long long rem = idx % per;
    const int t = static_cast<int>(rem / Co);
    co = static_cast<int>(rem % Co);
    tap = kk - 1 - t;
    dst = ci * ld + col_off + rem;
  }
  const float v = ci < Ci ? w[(static_cast<long long>(co) * Ci + ci) * kk + tap] : 0.0f;
  if (out_dtype == ISB_BF16) reinterpret_cast<__nv_bfloat16*>(out)[dst] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(out)[dst] = v;
}
__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// ---- plan data -------------------------------------------------------------------------------------------------
struct Raw {                 // a parameter as the host handed it over (fp32, device, handle-owned copy)
  float* p = nullptr;
  std::vector<int64_t> shape;
  size_t numel = 0;
};

struct Tens {                // a fp32 NHWC "stream" tensor (block input / output) and its gradient buffers
  size_t val = 0, grad = 0, grad_lo = 0, gn_part = 0;
  int N = 0, H = 0, W = 0, C = 0;
  int gn_slots = 0;          // what the producing conv can deliver (0: nothing)
  bool has_part = false;     // a consumer asked for the statistics: the producer fills gn_part
  bool has_grad = false;     // run-time state of the backward pass in flight
  size_t numel() const { return static_cast<size_t>(N) * H * W * C; }
};

enum Tag { T_A = 0, T_G, T_G2, T_GLO, T_GOLO, T_O, T_FADELTA, T_PTMP, T_COUNT };

struct Run {                 // one call's context
  char* base;
  cudaStream_t st;
  bool dry;                  // sizing pass: only query the conv workspace
  int slot;                  // conv workspace slot: 0 main, 1 forward tail, 2 backward side stream
  size_t conv_ws_need;
};

struct Plan;
struct Layer {
  std::string name;
  Tens* out = nullptr;
  virtual ~Layer() {}
  virtual int forward(Plan& p, Run& r) = 0;
  virtual int backward(Plan& p, Run& r) = 0;
};

struct Spec {                // one module of a TimestepEmbedSequential
  bool attn = false;
  int cin = 0, cout = 0, heads = 0;
  bool up = false, down = false;
};

struct Plan {
  isb_unet_cfg cfg;
  int lo = ISB_BF16;
  bool finalized = false;
  std::map<std::string, Raw> raw;
  std::vector<void*> owned;                 // packed panels and small fp32 parameter copies
  std::vector<std::unique_ptr<Layer>> layers;
  std::vector<std::unique_ptr<Tens>> tensors;
  std::vector<Tens*> block_out;
  Tens* h0 = nullptr;
  Tens* h_last = nullptr;
  // workspace layout
  size_t top = 0;
  size_t tag_bytes[T_COUNT] = {0}, tag_off[T_COUNT] = {0};
  size_t gn_scratch_off[2] = {0, 0};        // forward / backward kernels (they may run concurrently)
  size_t conv_ws_off[3] = {0, 0, 0}, conv_ws_bytes = 0;
  size_t ws_bytes = 0;
  // time embedding
  int film_rows = 0, hidden = 0, cin_pad = 0;
  float *te_w1 = nullptr, *te_b1 = nullptr, *te_w2 = nullptr, *te_b2 = nullptr, *w_all = nullptr, *b_all = nullptr,
        *freqs = nullptr;
  size_t te_scratch = 0, film_all = 0;
  // input conv / out layer
  void *w_in = nullptr, *w_in_d = nullptr, *w_out = nullptr, *w_out_d = nullptr;
  float *b_in = nullptr, *b_out = nullptr, *out_g = nullptr, *out_b = nullptr;
  size_t x_lo = 0, out_stats = 0, out_nhwc = 0;
  // run-time state
  size_t tail_from = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  ~Plan() {
    for (auto& kv : raw)
      if (kv.second.p) cudaFree(kv.second.p);
    for (void* p : owned) cudaFree(p);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
  }

  size_t alloc(size_t bytes) {
    const size_t off = top;
    top += (bytes + kAlign - 1) / kAlign * kAlign;
    return off;
  }
  void req(Tag t, size_t bytes) {
    if (bytes > tag_bytes[t]) tag_bytes[t] = bytes;
  }
  Tens* new_tens(int N, int H, int W, int C, int gn_slots) {
    tensors.emplace_back(new Tens());
    Tens* t = tensors.back().get();
    t->N = N; t->H = H; t->W = W; t->C = C; t->gn_slots = gn_slots;
    t->val = alloc(t->numel() * 4);
    if (cfg.want_backward) {
      t->grad = alloc(t->numel() * 4);
      t->grad_lo = alloc(t->numel() * esize(lo));
    }
    return t;
  }
  bool want_gn_part(Tens* t) {     // the consumer's GroupNorm reads this tensor alone: switch the producer's statistics on
    if (t->gn_slots <= 0) return false;
    if (!t->has_part) {
      t->gn_part = alloc(static_cast<size_t>(t->N) * 32 * t->gn_slots * 2 * 4);
      t->has_part = true;
    }
    return true;
  }
  template <typename T = void>
  T* at(const Run& r, size_t off) const { return reinterpret_cast<T*>(r.base + off); }
  template <typename T = void>
  T* scratch(const Run& r, Tag t) const { return reinterpret_cast<T*>(r.base + tag_off[t]); }

  int conv_gn_slots(int N, int H, int W, int Cin, int ks, int Cout, int Cin2 = 0) const {
    static const bool off = [] {
      const char* e = getenv("ISB_GN_FUSE");
      return e != nullptr && atoi(e) == 0;
    }();
    if (lo != ISB_BF16 || Cout % 32 != 0 || off) return 0;
    isb_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.a = d.w = reinterpret_cast<const void*>(256);
    d.out = reinterpret_cast<void*>(256);
    d.a_dtype = ISB_BF16;
    d.out_dtype = ISB_F32;
    d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.ksize = ks; d.Cout = Cout;
    if (Cin2) { d.a2 = reinterpret_cast<const void*>(256); d.Cin2 = Cin2; }
    d.gn_cg = Cout / 32;
    return isb_conv2d_gn_slots(&d);
  }

  // out = conv(a (+ a2 through the appended 1x1 columns)) + bias + residual; statistics of `out` to gn_part
  int conv(Run& r, const void* a, int N, int H, int W, int Cin, int ks, const void* w, const float* bias, void* out,
           int out_dtype, int Cout, const void* a2 = nullptr, int Cin2 = 0, const float* residual = nullptr,
           const Tens* stats_of = nullptr, float* part = nullptr, int part_slots = 0) {
    isb_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.a = a; d.a_dtype = lo; d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.ksize = ks;
    d.a2 = a2; d.Cin2 = Cin2;
    d.w = w; d.bias = bias; d.residual = residual;
    d.out = out; d.out_dtype = out_dtype; d.Cout = Cout;
    if (stats_of != nullptr && stats_of->has_part) {
      d.gn_partials = at<float>(r, stats_of->gn_part);
      d.gn_cg = Cout / 32;
      d.gn_slots = stats_of->gn_slots;
    } else if (part != nullptr) {
      d.gn_partials = part;
      d.gn_cg = Cout / 32;
      d.gn_slots = part_slots;
    }
    if (r.dry) {
      const size_t need = isb_conv2d_workspace(&d);
      if (need > r.conv_ws_need) r.conv_ws_need = need;
      return 0;
    }
    return isb_conv2d(&d, conv_ws_bytes ? r.base + conv_ws_off[r.slot] : nullptr, conv_ws_bytes, r.st);
  }

  void gn_desc(isb_gn_desc& d, const float* x1, int C1, const float* x2, int C2, int N, int H, int W,
               const float* gamma, const float* beta, const float* film, int silu, int resample, float* stats) const {
    memset(&d, 0, sizeof(d));
    d.x1 = x1; d.C1 = C1; d.x2 = x2; d.C2 = C2;
    d.N = N; d.H = H; d.W = W;
    d.groups = 32; d.eps = 1e-5f;
    d.gamma = gamma; d.beta = beta;
    d.film = film; d.film_stride = film ? film_rows : 0;
    d.silu = silu; d.resample = resample;
    d.stats = stats;
  }
  int gn_forward(Run& r, isb_gn_desc& d) {
    if (r.dry) return 0;
    return isb_gn_forward(&d, r.base + gn_scratch_off[0], r.st);
  }
  int gn_backward(Run& r, isb_gn_bwd_desc& b) {
    if (r.dry) return 0;
    return isb_gn_backward(&b, r.base + gn_scratch_off[1], r.st);
  }
  int cast_lo(Run& r, const float* src, void* dst, size_t n) {
    if (r.dry) return 0;
    if (lo == ISB_BF16) return isb_cast_f32_bf16(src, dst, n, r.st);
    ISB_CUDA(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, r.st));
    return 0;
  }

  // -- parameters ------------------------------------------------------------------------------------------------
  const Raw* find(const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = raw.find(name);
    if (it == raw.end()) {
      set_error("isb_unet_finalize: parameter '%s' was not loaded", name.c_str());
      return nullptr;
    }
    const Raw& w = it->second;
    size_t n = 1;
    for (int64_t s : shape) n *= static_cast<size_t>(s);
    if (w.numel != n) {
      set_error("isb_unet_finalize: parameter '%s' has %zu elements, the configuration needs %zu", name.c_str(),
                w.numel, n);
      return nullptr;
    }
    return &w;
  }
  // fp32 parameter used as is (GroupNorm affine, biases, Linear weights): ownership moves to `owned`
  float* take(const std::string& name, std::initializer_list<int64_t> shape) {
    const Raw* w = find(name, shape);
    if (!w) return nullptr;
    float* p = w->p;
    owned.push_back(p);
    raw[name].p = nullptr;
    return p;
  }
  void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      set_error("isb_unet_finalize: cudaMalloc(%zu) failed", bytes);
      return nullptr;
    }
    owned.push_back(p);
    return p;
  }
  // conv weight -> packed panel; `extra` (a 1x1 weight [Co][Ci2]) is appended to the K axis of the forward panel
  void* pack(const std::string& name, int Co, int Ci, int ks, int Ci_pad, bool dgrad, cudaStream_t st,
             const std::string* extra = nullptr, int Ci2 = 0) {
    const int kk = ks * ks;
    const Raw* w = find(name, {Co, Ci, kk});
    if (!w) return nullptr;
    const long long rows = dgrad ? Ci_pad : Co;
    const long long ld = dgrad ? static_cast<long long>(kk) * Co : static_cast<long long>(kk) * Ci_pad + Ci2;
    void* out = dev_alloc(static_cast<size_t>(rows) * ld * esize(lo));
    if (!out) return nullptr;
    const long long total = static_cast<long long>(Co) * kk * Ci_pad;
    pack_conv_kernel<<<cdiv(total, 256), 256, 0, st>>>(w->p, out, lo, Co, Ci, kk, Ci_pad, ld, 0, dgrad ? 1 : 0);
    count_launch();
    if (extra != nullptr) {
      const Raw* e = find(*extra, {Co, Ci2});
      if (!e) return nullptr;
      const long long t2 = static_cast<long long>(Co) * Ci2;
      pack_conv_kernel<<<cdiv(t2, 256), 256, 0, st>>>(e->p, out, lo, Co, Ci2, 1, Ci2, ld, kk * Ci_pad, 0);
      count_launch();
    }
    if (cudaPeekAtLastError() != cudaSuccess) {
      set_error("isb_unet_finalize: packing '%s' failed: %s", name.c_str(), cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
    return out;
  }
};

// ---- ResBlock (unet.py:143-256) -----------------------------------------------------------------------------------
struct ResLayer : Layer {
  Tens *s1 = nullptr, *s2 = nullptr;
  int N, H, W, Cin, Ho, Wo, Co, resample = 0, film_off = 0;
  bool has_skip = false, x_part = false;
  float *g1, *be1, *g2, *be2, *b1, *b2;
  void *w1, *w2, *w1_d = nullptr, *w2_d = nullptr, *wskip_d = nullptr;
  size_t stats1, stats2, xraw = 0, xres = 0, h1, h1_part = 0;
  int h1_slots = 0;

  int build(Plan& p, const Spec& sp, Tens* a, Tens* b, int film_offset, cudaStream_t st) {
    s1 = a; s2 = b;
    N = a->N; H = a->H; W = a->W;
    Cin = a->C + (b ? b->C : 0);
    ISB_CHECK_ARG(Cin == sp.cin, "%s: got %d input channels, the block expects %d", name.c_str(), Cin, sp.cin);
    Co = sp.cout;
    resample = sp.down ? 1 : (sp.up ? 2 : 0);
    Ho = sp.down ? H / 2 : (sp.up ? H * 2 : H);
    Wo = sp.down ? W / 2 : (sp.up ? W * 2 : W);
    has_skip = Cin != Co;
    ISB_CHECK_ARG(!(has_skip && resample), "%s: up/down ResBlocks keep the channel count", name.c_str());
    ISB_CHECK_ARG(b == nullptr || has_skip, "%s: two sources need a skip convolution", name.c_str());
    film_off = film_offset;
    g1 = p.take(name + ".in_layers.0.weight", {Cin});
    be1 = p.take(name + ".in_layers.0.bias", {Cin});
    g2 = p.take(name + ".out_layers.0.weight", {Co});
    be2 = p.take(name + ".out_layers.0.bias", {Co});
    b1 = p.take(name + ".in_layers.2.bias", {Co});
    b2 = p.take(name + ".out_layers.3.bias", {Co});
    if (!g1 || !be1 || !g2 || !be2 || !b1 || !b2) return ISB_ERR_ARG;
    w1 = p.pack(name + ".in_layers.2.weight", Co, Cin, 3, Cin, false, st);
    const std::string skip_w = name + ".skip_connection.weight";
    w2 = p.pack(name + ".out_layers.3.weight", Co, Co, 3, Co, false, st, has_skip ? &skip_w : nullptr,
                has_skip ? Cin : 0);
    if (!w1 || !w2) return ISB_ERR_ARG;
    if (has_skip) {      // the skip conv shares conv2's accumulator: its bias is folded into conv2's
      const float* bs = p.take(name + ".skip_connection.bias", {Co});
      float* sum = static_cast<float*>(p.dev_alloc(static_cast<size_t>(Co) * 4));
      if (!bs || !sum) return ISB_ERR_ARG;
      add_vec_kernel<<<cdiv(Co, 256), 256, 0, st>>>(b2, bs, sum, Co);
      count_launch();
      b2 = sum;
    }
    if (p.cfg.want_backward) {
      w1_d = p.pack(name + ".in_layers.2.weight", Co, Cin, 3, Cin, true, st);
      w2_d = p.pack(name + ".out_layers.3.weight", Co, Co, 3, Co, true, st);
      if (!w1_d || !w2_d) return ISB_ERR_ARG;
      if (has_skip) {
        wskip_d = p.pack(skip_w, Co, Cin, 1, Cin, true, st);
        if (!wskip_d) return ISB_ERR_ARG;
      }
    }
    stats1 = p.alloc(static_cast<size_t>(N) * 32 * 2 * 4);
    stats2 = p.alloc(static_cast<size_t>(N) * 32 * 2 * 4);
    const size_t px_in = static_cast<size_t>(N) * H * W, px_out = static_cast<size_t>(N) * Ho * Wo;
    p.req(T_A, px_out * Cin * esize(p.lo));
    if (has_skip) xraw = p.alloc(px_in * Cin * esize(p.lo));
    if (resample) xres = p.alloc(px_out * Cin * 4);
    h1 = p.alloc(px_out * Co * 4);
    h1_slots = p.conv_gn_slots(N, Ho, Wo, Cin, 3, Co);     // GN2 reads conv1's output: statistics from its epilogue
    if (h1_slots > 0) h1_part = p.alloc(static_cast<size_t>(N) * 32 * h1_slots * 2 * 4);
    p.req(T_A, px_out * Co * esize(p.lo));
    out = p.new_tens(N, Ho, Wo, Co, p.conv_gn_slots(N, Ho, Wo, Co, 3, Co, has_skip ? Cin : 0));
    // GN1 over a single, un-resampled source whose producer can deliver the statistics
    x_part = (b == nullptr && resample == 0) ? p.want_gn_part(a) : false;
    if (p.cfg.want_backward) {
      p.req(T_G, px_out * Co * 4);
      p.req(T_G, px_out * Cin * 4);
      p.req(T_GLO, px_out * Co * esize(p.lo));
      if (has_skip) p.req(T_G2, px_in * Cin * 4);
    }
    return 0;
  }

  int forward(Plan& p, Run& r) override {
    const float* film = p.at<float>(r, p.film_all) + film_off;
    const float* x1 = p.at<float>(r, s1->val);
    const float* x2 = s2 ? p.at<float>(r, s2->val) : nullptr;
    void* a = p.scratch(r, T_A);
    isb_gn_desc d;
    p.gn_desc(d, x1, s1->C, x2, s2 ? s2->C : 0, N, H, W, g1, be1, nullptr, 1, resample, p.at<float>(r, stats1));
    if (x_part) { d.partials = p.at<float>(r, s1->gn_part); d.partial_slots = s1->gn_slots; }
    d.y = a; d.y_dtype = p.lo;
    if (has_skip) { d.raw = p.at(r, xraw); d.raw_dtype = p.lo; }
    if (resample) d.xres = p.at<float>(r, xres);
    ISB_TRY(p.gn_forward(r, d));
    ISB_TRY(p.conv(r, a, N, Ho, Wo, Cin, 3, w1, b1, p.at(r, h1), ISB_F32, Co, nullptr, 0, nullptr, nullptr,
                   h1_slots > 0 ? p.at<float>(r, h1_part) : nullptr, h1_slots));
    p.gn_desc(d, p.at<float>(r, h1), Co, nullptr, 0, N, Ho, Wo, g2, be2, film, 1, 0, p.at<float>(r, stats2));
    if (h1_slots > 0) { d.partials = p.at<float>(r, h1_part); d.partial_slots = h1_slots; }
    d.y = a; d.y_dtype = p.lo;
    ISB_TRY(p.gn_forward(r, d));
    if (has_skip)
      return p.conv(r, a, N, Ho, Wo, Co, 3, w2, b2, p.at(r, out->val), ISB_F32, Co, p.at(r, xraw), Cin, nullptr, out);
    return p.conv(r, a, N, Ho, Wo, Co, 3, w2, b2, p.at(r, out->val), ISB_F32, Co, nullptr, 0,
                  resample ? p.at<float>(r, xres) : x1, out);
  }

  int backward(Plan& p, Run& r) override {
    const float* film = p.at<float>(r, p.film_all) + film_off;
    float* g_out = p.at<float>(r, out->grad);
    void* g_out_lo = p.at(r, out->grad_lo);
    float* g_a2 = p.scratch<float>(r, T_G);
    const float* gres = g_out;
    int at_input = 0;
    bool joined = false;
    if (has_skip) {      // d/dx through the 1x1 skip: only needs g_out, runs beside the conv2 -> GN2 -> conv1 chain
      float* g2buf = p.scratch<float>(r, T_G2);
      if (p.side != nullptr && !r.dry) {
        ISB_CUDA(cudaEventRecord(p.ev_fork, r.st));
        ISB_CUDA(cudaStreamWaitEvent(p.side, p.ev_fork, 0));
        Run rs = r;
        rs.st = p.side;
        rs.slot = 2;
        ISB_TRY(p.conv(rs, g_out_lo, N, H, W, Co, 1, wskip_d, nullptr, g2buf, ISB_F32, Cin));
        ISB_CUDA(cudaEventRecord(p.ev_join, p.side));
        joined = true;
      } else {
        ISB_TRY(p.conv(r, g_out_lo, N, H, W, Co, 1, wskip_d, nullptr, g2buf, ISB_F32, Cin));
      }
      gres = g2buf;
      at_input = 1;
    }
    ISB_TRY(p.conv(r, g_out_lo, N, Ho, Wo, Co, 3, w2_d, nullptr, g_a2, ISB_F32, Co));
    void* g_h1_lo = p.scratch(r, T_GLO);
    isb_gn_bwd_desc b;
    memset(&b, 0, sizeof(b));
    p.gn_desc(b.f, p.at<float>(r, h1), Co, nullptr, 0, N, Ho, Wo, g2, be2, film, 1, 0, p.at<float>(r, stats2));
    b.dy = g_a2;
    b.gx1_lo = g_h1_lo;
    b.lo_dtype = p.lo;
    ISB_TRY(p.gn_backward(r, b));
    float* g_a1 = p.scratch<float>(r, T_G);
    ISB_TRY(p.conv(r, g_h1_lo, N, Ho, Wo, Co, 3, w1_d, nullptr, g_a1, ISB_F32, Cin));
    if (joined) ISB_CUDA(cudaStreamWaitEvent(r.st, p.ev_join, 0));
    memset(&b, 0, sizeof(b));
    p.gn_desc(b.f, p.at<float>(r, s1->val), s1->C, s2 ? p.at<float>(r, s2->val) : nullptr, s2 ? s2->C : 0, N, H, W,
              g1, be1, nullptr, 1, resample, p.at<float>(r, stats1));
    b.dy = g_a1;
    b.gres = gres;
    b.gres_at_input = at_input;
    b.gx1 = p.at<float>(r, s1->grad); b.acc1 = s1->has_grad; b.gx1_lo = p.at(r, s1->grad_lo);
    if (s2) { b.gx2 = p.at<float>(r, s2->grad); b.acc2 = s2->has_grad; b.gx2_lo = p.at(r, s2->grad_lo); }
    b.lo_dtype = p.lo;
    ISB_TRY(p.gn_backward(r, b));
    s1->has_grad = true;
    if (s2) s2->has_grad = true;
    return 0;
  }
};

// ---- AttentionBlock + QKVAttentionLegacy (unet.py:259-354) ---------------------------------------------------------
struct AttnLayer : Layer {
  Tens* src = nullptr;
  int N, H, W, C, T, heads;
  bool flash = false, x_part = false;
  float *g, *be, *bqkv, *bproj;
  void *wqkv, *wproj, *wqkv_d = nullptr, *wproj_d = nullptr;
  size_t stats, qkv, lse = 0, o = 0, probs = 0;
  bool o_scratch = false;

  int build(Plan& p, const Spec& sp, Tens* s, cudaStream_t st) {
    src = s;
    N = s->N; H = s->H; W = s->W; C = s->C; T = H * W;
    ISB_CHECK_ARG(C == sp.cin, "%s: got %d channels, the block expects %d", name.c_str(), C, sp.cin);
    heads = sp.heads;
    g = p.take(name + ".norm.weight", {C});
    be = p.take(name + ".norm.bias", {C});
    bqkv = p.take(name + ".qkv.bias", {3 * C});
    bproj = p.take(name + ".proj_out.bias", {C});
    if (!g || !be || !bqkv || !bproj) return ISB_ERR_ARG;
    wqkv = p.pack(name + ".qkv.weight", 3 * C, C, 1, C, false, st);
    wproj = p.pack(name + ".proj_out.weight", C, C, 1, C, false, st);
    if (!wqkv || !wproj) return ISB_ERR_ARG;
    if (p.cfg.want_backward) {
      wqkv_d = p.pack(name + ".qkv.weight", 3 * C, C, 1, C, true, st);
      wproj_d = p.pack(name + ".proj_out.weight", C, C, 1, C, true, st);
      if (!wqkv_d || !wproj_d) return ISB_ERR_ARG;
    }
    const size_t px = static_cast<size_t>(N) * T;
    stats = p.alloc(static_cast<size_t>(N) * 32 * 2 * 4);
    p.req(T_A, px * C * esize(p.lo));
    // bf16 mode with 64-channel heads: fused attention, the [heads,T,T] probabilities are never materialised
    flash = p.lo == ISB_BF16 && C / heads == 64 && C % heads == 0 && T % 64 == 0;
    if (flash) {
      qkv = p.alloc(px * 3 * C * 2);
      lse = p.alloc(static_cast<size_t>(N) * heads * T * 4);
      if (p.cfg.want_backward) o = p.alloc(px * C * 2);
      else { o_scratch = true; p.req(T_O, px * C * 2); }
    } else {
      qkv = p.alloc(px * 3 * C * 4);
      probs = p.alloc(static_cast<size_t>(N) * heads * T * T * 4);
      o_scratch = true;
      p.req(T_O, px * C * esize(p.lo));
    }
    out = p.new_tens(N, H, W, C, p.conv_gn_slots(N, H, W, C, 1, C));
    x_part = p.want_gn_part(s);
    if (p.cfg.want_backward) {
      p.req(T_GLO, px * 3 * C * esize(p.lo));
      p.req(T_G, px * C * 4);
      if (flash) { p.req(T_GOLO, px * C * 2); p.req(T_FADELTA, static_cast<size_t>(N) * heads * T * 4); }
      else p.req(T_PTMP, static_cast<size_t>(N) * heads * T * T * 4);
    }
    return 0;
  }

  int forward(Plan& p, Run& r) override {
    const float* x = p.at<float>(r, src->val);
    void* a = p.scratch(r, T_A);
    void* ob = o_scratch ? p.scratch(r, T_O) : p.at(r, o);
    isb_gn_desc d;
    p.gn_desc(d, x, C, nullptr, 0, N, H, W, g, be, nullptr, 0, 0, p.at<float>(r, stats));
    if (x_part) { d.partials = p.at<float>(r, src->gn_part); d.partial_slots = src->gn_slots; }
    d.y = a; d.y_dtype = p.lo;
    ISB_TRY(p.gn_forward(r, d));
    ISB_TRY(p.conv(r, a, N, H, W, C, 1, wqkv, bqkv, p.at(r, qkv), flash ? ISB_BF16 : ISB_F32, 3 * C));
    if (!r.dry) {
      if (flash) ISB_TRY(isb_attention_flash_forward(p.at(r, qkv), N, T, heads, C / heads, ob, p.at<float>(r, lse), r.st));
      else ISB_TRY(isb_attention_forward(p.at<float>(r, qkv), N, T, heads, C / heads, p.at<float>(r, probs), ob, p.lo, r.st));
    }
    return p.conv(r, ob, N, H, W, C, 1, wproj, bproj, p.at(r, out->val), ISB_F32, C, nullptr, 0, x, out);
  }

  int backward(Plan& p, Run& r) override {
    void* g_qkv = p.scratch(r, T_GLO);
    if (flash) {
      void* g_o = p.scratch(r, T_GOLO);
      ISB_TRY(p.conv(r, p.at(r, out->grad_lo), N, H, W, C, 1, wproj_d, nullptr, g_o, ISB_BF16, C));
      if (!r.dry)
        ISB_TRY(isb_attention_flash_backward(p.at(r, qkv), p.at(r, o), g_o, p.at<float>(r, lse), N, T, heads, C / heads,
                                             p.scratch<float>(r, T_FADELTA), g_qkv, r.st));
    } else {
      float* g_o = p.scratch<float>(r, T_G);
      ISB_TRY(p.conv(r, p.at(r, out->grad_lo), N, H, W, C, 1, wproj_d, nullptr, g_o, ISB_F32, C));
      if (!r.dry)
        ISB_TRY(isb_attention_backward(p.at<float>(r, qkv), p.at<float>(r, probs), g_o, N, T, heads, C / heads,
                                       p.scratch<float>(r, T_PTMP), g_qkv, p.lo, r.st));
    }
    float* g_a = p.scratch<float>(r, T_G);
    ISB_TRY(p.conv(r, g_qkv, N, H, W, 3 * C, 1, wqkv_d, nullptr, g_a, ISB_F32, C));
    isb_gn_bwd_desc b;
    memset(&b, 0, sizeof(b));
    p.gn_desc(b.f, p.at<float>(r, src->val), C, nullptr, 0, N, H, W, g, be, nullptr, 0, 0, p.at<float>(r, stats));
    b.dy = g_a;
    b.gres = p.at<float>(r, out->grad);       // the residual `x + proj_out(...)`
    b.gres_at_input = 0;
    b.gx1 = p.at<float>(r, src->grad); b.acc1 = src->has_grad; b.gx1_lo = p.at(r, src->grad_lo);
    b.lo_dtype = p.lo;
    ISB_TRY(p.gn_backward(r, b));
    src->has_grad = true;
    return 0;
  }
};

// ---- the block structure of UNetModel.__init__ (unet.py:477-616) ----------------------------------------------------
static bool in_list(const int* v, int n, int x) {
  for (int i = 0; i < n; ++i)
    if (v[i] == x) return true;
  return false;
}

struct Structure {
  std::vector<std::vector<Spec>> input_blocks, output_blocks;   // input_blocks[0] is the plain conv (empty spec list)
  std::vector<Spec> middle;
  int ch_last = 0, input_ch = 0;
};

static int attn_heads(const isb_unet_cfg& c, int ch, int heads) {
  return c.num_head_channels == -1 ? heads : ch / c.num_head_channels;
}

static Structure make_structure(const isb_unet_cfg& c) {
  Structure s;
  const int mc = c.model_channels;
  const int heads_up = c.num_heads_upsample == -1 ? c.num_heads : c.num_heads_upsample;
  int ch = c.channel_mult[0] * mc;
  s.input_ch = ch;
  s.input_blocks.push_back({});
  std::vector<int> skip{ch};
  int ds = 1;
  auto res = [](int cin, int cout, bool up = false, bool down = false) {
    Spec sp; sp.cin = cin; sp.cout = cout; sp.up = up; sp.down = down; return sp;
  };
  auto attn = [&](int chn, int heads) {
    Spec sp; sp.attn = true; sp.cin = sp.cout = chn; sp.heads = attn_heads(c, chn, heads); return sp;
  };
  for (int level = 0; level < c.n_levels; ++level) {
    const int mult = c.channel_mult[level];
    for (int i = 0; i < c.num_res_blocks; ++i) {
      std::vector<Spec> seq{res(ch, mult * mc)};
      ch = mult * mc;
      if (in_list(c.attention_ds, c.n_attn, ds)) seq.push_back(attn(ch, c.num_heads));
      s.input_blocks.push_back(seq);
      skip.push_back(ch);
    }
    if (level != c.n_levels - 1) {
      s.input_blocks.push_back({res(ch, ch, false, true)});
      skip.push_back(ch);
      ds *= 2;
    }
  }
  s.middle = {res(ch, ch), attn(ch, c.num_heads), res(ch, ch)};
  for (int level = c.n_levels - 1; level >= 0; --level) {
    const int mult = c.channel_mult[level];
    for (int i = 0; i <= c.num_res_blocks; ++i) {
      const int ich = skip.back();
      skip.pop_back();
      std::vector<Spec> seq{res(ch + ich, mc * mult)};
      ch = mc * mult;
      if (in_list(c.attention_ds, c.n_attn, ds)) seq.push_back(attn(ch, heads_up));
      if (level && i == c.num_res_blocks) {
        seq.push_back(res(ch, ch, true, false));
        ds /= 2;
      }
      s.output_blocks.push_back(seq);
    }
  }
  s.ch_last = ch;
  return s;
}

static int check_cfg(const isb_unet_cfg& c) {
  ISB_CHECK_ARG(c.in_channels > 0 && c.model_channels > 0 && c.out_channels > 0 && c.num_res_blocks > 0,
                "isb_unet_create: channels / num_res_blocks must be positive");
  ISB_CHECK_ARG(c.model_channels % 32 == 0 && c.model_channels % 2 == 0, "isb_unet_create: model_channels %% 32 != 0");
  ISB_CHECK_ARG(c.n_levels >= 1 && c.n_levels <= 8 && c.n_attn >= 0 && c.n_attn <= 8, "isb_unet_create: n_levels / n_attn");
  ISB_CHECK_ARG(c.mode == ISB_BF16 || c.mode == ISB_F32, "isb_unet_create: mode must be ISB_BF16 or ISB_F32");
  ISB_CHECK_ARG(c.N >= 1 && c.N * 32 <= 4096, "isb_unet_create: batch N");
  const int down = 1 << (c.n_levels - 1);
  ISB_CHECK_ARG(c.H > 0 && c.W > 0 && c.H % down == 0 && c.W % down == 0,
                "isb_unet_create: H, W must be multiples of %d", down);
  ISB_CHECK_ARG(c.num_head_channels == -1 || c.num_head_channels > 0, "isb_unet_create: num_head_channels");
  ISB_CHECK_ARG(c.num_head_channels != -1 || c.num_heads > 0, "isb_unet_create: num_heads");
  ISB_CHECK_ARG(c.out_channels % 8 == 0, "isb_unet_create: out_channels %% 8 != 0");
  return 0;
}

static int finalize(Plan& p, cudaStream_t st) {
  const isb_unet_cfg& c = p.cfg;
  const Structure s = make_structure(c);
  const int N = c.N, H = c.H, W = c.W, mc = c.model_channels;
  p.hidden = 4 * mc;
  p.cin_pad = (c.in_channels + 63) / 64 * 64;

  // --- timestep-embedding path (unet.py:471-475 + every ResBlock's emb_layers Linear, row-concatenated) ---
  p.te_w1 = p.take("time_embed.0.weight", {p.hidden, mc});
  p.te_b1 = p.take("time_embed.0.bias", {p.hidden});
  p.te_w2 = p.take("time_embed.2.weight", {p.hidden, p.hidden});
  p.te_b2 = p.take("time_embed.2.bias", {p.hidden});
  if (!p.te_w1 || !p.te_b1 || !p.te_w2 || !p.te_b2) return ISB_ERR_ARG;
  if (p.raw.count("time_embed.freqs")) {       // the host's own table (bit-identical embeddings with that host)
    p.freqs = p.take("time_embed.freqs", {mc / 2});
    if (!p.freqs) return ISB_ERR_ARG;
  } else {                                     // nn.py:113-115
    std::vector<float> f(mc / 2);
    for (int i = 0; i < mc / 2; ++i)
      f[i] = static_cast<float>(exp(-log(10000.0) * static_cast<double>(static_cast<float>(i)) / (mc / 2)));
    p.freqs = static_cast<float*>(p.dev_alloc(f.size() * 4));
    if (!p.freqs) return ISB_ERR_ARG;
    ISB_CUDA(cudaMemcpyAsync(p.freqs, f.data(), f.size() * 4, cudaMemcpyHostToDevice, st));
    ISB_CUDA(cudaStreamSynchronize(st));
  }
  std::vector<std::pair<std::string, int>> res_names;      // (module name, out channels) in module order
  auto collect = [&](const std::vector<Spec>& seq, const std::string& prefix) {
    for (size_t li = 0; li < seq.size(); ++li)
      if (!seq[li].attn) res_names.emplace_back(prefix + "." + std::to_string(li), seq[li].cout);
  };
  for (size_t i = 1; i < s.input_blocks.size(); ++i) collect(s.input_blocks[i], "input_blocks." + std::to_string(i));
  collect(s.middle, "middle_block");
  for (size_t i = 0; i < s.output_blocks.size(); ++i) collect(s.output_blocks[i], "output_blocks." + std::to_string(i));
  std::map<std::string, int> film_offs;
  int rows = 0;
  for (auto& rn : res_names) { film_offs[rn.first] = rows; rows += 2 * rn.second; }
  p.film_rows = rows;
  p.w_all = static_cast<float*>(p.dev_alloc(static_cast<size_t>(rows) * p.hidden * 4));
  p.b_all = static_cast<float*>(p.dev_alloc(static_cast<size_t>(rows) * 4));
  if (!p.w_all || !p.b_all) return ISB_ERR_ARG;
  for (auto& rn : res_names) {
    const Raw* w = p.find(rn.first + ".emb_layers.1.weight", {2 * rn.second, p.hidden});
    const Raw* b = p.find(rn.first + ".emb_layers.1.bias", {2 * rn.second});
    if (!w || !b) return ISB_ERR_ARG;
    const int off = film_offs[rn.first];
    ISB_CUDA(cudaMemcpyAsync(p.w_all + static_cast<size_t>(off) * p.hidden, w->p, w->numel * 4, cudaMemcpyDeviceToDevice, st));
    ISB_CUDA(cudaMemcpyAsync(p.b_all + off, b->p, b->numel * 4, cudaMemcpyDeviceToDevice, st));
  }
  p.te_scratch = p.alloc(static_cast<size_t>(N) * (mc + 2 * p.hidden) * 4);
  p.film_all = p.alloc(static_cast<size_t>(N) * rows * 4);

  // --- input conv (input_blocks[0], unet.py:482): input channels zero-padded to a multiple of 64 ---
  p.w_in = p.pack("input_blocks.0.0.weight", s.input_ch, c.in_channels, 3, p.cin_pad, false, st);
  p.b_in = p.take("input_blocks.0.0.bias", {s.input_ch});
  if (!p.w_in || !p.b_in) return ISB_ERR_ARG;
  if (c.want_backward) {
    p.w_in_d = p.pack("input_blocks.0.0.weight", s.input_ch, c.in_channels, 3, p.cin_pad, true, st);
    if (!p.w_in_d) return ISB_ERR_ARG;
  }
  p.x_lo = p.alloc(static_cast<size_t>(N) * H * W * p.cin_pad * esize(p.lo));
  p.h0 = p.new_tens(N, H, W, s.input_ch, p.conv_gn_slots(N, H, W, p.cin_pad, 3, s.input_ch));

  auto add_block = [&](const std::vector<Spec>& seq, Tens* a, Tens* b, const std::string& prefix, Tens** result) -> int {
    Tens *cur = a, *cur2 = b;
    for (size_t li = 0; li < seq.size(); ++li) {
      const std::string name = prefix + "." + std::to_string(li);
      if (seq[li].attn) {
        ISB_CHECK_ARG(cur2 == nullptr, "%s: attention over two sources", name.c_str());
        std::unique_ptr<AttnLayer> l(new AttnLayer());
        l->name = name;
        ISB_TRY(l->build(p, seq[li], cur, st));
        cur = l->out;
        p.layers.push_back(std::move(l));
      } else {
        std::unique_ptr<ResLayer> l(new ResLayer());
        l->name = name;
        ISB_TRY(l->build(p, seq[li], cur, cur2, film_offs[name], st));
        cur = l->out;
        p.layers.push_back(std::move(l));
      }
      cur2 = nullptr;
    }
    *result = cur;
    return 0;
  };
  std::vector<Tens*> hs{p.h0};
  Tens* h = p.h0;
  for (size_t i = 1; i < s.input_blocks.size(); ++i) {
    ISB_TRY(add_block(s.input_blocks[i], h, nullptr, "input_blocks." + std::to_string(i), &h));
    hs.push_back(h);
  }
  ISB_TRY(add_block(s.middle, h, nullptr, "middle_block", &h));
  for (size_t i = 0; i < s.output_blocks.size(); ++i) {
    Tens* skip = hs.back();
    hs.pop_back();
    ISB_TRY(add_block(s.output_blocks[i], h, skip, "output_blocks." + std::to_string(i), &h));
    p.block_out.push_back(h);
  }
  p.h_last = h;
  p.want_gn_part(h);

  // --- out (unet.py:612-616) ---
  p.out_g = p.take("out.0.weight", {h->C});
  p.out_b = p.take("out.0.bias", {h->C});
  p.b_out = p.take("out.2.bias", {c.out_channels});
  p.w_out = p.pack("out.2.weight", c.out_channels, s.input_ch, 3, s.input_ch, false, st);
  if (!p.out_g || !p.out_b || !p.b_out || !p.w_out) return ISB_ERR_ARG;
  ISB_CHECK_ARG(h->C == s.input_ch, "isb_unet_finalize: last block has %d channels, out conv expects %d", h->C, s.input_ch);
  if (c.want_backward) {
    p.w_out_d = p.pack("out.2.weight", c.out_channels, s.input_ch, 3, s.input_ch, true, st);
    if (!p.w_out_d) return ISB_ERR_ARG;
    p.req(T_GLO, static_cast<size_t>(N) * H * W * c.out_channels * esize(p.lo));
    p.req(T_G, static_cast<size_t>(N) * H * W * h->C * 4);
    p.req(T_G, static_cast<size_t>(N) * H * W * p.cin_pad * 4);
  }
  p.out_stats = p.alloc(static_cast<size_t>(N) * 32 * 2 * 4);
  p.req(T_A, static_cast<size_t>(N) * H * W * h->C * esize(p.lo));
  p.out_nhwc = p.alloc(static_cast<size_t>(N) * H * W * c.out_channels * 4);

  for (int t = 0; t < T_COUNT; ++t) p.tag_off[t] = p.alloc(p.tag_bytes[t]);
  const size_t gs = isb_gn_scratch_bytes(N, 32);
  p.gn_scratch_off[0] = p.alloc(gs);
  p.gn_scratch_off[1] = p.alloc(gs);

  // the raw conv weights have been packed: release them (once the packing kernels have run)
  ISB_CUDA(cudaStreamSynchronize(st));
  for (auto& kv : p.raw)
    if (kv.second.p) { cudaFree(kv.second.p); kv.second.p = nullptr; }

  if (c.side_stream && c.want_backward) {
    int lo_pri = 0, hi_pri = 0;
    ISB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    ISB_CUDA(cudaStreamCreateWithPriority(&p.side, cudaStreamNonBlocking, hi_pri));
    ISB_CUDA(cudaEventCreateWithFlags(&p.ev_fork, cudaEventDisableTiming));
    ISB_CUDA(cudaEventCreateWithFlags(&p.ev_join, cudaEventDisableTiming));
  }
  return 0;
}

// ---- passes ------------------------------------------------------------------------------------------------------
static int forward_out_layer(Plan& p, Run& r, float* out, int out_nhwc) {
  const isb_unet_cfg& c = p.cfg;
  Tens* h = p.h_last;
  void* a = p.scratch(r, T_A);
  isb_gn_desc d;
  p.gn_desc(d, p.at<float>(r, h->val), h->C, nullptr, 0, h->N, h->H, h->W, p.out_g, p.out_b, nullptr, 1, 0,
            p.at<float>(r, p.out_stats));
  if (h->has_part) { d.partials = p.at<float>(r, h->gn_part); d.partial_slots = h->gn_slots; }
  d.y = a; d.y_dtype = p.lo;
  ISB_TRY(p.gn_forward(r, d));
  float* dst = (out != nullptr && out_nhwc) ? out : p.at<float>(r, p.out_nhwc);
  ISB_TRY(p.conv(r, a, h->N, h->H, h->W, h->C, 3, p.w_out, p.b_out, dst, ISB_F32, c.out_channels));
  if (out != nullptr && !out_nhwc && !r.dry)
    ISB_TRY(isb_nhwc_to_nchw(dst, ISB_F32, out, h->N, c.out_channels, h->H, h->W, c.out_channels, r.st));
  return 0;
}

static int forward(Plan& p, Run& r, const float* x_nchw, const float* t, int feat_layer, int stop_at_feat, float* out,
                   int out_nhwc) {
  const isb_unet_cfg& c = p.cfg;
  if (!r.dry) {
    ISB_TRY(isb_time_embed(t, p.freqs, c.N, c.model_channels, p.hidden, p.te_w1, p.te_b1, p.te_w2, p.te_b2, p.w_all,
                           p.b_all, p.film_rows, p.at<float>(r, p.te_scratch), p.at<float>(r, p.film_all), r.st));
    ISB_TRY(isb_nchw_to_nhwc(x_nchw, p.at(r, p.x_lo), p.lo, c.N, c.in_channels, c.H, c.W, p.cin_pad, r.st));
  }
  ISB_TRY(p.conv(r, p.at(r, p.x_lo), c.N, c.H, c.W, p.cin_pad, 3, p.w_in, p.b_in, p.at(r, p.h0->val), ISB_F32, p.h0->C,
                 nullptr, 0, nullptr, p.h0));
  const Tens* stop = (stop_at_feat && feat_layer >= 0) ? p.block_out[feat_layer] : nullptr;
  p.tail_from = p.layers.size();
  for (size_t li = 0; li < p.layers.size(); ++li) {
    ISB_TRY(p.layers[li]->forward(p, r));
    if (stop != nullptr && p.layers[li]->out == stop) {
      p.tail_from = li + 1;
      return 0;
    }
  }
  return forward_out_layer(p, r, out, out_nhwc);
}

static int forward_tail(Plan& p, Run& r, float* out, int out_nhwc) {
  for (size_t li = p.tail_from; li < p.layers.size(); ++li) ISB_TRY(p.layers[li]->forward(p, r));
  p.tail_from = p.layers.size();
  return forward_out_layer(p, r, out, out_nhwc);
}

static int backward(Plan& p, Run& r, int feat_layer, const float* d_feat, int feat_grad_in_place,
                    const float* d_out_nchw, float* dx_nchw) {
  const isb_unet_cfg& c = p.cfg;
  p.h0->has_grad = false;
  for (auto& l : p.layers) l->out->has_grad = false;
  if (d_out_nchw != nullptr || r.dry) {      // gradient arriving at the UNet output -> h_last (unet.py:612-616 backward)
    Tens* h = p.h_last;
    void* g_lo = p.scratch(r, T_GLO);
    float* g_a = p.scratch<float>(r, T_G);
    if (!r.dry) ISB_TRY(isb_nchw_to_nhwc(d_out_nchw, g_lo, p.lo, c.N, c.out_channels, c.H, c.W, c.out_channels, r.st));
    ISB_TRY(p.conv(r, g_lo, c.N, c.H, c.W, c.out_channels, 3, p.w_out_d, nullptr, g_a, ISB_F32, h->C));
    isb_gn_bwd_desc b;
    memset(&b, 0, sizeof(b));
    p.gn_desc(b.f, p.at<float>(r, h->val), h->C, nullptr, 0, h->N, h->H, h->W, p.out_g, p.out_b, nullptr, 1, 0,
              p.at<float>(r, p.out_stats));
    b.dy = g_a;
    b.gx1 = p.at<float>(r, h->grad); b.acc1 = 0; b.gx1_lo = p.at(r, h->grad_lo);
    b.lo_dtype = p.lo;
    ISB_TRY(p.gn_backward(r, b));
    h->has_grad = true;
  }
  if (feat_layer >= 0 && (d_feat != nullptr || feat_grad_in_place)) {
    Tens* f = p.block_out[feat_layer];
    ISB_CHECK_ARG(!f->has_grad, "isb_unet_backward_input: feat_layer is the last block; pass its gradient through d_out");
    if (d_feat != nullptr && !r.dry)
      ISB_CUDA(cudaMemcpyAsync(p.at(r, f->grad), d_feat, f->numel() * 4, cudaMemcpyDeviceToDevice, r.st));
    ISB_TRY(p.cast_lo(r, p.at<float>(r, f->grad), p.at(r, f->grad_lo), f->numel()));
    f->has_grad = true;
  }
  for (size_t li = p.layers.size(); li-- > 0;)
    if (p.layers[li]->out->has_grad || r.dry) ISB_TRY(p.layers[li]->backward(p, r));
  ISB_CHECK_ARG(p.h0->has_grad || r.dry, "isb_unet_backward_input: no gradient reached the input (no roots given)");
  float* g_x = p.scratch<float>(r, T_G);
  ISB_TRY(p.conv(r, p.at(r, p.h0->grad_lo), c.N, c.H, c.W, p.h0->C, 3, p.w_in_d, nullptr, g_x, ISB_F32, p.cin_pad));
  if (!r.dry) ISB_TRY(isb_nhwc_to_nchw(g_x, ISB_F32, dx_nchw, c.N, c.in_channels, c.H, c.W, p.cin_pad, r.st));
  return 0;
}

}  // namespace un
}  // namespace isb

struct isb_unet {
  isb::un::Plan plan;
};

using isb::un::Plan;
using isb::un::Run;

static int check_run(const isb_unet* h, const void* ws, size_t ws_bytes, const char* who) {
  ISB_CHECK_ARG(h != nullptr, "%s: null handle", who);
  ISB_CHECK_ARG(isb::is_initialised(), "%s: isb_init() has not been called", who);
  ISB_CHECK_ARG(h->plan.finalized, "%s: isb_unet_finalize() has not been called", who);
  ISB_CHECK_ARG(ws != nullptr && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  if (ws_bytes < h->plan.ws_bytes) {
    isb::set_error("%s: workspace of %zu bytes, isb_unet_workspace_bytes() says %zu", who, ws_bytes, h->plan.ws_bytes);
    return ISB_ERR_WORKSPACE;
  }
  return 0;
}

extern "C" {

int isb_unet_create(const isb_unet_cfg* cfg, isb_unet** out) {
  ISB_CHECK_ARG(cfg != nullptr && out != nullptr, "isb_unet_create: null argument");
  ISB_CHECK_ARG(isb::is_initialised(), "isb_unet_create: isb_init() has not been called");
  const int rc = isb::un::check_cfg(*cfg);
  if (rc) return rc;
  isb_unet* h = new isb_unet();
  h->plan.cfg = *cfg;
  h->plan.lo = cfg->mode;
  *out = h;
  return ISB_OK;
}

int isb_unet_load_weight(isb_unet* h, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim,
                         isb_stream_t stream) {
  ISB_CHECK_ARG(h != nullptr && name != nullptr && dev_ptr != nullptr && shape != nullptr, "isb_unet_load_weight: null argument");
  ISB_CHECK_ARG(!h->plan.finalized, "isb_unet_load_weight: the handle is finalized (create a new one for new weights)");
  ISB_CHECK_ARG(dtype == ISB_F32, "isb_unet_load_weight(%s): parameters are handed over as fp32 (the masters)", name);
  ISB_CHECK_ARG(ndim >= 1 && ndim <= 4, "isb_unet_load_weight(%s): ndim %d", name, ndim);
  isb::un::Raw r;
  r.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    ISB_CHECK_ARG(shape[i] > 0, "isb_unet_load_weight(%s): shape[%d] = %lld", name, i, static_cast<long long>(shape[i]));
    r.shape.push_back(shape[i]);
    r.numel *= static_cast<size_t>(shape[i]);
  }
  ISB_CUDA(cudaMalloc(reinterpret_cast<void**>(&r.p), r.numel * 4));
  cudaError_t e = cudaMemcpyAsync(r.p, dev_ptr, r.numel * 4, cudaMemcpyDeviceToDevice, isb::as_stream(stream));
  if (e != cudaSuccess) {
    cudaFree(r.p);
    isb::set_error("isb_unet_load_weight(%s): copy failed: %s", name, cudaGetErrorString(e));
    return ISB_ERR_CUDA;
  }
  auto it = h->plan.raw.find(name);
  if (it != h->plan.raw.end() && it->second.p) cudaFree(it->second.p);
  h->plan.raw[name] = r;
  return ISB_OK;
}

int isb_unet_finalize(isb_unet* h, isb_stream_t stream) {
  ISB_CHECK_ARG(h != nullptr, "isb_unet_finalize: null handle");
  ISB_CHECK_ARG(!h->plan.finalized, "isb_unet_finalize: already finalized");
  Plan& p = h->plan;
  int rc = isb::un::finalize(p, isb::as_stream(stream));
  if (rc) return rc;
  // sizing pass over the whole schedule: the largest split-K workspace any conv of the plan asks for
  Run r;
  r.base = reinterpret_cast<char*>(static_cast<uintptr_t>(1) << 30);
  r.st = nullptr; r.dry = true; r.slot = 0; r.conv_ws_need = 0;
  rc = isb::un::forward(p, r, nullptr, nullptr, -1, 0, nullptr, 0);
  if (rc) return rc;
  if (p.cfg.want_backward) {
    rc = isb::un::backward(p, r, -1, nullptr, 0, nullptr, nullptr);
    if (rc) return rc;
  }
  p.conv_ws_bytes = r.conv_ws_need ? (r.conv_ws_need + r.conv_ws_need / 4 + 1024) / isb::un::kAlign * isb::un::kAlign : 0;
  for (int s = 0; s < 3; ++s) p.conv_ws_off[s] = p.alloc(p.conv_ws_bytes);
  p.ws_bytes = p.top;
  p.finalized = true;
  return ISB_OK;
}

size_t isb_unet_workspace_bytes(const isb_unet* h) { return (h != nullptr && h->plan.finalized) ? h->plan.ws_bytes : 0; }

int isb_unet_workspace_init(isb_unet* h, void* workspace, size_t workspace_bytes, isb_stream_t stream) {
  const int rc = check_run(h, workspace, workspace_bytes, "isb_unet_workspace_init");
  if (rc) return rc;
  ISB_CUDA(cudaMemsetAsync(workspace, 0, h->plan.ws_bytes, isb::as_stream(stream)));
  return ISB_OK;
}

int isb_unet_num_blocks(const isb_unet* h) {
  return (h != nullptr && h->plan.finalized) ? static_cast<int>(h->plan.block_out.size()) : 0;
}

int isb_unet_forward(isb_unet* h, const float* x_nchw, const float* t, int feat_layer, int stop_at_feat, float* out,
                     int out_nhwc, void* workspace, size_t workspace_bytes, isb_stream_t stream) {
  const int rc = check_run(h, workspace, workspace_bytes, "isb_unet_forward");
  if (rc) return rc;
  Plan& p = h->plan;
  ISB_CHECK_ARG(x_nchw != nullptr && t != nullptr, "isb_unet_forward: null input");
  ISB_CHECK_ARG(feat_layer >= -1 && feat_layer < static_cast<int>(p.block_out.size()), "isb_unet_forward: feat_layer %d", feat_layer);
  Run r{static_cast<char*>(workspace), isb::as_stream(stream), false, 0, 0};
  return isb::un::forward(p, r, x_nchw, t, feat_layer, stop_at_feat, out, out_nhwc);
}

int isb_unet_forward_tail(isb_unet* h, float* out, int out_nhwc, void* workspace, size_t workspace_bytes,
                          isb_stream_t stream) {
  const int rc = check_run(h, workspace, workspace_bytes, "isb_unet_forward_tail");
  if (rc) return rc;
  Run r{static_cast<char*>(workspace), isb::as_stream(stream), false, 1, 0};
  return isb::un::forward_tail(h->plan, r, out, out_nhwc);
}

int isb_unet_feat(const isb_unet* h, void* workspace, int feat_layer, float** val, float** grad, int dims[4]) {
  ISB_CHECK_ARG(h != nullptr && h->plan.finalized, "isb_unet_feat: handle not finalized");
  const Plan& p = h->plan;
  ISB_CHECK_ARG(feat_layer >= 0 && feat_layer < static_cast<int>(p.block_out.size()), "isb_unet_feat: feat_layer %d", feat_layer);
  const isb::un::Tens* t = p.block_out[feat_layer];
  char* base = static_cast<char*>(workspace);
  if (val) *val = base ? reinterpret_cast<float*>(base + t->val) : nullptr;
  if (grad) *grad = (base && p.cfg.want_backward) ? reinterpret_cast<float*>(base + t->grad) : nullptr;
  if (dims) { dims[0] = t->N; dims[1] = t->H; dims[2] = t->W; dims[3] = t->C; }
  return ISB_OK;
}

int isb_unet_backward_input(isb_unet* h, int feat_layer, const float* d_feat_nhwc, int feat_grad_in_place,
                            const float* d_out_nchw, float* dx_nchw, void* workspace, size_t workspace_bytes,
                            isb_stream_t stream) {
  const int rc = check_run(h, workspace, workspace_bytes, "isb_unet_backward_input");
  if (rc) return rc;
  Plan& p = h->plan;
  ISB_CHECK_ARG(p.cfg.want_backward, "isb_unet_backward_input: the handle was created with want_backward = 0");
  ISB_CHECK_ARG(dx_nchw != nullptr, "isb_unet_backward_input: null dx");
  ISB_CHECK_ARG(feat_layer >= -1 && feat_layer < static_cast<int>(p.block_out.size()), "isb_unet_backward_input: feat_layer %d", feat_layer);
  Run r{static_cast<char*>(workspace), isb::as_stream(stream), false, 0, 0};
  return isb::un::backward(p, r, feat_layer, d_feat_nhwc, feat_grad_in_place, d_out_nchw, dx_nchw);
}

void isb_unet_destroy(isb_unet* h) { delete h; }

}  // extern "C"
