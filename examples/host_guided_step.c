/*
 * host_guided_step.c — a host with NO Python and NO PyTorch driving the drag-guided denoise step through the C ABI of
 * libishape_b200.so (include/ishape_b200.h).  It is the loop body of the reference's DragStuff.training
 * (drag_utils.py:336-398) written against the handle-level entry points:
 *
 *     isb_unet_forward(stop at inter_feat)  ->  isb_drag_loss_grad (reads the feature, writes its gradient in place)
 *     -> isb_unet_backward_input            ||  isb_unet_forward_tail on a second stream
 *     -> isb_ddpm_step (posterior + guidance update)
 *
 * Inputs (weights by their reference names, latent, per-step schedule rows, noise, cached origin features, the edit's
 * geometry) come from a blob written by tests/test_gpu_c_host.py, which also compares the result with the Python
 * host's GuidedStepper.  Build (the test does exactly this):
 *
 *     gcc -O2 -std=c99 -Iinclude -I/usr/local/cuda/include examples/host_guided_step.c \
 *         -o host_guided_step -Lishapediting_b200 -lishape_b200 -L/usr/local/cuda/lib64 -lcudart \
 *         -Wl,-rpath,$PWD/ishapediting_b200 -Wl,-rpath,/usr/local/cuda/lib64
 *     ./host_guided_step in.blob out.blob [graph]
 *
 * With "graph" the step is captured ONCE into a CUDA graph (two streams, fork / join by events) and replayed; the
 * per-step inputs live in static device buffers refreshed before each replay.
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ishape_b200.h"

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(e_));     \
      exit(2);                                                                                     \
    }                                                                                              \
  } while (0)
#define ISB(call)                                                                                  \
  do {                                                                                             \
    int rc_ = (call);                                                                              \
    if (rc_ != 0) {                                                                                \
      fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, isb_last_error());  \
      exit(3);                                                                                     \
    }                                                                                              \
  } while (0)

/* ---- blob: "ISB1", int32 count, then entries {char name[96]; int32 dtype; int32 ndim; int64 shape[4]; int64 nbytes;
 *      data padded to 8 bytes}.  dtype: 0 f32, 1 i32, 2 u8. ---- */
typedef struct {
  char name[96];
  int32_t dtype, ndim;
  int64_t shape[4], nbytes;
  void* data;
} entry;

static entry* g_entries;
static int g_count;

static void read_blob(const char* path) {
  FILE* f = fopen(path, "rb");
  char magic[4];
  int32_t n;
  if (!f || fread(magic, 1, 4, f) != 4 || memcmp(magic, "ISB1", 4) || fread(&n, 4, 1, f) != 1) {
    fprintf(stderr, "cannot read blob %s\n", path);
    exit(1);
  }
  g_entries = (entry*)calloc((size_t)n, sizeof(entry));
  g_count = n;
  for (int i = 0; i < n; ++i) {
    entry* e = &g_entries[i];
    if (fread(e->name, 1, 96, f) != 96 || fread(&e->dtype, 4, 1, f) != 1 || fread(&e->ndim, 4, 1, f) != 1 ||
        fread(e->shape, 8, 4, f) != 4 || fread(&e->nbytes, 8, 1, f) != 1) {
      fprintf(stderr, "truncated blob header\n");
      exit(1);
    }
    const size_t padded = ((size_t)e->nbytes + 7) / 8 * 8;
    e->data = malloc(padded ? padded : 8);
    if (fread(e->data, 1, padded, f) != padded) {
      fprintf(stderr, "truncated blob data (%s)\n", e->name);
      exit(1);
    }
  }
  fclose(f);
}
static const entry* get(const char* name) {
  for (int i = 0; i < g_count; ++i)
    if (!strcmp(g_entries[i].name, name)) return &g_entries[i];
  fprintf(stderr, "blob has no entry '%s'\n", name);
  exit(1);
}
/* Upload on the stream the library will be called with.  (A plain cudaMemcpy from pageable memory runs on the legacy
 * default stream and may return before its DMA has landed; work on a cudaStreamNonBlocking stream is NOT ordered
 * after it.) */
static cudaStream_t g_stream;
static void* to_device(const entry* e) {
  void* d = NULL;
  CK(cudaMalloc(&d, (size_t)e->nbytes ? (size_t)e->nbytes : 8));
  CK(cudaMemcpyAsync(d, e->data, (size_t)e->nbytes, cudaMemcpyHostToDevice, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return d;
}
static void write_entry(FILE* f, const char* name, int dtype, int64_t n0, const void* data, int64_t nbytes) {
  char nm[96];
  int32_t nd = 1;
  int64_t shape[4] = {n0, 1, 1, 1};
  const char pad[8] = {0};
  memset(nm, 0, sizeof(nm));
  strncpy(nm, name, 95);
  fwrite(nm, 1, 96, f);
  fwrite(&dtype, 4, 1, f);
  fwrite(&nd, 4, 1, f);
  fwrite(shape, 8, 4, f);
  fwrite(&nbytes, 8, 1, f);
  fwrite(data, 1, (size_t)nbytes, f);
  fwrite(pad, 1, (size_t)((8 - nbytes % 8) % 8), f);
}

/* ---- the step --------------------------------------------------------------------------------------------- */
typedef struct {
  isb_unet* net;
  void* ws;
  size_t ws_bytes;
  int feat_layer, C, H, W, out_ch, clip;
  isb_drag_desc drag;
  float *x, *x_next, *t, *coef, *noise, *origin, *dx, *out_nhwc;
  cudaStream_t main_st, tail_st;
  cudaEvent_t fork_ev, join_ev;
} step_ctx;

static void debug_sum(const char* what, const float* dev, size_t n, cudaStream_t st) {
  float* h = (float*)malloc(n * 4);
  double acc = 0;
  CK(cudaStreamSynchronize(st));
  CK(cudaMemcpy(h, dev, n * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) acc += h[i] < 0 ? -h[i] : h[i];
  fprintf(stderr, "debug %s: sum|v| = %.9g over %zu\n", what, acc, n);
  free(h);
}

static void step_body(step_ctx* s) {
  ISB(isb_unet_forward(s->net, s->x, s->t, s->feat_layer, 1, NULL, 0, s->ws, s->ws_bytes, s->main_st));
  if (getenv("ISB_HOST_DEBUG")) {
    debug_sum("x", s->x, (size_t)s->C * s->H * s->W, s->main_st);
    debug_sum("feat", s->drag.feat, (size_t)s->drag.S * s->drag.S * s->drag.Cf, s->main_st);
    debug_sum("origin", s->drag.origin, (size_t)3 * s->drag.S * s->drag.S * s->drag.Ca, s->main_st);
    debug_sum("patch_xy", s->drag.patch_xy, (size_t)3 * s->drag.npts * 2, s->main_st);
    for (int b = 0; b <= s->feat_layer; ++b) {
      float* v = NULL;
      int dm[4];
      char nm[32];
      ISB(isb_unet_feat(s->net, s->ws, b, &v, NULL, dm));
      snprintf(nm, sizeof(nm), "block %d", b);
      debug_sum(nm, v, (size_t)dm[0] * dm[1] * dm[2] * dm[3], s->main_st);
    }
  }
  /* the layers behind the feature only feed the DDPM update: run them beside the backward pass */
  CK(cudaEventRecord(s->fork_ev, s->main_st));
  CK(cudaStreamWaitEvent(s->tail_st, s->fork_ev, 0));
  ISB(isb_unet_forward_tail(s->net, s->out_nhwc, 1, s->ws, s->ws_bytes, s->tail_st));
  CK(cudaEventRecord(s->join_ev, s->tail_st));
  ISB(isb_drag_loss_grad(&s->drag, s->main_st));
  ISB(isb_unet_backward_input(s->net, s->feat_layer, NULL, 1, NULL, s->dx, s->ws, s->ws_bytes, s->main_st));
  CK(cudaStreamWaitEvent(s->main_st, s->join_ev, 0));
  isb_ddpm_desc d;
  memset(&d, 0, sizeof(d));
  d.x = s->x; d.model_out = s->out_nhwc; d.model_out_cstride = s->out_ch; d.model_out_nchw = 0;
  d.noise = s->noise; d.grad = s->dx; d.coef = s->coef;
  d.N = 1; d.C = s->C; d.H = s->H; d.W = s->W; d.clip_denoised = s->clip;
  d.x_next = s->x_next;
  ISB(isb_ddpm_step(&d, s->main_st));
  CK(cudaMemcpyAsync(s->x, s->x_next, (size_t)s->C * s->H * s->W * 4, cudaMemcpyDeviceToDevice, s->main_st));
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s in.blob out.blob [graph]\n", argv[0]);
    return 1;
  }
  const int use_graph = argc > 3 && !strcmp(argv[3], "graph");
  read_blob(argv[1]);
  const int32_t* c = (const int32_t*)get("cfg")->data;
  const float* sc = (const float*)get("scalars")->data;
  ISB(isb_init(0));

  isb_unet_cfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  int k = 0;
  cfg.in_channels = c[k++]; cfg.model_channels = c[k++]; cfg.out_channels = c[k++]; cfg.num_res_blocks = c[k++];
  cfg.n_levels = c[k++];
  for (int i = 0; i < 8; ++i) cfg.channel_mult[i] = c[k++];
  cfg.n_attn = c[k++];
  for (int i = 0; i < 8; ++i) cfg.attention_ds[i] = c[k++];
  cfg.num_heads = c[k++]; cfg.num_head_channels = c[k++]; cfg.num_heads_upsample = c[k++];
  cfg.N = c[k++]; cfg.H = c[k++]; cfg.W = c[k++]; cfg.mode = c[k++];
  const int feat_layer = c[k++], steps = c[k++], group_size = c[k++], mask_count = c[k++], loss_type = c[k++],
            clip = c[k++];
  cfg.want_backward = 1;
  cfg.side_stream = 1;

  step_ctx s;
  memset(&s, 0, sizeof(s));
  CK(cudaStreamCreateWithFlags(&s.main_st, cudaStreamNonBlocking));
  g_stream = s.main_st;
  CK(cudaStreamCreateWithFlags(&s.tail_st, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&s.fork_ev, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&s.join_ev, cudaEventDisableTiming));

  /* weights: every "w:<reference parameter name>" entry */
  ISB(isb_unet_create(&cfg, &s.net));
  for (int i = 0; i < g_count; ++i) {
    const entry* e = &g_entries[i];
    if (strncmp(e->name, "w:", 2)) continue;
    void* d = to_device(e);
    ISB(isb_unet_load_weight(s.net, e->name + 2, d, ISB_F32, e->shape, e->ndim, s.main_st));
    CK(cudaStreamSynchronize(s.main_st));
    CK(cudaFree(d));
  }
  ISB(isb_unet_finalize(s.net, s.main_st));
  s.ws_bytes = isb_unet_workspace_bytes(s.net);
  CK(cudaMalloc(&s.ws, s.ws_bytes));
  ISB(isb_unet_workspace_init(s.net, s.ws, s.ws_bytes, s.main_st));

  /* the edit: latent, schedule rows, noise and cached origin features per step, geometry */
  const entry *ex = get("x"), *et = get("t"), *ecoef = get("coef"), *enoise = get("noise"), *eorigin = get("origin");
  s.feat_layer = feat_layer; s.C = cfg.in_channels; s.H = cfg.H; s.W = cfg.W; s.out_ch = cfg.out_channels; s.clip = clip;
  const size_t img_bytes = (size_t)s.C * s.H * s.W * 4;
  s.x = (float*)to_device(ex);
  CK(cudaMalloc((void**)&s.x_next, img_bytes));
  CK(cudaMalloc((void**)&s.dx, img_bytes));
  CK(cudaMalloc((void**)&s.noise, img_bytes));
  CK(cudaMalloc((void**)&s.out_nhwc, (size_t)s.out_ch * s.H * s.W * 4));
  CK(cudaMalloc((void**)&s.t, 4));
  CK(cudaMalloc((void**)&s.coef, 32));
  const size_t origin_bytes = (size_t)eorigin->nbytes / (size_t)steps;
  CK(cudaMalloc((void**)&s.origin, origin_bytes));

  float *feat = NULL, *feat_grad = NULL;
  int fd[4];
  ISB(isb_unet_feat(s.net, s.ws, feat_layer, &feat, &feat_grad, fd));
  const entry* epatch = get("patch_xy");
  const int npts = (int)epatch->shape[1];
  const int Ca = (int)eorigin->shape[2];      /* origin: [steps][3*S*S][Ca] */
  isb_drag_desc* dd = &s.drag;
  dd->feat = feat; dd->S = fd[1]; dd->Cf = fd[3];
  dd->origin = s.origin; dd->Ca = Ca;
  dd->chan_map = (const int32_t*)to_device(get("chan_map"));
  dd->inv_map = (const int32_t*)to_device(get("inv_map"));
  dd->patch_xy = (const float*)to_device(epatch);
  dd->shift_xy = (const float*)to_device(get("shift_xy"));
  dd->weight = (const float*)to_device(get("weight"));
  dd->npts = npts; dd->group_size = group_size;
  dd->bbox = (const int32_t*)to_device(get("bbox"));
  dd->mask = (const uint8_t*)to_device(get("mask"));
  dd->mask_count = mask_count;
  dd->inv_count = sc[0]; dd->cof = sc[1]; dd->loss_type = loss_type;
  CK(cudaMalloc((void**)&dd->g, (size_t)3 * npts * Ca * 4));
  CK(cudaMalloc((void**)&dd->pt_info, (size_t)3 * npts * 4 * 4));
  dd->partial_len = (int)isb_drag_partial_len(dd->S, dd->Cf, npts);
  CK(cudaMalloc((void**)&dd->partial, (size_t)dd->partial_len * 8));
  CK(cudaMemsetAsync(dd->partial, 0, (size_t)dd->partial_len * 8, s.main_st));
  CK(cudaMalloc((void**)&dd->loss, 4));
  dd->d_feat = feat_grad;
  dd->dyn_scalars = (const float*)to_device(get("dyn"));   /* {inv_count, 1/(Ca*mask_count)} as the Python host passes them */

  cudaGraphExec_t exec = NULL;
  cudaEvent_t t0, t1;
  CK(cudaEventCreate(&t0));
  CK(cudaEventCreate(&t1));
  float ms_total = 0.f;
  const int first_timed = use_graph ? 2 : 1;      /* step 0 warms up, step 1 also pays the capture in graph mode */
  const uint64_t launches0 = isb_launch_count();
  for (int i = 0; i < steps; ++i) {
    CK(cudaMemcpyAsync(s.t, (const float*)et->data + i, 4, cudaMemcpyHostToDevice, s.main_st));
    CK(cudaMemcpyAsync(s.coef, (const float*)ecoef->data + 8 * i, 32, cudaMemcpyHostToDevice, s.main_st));
    CK(cudaMemcpyAsync(s.noise, (const char*)enoise->data + img_bytes * (size_t)i, img_bytes, cudaMemcpyHostToDevice, s.main_st));
    CK(cudaMemcpyAsync(s.origin, (const char*)eorigin->data + origin_bytes * (size_t)i, origin_bytes, cudaMemcpyHostToDevice, s.main_st));
    CK(cudaEventRecord(t0, s.main_st));
    if (use_graph && i >= 1) {     /* step 0 runs eagerly (warm-up), the graph is captured at step 1 and replayed after */
      if (exec == NULL) {
        cudaGraph_t graph;
        CK(cudaStreamBeginCapture(s.main_st, cudaStreamCaptureModeThreadLocal));
        step_body(&s);
        CK(cudaStreamEndCapture(s.main_st, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        CK(cudaGraphDestroy(graph));
      }
      CK(cudaGraphLaunch(exec, s.main_st));
    } else {
      step_body(&s);
    }
    CK(cudaEventRecord(t1, s.main_st));
    CK(cudaStreamSynchronize(s.main_st));
    float ms;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    if (i >= first_timed) ms_total += ms;
  }
  const uint64_t launches = isb_launch_count() - launches0;

  float* h_img = (float*)malloc(img_bytes);
  float* h_dx = (float*)malloc(img_bytes);
  float h_loss, h_ms = steps > first_timed ? ms_total / (float)(steps - first_timed) : 0.f;
  CK(cudaMemcpy(h_img, s.x, img_bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_dx, s.dx, img_bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&h_loss, dd->loss, 4, cudaMemcpyDeviceToHost));
  FILE* f = fopen(argv[2], "wb");
  if (!f) {
    fprintf(stderr, "cannot write %s\n", argv[2]);
    return 1;
  }
  const int32_t n_out = 4;
  fwrite("ISB1", 1, 4, f);
  fwrite(&n_out, 4, 1, f);
  write_entry(f, "img", 0, (int64_t)(img_bytes / 4), h_img, (int64_t)img_bytes);
  write_entry(f, "grad", 0, (int64_t)(img_bytes / 4), h_dx, (int64_t)img_bytes);
  write_entry(f, "loss", 0, 1, &h_loss, 4);
  write_entry(f, "ms_per_step", 0, 1, &h_ms, 4);
  fclose(f);
  printf("host_guided_step: %d steps (%s), %.3f ms/step after warm-up, %llu library launches, workspace %.1f MB\n",
         steps, use_graph ? "CUDA graph" : "eager", h_ms, (unsigned long long)launches, s.ws_bytes / 1048576.0);
  isb_unet_destroy(s.net);
  return 0;
}
