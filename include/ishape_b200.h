/*
 * ishape_b200.h — C ABI of libishape_b200.so
 *
 * B200 (sm_100a) kernels for the drag-guided triplane-diffusion editing step of
 * jinli99/iShapEditing.  The reference has no FFI of its own (it is pure
 * Python/PyTorch); every entry point below replaces one PyTorch library call
 * (or a fixed group of them) that the reference issues on its hot path, and the
 * comment on each entry point cites the reference call site it replaces.
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     it is named host_*; the caller owns every buffer, including workspaces;
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as
 *     void*), nothing synchronises, nothing allocates -> CUDA-graph capturable;
 *   - activations are NHWC (channels innermost); the "stream" tensors (block
 *     outputs, gradients) are fp32, GEMM operands are bf16 ("bf16 mode",
 *     tcgen05 tensor cores) or fp32 ("fp32 mode", FFMA);
 *   - return 0 on success, negative isb_status otherwise; the message is
 *     available through isb_last_error() (thread local);
 *   - callable from any host thread; no thread-affine state.
 */
#ifndef ISHAPE_B200_H
#define ISHAPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISB_ABI_VERSION 1

typedef void* isb_stream_t; /* cudaStream_t */

enum isb_dtype { ISB_F32 = 0, ISB_BF16 = 1 };

enum isb_status {
  ISB_OK = 0,
  ISB_ERR_ARG = -1,         /* bad argument / unsupported shape */
  ISB_ERR_CUDA = -2,        /* CUDA runtime/driver error */
  ISB_ERR_WORKSPACE = -3,   /* workspace too small */
  ISB_ERR_NOT_INIT = -4     /* isb_init() not called for this device */
};

/* ---- library ---------------------------------------------------------- */
int isb_abi_version(void);
/* Select device, resolve cuTensorMapEncodeTiled, raise kernel smem limits.
 * Must be called once per process per device before any other call. */
int isb_init(int device);
const char* isb_last_error(void);

/* ---- layout ----------------------------------------------------------- */
/* NCHW fp32 -> NHWC (fp32|bf16), channels zero-padded to c_pad >= C.
 * Replaces `h = x.type(self.dtype)` (unet.py:657) + cuDNN's internal layout. */
int isb_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int N, int C,
                     int H, int W, int c_pad, isb_stream_t stream);
/* NHWC (fp32|bf16) with c_stride channels per pixel -> NCHW fp32 (first C). */
int isb_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int N, int C,
                     int H, int W, int c_stride, isb_stream_t stream);
/* fp32 -> bf16 elementwise cast (n elements, n % 8 == 0). */
int isb_cast_f32_bf16(const float* src, void* dst, size_t n, isb_stream_t stream);

/* ---- convolution / linear as implicit GEMM ---------------------------- */
/* out[n,h,w,co] = (accumulate ? out : 0) + bias[co] + residual[n,h,w,co]
 *               + sum_{kh,kw,ci} a[n,h+kh-p,w+kw-p,ci] * w[co][(kh*ks+kw)*Cin+ci]
 *               + sum_{ci2}      a2[n,h,w,ci2]         * w[co][ks*ks*Cin+ci2]
 * Replaces nn.Conv2d 3x3/1x1 and nn.Conv1d k=1 (unet.py:185,211,222,286,294,
 * 482,615 via nn.py:21-31) and, with the transposed/flipped packing, their
 * cuDNN backward-data (autograd of drag_utils.py:383).
 * a_dtype == ISB_BF16 selects the TMA + tcgen05 kernel (fp32 accumulation in
 * TMEM); a_dtype == ISB_F32 selects the FFMA kernel ("fp32 mode"). */
typedef struct isb_conv_desc {
  const void* a;         /* [N,H,W,Cin], a_dtype */
  int a_dtype;
  int N, H, W, Cin;      /* bf16: Cin % 64 == 0; fp32: Cin % 16 == 0 */
  int ksize;             /* 1 or 3 (stride 1, zero pad ksize/2) */
  const void* a2;        /* optional second 1x1 source [N,H,W,Cin2] or NULL */
  int Cin2;
  const void* w;         /* packed [Cout][K], K = ksize*ksize*Cin + Cin2, a_dtype; or, if w_tiled,
                            the same matrix as [Cout/64][K/64][64][64] panels (bf16 path only) */
  const float* bias;     /* [Cout] or NULL */
  const float* residual; /* fp32 [N,H,W,Cout] or NULL */
  void* out;             /* [N,H,W,Cout] */
  int out_dtype;
  int Cout;              /* % 8 == 0 */
  int accumulate;        /* fp32 out only */
  /* tuning overrides for the tcgen05 kernel; 0 = heuristic */
  int block_n;           /* 64,128,192,256 */
  int split_k;
  int stages;
  int w_tiled;           /* 1: w is panel-tiled (see above); needs Cout % 64 == 0 */
  int two_cta;           /* 0 heuristic, 1 force the CTA-pair (cta_group::2) kernel, 2 forbid it */
  int debug_flags;       /* profiling only (results become garbage): bit0 skip the TMA loads, bit1 skip the MMAs */
  int min_smem_bytes;    /* 0, or a floor for the dynamic shared memory of the launch: a background-stream conv asks
                            for > half an SM (e.g. 116 KB) so that only one of its CTAs is resident per SM and the
                            latency-critical stream's CTAs always find room beside it */
  /* Optional GroupNorm statistics of the OUTPUT, fused into the epilogue (bf16 path, fp32 out): every CTA writes
   * the (sum, sum of squares) of its part of each group of gn_cg consecutive output channels; the consumer
   * (isb_gn_forward with `partials`) folds the gn_slots contributions per (image, group) in fixed order.
   * This removes the statistics pass of nn.py:16-18 GroupNorm32 over a conv output (unet.py:183,207,285). */
  float* gn_partials;    /* [N][Cout/gn_cg][gn_slots][2] fp32, or NULL */
  int gn_cg;             /* channels per group: 8, 16 or 32 (also set it for the isb_conv2d_gn_slots query) */
  int gn_slots;          /* must equal isb_conv2d_gn_slots(d) */
  /* gn_mode 2 (default 0/1 = the statistics above): this conv is the backward-data pass that produces dy of a
   * GroupNorm(+FiLM)(+SiLU) layer whose input gb_x [N,H,W,Cout] fp32 has the output's shape; gn_partials then
   * receives the two reduction terms of that GroupNorm's backward, sum(dz g') and sum(dz g' xhat) per (image, group)
   * (dz = dy * silu'(z)), so that isb_gn_backward (with `partials`) needs no reduction pass. */
  int gn_mode;
  const float* gb_x; const float* gb_gamma; const float* gb_beta;
  const float* gb_film; int gb_film_stride;   /* FiLM rows as in isb_gn_desc, or NULL */
  const float* gb_stats;                       /* [N,groups,2] mean,rstd of the forward pass */
  int gb_silu;
} isb_conv_desc;
/* Workspace (split-K partial tiles + arrival counters): must be ZERO-FILLED by the caller before its
 * first use; every launch leaves the counters at zero again, so one buffer serves all layers. */
size_t isb_conv2d_workspace(const isb_conv_desc* d);
/* Contributions per (image, group) the launch described by d writes to gn_partials; 0 = the statistics cannot be
 * fused for this configuration (the caller then runs the ordinary statistics pass). */
int isb_conv2d_gn_slots(const isb_conv_desc* d);
int isb_conv2d(const isb_conv_desc* d, void* workspace, size_t workspace_bytes,
               isb_stream_t stream);

/* ---- GroupNorm (+FiLM) (+SiLU) (+2x resample) ------------------------- */
/* Source tensor x is the channel concatenation of x1 [N,H,W,C1] and (optional)
 * x2 [N,H,W,C2], both fp32 NHWC — this fuses `th.cat([h, hs.pop()], dim=1)`
 * (unet.py:663) into the consumer.
 *   z   = GN_32groups(x) * gamma + beta                (nn.py:16-18, fp32)
 *   z   = z * (1 + scale[n,c]) + shift[n,c]            if film   (unet.py:248-252)
 *   a   = silu ? z*sigmoid(z) : z                       (unet.py:184,208,614)
 *   a   = avgpool2x2(a) | nearest_up2x(a)               resample 1 | 2 (unet.py:107,136)
 * Outputs (each optional unless noted):
 *   stats [N,groups,2] fp32 (mean,rstd)  — required, kept for the backward
 *   y     activation in y_dtype at output resolution — required
 *   raw   x itself (concatenated) cast to raw_dtype at input resolution (operand
 *         of the 1x1 skip convolution, unet.py:222)
 *   xres  x resampled like `a`, fp32 (the `x = self.x_upd(x)` branch, unet.py:240)
 * film points at [N, film_stride] fp32 with scale at [0,C) and shift at [C,2C). */
typedef struct isb_gn_desc {
  const float* x1; int C1;
  const float* x2; int C2;
  int N, H, W;            /* input resolution */
  int groups; float eps;
  const float* gamma; const float* beta;   /* [C1+C2] */
  const float* film; int film_stride;      /* NULL = no FiLM */
  int silu;
  int resample;           /* 0 none, 1 down (avgpool 2x2), 2 up (nearest 2x) */
  float* stats;           /* [N,groups,2] */
  void* y; int y_dtype;
  void* raw; int raw_dtype;
  float* xres;
  /* statistics already accumulated by the producer conv (isb_conv_desc.gn_partials): [N][groups][partial_slots][2];
   * needs x2 == NULL and resample == 0.  NULL = compute them here. */
  const float* partials; int partial_slots;
} isb_gn_desc;
/* scratch: device buffer of isb_gn_scratch_bytes(N, groups) bytes, must be
 * zero-initialised once by the caller and is left zeroed by each call.  The
 * arrival counters occupy a fixed 16 KiB head (N*groups <= 4096), so one buffer
 * sized for the largest N may serve launches of every smaller N; launches that
 * may run CONCURRENTLY (different streams) need separate buffers. */
size_t isb_gn_scratch_bytes(int N, int groups);
int isb_gn_forward(const isb_gn_desc* d, void* scratch, isb_stream_t stream);

/* Backward of the above w.r.t. x (input gradient only; the reference's weight
 * gradients — drag_utils.py:383 computes them and never reads them — are not
 * produced).
 *   dy    [N,Ho,Wo,C] fp32: gradient w.r.t. `y` (output resolution)
 *   gres  optional [N,H,W,C] fp32 at INPUT resolution if gres_at_input, else at
 *         output resolution and pulled back through the resample (the identity
 *         / x_upd skip path): added to dx
 *   gx1/gx2 fp32 gradient of x1/x2, overwritten or accumulated (acc1/acc2)
 *   gx1_lo/gx2_lo optional copies of the final gx in lo_dtype (operand of the
 *         upstream dgrad GEMM)
 */
typedef struct isb_gn_bwd_desc {
  isb_gn_desc f;          /* forward description (stats is an INPUT here; y/raw/xres unused) */
  const float* dy;
  const float* gres; int gres_at_input;
  float* gx1; int acc1; void* gx1_lo;
  float* gx2; int acc2; void* gx2_lo;
  int lo_dtype;
  /* reduction terms already accumulated by the dgrad conv that produced dy (isb_conv_desc.gn_mode 2):
   * [N][groups][partial_slots][2]; needs a single source and no resample.  NULL = reduce here. */
  const float* partials; int partial_slots;
} isb_gn_bwd_desc;
int isb_gn_backward(const isb_gn_bwd_desc* d, void* scratch, isb_stream_t stream);

/* ---- attention core (QKVAttentionLegacy, unet.py:337-354) -------------- */
/* qkv [N,T,3*C] fp32 NHWC-flattened, channel layout per head h:
 * [h*3*ch, +ch) = q, next ch = k, next ch = v  (legacy interleave, unet.py:346).
 * out [N,T,C] (out_dtype).  P = softmax_fp32(q k^T / sqrt(ch)) is written to
 * probs [N,heads,T,T] fp32 and kept for the backward. */
int isb_attention_forward(const float* qkv, int N, int T, int heads, int ch,
                          float* probs, void* out, int out_dtype,
                          isb_stream_t stream);
/* d_out [N,T,C] fp32 -> d_qkv [N,T,3C] (lo_dtype); tmp [N,heads,T,T] fp32. */
int isb_attention_backward(const float* qkv, const float* probs,
                           const float* d_out, int N, int T, int heads, int ch,
                           float* tmp, void* d_qkv, int lo_dtype,
                           isb_stream_t stream);

/* Fused (flash) form of the same attention for the bf16 mode, 64 channels per head: the [heads,T,T]
 * probabilities are never written.  qkv, out, d_out, d_qkv are bf16 with the layouts above; the forward keeps
 * only lse [N,heads,T] fp32 (row log-sum-exp of the scaled scores, log2 domain); the backward recomputes P from
 * it, writes delta [N,heads,T] = rowsum(d_out o out) as scratch and d_qkv.  T % 64 == 0, ch == 64. */
int isb_attention_flash_forward(const void* qkv, int N, int T, int heads, int ch,
                                void* out, float* lse, isb_stream_t stream);
int isb_attention_flash_backward(const void* qkv, const void* out, const void* d_out,
                                 const float* lse, int N, int T, int heads, int ch,
                                 float* delta, void* d_qkv, isb_stream_t stream);

/* ---- timestep embedding path (nn.py:102-120, unet.py:471-475,199-205) --- */
/* film_all[n, :] = W_all @ silu(W2 @ silu(W1 @ sinus(t[n]) + b1) + b2) + b_all
 * where W_all/b_all is the row-concatenation of every ResBlock's emb_layers
 * Linear (rows_all = sum 2*Cout).  All weights fp32 row-major [out,in].
 * t is a device fp32 array of the timestep VALUES the UNet sees (after respace.py:122-126 and, with
 * rescale_timesteps, gaussian_diffusion.py:171-174: possibly fractional — the reference embeds floats, nn.py:116);
 * freqs [model_ch/2] is the reference's exp(-ln(1e4)*i/half) table (nn.py:113-115)
 * computed once by the host.  scratch: N * (model_ch + 2*hidden) floats. */
int isb_time_embed(const float* t, const float* freqs, int N, int model_ch,
                   int hidden, const float* w1, const float* b1, const float* w2,
                   const float* b2, const float* w_all, const float* b_all,
                   int rows_all, float* scratch, float* film_all,
                   isb_stream_t stream);

/* ---- fused DDPM posterior / guidance update ---------------------------- */
/* One kernel for gaussian_diffusion.py:265-279,296-317,333-338,208-230,490-506
 * and drag_utils.py:384-392.  Per-step schedule scalars are read from DEVICE
 * memory (coef[8], see isb_sched_coef) so that one captured CUDA graph serves
 * every step index.  model_out is NHWC fp32 with 2*C channels (eps | v).
 * All image tensors are NCHW fp32 [N,C,H,W] except model_out.
 *   x_next = mean + nonzero*sqrt(var)*noise + (grad ? var*scale*grad : 0)
 * Optional outputs (NULL to skip): sample (without guidance), mean, var, x0, eps. */
enum isb_sched_coef {
  ISB_SC_SQRT_RECIP_ACP = 0, ISB_SC_SQRT_RECIPM1_ACP = 1,
  ISB_SC_POST_COEF1 = 2, ISB_SC_POST_COEF2 = 3,
  ISB_SC_MIN_LOG = 4, ISB_SC_MAX_LOG = 5, ISB_SC_NONZERO = 6, ISB_SC_GUIDE_SCALE = 7
};
typedef struct isb_ddpm_desc {
  const float* x; const float* model_out; int model_out_cstride;
  int model_out_nchw;   /* 1: model_out is NCHW [N,2C,H,W] (generic API route) */
  const float* noise; const float* grad;
  const float* coef;  /* device [8] */
  int N, C, H, W; int clip_denoised;
  float* x_next; float* sample; float* mean; float* var; float* x0; float* eps;
  int coef_per_sample;  /* 1: coef is [N][8], one row per batch element (batched DDPM inversion) */
} isb_ddpm_desc;
int isb_ddpm_step(const isb_ddpm_desc* d, isb_stream_t stream);

/* ---- drag guidance: motion supervision + mask regulariser -------------- */
/* Replaces resize_feat_align + 2x F.grid_sample + masked gathers + the autograd
 * backward of the loss down to inter_feat (drag_utils.py:141-159,351-383).
 * feat     : inter_feat of the current step, NHWC fp32 [1,S,S,Cf]
 * origin   : cached resize_feat_align(inter_feat) of the unedited trajectory,
 *            plane-major channels-last fp32 [3,S,S,Ca]
 * chan_map : [3*Ca] int32, feat channel that feeds aligned (plane, a)
 * inv_map  : [Cf] int32, plane*Ca + a for a feat channel, -1 if resize_feat_align
 *            drops it (drag_utils.py:146-151)
 * patch_xy / shift_xy : [3,npts,2] fp32, per-plane 2-D sample points (x->W, y->H)
 *            in [-1,1]; the (2r+1)^3 lattice of drag_utils.py:316-321 projects to
 *            (2r+1)^2 distinct points per plane and handle, `weight[npts]` holds
 *            each point's multiplicity
 * group_size : points per handle (contiguous), bbox [3,npts/group_size,4] int32 =
 *            (x0min,x0max,y0min,y0max) over floor() pixel indices of the group's
 *            shift points (may be conservative)
 * mask     : [3,S,S] uint8, 1 where the pixel is OUTSIDE the content set
 *            (drag_utils.py:322-334); mask_count = number of ones
 * inv_count: 1 / (3*Ca*B*N1)   loss_type 0 = l2, 1 = l1
 * scratch  : g [3,npts,Ca] f32, pt_info [3,npts,4] f32,
 *            partial [isb_drag_partial_len()] f64
 * Writes loss (1 float) and d_feat = dLoss/dfeat NHWC fp32 [1,S,S,Cf] (fully
 * overwritten, deterministic: gather, no atomics). */
typedef struct isb_drag_desc {
  const float* feat; int S; int Cf;
  const float* origin; int Ca;
  const int32_t* chan_map; const int32_t* inv_map;
  const float* patch_xy; const float* shift_xy; const float* weight;
  int npts; int group_size;
  const int32_t* bbox;
  const uint8_t* mask; int mask_count;
  float inv_count; float cof; int loss_type;
  float* g; float* pt_info; double* partial; int partial_len;
  float* loss; float* d_feat;
  const float* dyn_scalars;  /* optional DEVICE [2] = {inv_count, 1/(Ca*mask_count)}: overrides the two by-value
                                scalars so that one captured CUDA graph serves edits with different handles */
} isb_drag_desc;
size_t isb_drag_partial_len(int S, int Cf, int npts);
int isb_drag_loss_grad(const isb_drag_desc* d, isb_stream_t stream);
/* resize_feat_align as a gather: NHWC fp32 [1,S,S,Cf] -> [3,S,S,Ca]. */
int isb_resize_feat_align(const float* feat, int S, int Cf, const int32_t* chan_map,
                          float* out, int Ca, isb_stream_t stream);

/* ---- point tracking (OPT-IN EXTENSION; parity unpinned) ----------------- */
/* BASELINE.json north_star names "point tracking ... nearest-feature search with a warp-level argmin"; the
 * reference has no such function (handles and targets are fixed for a whole edit, drag_utils.py:305-321), so there
 * is no reference interface this replaces.  For each of B handle points `center` [B,3] the (2r+1)^3 voxel lattice
 * around it (offsets in drag_utils.make_offsets order, index = ((i+r)*side + (j+r))*side + (k+r), spacing `voxel`)
 * is searched for the point whose aligned triplane feature (bilinear samples of the three planes, as in
 * drag_utils.py:318-321,355-358) is nearest in L1 to f0 [B,3,Ca].  feat is the raw NHWC intermediate feature
 * [1,S,S,Cf]; chan_map as for isb_drag_loss_grad.  table: scratch [B,3,(2r+1)^2] fp32 (the three separable
 * distance tables, left filled for inspection).  Outputs: out_idx [B] int32 lattice index (ties -> lowest index),
 * out_dist [B] fp32, out_pts [B,3] fp32 = center + voxel * offset. */
typedef struct isb_track_desc {
  const float* feat; int S; int Cf; int Ca;
  const int32_t* chan_map;
  const float* f0;
  const float* center;
  int B; int r; float voxel;
  float* table;
  int32_t* out_idx; float* out_dist; float* out_pts;
} isb_track_desc;
int isb_track_points(const isb_track_desc* d, isb_stream_t stream);

/* ---- triplane decoder (axisnetworks.py:537-562, visualize.py:79-98) ---- */
typedef struct isb_triplane_mlp {
  const float* fourier_B;      /* [32,64] */
  const float* w1; const float* b1;  /* [128,128],[128] */
  const float* w2; const float* b2;  /* [128,128],[128] */
  const float* w3; const float* b3;  /* [1,128],[1] */
  /* Upper bound of the layer-1 pre-activations, max_r(sum_k |w1[r,k]| + |b1[r]|) (its inputs are sines / cosines),
   * computed once by the host.  > 0 and fp16-safe (<= 30000): the two 128x128 layers run on tcgen05 as split-fp16
   * (hi + lo) products with fp32 TMEM accumulation (decode_tc.cu, fp32-grade results).  0 = unknown: the 3xTF32
   * mma.sync kernel (decode.cu) is used.  Both are GPU paths with the same results to ~1e-6. */
  float h1_bound;
} isb_triplane_mlp;
/* planes_hwc: [3,R,R,32] fp32 channels-last (isb_nchw_to_nhwc of the reference's
 * three (1,32,R,R) embeddings).  Dense grid query of lin^3 where lin[res] is the
 * host's torch.linspace(-1,1,res) table (bit-identical coordinates), index =
 * x*res^2 + y*res + z (visualize.py:79-86), for x in [x_begin, x_end):
 * out[(x-x_begin)*res^2 + y*res + z]. */
int isb_triplane_decode_grid(const float* planes_hwc, int R, const isb_triplane_mlp* w,
                             const float* lin, int res, int x_begin, int x_end,
                             float* out, isb_stream_t stream);
/* Arbitrary points: coords [npts,3] -> out [npts]  (MultiTriplane.forward). */
int isb_triplane_decode_points(const float* planes_hwc, int R, const isb_triplane_mlp* w,
                               const float* coords, int64_t npts, float* out,
                               isb_stream_t stream);

/* Gradient of sum_i d_logits[i] * logit_i w.r.t. the three planes, ADDED to d_planes_hwc [3,R,R,32] (the caller
 * zero-fills it): the autograd backward of MultiTriplane.forward that the reference's reconstruction guidance
 * runs (drag_utils.py:443-463, loss.backward() through axisnetworks.py:546-562 into pred_xstart).  The forward is
 * recomputed in fp32; the bilinear scatter uses float atomics (summation order not fixed). */
int isb_triplane_decode_points_backward(const float* planes_hwc, int R, const isb_triplane_mlp* w,
                                        const float* coords, int64_t npts, const float* d_logits,
                                        float* d_planes_hwc, isb_stream_t stream);

/* ---- meshing tail of the decode path -------------- */
/* Marching cubes over the logit volume vol[res][res][res] (x-major, as isb_triplane_decode_grid writes it), replacing
 * `mcubes.marching_cubes(pred, 0)` + `vertices / res * 2 - 1` (triplane_decoder/visualize.py:100-101).  Two calls,
 * because the output size is data-dependent:
 *   isb_mc_count  classifies every cell / grid edge and scans; counts[0] = vertices, counts[1] = triangles (DEVICE
 *                 int64[2]; the caller reads them back and allocates);
 *   isb_mc_emit   writes verts [nv,3] fp32 and tris [nt,3] int32 using the SAME workspace.
 * One shared vertex per crossed grid edge (linear interpolation, IEEE fp32 ops); vertex order = grid point (x-major)
 * then edge axis, triangle order = cell (x-major) then case-table order: deterministic, no atomics.
 * scale_div > 0: vertices are written as v / scale_div * 2 - 1 (the reference passes res); 0: index coordinates as
 * PyMCubes returns them.  Case bit i = value(corner i) < iso, Bourke corner / edge numbering; the table is derived in
 * ishapediting_b200/triplane_decoder/mc_table.py (PyMCubes is an un-pinned dependency not installed offline: parity
 * with that binary is unpinned; the CPU checker is oracle/mcubes_oracle.py). */
size_t isb_mc_workspace_bytes(int res);
int isb_mc_count(const float* vol, int res, float iso, void* workspace, size_t workspace_bytes, int64_t* counts,
                 isb_stream_t stream);
int isb_mc_emit(const float* vol, int res, float iso, void* workspace, size_t workspace_bytes, float scale_div,
                float* verts, int32_t* tris, isb_stream_t stream);
/* Open3D TriangleMesh::FilterSmoothSimple (`filter_smooth_simple(number_of_iterations=10)`, drag_utils.py:300):
 * `iterations` times v_i <- (v_i + sum_{j in N(i)} v_j) / (1 + |N(i)|), N(i) = vertices sharing a triangle edge with i,
 * every vertex updated from the previous iterate, float64 arithmetic (as Open3D), verts [nv,3] fp32 in place. */
size_t isb_mesh_smooth_workspace_bytes(int64_t nv, int64_t nt);
int isb_mesh_smooth_simple(float* verts, int64_t nv, const int32_t* tris, int64_t nt, int iterations,
                           void* workspace, size_t workspace_bytes, isb_stream_t stream);

/* ---- L2 prefetch -------------------------------------------------------- */
/* Start pulling [ptr, ptr+bytes) into L2 (cp.async.bulk.prefetch.L2, one per 16 KiB) and return; the host layer
 * uses it on a side stream to stream the weight panels of the NEXT layers while the current, latency-bound layer
 * runs (HBM is idle ~95 % of a batch-1 step).  The reference has no counterpart: PyTorch eager issues no prefetch. */
int isb_prefetch_l2(const void* ptr, size_t bytes, isb_stream_t stream);

/* ---- the UNet behind a handle ---------------------------- */
/* Replaces `UNetModel.__init__` + `load_state_dict` + `forward(x, timesteps, feat_layer)` (neural_field_diffusion/
 * guided_diffusion/unet.py:396-671) and the input-gradient half of `loss.backward()` (drag_utils.py:383) for hosts
 * that are not Python: the block structure, the weight packing (forward and backward-data panels), the activation
 * layout and the launch schedule live in the library.  The per-operator entry points above are what it calls, in
 * the same order as the Python host's plan, so both hosts produce bit-identical results.
 *   create -> load_weight (every entry of the reference's state_dict, by its own name, fp32 device pointers)
 *          -> finalize (packs; afterwards the handle owns only packed weights)
 *          -> workspace_bytes / workspace_init (ONE caller-owned buffer: activations, gradients, scratch)
 *          -> forward / forward_tail / backward_input on the caller's stream (no sync, no allocation: capturable).
 * One handle serves one pass at a time (it records which tensors carry a gradient).  Dropout is the identity
 * (eval mode, drag_utils.py:187); class conditioning, resblock_updown=False and use_new_attention_order=True are
 * outside the NFD configuration, as in the Python host. */
typedef struct isb_unet_cfg {
  int in_channels, model_channels, out_channels, num_res_blocks;
  int n_levels; int channel_mult[8];
  int n_attn; int attention_ds[8];   /* script_util.py:139-141: image_size // attention_resolution, e.g. {4,8,16} */
  int num_heads, num_head_channels, num_heads_upsample;   /* as the reference's arguments (-1 as there) */
  int N, H, W;                       /* batch and latent resolution this handle is planned for */
  int mode;                          /* ISB_BF16: tcgen05 path (bf16 operands, fp32 accumulate); ISB_F32: FFMA path */
  int want_backward;                 /* 0: no backward-data panels, no gradient buffers */
  int side_stream;                   /* 1: each ResBlock's 1x1-skip backward runs on an internal high-priority stream
                                        beside the conv2 -> GN2 -> conv1 chain (event fork / join: capturable) */
} isb_unet_cfg;
typedef struct isb_unet isb_unet;
int isb_unet_create(const isb_unet_cfg* cfg, isb_unet** out);
/* name: the reference's parameter name ("input_blocks.3.0.in_layers.2.weight", ...); the tensor is copied.  Optional
 * extra entry "time_embed.freqs" [model_channels/2]: the host's own frequency table (nn.py:113-115). */
int isb_unet_load_weight(isb_unet* h, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim,
                         isb_stream_t stream);
/* Builds the plan and packs the weights; synchronises `stream`; names the first missing / mis-shaped parameter. */
int isb_unet_finalize(isb_unet* h, isb_stream_t stream);
size_t isb_unet_workspace_bytes(const isb_unet* h);
/* Zero-fills the workspace (arrival counters, statistics partials) — once, before the first pass. */
int isb_unet_workspace_init(isb_unet* h, void* workspace, size_t workspace_bytes, isb_stream_t stream);
int isb_unet_num_blocks(const isb_unet* h);   /* len(output_blocks): valid feat_layer values are [0, n) */
/* x_nchw [N,in_channels,H,W] fp32, t [N] fp32 timestep VALUES (after the respacing map).  out: [N,out_channels,H,W]
 * (out_nhwc = 0) or [N,H,W,out_channels] (out_nhwc = 1), may be NULL.  stop_at_feat = 1 (needs feat_layer >= 0):
 * return after output_blocks[feat_layer] — the guided step then runs isb_unet_forward_tail on a second stream beside
 * the backward pass (separate scratch and split-K workspace slots are reserved for exactly that). */
int isb_unet_forward(isb_unet* h, const float* x_nchw, const float* t, int feat_layer, int stop_at_feat, float* out,
                     int out_nhwc, void* workspace, size_t workspace_bytes, isb_stream_t stream);
int isb_unet_forward_tail(isb_unet* h, float* out, int out_nhwc, void* workspace, size_t workspace_bytes,
                          isb_stream_t stream);
/* The intermediate feature of unet.py:667-668 inside the workspace: val / grad are fp32 NHWC [dims] (grad NULL
 * without want_backward).  isb_drag_loss_grad reads val and writes grad in place. */
int isb_unet_feat(const isb_unet* h, void* workspace, int feat_layer, float** val, float** grad, int dims[4]);
/* d(sum(d_feat o inter_feat) + sum(d_out o out)) / dx -> dx_nchw [N,in_channels,H,W] fp32.  d_feat_nhwc: gradient of
 * the feature (copied in), or NULL with feat_grad_in_place = 1 when the caller already wrote it to isb_unet_feat's
 * grad; d_out_nchw [N,out_channels,H,W] or NULL.  Must follow the forward it differentiates. */
int isb_unet_backward_input(isb_unet* h, int feat_layer, const float* d_feat_nhwc, int feat_grad_in_place,
                            const float* d_out_nchw, float* dx_nchw, void* workspace, size_t workspace_bytes,
                            isb_stream_t stream);
void isb_unet_destroy(isb_unet* h);

/* ---- introspection ---------------------------------------------------- */
/* Number of kernels this library has launched in this process (all threads). */
uint64_t isb_launch_count(void);
/* Profiling only (tools/trace_conv.py): conv launches made with isb_conv_desc.debug_flags bit 2 write per-CTA
 * %globaltimer phase stamps ([cta][16] uint64 + a per-iteration area, 4096*16*8 bytes) to this device buffer. */
void isb_debug_set_trace(void* device_buffer);
/* Profiling only (tools/launch_floor.py): enqueue n dependent empty kernels (ctas x threads), optionally as
 * programmatic dependents — measures the floor of one dependent launch inside a stream / CUDA graph. */
int isb_debug_launch_chain(int n, int ctas, int threads, int pdl, isb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ISHAPE_B200_H */
