#!/usr/bin/env python
"""bench.py — drag-guided denoise steps/s (96x128^2 triplane latent) and edits/s on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

A "step" is one iteration of DragStuff.training's loop body (reference drag_utils.py:340-392):
UNet forward (feat_layer=8) -> drag loss + gradient -> UNet input-gradient backward -> fused DDPM
posterior/guidance update, batch 1, NFD architecture, 4 handles with r=12 (BASELINE.json configs[1]).
N>1 runs one independent replica per GPU (weak scaling: edits are independent).
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "drag-guided denoise steps/s (96x128^2 triplane, batch 1, NFD UNet, 4 handles r=12)"
UNIT = "steps/s"
W_TIME = 50          # guided steps per edit (BASELINE.json configs[1])
DECODE_RES = 256
# algorithmic FLOPs (2*MAC) of one guided step, SURVEY.md §8d: fwd 634.9 + dgrad 330.9 + attention bwd 23.7
STEP_GFLOP = 989.5


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:  # noqa: BLE001
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _problem(seed):
    """Synthetic edit: 4 handles, r=12 lattice, voxel 2/256."""
    import numpy as np
    rng = np.random.RandomState(seed)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    return src, tgt


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's step on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_steps(n_steps, n_warm, budget_s=150.0):
    import torch
    from oracle import nfd_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.NFD_CFG
    sd = O.synth_state_dict(cfg)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 96, 128, 128, generator=g)
    noise = torch.randn(1, 96, 128, 128, generator=g)
    origin = torch.randn(3, 170, 64, 64, generator=g)
    src, tgt = _problem(4)
    pg, sg, masks = O.drag_setup(src, tgt, 12, 2.0 / 256, 64)
    times = []
    t_start = time.time()
    i = W_TIME - 1
    for k in range(n_warm + n_steps):
        t0 = time.time()
        out = O.guided_step(sd, cfg, sched, x, i, origin, noise, pg, sg, masks, scale=600.0, cof=0.2)
        x = out["img"]
        dt = time.time() - t0
        if k >= n_warm:
            times.append(dt)
        i = i - 1 if i > 0 else W_TIME - 1
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    ms = 1e3 * sum(times) / len(times)
    return dict(value=1e3 / ms, ms_per_step=ms, cores=cores, done=len(times),
                sample=f"{len(times)} guided steps of the NFD 96x128x128 step (4 handles, r=12, fp32 torch CPU, "
                       f"input-gradient only) after {n_warm} warm-up")


def torch_eager_gpu_steps(dev, n_steps=3):
    """Informational comparator: the same oracle restatement (stock PyTorch ops + autograd, fp32, TF32 off)
    executed on the B200 — i.e. what the reference's own code path costs on this GPU (the
    reference has no Blackwell kernel; torch eager is the practical comparator)."""
    import torch
    from oracle import nfd_oracle as O

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = O.NFD_CFG
    sd = {k: v.to(dev) for k, v in O.synth_state_dict(cfg).items()}
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 96, 128, 128, generator=g).to(dev)
    noise = torch.randn(1, 96, 128, 128, generator=g).to(dev)
    origin = torch.randn(3, 170, 64, 64, generator=g).to(dev)
    src, tgt = _problem(4)
    pg, sg, masks = O.drag_setup(src, tgt, 12, 2.0 / 256, 64)
    pg, sg = pg.to(dev), sg.to(dev)
    ts = []
    for k in range(n_steps + 1):
        torch.cuda.synchronize()
        t0 = time.time()
        out = O.guided_step(sd, cfg, sched, x, W_TIME - 1 - k, origin, noise, pg, sg, masks, scale=600.0, cof=0.2)
        x = out["img"]
        torch.cuda.synchronize()
        if k > 0:
            ts.append(time.time() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    return {"value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms,
            "what": "oracle restatement (stock torch ops + autograd, fp32, TF32 off) run eagerly on the same B200"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["done"], "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "single drag-guided DDPM step, NFD UNet 96x128x128, batch 1, 4 handles (configs[1] step)",
                       "note": "reference is pure PyTorch and cannot travel to the GPU box; this is the oracle port "
                               "(oracle/nfd_oracle.py, pinned to the reference's outputs) on the host cores"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from oracle import nfd_oracle as O   # weights only (synth_state_dict); the oracle never computes here
    from ishapediting_b200 import _lib
    from ishapediting_b200.drag_utils import DragGeometry, DragStuff, GuidedStepper, get_args

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed): model, feature cache, geometry ----
    a = get_args(["--num_steps", "200", "--w_time", str(W_TIME), "--shape_resolution", str(DECODE_RES)])
    a.use_fp16 = (args.mode == "bf16")
    ds = DragStuff(args=a, device=dev, use_graph=not args.no_graph)
    ds.model.load_state_dict(O.synth_state_dict(O.NFD_CFG))
    ds.model.to(dev).eval()
    w, planes = O.synth_decoder()
    ds.decoder.net[0]._B.data.copy_(w["B"])
    for idx, k in ((1, "1"), (3, "2"), (5, "3")):
        ds.decoder.net[idx].weight.data.copy_(w["w" + k])
        ds.decoder.net[idx].bias.data.copy_(w["b" + k])
    torch.manual_seed(100 + rank)
    t0 = time.time()
    ds.update_latent_params(torch.randn(1, 96, 128, 128, device=dev))       # 200 no-grad steps + feature cache
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    src, tgt = _problem(4 + rank)
    S, Ca = ds.feature_guidance[0].shape[1], ds.feature_guidance[0].shape[3]
    geo = DragGeometry(src, tgt, ds.r1, ds.voxel_size, S, Ca)
    st = GuidedStepper(ds.model, ds.diffusion, geo, a.feat_layer, 0.2, "l2", 600.0, use_graph=not args.no_graph)
    st.img.copy_(ds.w)

    def step_index(k):
        return W_TIME - 1 - (k % W_TIME)

    # ---- warm-up (also: eager step -> launches per step, graph capture) ----
    lc0 = _lib.launch_count()
    st.step(step_index(0), ds.feature_guidance[0])
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - lc0
    for k in range(1, max(args.warmup, 3)):
        st.step(step_index(k), ds.feature_guidance[k % W_TIME])
    barrier()

    # ---- timed: K steps, inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        st.step(step_index(k), ds.feature_guidance[k % W_TIME])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: same step through host buffers (pinned), H2D of the step's inputs + D2H of its result ----
    # HostStepPipeline (the product's host-buffer driver): step k+1's inputs travel while step k computes, step k's
    # latent + loss travel back while step k+1 computes; the caller reads every step's result (one step late).
    from ishapediting_b200.drag_utils import HostStepPipeline

    n_e2e = args.steps
    origin_host = [f.cpu().pin_memory() for f in ds.feature_guidance[:min(W_TIME, n_e2e)]]
    noise_host = torch.randn(1, 96, 128, 128).pin_memory()
    pipe = HostStepPipeline(st)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    checksum = 0.0
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    pipe.prefetch(0, origin_host[0], noise_host)
    for k in range(n_e2e):
        if k + 1 < n_e2e:
            pipe.prefetch(k + 1, origin_host[(k + 1) % len(origin_host)], noise_host)
        pipe.run(k, step_index(k))
        if k:
            checksum += float(pipe.result(k - 1)[1])        # the caller reads the step's result
    checksum += float(pipe.result(n_e2e - 1)[1])
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    # ---- one full edit: 50 guided steps + 256^3 decode through DragStuff.training (edits/s) ----
    ds.use_graph = not args.no_graph
    for _ in ds.training(src, tgt, scale=600, cof=0.2):     # first edit of a session: captures the step graph
        pass
    src2, tgt2 = _problem(1000 + rank)                       # timed: a NEW edit (other handles), graph reused
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for _ in ds.training(src2, tgt2, scale=600, cof=0.2):
        pass
    e5.record()
    barrier()
    ms_edit = e4.elapsed_time(e5)

    # ---- throughput mode: B independent edits advanced as one batch (BASELINE configs[4], "batched" variant) ----
    batched = None
    if args.batch > 1:
        Bn = args.batch
        geos = [DragGeometry(*_problem(2000 + rank * 64 + b), ds.r1, ds.voxel_size, S, Ca) for b in range(Bn)]
        stb = GuidedStepper(ds.model, ds.diffusion, geos, a.feat_layer, 0.2, "l2", 600.0, use_graph=not args.no_graph)
        stb.img.copy_(ds.w.expand(Bn, -1, -1, -1))
        origin_b = [torch.stack([ds.feature_guidance[k]] * Bn) for k in range(min(W_TIME, 8))]
        for k in range(3):
            stb.step(step_index(k), origin_b[k % len(origin_b)])
        barrier()
        e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nb = max(10, args.steps // 2)
        e6.record()
        for k in range(nb):
            stb.step(step_index(k), origin_b[k % len(origin_b)])
        e7.record()
        barrier()
        ms_b = e6.elapsed_time(e7)
        tb = torch.tensor([ms_b], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        ms_b = float(tb.item())
        batched = {"batch_per_gpu": Bn, "edit_steps_per_s": world * Bn * nb / (ms_b / 1e3), "ms_per_batched_step": ms_b / nb,
                   "what": f"{Bn} independent edits per GPU advanced as one batch-{Bn} guided step"}
        del stb, origin_b
        torch.cuda.empty_cache()

    # ---- instrumented pass: CUDA events around every conv launch (roofline of the dominant kernel) ----
    roof = None
    if rank == 0:
        roof = conv_roofline(st, ds, step_index, geo, a.feat_layer, not args.no_graph, ms / args.steps)

    # ---- reduce over ranks: max time ----
    times = torch.tensor([ms, ms_e2e, ms_edit], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_edit = (float(v) for v in times.tolist())

    if rank == 0:
        peaks = _peaks()
        value = world * args.steps / (ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
            "config": {"workload": "guided DDPM step of the 50-step drag edit (BASELINE configs[1]): NFD UNet "
                                   "(421M params) on a 1x96x128x128 latent, feat_layer=8, 4 handles r=12, l2 loss",
                       "replicas": world, "cuda_graph": not args.no_graph,
                       "l2_policy": "working set > L2: 1.57 GB of bf16 weight panels streamed per step vs 126 MB L2",
                       "weights": "synthetic seeded (oracle.synth_state_dict)", "setup_s": round(setup_s, 2)},
            "e2e": {"value": world * n_e2e / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / n_e2e,
                    "how": "HostStepPipeline: pinned host buffers; every step's inputs are copied H2D and its latent + "
                           "loss D2H inside the timed region and read by the host; the copies of step k+1 / k-1 run on "
                           "copy streams while step k computes (synchronous copies cost +0.39 ms/step)",
                    "loss_checksum": checksum},
            "edit": {"edits_per_s": world / (ms_edit / 1e3), "ms_per_edit": ms_edit,
                     "what": f"{W_TIME} guided steps + {DECODE_RES}^3 occupancy decode via DragStuff.training"},
            "step_tflops": STEP_GFLOP / (ms / args.steps),
            "step_frac_of_bf16_sustained": STEP_GFLOP / (ms / args.steps) / peaks["tf_sustained"],
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "roofline": roof,
            "batched": batched,
        }
        if world == 1 and args.torch_eager:      # opt-in, informational: the oracle restatement run eagerly on the GPU
            try:
                line["torch_eager_b200"] = torch_eager_gpu_steps(dev)
            except Exception as e:  # noqa: BLE001
                line["torch_eager_b200"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(2, 1, budget_s=60.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def conv_roofline(st, ds, step_index, geo, feat_layer, use_graph, ms_full_step):
    """Roofline of the dominant kernel family (conv_tc_kernel / conv_tc2_kernel: every 3x3 / 1x1 convolution,
    linear layer and their backward-data — 205 launches, 958.8 algorithmic GFLOP per step).

    Time in the kernel is measured live, with CUDA events on the launch stream, as the difference between the
    timed graph-replayed step and the same step re-captured with the conv launches removed (everything else
    identical; the values it then computes are garbage and are discarded).  Per-launch event pairs in eager mode
    were tried first and rejected: the host needs longer to encode three tensor maps and enqueue a launch than the
    small layers run, so eager event pairs mostly time host gaps.  FLOPs = sum over launches of 2*M*Cout*K
    (K = k*k*Cin (+Cin_skip)), counted from the shapes of one recorded step."""
    import torch
    from ishapediting_b200.drag_utils import GuidedStepper

    ops = st.ops
    orig = ops.conv
    rec = []

    def count_conv(a, w, bias, ksize, out, a2=None, residual=None, accumulate=False, tune=None, **kw):
        M = a.shape[0] * a.shape[1] * a.shape[2]
        rec.append(2.0 * M * w.shape[0] * w.shape[1])

    ops.conv = count_conv
    try:
        st2 = GuidedStepper(ds.model, ds.diffusion, geo, feat_layer, 0.2, "l2", 600.0, use_graph=use_graph)
        st2.img.copy_(ds.w)
        st2.step(step_index(0), ds.feature_guidance[0])
        flops, launches = sum(rec), len(rec)
        for k in range(1, 4):
            st2.step(step_index(k), ds.feature_guidance[k % W_TIME])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for k in range(reps):
            st2.step(step_index(k), ds.feature_guidance[k % W_TIME])
        e1.record()
        torch.cuda.synchronize()
        ms_noconv = e0.elapsed_time(e1) / reps
    finally:
        ops.conv = orig
    t_ms = max(ms_full_step - ms_noconv, 1e-6)
    peaks = _peaks()
    ach = flops / (t_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "conv_tc_kernel + conv_tc2_kernel (tcgen05 implicit-GEMM conv / linear / dgrad)",
            "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sustained"],
            "peak_source": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a step)",
            "launches": launches, "gflop_per_step": flops / 1e9, "ms_in_kernel_per_step": t_ms,
            "avg_launch_us": 1e3 * t_ms / max(launches, 1), "ms_step_without_kernel": ms_noconv,
            "method": "graph-replayed step minus the same step captured without the conv launches (CUDA events)",
            # dram__bytes_read+write of one captured launch (profiles/r01_ncu_full_conv_tc.md): the 3x3 256->256 @128x128
            # layer moves 9.63 MB = its algorithmic bytes (8.39 MB bf16 activations + 1.18 MB weights; the fp32 output
            # stays in L2)
            "traffic": 9.61e6, "traffic_launch": "3x3 conv 256->256 @128x128 (19.3 GFLOP), profiles/r01_ncu_full_conv_tc_v3.md"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--batch", type=int, default=8, help="extra throughput leg: edits per GPU advanced as one batch (0/1 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-eager", action="store_true",
                    help="also time the oracle restatement (stock torch ops, fp32) eagerly on the GPU: an informational "
                         "comparator, 14.5 steps/s on B200 (profiles/r01_bench_n1_latest.json)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
