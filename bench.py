#!/usr/bin/env python
"""bench.py — drag-guided denoise steps/s (96x128^2 triplane latent) and edits/s on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own loop on the host cores

A "step" is one iteration of DragStuff.training's loop body (reference drag_utils.py:340-392):
UNet forward (feat_layer=8) -> drag loss + gradient -> UNet input-gradient backward -> fused DDPM
posterior/guidance update, batch 1, NFD architecture, 4 handles with r=12 (BASELINE.json configs[1]).
N>1 runs one independent replica per GPU (weak scaling: edits are independent), plus the two
sharded legs of BASELINE configs[3] (x-slab decode sweep + NCCL all-gather) and configs[4] (edits dealt round-robin,
results gathered over NCCL).  One JSON line is printed by rank 0.
"""
import argparse
import csv
import glob
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
WORKLOAD = ("guided DDPM step of the 50-step drag edit (BASELINE configs[1]): NFD UNet (421M params) on a 1x96x128x128 "
            "latent, feat_layer=8, 4 handles r=12, l2 loss")
W_TIME = 50          # guided steps per edit (BASELINE.json configs[1])
DECODE_RES = 256
# algorithmic FLOPs (2*MAC) of one guided step, SURVEY.md §8d: fwd 634.9 + dgrad 330.9 + attention bwd 23.7
STEP_GFLOP = 989.5
DECODE_FLOP_PER_POINT = 69888.0      # MLP 2*MAC per query point


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:  # noqa: BLE001
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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


def _step_inputs():
    """The seeded latent / noise / origin feature shared by every host-side and eager comparator."""
    import torch
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 96, 128, 128, generator=g)
    noise = torch.randn(1, 96, 128, 128, generator=g)
    origin = torch.randn(3, 170, 64, 64, generator=g)
    return x, noise, origin


# ------------------------------------------------------------------------------------------------------
# the reference's own loop (oracle/_ref snapshot of /root/reference, unmodified) and the oracle port
# ------------------------------------------------------------------------------------------------------
def reference_training_steps(device, n_steps, n_warm, budget_s=150.0, fp16=False, tf32=False, threads=None):
    """Time the UNMODIFIED reference's DragStuff.training loop body (drag_utils.py:336-398: p_sample_guidance,
    grid_sample motion loss, loss.backward() with weight gradients, guided update) on `device`.  The harness only
    supplies what the reference reads from disk / the GUI: seeded synthetic weights, latent, cached origin features
    and handles, and a no-op get_mesh.  fp16=True is the reference's GPU default (use_fp16 + convert_to_fp16,
    drag_utils.py:52,231-232)."""
    import torch
    import torch.distributed as dist
    from oracle import nfd_oracle as O
    from oracle import ref_import as R

    if not R.available():
        raise RuntimeError("reference snapshot unavailable (neither /root/reference nor oracle/_ref)")
    if threads:
        torch.set_num_threads(threads)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    if not dist.is_initialized():
        # The reference's setup_dist() would resolve the container hostname and rendezvous over TCP; under torchrun a
        # tcp:// rendezvous is redirected to the elastic agent's store (TORCHELASTIC_USE_AGENT_STORE) and never
        # completes for a private single-rank group.  An in-process HashStore needs no network at all.
        dist.init_process_group("gloo", store=dist.HashStore(), rank=0, world_size=1)
    ns = R.import_reference()
    du = ns.drag_utils
    if du is None:
        raise RuntimeError("reference drag_utils not importable: " + getattr(ns, "drag_utils_error", "?"))
    dev = torch.device(device)
    du.dist_util.dev = lambda: dev
    a = du.DragStuff.args
    a.use_fp16 = bool(fp16)
    a.num_steps, a.timestep_respacing, a.w_time = 200, "200", n_warm + n_steps
    a.feat_layer, a.loss_type = 8, "l2"
    ds = du.DragStuff()
    ds.model.load_state_dict(O.synth_state_dict(O.NFD_CFG), strict=True)
    if fp16:
        ds.model.convert_to_fp16()
    ds.model.eval()
    x, _, origin = _step_inputs()
    ds.w = x.to(dev)
    ds.feature_guidance = [origin] * a.w_time            # host tensors, moved per step like the reference (:352)
    ds.r1, ds.offset1, ds.voxel_size = 12, du.make_offsets(12, dev), 2.0 / 256
    ds.get_mesh = lambda **kw: None
    src, tgt = _problem(4)
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    times, t_start = [], time.time()
    sync()
    t0 = time.time()
    for k, _ in enumerate(ds.training(src, tgt, scale=600, cof=0.2)):
        sync()
        now = time.time()
        if k >= n_warm:          # step 0 also carries the one-time Python-set mask setup (:322-334): warm-up
            times.append(now - t0)
        t0 = now
        if now - t_start > budget_s and len(times) >= 1:
            ds.train_flag = False
    ms = 1e3 * sum(times) / len(times)
    return dict(value=1e3 / ms, ms_per_step=ms, done=len(times), kind="reference",
                sample=f"{len(times)} iterations of the reference's own DragStuff.training loop (oracle/_ref, unmodified; "
                       f"NFD 96x128x128, 4 handles, r=12, {'fp16 torso' if fp16 else 'fp32'}, autograd incl. weight "
                       f"gradients) after {n_warm} warm-up")


def port_steps(device, n_steps, n_warm, budget_s=150.0, threads=None, autocast=False, tf32=False):
    """The oracle restatement (oracle/nfd_oracle.py: stock torch ops + autograd, input gradient only)."""
    import torch
    from oracle import nfd_oracle as O

    if threads:
        torch.set_num_threads(threads)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    dev = torch.device(device)
    cfg = O.NFD_CFG
    sd = {k: v.to(dev) for k, v in O.synth_state_dict(cfg).items()}
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    x, noise, origin = (t.to(dev) for t in _step_inputs())
    src, tgt = _problem(4)
    pg, sg, masks = O.drag_setup(src, tgt, 12, 2.0 / 256, 64)
    pg, sg = pg.to(dev), sg.to(dev)
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    times, t_start, i = [], time.time(), W_TIME - 1
    for k in range(n_warm + n_steps):
        sync()
        t0 = time.time()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and dev.type == "cuda"):
            out = O.guided_step(sd, cfg, sched, x, i, origin, noise, pg, sg, masks, scale=600.0, cof=0.2)
        x = out["img"].float()
        sync()
        if k >= n_warm:
            times.append(time.time() - t0)
        i = i - 1 if i > 0 else W_TIME - 1
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    ms = 1e3 * sum(times) / len(times)
    return dict(value=1e3 / ms, ms_per_step=ms, done=len(times), kind="port",
                sample=f"{len(times)} guided steps of the oracle port (NFD 96x128x128, 4 handles, r=12, fp32 torch, "
                       f"input-gradient only) after {n_warm} warm-up")


def cpu_reference_steps(n_steps, n_warm, budget_s=150.0):
    """Host-core baseline: the reference's own loop when its snapshot travelled here, else the oracle port."""
    cores = os.cpu_count() or 1
    try:
        r = reference_training_steps("cpu", n_steps, n_warm, budget_s, threads=cores)
    except Exception as e:  # noqa: BLE001
        r = port_steps("cpu", n_steps, n_warm, budget_s, threads=cores)
        r["fallback_reason"] = repr(e)[:160]
    r["cores"] = cores
    return r


def torch_eager_gpu(dev):
    """What the reference's code path costs on THIS GPU (the reference has no Blackwell kernel;
    stock PyTorch eager on the same box is the practical comparator).  Two numbers: strict fp32 (TF32 off), and the
    reference's own GPU default — fp16 torso via convert_to_fp16() with TF32 on."""
    import torch
    out = {}
    for name, kw in (("fp32_tf32_off", dict(fp16=False, tf32=False)), ("fp16_torso_tf32_on", dict(fp16=True, tf32=True))):
        try:
            r = reference_training_steps(str(dev), 4, 2, budget_s=30.0, **kw)
            out[name] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "kind": "reference",
                         "what": "the reference's own DragStuff.training loop, stock PyTorch eager on the same B200"}
        except Exception as e:  # noqa: BLE001
            try:
                r = port_steps(str(dev), 4, 2, budget_s=30.0, autocast=kw["fp16"], tf32=kw["tf32"])
                out[name] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "kind": "port",
                             "what": "oracle port (stock torch ops + autograd" + (", bf16 autocast" if kw["fp16"] else "")
                                     + ") eagerly on the same B200", "reference_error": repr(e)[:160]}
            except Exception as e2:  # noqa: BLE001
                out[name] = {"error": repr(e2)[:200]}
        torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every worker; the reference arm is ONE process that owns all host cores.
    # Must happen before torch is imported (OpenMP / MKL read it at load time).
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count() or 1)
    r = cpu_reference_steps(args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["done"], "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "note": "host cores only; kind=reference: the unmodified reference modules (oracle/_ref) drive "
                               "their own DragStuff.training loop; kind=port: oracle/nfd_oracle.py"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# ncu evidence committed under profiles/: per-kernel DRAM traffic of one guided step
# ------------------------------------------------------------------------------------------------------
def ncu_traffic():
    """{family: {launches, dram_bytes_per_launch, ...}} from the newest profiles/r*_ncu_step_traffic.json (written by
    tools/ncu_summarize.py from the ncu CSV of one eager guided step, committed next to it)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_step_traffic.json")))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            return json.load(f), os.path.relpath(files[-1], ROOT)
    except Exception:  # noqa: BLE001
        return None, None


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    from oracle import nfd_oracle as O   # weights only (synth_state_dict); the oracle never computes here
    from ishapediting_b200 import _lib, parallel
    from ishapediting_b200.drag_utils import DragGeometry, DragStuff, GuidedStepper, HostStepPipeline, get_args
    from ishapediting_b200.triplane_decoder.visualize import query_volume

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

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def timed(fn):
        """ms of fn() on this rank, CUDA events on the current stream, barrier + synchronize on both sides."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        r = fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1), r

    # ---- setup (untimed): model, feature cache, geometry ----
    a = get_args(["--num_steps", "200", "--w_time", str(W_TIME), "--shape_resolution", str(DECODE_RES)])
    a.use_fp16 = (args.mode == "bf16")
    ds = DragStuff(args=a, device=dev, use_graph=not args.no_graph)
    ds.mesh_on_device = False        # configs[1] times the steps + the logit volume; meshing is its own leg below
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
    n_warm = max(args.warmup, 3)
    for k in range(1, n_warm):
        st.step(step_index(k), ds.feature_guidance[k % W_TIME])
    barrier()

    # ---- timed: K steps, inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    def k_steps():
        for k in range(args.steps):
            st.step(step_index(k), ds.feature_guidance[k % W_TIME])

    ms, _ = timed(k_steps)
    if rank == 0 and ms < 400.0:          # nvidia-smi samples every 50 ms: keep the GPU under the same load a little longer
        for k in range(int(400.0 / max(ms / args.steps, 1e-3)) + 1):
            st.step(step_index(k), ds.feature_guidance[k % W_TIME])
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: same step through host buffers (pinned), H2D of the step's inputs + D2H of its result ----
    n_e2e = args.steps
    origin_host = [f.cpu().pin_memory() for f in ds.feature_guidance[:min(W_TIME, n_e2e)]]
    noise_host = torch.randn(1, 96, 128, 128).pin_memory()
    pipe = HostStepPipeline(st)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    box = {"checksum": 0.0}

    def e2e_steps():
        pipe.prefetch(0, origin_host[0], noise_host)
        for k in range(n_e2e):
            if k + 1 < n_e2e:
                pipe.prefetch(k + 1, origin_host[(k + 1) % len(origin_host)], noise_host)
            pipe.run(k, step_index(k))
            if k:
                box["checksum"] += float(pipe.result(k - 1)[1])        # the caller reads the step's result
        box["checksum"] += float(pipe.result(n_e2e - 1)[1])

    ms_e2e, _ = timed(e2e_steps)

    # ---- one full edit: 50 guided steps + 256^3 decode through DragStuff.training (edits/s) ----
    ds.use_graph = not args.no_graph
    for _ in ds.training(src, tgt, scale=600, cof=0.2):     # first edit of a session: captures the step graph
        pass
    src2, tgt2 = _problem(1000 + rank)                       # timed: a NEW edit (other handles), graph reused

    def one_edit():
        for _ in ds.training(src2, tgt2, scale=600, cof=0.2):
            pass

    ms_edit, _ = timed(one_edit)

    # ---- BASELINE configs[4]: 8 edits per GPU dealt round-robin, final latents gathered over NCCL ----
    n_edits = 8 * world
    mine = parallel.assign_edits(n_edits, rank, world)
    lat = torch.empty((len(mine), 96, 128, 128), device=dev)

    def many_edits():
        for j, e in enumerate(mine):
            s_e, t_e = _problem(3000 + e)
            for _ in ds.training(s_e, t_e, scale=600, cof=0.2):
                pass
            lat[j].copy_(ds.stepper.img[0])
        torch.cuda.current_stream().synchronize()
        tg = time.time()
        allr = parallel.gather_results(lat, n_edits)
        torch.cuda.current_stream().synchronize()
        return allr, 1e3 * (time.time() - tg)

    ms_many, (all_lat, ms_gather_edits) = timed(many_edits)
    edits_ok = bool(all_lat.shape[0] == n_edits and torch.isfinite(all_lat).all()
                    and torch.equal(all_lat[rank], lat[0]))
    del all_lat

    # ---- BASELINE configs[3]: x-slab sharded decode sweep 128^3 / 256^3 / 512^3 + in-place all-gather ----
    for p in range(3):
        ds.decoder.embeddings[p] = planes[[p]].to(dev)
    decode = {}
    for res in (128, 256, 512):
        vol = torch.empty((res, res, res), device=dev)
        b, e = parallel.slab_range(res, rank, world)
        query_volume(ds.decoder, 0, res, b, e, out=vol[b:e])      # warm-up (planes cache, clocks)
        ms_slab, _ = timed(lambda: query_volume(ds.decoder, 0, res, b, e, out=vol[b:e]))
        ms_gather, _ = timed(lambda: parallel.complete_volume_(vol))
        ms_total, _ = timed(lambda: parallel.decode_volume_sharded(
            lambda xb, xe, o: query_volume(ds.decoder, 0, res, xb, xe, out=o), res, dev, out=vol))
        exact = None
        if res <= 256:           # bit-exactness of the sharded volume against the un-sharded decode
            full = query_volume(ds.decoder, 0, res).reshape(res, res, res)
            exact = bool(torch.equal(full, vol))
            del full
        m = max_over_ranks([ms_slab, ms_gather, ms_total])
        decode[str(res)] = {"ms_slab_decode": m[0], "ms_gather": m[1], "ms_total": m[2],
                            "gpoints_per_s": res ** 3 / (m[2] * 1e-3) / 1e9, "bit_exact_vs_single_gpu": exact,
                            "occupancy": float((vol > 0).float().mean())}
        del vol
        torch.cuda.empty_cache()

    # ---- meshing tail: marching cubes + 10 Laplacian iterations on the 256^3 volume ----
    from ishapediting_b200.triplane_decoder.visualize import create_obj_o3d
    vol256 = query_volume(ds.decoder, 0, DECODE_RES)
    create_obj_o3d(ds.decoder, 0, res=DECODE_RES, volume=vol256).filter_smooth_simple(10)      # warm-up
    ms_mc, mesh_dev = timed(lambda: create_obj_o3d(ds.decoder, 0, res=DECODE_RES, volume=vol256))
    ms_smooth, _ = timed(lambda: mesh_dev.filter_smooth_simple(10))
    mesh_leg = {"res": DECODE_RES, "vertices": int(mesh_dev.vertices.shape[0]), "triangles": int(mesh_dev.triangles.shape[0]),
                "ms_marching_cubes": ms_mc, "ms_smooth_10_iterations": ms_smooth,
                "what": "device marching cubes (incl. the read-back of the two counts) and filter_smooth_simple(10) on the "
                        "synthetic 24 %-occupancy field; the reference runs PyMCubes / Open3D on the CPU here"}
    del vol256, mesh_dev
    torch.cuda.empty_cache()

    # ---- throughput mode: B independent edits advanced as one batch (configs[4], "batched" variant) ----
    batched = None
    if args.batch > 1:
        Bn = args.batch
        geos = [DragGeometry(*_problem(2000 + rank * 64 + b), ds.r1, ds.voxel_size, S, Ca) for b in range(Bn)]
        stb = GuidedStepper(ds.model, ds.diffusion, geos, a.feat_layer, 0.2, "l2", 600.0, use_graph=not args.no_graph)
        stb.img.copy_(ds.w.expand(Bn, -1, -1, -1))
        origin_b = [torch.stack([ds.feature_guidance[k]] * Bn) for k in range(min(W_TIME, 8))]
        for k in range(3):
            stb.step(step_index(k), origin_b[k % len(origin_b)])
        nb = max(10, args.steps // 2)

        def b_steps():
            for k in range(nb):
                stb.step(step_index(k), origin_b[k % len(origin_b)])

        ms_b, _ = timed(b_steps)
        ms_b = max_over_ranks([ms_b])[0]
        batched = {"batch_per_gpu": Bn, "edit_steps_per_s": world * Bn * nb / (ms_b / 1e3), "ms_per_batched_step": ms_b / nb,
                   "step_tflops": Bn * STEP_GFLOP / (ms_b / nb),
                   "what": f"{Bn} independent edits per GPU advanced as one batch-{Bn} guided step"}
        del stb, origin_b
        torch.cuda.empty_cache()
        # the same as whole edits through the product API: Bn x (50 steps + 256^3 decode) per GPU, latents gathered
        edits_b = [_problem(5000 + rank * 64 + b) for b in range(Bn)]
        ds.training_batch(edits_b, scale=600, cof=0.2)                      # warm-up: captures the batch-Bn graph
        edits_b = [_problem(6000 + rank * 64 + b) for b in range(Bn)]

        def batch_edits():
            lat_b, vols = ds.training_batch(edits_b, scale=600, cof=0.2)
            return parallel.gather_results(lat_b, Bn * world)

        ms_be, all_b = timed(batch_edits)
        ms_be = max_over_ranks([ms_be])[0]
        batched.update(edits_per_s=world * Bn / (ms_be / 1e3), ms_per_batch_of_edits=ms_be,
                       edits_what=f"BASELINE configs[4], batched variant: {Bn} whole edits per GPU ({W_TIME} batch-{Bn} steps + "
                                  f"{Bn} x {DECODE_RES}^3 decodes) via DragStuff.training_batch, latents all-gathered",
                       gathered_ok=bool(all_b.shape[0] == Bn * world and torch.isfinite(all_b).all()))
        del all_b
        ds.batch_stepper = None
        torch.cuda.empty_cache()

    # ---- rooflines (rank 0): ablation of one kernel family out of the graph-replayed step ----
    roof = roof_gn = roof_dec = None
    if rank == 0:
        ms_step = ms / args.steps
        traffic, traffic_src = ncu_traffic()
        roof, roof_gn = family_rooflines(st, ds, step_index, geo, a.feat_layer, not args.no_graph, ms_step, traffic,
                                         traffic_src)
        roof_dec = decode_roofline(ds, query_volume, dev, traffic, traffic_src)

    # ---- reduce over ranks: max time ----
    ms, ms_e2e, ms_edit, ms_many, ms_gather_edits = max_over_ranks([ms, ms_e2e, ms_edit, ms_many, ms_gather_edits])

    if rank == 0:
        peaks = _peaks()
        value = world * args.steps / (ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "replicas": world, "cuda_graph": not args.no_graph,
                       "l2_policy": "working set > L2: 1.57 GB of bf16 weight panels streamed per step vs 126 MB L2",
                       "weights": "synthetic seeded (oracle.synth_state_dict)", "setup_s": round(setup_s, 2)},
            "e2e": {"value": world * n_e2e / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / n_e2e,
                    "how": "HostStepPipeline: pinned host buffers; every step's inputs are copied H2D and its latent + "
                           "loss D2H inside the timed region and read by the host; the copies of step k+1 / k-1 run on "
                           "copy streams while step k computes",
                    "loss_checksum": box["checksum"]},
            "edit": {"edits_per_s": world / (ms_edit / 1e3), "ms_per_edit": ms_edit,
                     "what": f"{W_TIME} guided steps + {DECODE_RES}^3 occupancy decode via DragStuff.training"},
            "edits_sharded": {"n_edits": n_edits, "edits_per_s": n_edits / (ms_many / 1e3), "ms_total": ms_many,
                              "ms_gather": ms_gather_edits, "gathered_ok": edits_ok,
                              "what": "BASELINE configs[4]: 8 edits per GPU dealt round-robin (parallel.assign_edits), "
                                      "final latents collected with one all_gather_into_tensor (parallel.gather_results)"},
            "mesh": mesh_leg,
            "decode_sharded": dict(decode, what="BASELINE configs[3]: dense decode, x-slabs over the ranks written "
                                                "straight into the full volume + one in-place all_gather_into_tensor; "
                                                "times are max over ranks"),
            "step_tflops": STEP_GFLOP / (ms / args.steps),
            "step_frac_of_bf16_sustained": STEP_GFLOP / (ms / args.steps) / peaks["tf_sustained"],
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "roofline": roof,
            "roofline_hbm": roof_gn,
            "roofline_decode": roof_dec,
            "batched": batched,
        }
        if world == 1 and not args.no_eager:
            line["torch_eager_b200"] = torch_eager_gpu(dev)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(2, 1, budget_s=40.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _family_of(name):
    n = name.lower()
    if "conv_tc" in n:
        return "conv"
    if n.startswith("gn_") or "::gn_" in n:
        return "groupnorm"
    if "fa_" in n:
        return "attention"
    if "triplane_decode" in n or "decode_tc" in n:
        return "decode"
    if "drag_" in n:
        return "drag"
    return "other"


def family_rooflines(st, ds, step_index, geo, feat_layer, use_graph, ms_full_step, traffic, traffic_src):
    """Rooflines of the two dominant kernel families of the guided step.

    conv (conv_tc_kernel / conv_tc2_kernel: every 3x3 / 1x1 convolution, linear layer and their backward-data),
    tensor-bound: FLOPs = sum over launches of 2*M*Cout*K (K = k*k*Cin (+Cin_skip)), counted from the shapes of one
    recorded step.  GroupNorm family (gn_* kernels: GroupNorm + FiLM + SiLU + resample + cat and their backward),
    HBM-bound: algorithmic bytes = every input tensor read once + every output tensor written once at its storage
    dtype, counted the same way.
    Time in a family is measured live, with CUDA events on the launch stream, as the difference between the timed
    graph-replayed step and the same step re-captured with that family's launches removed (everything else identical;
    the values it then computes are garbage and are discarded).  Per-launch event pairs in eager mode were rejected
    in round 1: the host needs longer to enqueue a launch than the small layers run.  `traffic` = measured
    dram__bytes_read+write per launch of the family, from the committed ncu capture of one eager step."""
    import torch
    from ishapediting_b200.drag_utils import GuidedStepper

    ops = st.ops
    peaks = _peaks()
    bpe = lambda t: 0 if t is None else t.numel() * t.element_size()  # noqa: E731

    def ablate(patches, record):
        """ms per step with `patches` applied; record[:] is cut to what ONE step recorded."""
        saved = {k: getattr(ops, k) for k in patches}
        for k, v in patches.items():
            setattr(ops, k, v)
        try:
            st2 = GuidedStepper(ds.model, ds.diffusion, geo, feat_layer, 0.2, "l2", 600.0, use_graph=use_graph)
            st2.img.copy_(ds.w)
            st2.step(step_index(0), ds.feature_guidance[0])
            one_step = len(record)
            counted = True
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
            del record[one_step:]
            return e0.elapsed_time(e1) / reps if counted else None
        finally:
            for k, v in saved.items():
                setattr(ops, k, v)

    # -- conv --
    rec = []

    def count_conv(a, w, bias, ksize, out, a2=None, residual=None, accumulate=False, tune=None, **kw):
        rec.append(2.0 * a.shape[0] * a.shape[1] * a.shape[2] * w.shape[0] * w.shape[1])

    ms_noconv = ablate({"conv": count_conv}, rec)
    n_launch_rec = len(rec)
    flops = sum(rec)
    t_ms = ms_full_step - ms_noconv
    tr = (traffic or {}).get("conv")
    conv = {"bound": "tensor", "kernel": "conv_tc_kernel + conv_tc2_kernel (tcgen05 implicit-GEMM conv / linear / dgrad)",
            "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "peak_source": f"{peaks['src']} bf16_tflops_sustained (kernel timed inside a step)",
            "launches": n_launch_rec, "gflop_per_step": flops / 1e9, "ms_step_without_kernel": ms_noconv,
            "method": "graph-replayed step minus the same step captured without the conv launches (CUDA events)",
            "traffic": tr["dram_bytes_per_launch"] if tr else None,
            "traffic_source": (f"{traffic_src}: mean dram__bytes_read.sum + dram__bytes_write.sum over the "
                               f"{tr['launches']} conv launches of one eager step") if tr else None,
            "algorithmic_bytes_per_launch": tr.get("algorithmic_bytes_per_launch") if tr else None}
    if t_ms > 0.05 * ms_full_step:          # guard: a noisy ablation (t <= 0) must not print a fantasy fraction
        ach = flops / (t_ms * 1e-3) / 1e12
        conv.update(achieved=ach, frac=ach / peaks["tf_sustained"], ms_in_kernel_per_step=t_ms,
                    avg_launch_us=1e3 * t_ms / max(n_launch_rec, 1))
    else:
        conv.update(achieved=None, frac=None, ms_in_kernel_per_step=None, note="ablation not resolvable on this run")

    # -- GroupNorm family --
    gbytes = []

    def count_gn_fwd(x1, x2, gamma, beta, film, film_off, silu, resample, stats, y, raw=None, xres=None, partials=None):
        gbytes.append(bpe(x1) + bpe(x2) + bpe(y) + bpe(raw) + bpe(xres))
        return y

    def count_gn_bwd(x1, x2, gamma, beta, film, film_off, silu, resample, stats, dy, gres, gres_at_input,
                     gx1, acc1, gx1_lo, gx2, acc2, gx2_lo, partials=None):
        gbytes.append(bpe(x1) + bpe(x2) + bpe(dy) + bpe(gres) + bpe(gx1) + bpe(gx1_lo) + bpe(gx2) + bpe(gx2_lo)
                      + (bpe(gx1) if acc1 else 0) + (bpe(gx2) if acc2 else 0))

    ms_nogn = ablate({"gn_forward": count_gn_fwd, "gn_backward": count_gn_bwd}, gbytes)
    n_gn = len(gbytes)
    gb = sum(gbytes)
    t_gn = ms_full_step - ms_nogn
    trg = (traffic or {}).get("groupnorm")
    gn = {"bound": "hbm", "kernel": "gn_* (GroupNorm + FiLM + SiLU + resample + concat, forward and backward)",
          "peak": peaks["hbm"], "unit": "GB/s", "peak_source": f"{peaks['src']} hbm_gbs",
          "calls": n_gn, "algorithmic_mb_per_step": gb / 1e6, "ms_step_without_kernel": ms_nogn,
          "method": "graph-replayed step minus the same step captured without the GroupNorm launches (CUDA events); "
                    "algorithmic bytes = inputs read once + outputs written once at their storage dtype",
          "traffic": trg["dram_bytes_per_launch"] if trg else None,
          "traffic_source": (f"{traffic_src}: mean dram bytes over the {trg['launches']} gn_* launches of one eager step")
          if trg else None}
    if t_gn > 0.02 * ms_full_step:
        ach = gb / (t_gn * 1e-3) / 1e9
        gn.update(achieved=ach, frac=ach / peaks["hbm"], ms_in_kernel_per_step=t_gn)
    else:
        gn.update(achieved=None, frac=None, ms_in_kernel_per_step=None, note="ablation not resolvable on this run")
    return conv, gn


def decode_roofline(ds, query_volume, dev, traffic, traffic_src):
    """The 256^3 dense decode, timed alone (CUDA events, 5 launches).  Compute-bound (69 888 FLOP per point against
    4 bytes written): reported against the tensor peak, with the HBM fraction beside it as SURVEY.md §8d asks."""
    import torch
    peaks = _peaks()
    res = DECODE_RES
    vol = torch.empty((res, res, res), device=dev)
    query_volume(ds.decoder, 0, res, out=vol)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        query_volume(ds.decoder, 0, res, out=vol)
    e1.record()
    torch.cuda.synchronize()
    t_ms = e0.elapsed_time(e1) / 5
    flops = DECODE_FLOP_PER_POINT * res ** 3
    byts = 4.0 * res ** 3 + 3 * 32 * 128 * 128 * 4
    ach = flops / (t_ms * 1e-3) / 1e12
    tr = (traffic or {}).get("decode")
    return {"bound": "tensor", "kernel": "triplane decode (bilinear plane samples + Fourier features + 3-layer MLP per point)",
            "achieved": ach, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": ach / peaks["tf_burst"],
            "peak_source": f"{peaks['src']} bf16_tflops (burst: kernel timed alone)", "ms_per_launch": t_ms,
            "points": res ** 3, "flop_per_point": DECODE_FLOP_PER_POINT,
            "hbm": {"algorithmic_bytes": byts, "achieved_gbs": byts / (t_ms * 1e-3) / 1e9,
                    "frac": byts / (t_ms * 1e-3) / 1e9 / peaks["hbm"]},
            "traffic": tr["dram_bytes_per_launch"] if tr else None,
            "traffic_source": traffic_src if tr else None}


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
    ap.add_argument("--no-eager", action="store_true", help="skip the stock-PyTorch-eager-on-B200 comparators")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
