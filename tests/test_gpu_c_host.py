"""A host WITHOUT Python or PyTorch (examples/host_guided_step.c, plain C99 + the CUDA runtime) drives the guided
denoise step through the C ABI alone — isb_unet_* handle, isb_drag_loss_grad, isb_ddpm_step — eagerly and from a CUDA
graph it captures itself, and must reproduce the Python host's GuidedStepper bit for bit (same kernels, same order)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest
import torch

from oracle import nfd_oracle as O
from tests.conftest import ROOT, rel_l2
from tests.helpers import build_model, drag_problem, seeded_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_DT = {np.dtype("float32"): 0, np.dtype("int32"): 1, np.dtype("uint8"): 2}


def write_blob(path, entries):
    with open(path, "wb") as f:
        f.write(b"ISB1" + struct.pack("<i", len(entries)))
        for name, arr in entries:
            arr = np.ascontiguousarray(arr)
            assert arr.ndim <= 4 and len(name) < 96, name
            shape = list(arr.shape) + [1] * (4 - arr.ndim)
            f.write(name.encode().ljust(96, b"\0"))
            f.write(struct.pack("<ii4qq", _DT[arr.dtype], arr.ndim, *shape, arr.nbytes))
            f.write(arr.tobytes())
            f.write(b"\0" * ((8 - arr.nbytes % 8) % 8))


def read_blob(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(4) == b"ISB1"
        (n,) = struct.unpack("<i", f.read(4))
        for _ in range(n):
            name = f.read(96).rstrip(b"\0").decode()
            dt, nd, s0, s1, s2, s3, nbytes = struct.unpack("<ii4qq", f.read(48))
            data = f.read((nbytes + 7) // 8 * 8)[:nbytes]
            out[name] = np.frombuffer(data, dtype={0: np.float32, 1: np.int32, 2: np.uint8}[dt]).copy()
    return out


def build_host(tmp):
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA runtime headers not available on this box")
    exe = os.path.join(tmp, "host_guided_step")
    libdir = os.path.join(ROOT, "ishapediting_b200")
    cmd = ["gcc", "-O2", "-std=c99", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "host_guided_step.c"), "-o", exe, "-L" + libdir, "-lishape_b200",
           "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + libdir,
           "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_c_host_reproduces_python_stepper(tmp_path, mode):
    from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper, align_maps

    exe = build_host(str(tmp_path))
    cfg = O.mid_cfg()
    sd = O.synth_state_dict(cfg)
    model, diff = build_model(cfg, sd, mode, DEV)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    g, x, x2, noise0 = seeded_inputs(cfg)
    fl = cfg["feat_layer"]
    steps_i = [49, 48, 47, 46]
    origin, src, tgt, r1, voxel, *_ = drag_problem(cfg, sd, sched, x2, noise0, steps_i[0], g, r1=4, voxel=2.0 / 64)
    S, Ca = origin.shape[-1], origin.shape[1]
    geo = DragGeometry(src, tgt, r1, voxel, S, Ca)
    origin_cl = origin.permute(0, 2, 3, 1).contiguous()                       # (3,S,S,Ca) channels-last
    origins = [origin_cl * (1.0 + 0.05 * k) for k in range(len(steps_i))]     # a different cached feature per step
    noises = [torch.randn(x.shape, generator=g) for _ in steps_i]

    # --- the Python host ---
    st = GuidedStepper(model, diff, geo, fl, 0.2, "l2", 600.0, use_graph=False)
    st.img.copy_(x.to(DEV))
    for k, i in enumerate(steps_i):
        st.step(i, origins[k].to(DEV), noises[k].to(DEV))
    torch.cuda.synchronize()
    if os.environ.get("ISB_HOST_DEBUG"):
        print("py debug: x", float(x.abs().double().sum()), "feat", float(st.plan.block_out[fl].val.abs().double().sum()),
              "origin", float(st.origin.abs().double().sum()), "patch", float(geo.patch_xy.abs().double().sum()),
              "loss", float(st.loss))
    ref = dict(img=st.img.cpu().numpy().ravel(), grad=st.grad.cpu().numpy().ravel(), loss=float(st.loss))

    # --- the C host: the same edit from a blob ---
    chan_map, inv_map, _ = align_maps(st.Cf)
    mult = [int(m) for m in cfg["channel_mult"]]
    ds = sorted(int(a) for a in model.attention_resolutions)
    R = cfg["image_size"]
    ci = ([model.in_channels, model.model_channels, model.out_channels, model.num_res_blocks, len(mult)]
          + mult + [0] * (8 - len(mult)) + [len(ds)] + ds + [0] * (8 - len(ds))
          + [model.num_heads, model.num_head_channels, model.num_heads_upsample, 1, R, R, 1 if mode == "bf16" else 0,
             fl, len(steps_i), geo.group_size, geo.mask_count, 0, 1])
    coef = st.coef_table.cpu().numpy()[steps_i].astype(np.float32)
    tvals = st.t_table.cpu().numpy()[steps_i].astype(np.float32)
    entries = [("cfg", np.array(ci, dtype=np.int32)),
               ("scalars", np.array([geo.inv_count, 0.2], dtype=np.float32)), ("dyn", st.dyn.cpu().numpy()),
               ("w:time_embed.freqs", st.plan.freqs.cpu().numpy()),
               ("x", x.numpy()), ("t", tvals), ("coef", coef),
               ("noise", torch.stack(noises).numpy().reshape(len(steps_i), -1)),
               ("origin", torch.stack(origins).numpy().reshape(len(steps_i), 3 * S * S, Ca)),
               ("chan_map", chan_map.numpy()), ("inv_map", inv_map.numpy()),
               ("patch_xy", geo.patch_xy.numpy()), ("shift_xy", geo.shift_xy.numpy()), ("weight", geo.weight.numpy()),
               ("bbox", geo.bbox.numpy().reshape(3, -1, 4)), ("mask", geo.mask.numpy())]
    for name, t in model.state_dict().items():
        entries.append(("w:" + name, t.detach().float().cpu().numpy()))
    blob = os.environ.get("ISB_KEEP_BLOB") or str(tmp_path / "in.blob")
    write_blob(blob, entries)
    for variant in ([], ["graph"]):
        out = str(tmp_path / "out.blob")
        r = subprocess.run([exe, blob, out] + variant, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        print(r.stdout.strip(), r.stderr.strip())
        got = read_blob(out)
        errs = dict(img=rel_l2(torch.from_numpy(got["img"]), torch.from_numpy(ref["img"])),
                    grad=rel_l2(torch.from_numpy(got["grad"]), torch.from_numpy(ref["grad"])),
                    loss=abs(float(got["loss"][0]) - ref["loss"]) / abs(ref["loss"]))
        print("c host", mode, variant or "eager", errs)
        assert np.array_equal(got["img"], ref["img"]) and np.array_equal(got["grad"], ref["grad"]), errs
        assert float(got["loss"][0]) == ref["loss"]
