"""Import the UNMODIFIED reference modules from /root/reference as the pinning oracle.

TEST INFRASTRUCTURE ONLY (see oracle/nfd_oracle.py header).  /root/reference exists only in the
build container; elsewhere (the GPU box) the git-ignored snapshot oracle/_ref/ made by
oracle/build_ref.py is used, and without either `available()` is False and callers fall back to the
committed fixtures under tests/golden/.  Third-party modules the reference imports at module scope
but that are not installable offline are stubbed: mpi4py, blobfile, open3d, mcubes,
matplotlib.  The reference modules are executed in place, unmodified.
"""
from __future__ import annotations

import os
import sys
import types

def _find_root():
    env = os.environ.get("ISB_REFERENCE_ROOT")
    cands = [env] if env else ["/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if os.path.isdir(os.path.join(c, "neural_field_diffusion", "guided_diffusion")):
            return c
    return cands[0]


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "neural_field_diffusion", "guided_diffusion"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    class _Comm:
        rank, size = 0, 1

        def Get_rank(self):
            return 0

        def Get_size(self):
            return 1

        def bcast(self, x, root=0):
            return x

    _stub("mpi4py", MPI=types.SimpleNamespace(COMM_WORLD=_Comm()))
    _stub("blobfile", BlobFile=open)
    _stub("mcubes")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    o3d = _stub("open3d")
    o3d.geometry = types.SimpleNamespace()
    o3d.utility = types.SimpleNamespace()
    o3d.io = types.SimpleNamespace()
    o3d.t = types.SimpleNamespace()
    _stub("skimage")


def import_reference():
    """Returns a namespace with the reference's unet / script_util / gaussian_diffusion / respace /
    axisnetworks modules (and drag_utils when importable)."""
    if not available():
        raise ImportError(f"reference tree not found under {REF_ROOT}")
    install_stubs()
    for p in (REF_ROOT, os.path.join(REF_ROOT, "neural_field_diffusion")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib

    ns = types.SimpleNamespace()
    ns.unet = importlib.import_module("neural_field_diffusion.guided_diffusion.unet")
    ns.script_util = importlib.import_module("neural_field_diffusion.guided_diffusion.script_util")
    ns.gaussian_diffusion = importlib.import_module("neural_field_diffusion.guided_diffusion.gaussian_diffusion")
    ns.respace = importlib.import_module("neural_field_diffusion.guided_diffusion.respace")
    ns.axisnetworks = importlib.import_module("triplane_decoder.axisnetworks")
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        ns.drag_utils = importlib.import_module("drag_utils")
    except Exception as e:  # noqa: BLE001 - optional: needs more of open3d than the stub offers
        ns.drag_utils = None
        ns.drag_utils_error = repr(e)
    finally:
        sys.argv = argv
    return ns


def reference_model_and_diffusion(cfg):
    """create_model_and_diffusion with the NFD arguments of drag_utils.py:44-57 (fp32)."""
    ns = import_reference()
    su = ns.script_util
    kw = su.model_and_diffusion_defaults()
    kw.update(image_size=cfg["image_size"], num_channels=cfg["num_channels"], num_res_blocks=cfg["num_res_blocks"],
              num_heads=4, num_heads_upsample=-1, num_head_channels=cfg["num_head_channels"],
              attention_resolutions=cfg["attention_resolutions"], channel_mult="", dropout=0.1, class_cond=False,
              use_checkpoint=False, use_scale_shift_norm=True, resblock_updown=True, use_fp16=False,
              use_new_attention_order=False, in_out_channels=cfg["in_out_channels"], learn_sigma=cfg["learn_sigma"],
              diffusion_steps=cfg["diffusion_steps"], noise_schedule="linear",
              timestep_respacing=cfg["timestep_respacing"], use_kl=False, predict_xstart=False,
              rescale_timesteps=False, rescale_learned_sigmas=False)
    if cfg["image_size"] != 128:
        kw["channel_mult"] = ",".join(str(int(m)) for m in cfg["channel_mult"])
    model, diffusion = su.create_model_and_diffusion(**kw)
    model.eval()
    return ns, model, diffusion
