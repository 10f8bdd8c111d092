"""Recipe for oracle/_ref: a byte-for-byte snapshot of the reference's Python modules for the hot path.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (jinli99/iShapEditing) is pure Python: there is nothing to
compile, and /root/reference does not exist on the GPU box.  `python oracle/build_ref.py` (run by
__graft_entry__.build() in the build container) copies the modules listed below, unmodified, from /root/reference
into oracle/_ref/ — a directory that is git-ignored (never enters history) but travels to the GPU box with the
snapshot, exactly like a compiled oracle/_ref/*.so would.  There `oracle/ref_import.py` imports them in place, so
that
  * `bench.py --impl reference` and the `cpu_baseline` leg time the reference's OWN DragStuff.training loop body
    (drag_utils.py:336-398: autograd with weight gradients and all) on the host cores, kind = "reference";
  * the live-reference checks of tests/test_oracle.py also run on the GPU box.
Nothing under ishapediting_b200/ ever imports from here.  SOURCE.txt records the sha256 of every file copied.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ISB_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "drag_utils.py",
    "meshProcess.py",
    "neural_field_diffusion/guided_diffusion/dist_util.py",
    "neural_field_diffusion/guided_diffusion/fp16_util.py",
    "neural_field_diffusion/guided_diffusion/gaussian_diffusion.py",
    "neural_field_diffusion/guided_diffusion/logger.py",
    "neural_field_diffusion/guided_diffusion/losses.py",
    "neural_field_diffusion/guided_diffusion/nn.py",
    "neural_field_diffusion/guided_diffusion/respace.py",
    "neural_field_diffusion/guided_diffusion/script_util.py",
    "neural_field_diffusion/guided_diffusion/unet.py",
    "triplane_decoder/axisnetworks.py",
    "triplane_decoder/dataset_3d.py",
    "triplane_decoder/visualize.py",
]


def build(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "neural_field_diffusion", "guided_diffusion")):
        if verbose:
            print(f"oracle/_ref: {SRC} not present (GPU box?) - keeping whatever snapshot travelled here")
        return os.path.isdir(os.path.join(DST, "neural_field_diffusion", "guided_diffusion"))
    lines = [f"snapshot of {SRC} made by oracle/build_ref.py (unmodified copies; test/baseline infrastructure only)"]
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(src, "rb") as f:
            lines.append(f"{hashlib.sha256(f.read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "SOURCE.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference modules snapshotted")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
