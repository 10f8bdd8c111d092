"""GPU marching cubes + Laplacian smoothing (csrc/mc.cu) against the CPU restatement oracle/mcubes_oracle.py
(reference visualize.py:100-105, drag_utils.py:300).  Index / integer work is compared
bit-exactly: same vertex order, same triangles, vertex coordinates equal to the last bit (IEEE fp32 ops on both sides)."""
import numpy as np
import pytest
import torch

from oracle import mcubes_oracle as M
from oracle import nfd_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _grid(res):
    g = np.linspace(-1, 1, res, dtype=np.float32)
    return np.meshgrid(g, g, g, indexing="ij")


def _volumes():
    x, y, z = _grid(33)
    yield "sphere", (0.6 - np.sqrt(x * x + y * y + z * z)).astype(np.float32)
    yield "torus", (0.22 - np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z)).astype(np.float32)
    rng = np.random.default_rng(3)
    yield "noise", rng.standard_normal((21, 21, 21)).astype(np.float32)            # surface runs into the border
    yield "empty", np.full((8, 8, 8), -1.0, dtype=np.float32)
    yield "full", np.full((8, 8, 8), 1.0, dtype=np.float32)
    w, planes = O.synth_decoder(R=128)
    yield "decoder48", O.decode_grid(w, planes, 48).reshape(48, 48, 48).numpy()


@pytest.mark.parametrize("name,vol", list(_volumes()), ids=[n for n, _ in _volumes()])
def test_marching_cubes_equals_oracle(name, vol):
    from ishapediting_b200.triplane_decoder.marching_cubes import marching_cubes, mesh_from_volume

    v_ref, t_ref = M.marching_cubes(vol, 0.0)
    v, t = marching_cubes(torch.from_numpy(vol).to(DEV), 0.0)
    assert v.shape == (len(v_ref), 3) and t.shape == (len(t_ref), 3)
    assert len(v_ref) == M.crossed_edge_count(vol)
    assert torch.equal(t.cpu(), torch.from_numpy(t_ref))                  # indices: bit-exact
    assert torch.equal(v.cpu(), torch.from_numpy(v_ref))                  # IEEE fp32 interpolation: bit-exact
    if len(t_ref):
        res = vol.shape[0]
        m = mesh_from_volume(torch.from_numpy(vol).to(DEV), res)          # the reference's scaling v / res * 2 - 1
        want = (v_ref / np.float32(res) * np.float32(2) - np.float32(1)).astype(np.float32)
        assert torch.equal(m.vertices.cpu(), torch.from_numpy(want))
        assert torch.equal(m.triangles.cpu(), torch.from_numpy(t_ref))


def test_smoothing_equals_oracle():
    from ishapediting_b200.triplane_decoder.marching_cubes import mesh_from_volume

    w, planes = O.synth_decoder(R=128)
    vol = O.decode_grid(w, planes, 64).reshape(64, 64, 64).numpy()
    m = mesh_from_volume(torch.from_numpy(vol).to(DEV), 64)
    v0, t0 = m.vertices.cpu().numpy(), m.triangles.cpu().numpy()
    for it in (1, 10):
        want = M.filter_smooth_simple(v0, t0, it)
        got = m.filter_smooth_simple(number_of_iterations=it).vertices.cpu().numpy()
        assert np.abs(got.astype(np.float64) - want).max() <= 1e-7 * max(1.0, np.abs(want).max())     # fp32 store only
    assert torch.equal(m.vertices.cpu(), torch.from_numpy(v0))            # filter returns a NEW mesh (Open3D semantics)


def test_get_mesh_decodes_once_and_returns_a_smoothed_mesh():
    """DragStuff.get_mesh (drag_utils.py:282-300): decode -> marching cubes -> 10 smoothing iterations, one decode."""
    from ishapediting_b200 import _lib
    from ishapediting_b200.drag_utils import DragStuff, get_args
    from ishapediting_b200.triplane_decoder import visualize

    a = get_args(["--num_steps", "20", "--w_time", "4", "--shape_resolution", "48", "--feat_layer", "5", "--resolution", "32"])
    a.channel_mult, a.attention_resolutions, a.use_fp16 = "1,2,4", "16,8", True
    ds = DragStuff(args=a, device=DEV, use_graph=False)
    w, planes = O.synth_decoder(R=32)
    ds.decoder.net[0]._B.data.copy_(w["B"])
    for idx, k in ((1, "1"), (3, "2"), (5, "3")):
        ds.decoder.net[idx].weight.data.copy_(w["w" + k])
        ds.decoder.net[idx].bias.data.copy_(w["b" + k])
    calls = []
    orig = visualize.query_volume

    def counting(*args, **kw):
        calls.append(1)
        return orig(*args, **kw)

    import ishapediting_b200.drag_utils as du
    du.query_volume, visualize.query_volume = counting, counting
    try:
        mesh = ds.get_mesh(tri_feat=planes.reshape(1, 96, 32, 32).to(DEV))
    finally:
        du.query_volume, visualize.query_volume = orig, orig
    assert len(calls) == 1, "get_mesh decoded the volume more than once"
    vol = ds.last_volume.cpu().numpy()
    v_ref, t_ref = M.marching_cubes(vol, 0.0)
    want = M.filter_smooth_simple(v_ref / np.float32(48) * np.float32(2) - np.float32(1), t_ref, 10)
    assert torch.equal(mesh.triangles.cpu(), torch.from_numpy(t_ref))
    assert np.abs(mesh.vertices.cpu().numpy() - want).max() < 1e-6
    assert _lib.launch_count() > 0
