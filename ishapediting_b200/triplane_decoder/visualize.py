"""Dense-grid occupancy query with the reference's surface (triplane_decoder/visualize.py:76-105).

The reference builds the (res^3, 3) coordinate tensor on the CPU, loops over 50 000-point chunks with
a host<->device round trip per chunk, and meshes with PyMCubes.  Here the grid is generated inside
the decode kernel, a contiguous x-slab [x_begin, x_end) can be requested (the multi-GPU shard, index =
x*res^2 + y*res + z as in visualize.py:83-86) and the logit volume stays on the device, where
marching cubes and the Laplacian smoothing run too (marching_cubes.py, csrc/mc.cu).
"""
import torch


def query_volume(model, obj_idx, res=128, x_begin=0, x_end=None, out=None):
    """Logit volume (x_end-x_begin, res, res) fp32 on the device."""
    model.eval()
    x_end = res if x_end is None else x_end
    ops = model._get_ops()
    lin = torch.linspace(-1, 1, res).to(ops.device)        # same table as visualize.py:79-81
    n = (x_end - x_begin) * res * res
    if out is None:
        out = ops.empty((n,))
    with torch.no_grad():
        ops.decode_grid(model.planes_hwc(obj_idx), model.mlp_weights(), lin, x_begin, x_end, out)
    return out.view(x_end - x_begin, res, res)


def create_obj_o3d(model, obj_idx, res=128, max_batch_size=50000, volume=None):
    """Reference signature (visualize.py:76-105); `max_batch_size` is accepted and ignored (no chunking is needed).
    Dense query -> marching cubes at 0 -> vertices / res * 2 - 1, all on the device; returns a
    marching_cubes.TriangleMesh (`.vertices`, `.triangles`, `.filter_smooth_simple`, `.to_open3d()`).
    `volume`: an already decoded (res,res,res) logit volume to mesh (get_mesh passes the one it just computed
    instead of decoding twice)."""
    from .marching_cubes import mesh_from_volume

    vol = query_volume(model, obj_idx, res) if volume is None else volume
    return mesh_from_volume(vol.reshape(res, res, res), res, 0.0, model._get_ops())
