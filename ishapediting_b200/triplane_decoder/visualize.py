"""Dense-grid occupancy query with the reference's surface (triplane_decoder/visualize.py:76-105).

The reference builds the (res^3, 3) coordinate tensor on the CPU, loops over 50 000-point chunks with
a host<->device round trip per chunk, and meshes with PyMCubes.  Here the grid is generated inside
the decode kernel, a contiguous x-slab [x_begin, x_end) can be requested (the multi-GPU shard, index =
x*res^2 + y*res + z as in visualize.py:83-86) and the logit volume stays on the device.  Marching
cubes / Open3D smoothing are CPU third-party code outside the hot path: if
`mcubes` and `open3d` are installed the reference's meshing tail is reproduced, otherwise the volume
is returned.
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


def create_obj_o3d(model, obj_idx, res=128, max_batch_size=50000):
    """Reference signature; `max_batch_size` is accepted and ignored (no chunking is needed).
    Returns an open3d TriangleMesh when mcubes+open3d are importable, else the logit volume."""
    vol = query_volume(model, obj_idx, res)
    try:
        import mcubes
        import open3d as o3d
    except ImportError:
        return vol
    vertices, triangles = mcubes.marching_cubes(vol.cpu().numpy(), 0)
    vertices = vertices / res * 2 - 1
    mesh = o3d.geometry.TriangleMesh()
    mesh.vertices = o3d.utility.Vector3dVector(vertices)
    mesh.triangles = o3d.utility.Vector3iVector(triangles)
    return mesh
