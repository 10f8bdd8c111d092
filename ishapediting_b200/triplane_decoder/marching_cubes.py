"""Device-side meshing with the reference's call shapes.

`marching_cubes(volume, isovalue)` stands in for `mcubes.marching_cubes` (triplane_decoder/visualize.py:100) and
`TriangleMesh.filter_smooth_simple` for Open3D's (drag_utils.py:300), both on the GPU through the C ABI
(isb_mc_count / isb_mc_emit / isb_mesh_smooth_simple).  PyMCubes and Open3D are third-party CPU packages that are not
installed offline: parity with those binaries is unpinned; the checker is oracle/mcubes_oracle.py.
"""
from __future__ import annotations

import torch


def _ops_for(t):
    from ..ops import CudaOps
    if not t.is_cuda:
        raise RuntimeError("marching cubes runs on the B200 path only (no CPU fallback): pass a CUDA tensor")
    return CudaOps(t.device, "fp32")


def marching_cubes(volume, isovalue=0.0, ops=None):
    """volume (res,res,res) CUDA fp32 -> (vertices (V,3) fp32 in INDEX coordinates, triangles (T,3) int32), like
    `mcubes.marching_cubes(volume, isovalue)`."""
    vol = volume.detach().to(torch.float32).contiguous()
    ops = ops or _ops_for(vol)
    return ops.marching_cubes(vol, isovalue, 0.0)


class TriangleMesh:
    """The slice of open3d.geometry.TriangleMesh the editor uses: `.vertices`, `.triangles` (device tensors here)
    and `filter_smooth_simple(number_of_iterations)` returning a new mesh (drag_utils.py:300)."""

    def __init__(self, vertices, triangles, ops=None):
        self.vertices, self.triangles, self._ops = vertices, triangles, ops

    def filter_smooth_simple(self, number_of_iterations=1):
        ops = self._ops or _ops_for(self.vertices)
        v = ops.smooth_simple(self.vertices.clone(), self.triangles, number_of_iterations)
        return TriangleMesh(v, self.triangles, ops)

    def clone(self):
        return TriangleMesh(self.vertices.clone(), self.triangles.clone(), self._ops)

    def to_open3d(self):
        """An open3d.geometry.TriangleMesh (needs open3d; for the GUI)."""
        import open3d as o3d
        m = o3d.geometry.TriangleMesh()
        m.vertices = o3d.utility.Vector3dVector(self.vertices.double().cpu().numpy())
        m.triangles = o3d.utility.Vector3iVector(self.triangles.cpu().numpy())
        return m


def mesh_from_volume(volume, res=None, isovalue=0.0, ops=None):
    """The reference's meshing tail (visualize.py:100-105): marching cubes at `isovalue`, vertices / res * 2 - 1."""
    vol = volume.detach().to(torch.float32).contiguous()
    ops = ops or _ops_for(vol)
    res = vol.shape[0] if res is None else res
    v, t = ops.marching_cubes(vol, isovalue, float(res))
    return TriangleMesh(v, t, ops)
