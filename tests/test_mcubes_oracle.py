"""CPU checks of the meshing oracle (oracle/mcubes_oracle.py) and of the derived marching-cubes case table
(ishapediting_b200/triplane_decoder/mc_table.py).  PyMCubes / Open3D are not installed offline and the reference
ships no mesh fixture, so parity with those binaries is unpinned; what is pinned here are the algorithm's invariants."""
import os

import numpy as np

from oracle import mcubes_oracle as M


def _grid(res):
    g = np.linspace(-1, 1, res, dtype=np.float32)
    return np.meshgrid(g, g, g, indexing="ij")


def test_case_table_invariants():
    from ishapediting_b200.triplane_decoder.mc_table import TRI_COUNT, TRI_TABLE

    assert TRI_TABLE.shape == (256, 15) and int(TRI_COUNT.max()) == 5
    assert int(TRI_COUNT[0]) == 0 and int(TRI_COUNT[255]) == 0
    assert M.check_table()
    # complementary cases cut the same edges
    for c in range(256):
        assert set(TRI_TABLE[c][TRI_TABLE[c] >= 0]) == set(TRI_TABLE[255 - c][TRI_TABLE[255 - c] >= 0])


def test_generated_include_is_current():
    import tools.gen_mc_table as G

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ishapediting_b200", "csrc", "mc_table.inc")
    assert open(path).read() == G.render(), "run python tools/gen_mc_table.py"


def test_sphere_and_torus_are_closed_oriented_manifolds():
    x, y, z = _grid(40)
    sphere = (0.6 - np.sqrt(x * x + y * y + z * z)).astype(np.float32)          # positive inside, like occupancy logits
    v, t = M.marching_cubes(sphere, 0.0)
    rep = M.mesh_report(v, t)
    assert rep["closed"] and rep["oriented"] and rep["euler"] == 2
    assert len(v) == M.crossed_edge_count(sphere)
    # outward normals (towards smaller values): positive signed volume, close to the ball's (index units)
    scale = (40 - 1) / 2.0
    assert abs(rep["volume"] / (4 / 3 * np.pi * (0.6 * scale) ** 3) - 1) < 0.02
    # vertices lie on the surface: |r - 0.6| small after mapping back to [-1,1]
    r = np.linalg.norm(v / scale - 1.0, axis=1)
    assert np.abs(r - 0.6).max() < 2e-3
    torus = (0.22 - np.sqrt((np.sqrt(x * x + y * y) - 0.55) ** 2 + z * z)).astype(np.float32)
    v, t = M.marching_cubes(torus, 0.0)
    rep = M.mesh_report(v, t)
    assert rep["closed"] and rep["oriented"] and rep["euler"] == 0


def test_noise_volumes_are_watertight():
    """Ambiguous faces everywhere: the face rule must keep every surface closed and consistently oriented."""
    for seed in range(4):
        rng = np.random.default_rng(seed)
        vol = rng.standard_normal((20, 20, 20)).astype(np.float32)
        vol[0] = vol[-1] = -1
        vol[:, 0] = vol[:, -1] = -1
        vol[:, :, 0] = vol[:, :, -1] = -1
        for sign in (1.0, -1.0):
            v, t = M.marching_cubes(sign * vol - (0 if sign > 0 else 2.5), 0.0)
            if len(t) == 0:
                continue
            rep = M.mesh_report(v, t)
            assert rep["closed"] and rep["oriented"], (seed, sign)
            assert len(v) == M.crossed_edge_count(sign * vol - (0 if sign > 0 else 2.5))


def test_smoothing_matches_its_definition():
    """filter_smooth_simple against a literal per-vertex loop over neighbour sets (Open3D's FilterSmoothSimple)."""
    x, y, z = _grid(16)
    vol = (0.5 - np.sqrt(x * x + 1.3 * y * y + z * z)).astype(np.float32)
    v, t = M.marching_cubes(vol, 0.0)
    nb = [set() for _ in range(len(v))]
    for a, b, c in t:
        nb[a] |= {b, c}
        nb[b] |= {a, c}
        nb[c] |= {a, b}
    cur = v.astype(np.float64)
    for _ in range(3):
        nxt = np.empty_like(cur)
        for i in range(len(cur)):
            s = cur[i].copy()
            for j in sorted(nb[i]):
                s += cur[j]
            nxt[i] = s / (1 + len(nb[i]))
        cur = nxt
    got = M.filter_smooth_simple(v, t, 3)
    assert np.abs(got - cur).max() < 1e-12
    # smoothing shrinks a convex blob
    assert M.mesh_report(got, t)["volume"] < M.mesh_report(v, t)["volume"]
