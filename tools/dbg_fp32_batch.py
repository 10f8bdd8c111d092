"""Debug: fp32-mode batched forward at the NFD size — history dependence of a reused batch-8 plan."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nfd_oracle as O
from tests.helpers import build_model
from tests.conftest import rel_l2

DEV = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
cfg = O.NFD_CFG
sd = O.synth_state_dict(cfg)
A, _ = build_model(cfg, sd, mode, DEV)
B, _ = build_model(cfg, sd, mode, DEV)
g = torch.Generator().manual_seed(3)
xs = torch.randn(8, 96, 128, 128, generator=g).to(DEV)
ts = torch.tensor([246, 241, 236, 231, 226, 221, 216, 211], device=DEV)
ts2 = ts - 120
with torch.no_grad():
    pa = A.plan(8, 128, 128, True)
    pb = B.plan(8, 128, 128, True)
    A(xs, ts, feat_layer=8)
    oa = A(xs, ts2, feat_layer=8)[0]
    ob = B(xs, ts2, feat_layer=8)[0]
    print("film", [f"{rel_l2(pa.film_all[k], pb.film_all[k]):.1e}" for k in range(8)])
    print("out", [f"{rel_l2(oa[k], ob[k]):.1e}" for k in range(8)])
    va = [pa.h0.val] + [l.out.val for l in pa.layers]
    vb = [pb.h0.val] + [l.out.val for l in pb.layers]
    names = ["h0"] + [l.name for l in pb.layers]
    shown = 0
    for n, a, b in zip(names, va, vb):
        e = [rel_l2(a[k], b[k]) for k in range(8)]
        if max(e) > 1e-5 and shown < 4:
            print("DIVERGE", n, tuple(a.shape), [f"{v:.1e}" for v in e])
            shown += 1
