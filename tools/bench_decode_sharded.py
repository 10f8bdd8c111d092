"""BASELINE config 4: dense triplane-decoder sweep (128^3 / 256^3 / 512^3), x-slab sharded over the ranks of
one node and all-gathered over NCCL (parallel.slab_range / gather_volume).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_decode_sharded.py
(or plain `python tools/bench_decode_sharded.py` for one GPU).  Rank 0 prints one JSON line per resolution."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from oracle import nfd_oracle as O          # synthetic decoder weights/planes + the checker on a small grid
from ishapediting_b200.parallel import gather_volume, slab_range
from ishapediting_b200.triplane_decoder.axisnetworks import MultiTriplane
from ishapediting_b200.triplane_decoder.visualize import query_volume


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, planes = O.synth_decoder()
    dec = MultiTriplane(1).to(dev)
    dec.net[0]._B.data.copy_(w["B"])
    for idx, k in ((1, "1"), (3, "2"), (5, "3")):
        dec.net[idx].weight.data.copy_(w["w" + k])
        dec.net[idx].bias.data.copy_(w["b" + k])
    for p in range(3):
        dec.embeddings[p] = planes[[p]].to(dev)
    # correctness of the sharded + gathered volume on a grid the oracle can afford
    b, e = slab_range(48, rank, world)
    vol = gather_volume(query_volume(dec, 0, res=48, x_begin=b, x_end=e), 48)
    if rank == 0:
        ref = O.decode_grid(w, planes, 48).view(48, 48, 48)
        err = float((vol.cpu() - ref).abs().max())
        assert err < 1e-4, err
    for res in (128, 256, 512):
        b, e = slab_range(res, rank, world)
        for _ in range(2):
            full = gather_volume(query_volume(dec, 0, res=res, x_begin=b, x_end=e), res)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 3
        t_dec = t_all = 0.0
        for _ in range(reps):
            e0.record()
            slab = query_volume(dec, 0, res=res, x_begin=b, x_end=e)
            e1.record()
            full = gather_volume(slab, res)
            e2.record()
            torch.cuda.synchronize()
            t_dec += e0.elapsed_time(e1)
            t_all += e0.elapsed_time(e2)
        t = torch.tensor([t_dec / reps, t_all / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            pts = res ** 3
            print(json.dumps({"metric": "triplane decode points/s", "res": res, "n_gpus": world,
                              "ms_decode_max_rank": float(t[0]), "ms_decode_plus_gather": float(t[1]),
                              "gpts_per_s": pts / float(t[1]) / 1e6, "tflops_algorithmic": pts * 69888 / float(t[1]) / 1e9,
                              "gather_bytes": pts * 4, "parity_max_abs_48": err}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
