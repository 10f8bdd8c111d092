"""Summarise an ncu CSV (long format: `ncu --csv --log-file X --print-units base --metrics ...`) of one eager guided
step into per-kernel and per-family tables:

    python tools/ncu_summarize.py gpurun_out/r02_ncu_step_metrics.csv profiles/r02_ncu_step_traffic.json \
        [gpurun_out/r02_step_algbytes.json]

The JSON (read by bench.py: `roofline.traffic`) holds, per kernel family, the number of launches, the mean
dram__bytes_read.sum + dram__bytes_write.sum per launch, the summed device time, the mean L2 hit rate and (conv) the
time-weighted tensor-pipe utilisation; a markdown table with the same numbers per kernel NAME is written next to it."""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _family_of  # noqa: E402

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
         "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}


def short(name):
    n = name.split("(")[0]
    return n.replace("isb::", "").replace("void ", "").strip()


def main():
    src, dst = sys.argv[1], sys.argv[2]
    alg = json.load(open(sys.argv[3])) if len(sys.argv) > 3 and os.path.exists(sys.argv[3]) else {}
    rows = {}
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= SCALE.get(r["Metric Unit"], 1.0)
        rows.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"]), "grid": r["Grid Size"]})[r["Metric Name"]] = v
    per_kernel, per_family = {}, {}
    for _, r in sorted(rows.items()):
        t = r.get("gpu__time_duration.sum", 0.0)
        d = r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
        for key, table in ((r["name"], per_kernel), (_family_of(r["name"]), per_family)):
            e = table.setdefault(key, {"launches": 0, "time_ns": 0.0, "dram_bytes": 0.0, "l2_hit_w": 0.0, "tensor_w": 0.0,
                                       "xbar_bytes": 0.0})
            e["launches"] += 1
            e["time_ns"] += t
            e["dram_bytes"] += d
            e["l2_hit_w"] += r.get("lts__t_sector_hit_rate.pct", 0.0) * t
            e["tensor_w"] += r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
            e["xbar_bytes"] += r.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0.0)
    total_t = sum(e["time_ns"] for e in per_family.values()) or 1.0

    def finish(e):
        t = e["time_ns"] or 1.0
        return {"launches": e["launches"], "time_us": e["time_ns"] / 1e3, "share_of_step": e["time_ns"] / total_t,
                "dram_bytes_per_launch": e["dram_bytes"] / e["launches"], "dram_bytes": e["dram_bytes"],
                "dram_gbs_cold": e["dram_bytes"] / t, "l2_hit_pct": e["l2_hit_w"] / t,
                "tensor_pipe_pct": e["tensor_w"] / t, "l2_to_sm_bytes": e["xbar_bytes"]}

    out = {k: finish(e) for k, e in per_family.items()}
    for fam, a in alg.items():
        if fam in out and out[fam]["launches"]:
            out[fam]["algorithmic_bytes_per_launch"] = a["bytes"] / max(a["launches"], 1)
            out[fam]["algorithmic_launches"] = a["launches"]
    out["_source"] = os.path.basename(src)
    out["_note"] = ("one eager guided step (tools/profile_step.py) under ncu: cold-cache, serialised launches — "
                    "shares and bytes are meaningful, absolute times are not")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    md = os.path.splitext(dst)[0] + ".md"
    with open(md, "w") as f:
        f.write(f"# ncu per-kernel summary of one eager guided step ({os.path.basename(src)})\n\n")
        f.write("Cold-cache, serialised launches: compare shares and bytes, not absolute times.\n\n")
        f.write("| kernel | launches | time us | share | DRAM MB/launch | DRAM GB/s (cold) | L2 hit % | tensor pipe % | L2->SM MB |\n")
        f.write("|---|---|---|---|---|---|---|---|---|\n")
        for k, e in sorted(per_kernel.items(), key=lambda kv: -kv[1]["time_ns"]):
            r = finish(e)
            f.write(f"| {k} | {r['launches']} | {r['time_us']:.1f} | {100 * r['share_of_step']:.1f}% | "
                    f"{r['dram_bytes_per_launch'] / 1e6:.2f} | {r['dram_gbs_cold']:.0f} | {r['l2_hit_pct']:.1f} | "
                    f"{r['tensor_pipe_pct']:.1f} | {r['l2_to_sm_bytes'] / 1e6:.1f} |\n")
        f.write("\n| family | launches | time us | share | DRAM MB/launch | algorithmic MB/launch |\n|---|---|---|---|---|---|\n")
        for k, r in sorted(((k, v) for k, v in out.items() if not k.startswith("_")), key=lambda kv: -kv[1]["time_us"]):
            ab = r.get("algorithmic_bytes_per_launch")
            f.write(f"| {k} | {r['launches']} | {r['time_us']:.1f} | {100 * r['share_of_step']:.1f}% | "
                    f"{r['dram_bytes_per_launch'] / 1e6:.2f} | {'' if ab is None else f'{ab / 1e6:.2f}'} |\n")
    print("wrote", dst, md)


if __name__ == "__main__":
    main()
