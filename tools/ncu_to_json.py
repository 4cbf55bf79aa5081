"""Per-launch ncu metric CSVs -> profiles/ncu_metrics.json entries (read by bench.py for roofline.traffic and the submetrics).
  python tools/ncu_to_json.py <round> <full_gemm.csv> <backbone_gemm.csv> <roi.csv> [out.json]
Each GEMM csv: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:conv_gemm -s <one step> -c <launches of a step>."""
import csv
import json
import sys
from collections import defaultdict


def per_launch(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    by_id = defaultdict(dict)
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if unit in ("Mbyte",):
            v *= 1e6
        elif unit in ("Kbyte",):
            v *= 1e3
        elif unit in ("Gbyte",):
            v *= 1e9
        elif unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        by_id[r["ID"]][r["Metric Name"]] = v
        by_id[r["ID"]]["kernel"] = r["Kernel Name"]
    return list(by_id.values())


def gemm_summary(path, source):
    rows = per_launch(path)
    t = sum(r["gpu__time_duration.sum"] for r in rows)
    dram = sum(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"] for r in rows)
    key = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tw = sum(r[key] * r["gpu__time_duration.sum"] for r in rows) / t
    busy = [r for r in rows if r[key] >= 60.0]
    tb = sum(r["gpu__time_duration.sum"] for r in busy)
    return {"kernel": f"conv_gemm_kernel, all {len(rows)} launches of one step", "launches": len(rows),
            "dram_bytes_per_step": int(dram), "gemm_ms_under_ncu": t / 1e6,
            "time_weighted_tensor_pipe_active_pct": round(tw, 1),
            "tensor_pipe_active_pct_mma_bound_layers": round(sum(r[key] * r["gpu__time_duration.sum"] for r in busy) / tb, 1) if busy else None,
            "mma_bound_layers_share_of_gemm_time": round(tb / t, 3), "source": source}


def main():
    rnd, full, bb, roi = sys.argv[1:5]
    out = sys.argv[5] if len(sys.argv) > 5 else "profiles/ncu_metrics.json"
    res = {"_comment": "profiler-only numbers read by bench.py (roofline.traffic, submetrics.*.tensor_pipe_active_pct / dram "
                       "bytes); each entry names the ncu pass it came from (cold-cache, serialised launches)"}
    res["full_bs4"] = dict(gemm_summary(full, f"profiles/{rnd}_full_bs4_gemm_metrics.csv"), round=rnd)
    res["backbone_bs8"] = dict(gemm_summary(bb, f"profiles/{rnd}_backbone_bs8_gemm_metrics.csv"), round=rnd)
    r = per_launch(roi)
    r = [x for x in r if "roi_align" in x["kernel"]]
    res["roialign_512"] = {"kernel": "roi_align_rotated_split8_kernel<2>", "round": rnd,
                           "dram_bytes_per_launch": int(sum(x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"] for x in r) / len(r)),
                           "launches_averaged": len(r), "source": f"profiles/{rnd}_roialign_metrics.csv"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
