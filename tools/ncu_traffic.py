"""profiles/conv_dram_traffic.json from an ncu launch list (csv with dram__bytes_read.sum / dram__bytes_write.sum /
gpu__time_duration.sum per launch, as `ncu --metrics ... --csv --log-file` writes it for `bench.py --no-graph
--ncu-range --steps 1`).  bench.py reads the file for `roofline.traffic`: the number is measured on the model and
build it is printed for, never typed in.

    python tools/ncu_traffic.py <launches.csv> <model> <batch> [source label]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, model, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
    label = sys.argv[4] if len(sys.argv) > 4 else os.path.relpath(path, ROOT)
    rows = {}
    for r in csv.DictReader(l for l in open(path) if l.startswith('"')):
        e = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        e[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    conv = [e for e in rows.values() if "conv_tap_gemm_kernel" in e["name"]]
    tot = sum(e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0) for e in conv)
    allb = sum(e.get("dram__bytes_read.sum", 0) + e.get("dram__bytes_write.sum", 0) for e in rows.values())
    t_conv = sum(e.get("gpu__time_duration.sum", 0) for e in conv)
    t_all = sum(e.get("gpu__time_duration.sum", 0) for e in rows.values())
    out_p = os.path.join(ROOT, "profiles", "conv_dram_traffic.json")
    d = json.load(open(out_p)) if os.path.isfile(out_p) else {}
    d["%s_bs%d" % (model, batch)] = {"bytes_per_launch": round(tot / max(len(conv), 1)), "bytes_per_step": round(tot),
                                     "launches": len(conv), "all_kernels_bytes_per_step": round(allb),
                                     "conv_share_of_step_time": round(t_conv / t_all, 4) if t_all else None,
                                     "source": label}
    json.dump(d, open(out_p, "w"), indent=1, sort_keys=True)
    print(json.dumps(d["%s_bs%d" % (model, batch)]))


if __name__ == "__main__":
    main()
