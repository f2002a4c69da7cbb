#!/usr/bin/env python
"""Turns the scratch output of `gpurun -- profiles/measure_r2.sh <tag>` (gpurun_out/<tag>_*) into the tracked round-2
files under profiles/: ncu summaries (ncu_summary.py), the condensed launch list, the FP64 instruction counts and
eval_traffic.json (what bench.py reports as roofline.traffic / roofline.fp64.kernel_*), the bench lines.

    python profiles/refresh.py <tag>
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "gpurun_out")

CAPTURES = {  # tracked name -> scratch name
    "r2_eval_talos": "eval_talos", "r2_eval_solo": "eval_solo", "r2_eval_panda": "eval_panda",
    "r2_aba2_talos": "aba2_talos", "r2_minv_talos": "minv_talos", "r2_rows_talos": "rows_talos",
    "r2_soa_to_aos": "soa_to_aos", "r2_drnea_talos": "drnea_talos", "r2_applyJT_panda": "applyJT_panda",
}
EVALS = {"talos_reduced:1048576:tile32": ("eval_talos", "talos", 1048576),
         "solo12:262144:tile32": ("eval_solo", "solo", 262144),
         "panda7:65536:tile32": ("eval_panda", "panda", 65536)}


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}


def to_bytes(u, v):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return float(v) * scale


def main():
    tag = sys.argv[1]
    for name, src in CAPTURES.items():
        raw = os.path.join(OUT, f"{tag}_{src}_raw.csv"); sass = os.path.join(OUT, f"{tag}_{src}_sass.csv")
        if not os.path.exists(raw) or os.path.getsize(raw) < 1000:
            print("missing", raw); continue
        r = subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), raw, sass], capture_output=True, text=True)
        kern = raw_metrics(raw).get("Kernel Name", ("", "?"))[1]
        with open(os.path.join(HERE, f"{name}_ncu_full.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, one launch of {kern} ({tag}); condensed by ncu_summary.py\n")
            f.write(r.stdout)
    # launch list
    p = os.path.join(OUT, f"{tag}_launches.csv")
    if os.path.exists(p):
        text = open(p).read()
        text = text[text.index('"ID"'):]
        agg = collections.OrderedDict()
        for row in csv.DictReader(io.StringIO(text)):
            if row["Metric Name"] != "gpu__time_duration.sum":
                continue
            k = row["Kernel Name"].split("(")[0]
            t = float(row["Metric Value"]) / (1e3 if row["Metric Unit"] in ("ns", "nsecond") else 1.0)
            c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += t
        with open(os.path.join(HERE, "r2_launches_condensed.txt"), "w") as f:
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 3 --warmup 3 --no-e2e "
                    "--no-cpu --secondary-steps 2` (first 400 launches; per-launch times under ncu are cold-cache and serialised)\n")
            f.write(f"{'kernel':72s} {'count':>6s} {'mean us':>12s} {'total us':>12s}\n")
            for k, (n, t) in agg.items():
                f.write(f"{k[:70]:72s} {n:6d} {t / n:12.1f} {t:12.1f}\n")
    # FP64 instruction counts + DRAM traffic of the three eval instantiations
    traffic = {}
    lines = []
    for key, (cap, w, n) in EVALS.items():
        raw = os.path.join(OUT, f"{tag}_{cap}_raw.csv"); fp = os.path.join(OUT, f"{tag}_fp64_{w}.csv")
        if not (os.path.exists(raw) and os.path.exists(fp)):
            continue
        m = raw_metrics(raw)
        rd, wr = to_bytes(*m["dram__bytes_read.sum"]), to_bytes(*m["dram__bytes_write.sum"])
        text = open(fp).read(); text = text[text.index('"ID"'):]
        cnt = {}
        for row in csv.DictReader(io.StringIO(text)):
            lines.append(row)
            for op in ("dfma", "dadd", "dmul"):
                if f"op_{op}_pred_on" in row["Metric Name"]:
                    cnt[op] = float(row["Metric Value"]) / n
        kern = m.get("Kernel Name", ("", "?"))[1].replace("void ", "").split("(")[0]
        traffic[key] = {
            "dram_bytes_per_launch": rd + wr, "read": rd, "write": wr,
            "source": f"profiles/r2_{cap}_ncu_full.txt (ncu --set full --clock-control none, one launch of rbd::{kern})",
            "fp64_flops_per_eval": round(2 * cnt["dfma"] + cnt["dadd"] + cnt["dmul"]),
            "fp64_inst_per_eval": {k: round(v) for k, v in cnt.items()},
            "fp64_source": "ncu smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum of one launch (2 flops per DFMA), "
                           "profiles/r2_fp64_inst_counts.csv"}
    if traffic:
        with open(os.path.join(HERE, "eval_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        with open(os.path.join(HERE, "r2_fp64_inst_counts.csv"), "w") as f:
            w = csv.writer(f); w.writerow(["kernel", "metric", "unit", "value"])
            for row in lines:
                w.writerow([row["Kernel Name"].split("(")[0], row["Metric Name"], row["Metric Unit"], row["Metric Value"]])
    for src, dst in ((f"{tag}_bench_1gpu.json", "r2_bench_1gpu.json"), (f"{tag}_bench_reference_arm.json", "r2_bench_reference_arm.json"),
                     (f"{tag}_time_aba.json", "r2_time_aba.json"), (f"{tag}_time_minv.json", "r2_time_minv.json"),
                     (f"{tag}_time_drnea.json", "r2_time_drnea.json")):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(HERE, dst))
    print("refreshed from", tag)


if __name__ == "__main__":
    main()
