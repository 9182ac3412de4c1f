#!/usr/bin/env python
"""profiles/r02_traffic.json from `ncu --set full` raw pages (ncu -i X.ncu-rep --page raw --csv > X_raw.csv):
per step-kernel DRAM traffic per launch = dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured
launches, stamped with the source digest of the build that is in the tree NOW (run it right after the capture).
bench.py reports these as `roofline.traffic` and says so when the digest no longer matches.

usage: python tools/ncu_traffic.py gpurun_out/r02_*_raw.csv [...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# step-kernel index of bench.py STEP_KERNELS <- substring of the ncu kernel name
TIME_US = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
MATCH = [(0, "edge_chain_kernel"), (0, "gemm_tc_prod_kernel"), (1, "attn_aggregate_kernel"), (2, "gemm_tc_kernel<128, 1,"), (3, "gemm_tc_kernel<128, 5,"),
         (4, "graph_build_kernel"), (5, "rollout_persist_kernel")]


def main(paths):
    from gcbfplus_b200 import build as _b
    out = {"source_digest": _b._digest(), "kernels": {}, "files": [os.path.basename(p) for p in paths]}
    acc = {}
    for p in paths:
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        ix = {n: i for i, n in enumerate(hdr)}
        for r in rows[2:]:
            name = r[ix["Kernel Name"]]
            k = next((i for i, sub in MATCH if sub in name), None)
            if k is None:
                continue
            rd = float(r[ix["dram__bytes_read.sum"]]) * UNIT[units[ix["dram__bytes_read.sum"]]]
            wr = float(r[ix["dram__bytes_write.sum"]]) * UNIT[units[ix["dram__bytes_write.sum"]]]
            dur = float(r[ix["gpu__time_duration.sum"]]) * TIME_US.get(units[ix["gpu__time_duration.sum"]], 1.0)
            a = acc.setdefault(k, {"name": name.split("(")[0], "rd": [], "wr": [], "us": []})
            a["rd"].append(rd), a["wr"].append(wr), a["us"].append(dur)
    for k, a in acc.items():
        n = len(a["rd"])
        out["kernels"][str(k)] = {"ncu_kernel": a["name"], "launches": n, "dram_read_bytes": sum(a["rd"]) / n,
                                  "dram_write_bytes": sum(a["wr"]) / n,
                                  "dram_bytes_per_launch": (sum(a["rd"]) + sum(a["wr"])) / n,
                                  "ncu_duration_us_cold": sum(a["us"]) / n}
    dst = os.path.join(ROOT, "profiles", "r02_traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1:])
