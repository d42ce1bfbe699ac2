#!/usr/bin/env python3
"""Writes the record of one workload into profiles/traffic.json from a tools/ncu_summary.py JSON:
python tools/update_traffic.py <key> <selected.json> <capture description>"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key, sel_path, capture = sys.argv[1], sys.argv[2], sys.argv[3]
sel = json.load(open(sel_path))
mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def num(k):
    return float(sel[k][0].replace(",", ""))


dram = num("dram__bytes_read.sum") * mult[sel["dram__bytes_read.sum"][1]] + num("dram__bytes_write.sum") * mult[sel["dram__bytes_write.sum"][1]]
p = os.path.join(ROOT, "profiles", "traffic.json")
t = json.load(open(p))
t[key] = {"dram_bytes_per_launch": dram,
          "lsu_wavefronts_pct_of_peak": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
          "dram_throughput_pct_of_peak": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
          "gpu_time_us_under_ncu": num("gpu__time_duration.sum"), "capture": capture}
json.dump(t, open(p, "w"), indent=1)
print(key, json.dumps(t[key]))
