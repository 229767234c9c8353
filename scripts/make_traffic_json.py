"""profiles/traffic.json (read by bench.py for `roofline.traffic`) from an ncu step capture with DRAM byte counters:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file X.csv python scripts/profile_step.py
    python scripts/make_traffic_json.py X.csv [profiles/traffic.json]
"""
import collections
import csv
import json
import os
import sys

src = sys.argv[1]
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
rows = list(csv.reader(open(src)))
hdr = next(r for r in rows if r and r[0] == "ID")
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = collections.defaultdict(lambda: collections.defaultdict(float))
ids = collections.defaultdict(set)
for r in rows:
    if len(r) != len(hdr) or r[0] == "ID":
        continue
    name = r[ik]
    for key, pat in (("mlp_pair_kernel", "mlp_pair_kernel"), ("gemm_pair_kernel<1,2,0>", "gemm_pair_kernel<1, 2, 0>"),
                     ("attn_tc_fwd_kernel", "attn_tc_fwd_kernel")):
        if pat in name:
            per[key][r[im]] += float(r[iv].replace(",", ""))
            ids[key].add(r[0])
out = {}
for key, m in per.items():
    rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    out[key] = {"launches_per_step": len(ids[key]), "dram_bytes_per_step": rd + wr, "dram_read": rd, "dram_write": wr,
                "time_us_per_step": m["gpu__time_duration.sum"] / 1e3,
                "source": f"profiles/{os.path.basename(src)} (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one eager step, B=1024)"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
