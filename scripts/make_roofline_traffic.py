"""profiles/roofline_traffic.json from the ncu per-kernel table of the scene step
(gpurun_out/r02_scene_kernels.csv, written by scripts/profile_r2.sh): DRAM bytes, duration and tensor-pipe activity
per kernel.  bench.py copies `stage_dram_bytes` into its roofline line."""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_scene_kernels.csv")
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        d.setdefault((r[ii], name), {})[r[mi]] = float(r[vi].replace(",", ""))
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none on scripts/profile_scene.py "
                 "(PaviaU shape, third of four scenes), round 2: scripts/profile_r2.sh -> profiles/r02_scene_kernels.csv",
       "stage_dram_bytes": {}, "stage_detail": {}}
for (_, k), m in d.items():
    rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    out["stage_dram_bytes"][k] = int(rd + wr)
    out["stage_detail"][k] = {"dram_read": int(rd), "dram_write": int(wr), "ncu_us": m["gpu__time_duration.sum"] / 1e3,
                              "tensor_pipe_pct": m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]}
out["conv2_scene_dram_bytes_per_launch"] = out["stage_dram_bytes"]["conv2_scene_kernel"]
out["step_dram_bytes"] = sum(out["stage_dram_bytes"].values())
json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
shutil.copy(src, os.path.join(ROOT, "profiles", "r02_scene_kernels.csv"))
print(json.dumps(out["stage_dram_bytes"]), out["step_dram_bytes"] / 1e9, "GB")
