"""Turns the .ncu-rep captures and the launch list of scripts/gpu_bench.sh into the text summaries
committed under profiles/ (gpurun_out/ is scratch). Usage: python scripts/ncu_summary.py <round tag>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
src = os.path.join(ROOT, "gpurun_out")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed_op_ldgsts.sum", "smsp__inst_executed_pipe_uniform.sum"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]


traffic = {}
for name in sorted(os.listdir(src)):
    if not name.endswith(".ncu-rep"):
        continue
    hdr, units, rows = raw(os.path.join(src, name))
    lines = [f"# ncu --set full --clock-control none --import-source on  ({name}; raw page, selected metrics)"]
    for r in rows:
        kn = r[hdr.index("Kernel Name")]
        lines.append(f"kernel: {kn}")
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:60s} {r[i]:>16s} {units[i]}")
                vals[w] = (r[i], units[i])
        try:
            def tob(v, u):
                f = float(v.replace(",", ""))
                return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            t = tob(*vals["dram__bytes_read.sum"]) + tob(*vals["dram__bytes_write.sum"])
            lines.append(f"  dram traffic (read+write) per launch: {t / 1e6:.1f} MB")
            key = "spmv_dot" if "EpiDot," in kn else "cg_update" if "OpCgUpdateEigen" in kn else "cg_dir" if "OpCgDirEigen" in kn else None
            if key:
                traffic.setdefault(key + "_dram_bytes_per_launch", t)
        except Exception as e:  # noqa: BLE001
            lines.append(f"  (traffic not computed: {e})")
    with open(os.path.join(out_dir, f"{tag}_{name.replace('.ncu-rep', '')}_ncu_summary.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")

lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = list(csv.DictReader(l for l in open(lp) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python scripts/profile_target.py (C2, 12 CG iterations + setup)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[0]:5d} launches {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%  {k}\n")
    os.replace(lp, os.path.join(out_dir, f"{tag}_launches.csv")) if False else None
    import shutil
    shutil.copy(lp, os.path.join(out_dir, f"{tag}_launches.csv"))
if traffic:
    json.dump(traffic, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)
print("wrote summaries to", out_dir, traffic)
