"""Turn one kernel of an .ncu-rep into the small JSON + markdown summary committed under profiles/.
usage: ncu_to_profile.py rep out_prefix width height [kernel_index [bytes_per_px]]"""
import csv, io, json, subprocess, sys
rep, out, W, H = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
ki = int(sys.argv[5]) if len(sys.argv) > 5 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + ki]
g = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def num(k):
    v, u = g[k]
    v = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}.get(u, 1)
    return v * scale


d = {
    "kernel": g["Kernel Name"][0], "width": W, "height": H,
    "gpu_time_us": num("gpu__time_duration.sum"),
    "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
    "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "registers_per_thread": num("launch__registers_per_thread"),
    "grid": g["launch__grid_size"][0], "block": g["launch__block_size"][0],
    "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "inst_executed": num("smsp__inst_executed.sum"),
    "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
    "stalls_per_issue": {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(vals[i]), 3)
                         for i, h in enumerate(hdr) if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h and float(vals[i] or 0) >= 0.05},
    "command": "ncu --set full --clock-control none --import-source on (see the .md beside this file)",
}
alg = (float(sys.argv[6]) if len(sys.argv) > 6 else 88.0) * W * H
d["algorithmic_bytes_per_launch"] = alg
d["algorithmic_gbs_this_launch"] = alg / d["gpu_time_us"] / 1e3
json.dump(d, open(out + ".json", "w"), indent=1)
print(json.dumps(d, indent=1))
