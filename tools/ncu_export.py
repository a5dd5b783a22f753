"""profiles/ncu_metrics_r2.csv from an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv`:
one row per kernel CLASS of the Newton step (average over the captured launches of that class): duration, DRAM
read / write bytes, DRAM throughput %, fp64 pipe %, warps active %, registers, and the largest issue-stall reasons.
usage: python tools/ncu_export.py raw.csv > profiles/ncu_metrics_r2.csv"""
import csv
import re
import sys
from collections import defaultdict


def classify(name, grid, block):
    g = int(re.sub(r"[^0-9,]", "", grid).split(",")[0] or 0)
    b = int(re.sub(r"[^0-9,]", "", block).split(",")[0] or 0)
    if "zebra_ring_kernel" in name:
        m = re.search(r"zebra_ring_kernel<\s*(?:\(bool\))?\s*(\w+)", name)   # first template argument: COUPLED
        coupled = bool(m) and m.group(1) in ("1", "true")
        m2 = re.search(r"zebra_ring_kernel<[^,>]*,\s*(?:\(bool\))?\s*(\w+)", name)   # second: F32 operator pack
        sfx = "_f32op" if (m2 and m2.group(1) in ("1", "true")) else ""
        if b >= 800:
            return ("zebra_fine_coupled" if coupled else "zebra_fine_zero") + sfx
        return (("zebra_level1_coupled" if coupled else "zebra_level1_zero") if b >= 400 else (
            "zebra_level2_coupled" if coupled else "zebra_level2_zero")) + sfx
    if "zebra_mid_kernel" in name:
        m = re.search(r"zebra_mid_kernel<\s*(?:\(int\))?\s*(\d)", name)
        return "zebra_mid_mode" + (m.group(1) if m else "?")
    if "coarse_smem_kernel" in name:
        return "coarse_smem_cycle"
    if "pack_f32" in name:
        return "pack_f32"
    if "pq_stencil_spmv" in name:
        return "pq_stencil_spmv"
    if "gd_dot_kernel" in name:
        return "gmres_dot"
    if "gd_update_kernel" in name:
        return "gmres_update"
    if "pq_sweep_kernel" in name:
        return "pq_sweep_J_and_F" if re.search(r", (1|true|\(bool\)1)>", name) else "pq_sweep_F_only"
    if "extract_blocks" in name:
        return "extract_blocks"
    if "resid_restrict" in name:
        return "resid_restrict_fine" if g > 10000 else "resid_restrict_coarser"
    if "unscale_dev" in name:
        return "unscale"
    if "prolong_odd" in name:
        return "prolong_odd_fine" if g > 1000 else "prolong_odd_coarser"
    if "line_factor" in name:
        return "line_factor_scan"
    return None


WANT = [
    ("duration_us", r"^gpu__time_duration\.sum$", 1e-3),           # ns -> us
    ("dram_read_bytes", r"^dram__bytes_read\.sum$", None),
    ("dram_write_bytes", r"^dram__bytes_write\.sum$", None),
    ("dram_throughput_pct", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", 1.0),
    ("fp64_pipe_pct", r"^sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active$|^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active$", 1.0),
    ("warps_active_pct", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", 1.0),
    ("issue_active_pct", r"^sm__inst_issued\.avg\.pct_of_peak_sustained_active$|^smsp__issue_active\.avg\.pct$", 1.0),
    ("registers_per_thread", r"^launch__registers_per_thread$", 1.0),
    ("l2_hit_pct", r"^lts__t_sector_hit_rate\.pct$", 1.0),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6,
              "msecond": 1e6, "nsecond": 1.0, "second": 1e9}


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [(n, i) for n, i in col.items()
                  if re.match(r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", n)]
    agg = defaultdict(lambda: defaultdict(float))
    cnt = defaultdict(int)
    for r in data:
        if len(r) < len(hdr):
            continue
        cls = classify(r[col["Kernel Name"]], r[col["Grid Size"]], r[col["Block Size"]])
        if cls is None:
            continue
        cnt[cls] += 1
        for key, rx, _ in WANT:
            for n, i in col.items():
                if re.match(rx, n):
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    u = units[i]
                    if key.endswith("_bytes"):
                        v *= UNIT_SCALE.get(u, 1.0)
                    elif key == "duration_us":
                        v *= UNIT_SCALE.get(u, 1.0) * 1e-3
                    agg[cls][key] += v
                    break
        for n, i in stall_cols:
            try:
                agg[cls]["stall:" + n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] += float(
                    r[i].replace(",", ""))
            except ValueError:
                pass
    keys = [k for k, _, _ in WANT]
    w = csv.writer(sys.stdout)
    w.writerow(["class", "launches_captured"] + keys + ["top_stalls_per_issue"])
    for cls in sorted(cnt):
        n = cnt[cls]
        st = sorted(((k[6:], v / n) for k, v in agg[cls].items() if k.startswith("stall:")), key=lambda t: -t[1])[:4]
        w.writerow([cls, n] + [f"{agg[cls][k] / n:.6g}" if k in agg[cls] else "" for k in keys] +
                   ["; ".join(f"{k} {v:.2f}" for k, v in st)])


if __name__ == "__main__":
    main(sys.argv[1])
