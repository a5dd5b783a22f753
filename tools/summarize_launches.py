"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share over ONE Newton step (the launches between two consecutive Jacobian assemblies).
usage: python tools/summarize_launches.py gpurun_out/launches.csv "title" > profiles/xxx.csv"""
import collections
import csv
import re
import sys


def main(path, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi, bi = (hdr.index(n) for n in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
    names = [r[ki] for r in data]
    # Jacobian assembly = sweep/rows kernel instantiated with WANTJ = true (last template argument 1)
    starts = [i for i, n in enumerate(names) if re.search(r"pq_(sweep|rows)_kernel<[^>]*, 1>", n)]
    if len(starts) < 2:
        raise SystemExit(f"need two Jacobian assemblies in the capture, found {len(starts)}")
    s, e = starts[0], starts[1]

    def short(n):
        n = n.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        n = re.sub(r"\(.*", "", n)
        m = re.match(r"(void )?(cm::)?([A-Za-z_0-9:]+)(<[^>]*>)?", n)
        return (m.group(3) + (m.group(4) or "")) if m else n

    agg = collections.defaultdict(lambda: [0, 0.0])
    big = collections.defaultdict(lambda: [0, 0.0])
    for r in data[s:e]:
        v = float(r[vi].replace(",", ""))
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += v
        b = big[(short(r[ki]), r[gi], r[bi])]
        b[0] += 1
        b[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {title}")
    print("# one Newton step = launches between two consecutive Jacobian assemblies; durations are "
          "cold-cache and serialised under ncu: compare SHARES, not absolutes")
    print(f"# total {tot / 1e6:.2f} ms over {e - s} launches")
    print("kernel,launches,total_ms,share")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'"{k}",{c},{t / 1e6:.3f},{t / tot:.4f}')
    print("# largest (kernel, grid, block) groups: avg duration per launch")
    print("kernel,grid,block,launches,avg_us")
    for (k, g, b), (c, t) in sorted(big.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f'"{k}","{g}","{b}",{c},{t / c / 1e3:.1f}')


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
