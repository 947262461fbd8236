"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total, average, share).

    python scripts/ncu_summary.py gpurun_out/launches.csv [first_launch [last_launch]]

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import sys
from collections import OrderedDict


def rows(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {n: i for i, n in enumerate(hdr)}
    for r in rd:
        if len(r) != len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[ix["Metric Unit"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        yield int(r[ix["ID"]]), r[ix["Kernel Name"]], v * scale


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    agg = OrderedDict()
    total = 0.0
    n = 0
    for i, name, us in rows(path):
        if i < lo or i >= hi:
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    print(f"{n} launches, {total / 1e3:.3f} ms serialised\n")
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---|---|---|---|")
    for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {c} | {us / 1e3:.3f} | {us / c:.1f} | {100 * us / total:.1f}% |")


if __name__ == "__main__":
    main()
