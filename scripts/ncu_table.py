"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active...,gpu__dram_throughput... --csv` launch list.

    python scripts/ncu_table.py launches.csv [first_index [last_index]]     (indices into the launch order, not ncu IDs)

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import sys

SCALE_T = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
SCALE_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {n: i for i, n in enumerate(hdr)}
    rows = collections.OrderedDict()
    for r in rd:
        if len(r) != len(hdr):
            continue
        d = rows.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
        m, u, v = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
        if m == "gpu__time_duration.sum":
            d["us"] = v * SCALE_T.get(u, 1.0)
        elif m.startswith("dram__bytes_read"):
            d["rd"] = v * SCALE_B.get(u, 1)
        elif m.startswith("dram__bytes_write"):
            d["wr"] = v * SCALE_B.get(u, 1)
        elif m.startswith("sm__pipe_tensor"):
            d["tc"] = v
        elif m.startswith("gpu__dram_throughput"):
            d["dram"] = v
    return list(rows.values())


def short(name):
    return name.split("(")[0].replace("void ", "").replace("tacorl::", "")[:70]


def main():
    rows = load(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
    sel = rows[lo:hi]
    tot = sum(r.get("us", 0) for r in sel)
    trf = sum(r.get("rd", 0) + r.get("wr", 0) for r in sel)
    print(f"{len(sel)} launches (of {len(rows)}), {tot / 1e3:.3f} ms serialised, DRAM read+write {trf / 1e6:.0f} MB\n")
    print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB | GB/s | tensor pipe % (max) | DRAM % (max) |")
    print("|---|---|---|---|---|---|---|---|---|")
    agg = collections.OrderedDict()
    for r in sel:
        a = agg.setdefault(short(r["name"]), [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += r.get("us", 0)
        a[2] += r.get("rd", 0)
        a[3] += r.get("wr", 0)
        a[4] = max(a[4], r.get("tc", 0))
        a[5] = max(a[5], r.get("dram", 0))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] < 0.002 * tot:
            continue
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.1f} | {a[3] / 1e6:.1f} | "
              f"{(a[2] + a[3]) / a[1] / 1e3 if a[1] else 0:.0f} | {a[4]:.1f} | {a[5]:.1f} |")


if __name__ == "__main__":
    main()
