"""Per-kernel summary of an ncu launch list (CSV written by `ncu --csv --log-file ... --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]`).

    python tools/summarize_launches.py LAUNCHES.csv "title line" > profiles/NAME_summary.txt

Per-launch times under ncu are cold-cache and serialised: the kernels' SHARES of the step are the comparable quantity.
"""
import collections
import csv
import sys

TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}
BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        m = d["Metric Name"]
        scale = TIME[d["Metric Unit"]] if m.startswith("gpu__time") else BYTES.get(d["Metric Unit"], 0.0)
        agg.setdefault((d["ID"], d["Kernel Name"]), {})[m] = v * scale
    tot = collections.OrderedDict()
    for (_, k), m in agg.items():
        name = k.split("(")[0].replace("void ", "").split("::")[-1]
        t = tot.setdefault(name, [0, 0.0, 0.0, 0.0])
        t[0] += 1
        t[1] += m.get("gpu__time_duration.sum", 0.0)
        t[2] += m.get("dram__bytes_read.sum", 0.0)
        t[3] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(v[1] for v in tot.values())
    print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    print("(per-launch times are cold-cache and serialised: compare shares, not absolutes)\n")
    for name, v in tot.items():
        line = f"{name:<40} n={v[0]:3d} time {v[1]:8.3f} ms ({100 * v[1] / total:5.1f}%)"
        if v[2] or v[3]:
            line += f"  dram rd {v[2]:9.1f} MB wr {v[3]:9.1f} MB"
        print(line)
    print(f"{'total':<40} n={sum(v[0] for v in tot.values()):3d} time {total:8.3f} ms" +
          (f"  dram rd {sum(v[2] for v in tot.values()):9.1f} MB wr {sum(v[3] for v in tot.values()):9.1f} MB" if any(v[2] for v in tot.values()) else ""))


if __name__ == "__main__":
    main()
