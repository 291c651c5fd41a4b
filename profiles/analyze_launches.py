"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of one
warm step (between two consecutive patchify16 launches) and the individually slow launches."""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e3 if u == "ns" else v if u == "us" else v * 1e3


def main(path, step=3, slow_us=400.0):
    rows = load(path)
    idx = [i for i, r in enumerate(rows) if "patchify" in r["Kernel Name"]]
    seg = rows[idx[step]:idx[step + 1]]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        n = re.sub(r"\(.*", "", re.sub(r"void |rsp::", "", r["Kernel Name"]))
        n = re.sub(r"<.*", "", n) if "gemm" not in n else n[:60]
        agg[n][0] += 1
        agg[n][1] += us(r)
    tot = sum(v[1] for v in agg.values())
    print(f"step {step}: {len(seg)} launches, {tot / 1e3:.2f} ms of kernel time")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{c:4d}  {n[:80]}")
    print("-- launches slower than", slow_us, "us")
    for i, r in enumerate(seg):
        t = us(r)
        if t > slow_us:
            n = re.sub(r"\(.*", "", re.sub(r"void |rsp::", "", r["Kernel Name"]))[:70]
            print(f"{i:4d} {t:9.1f} us grid={r['Grid Size']:16s} {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3, float(sys.argv[3]) if len(sys.argv) > 3 else 400.0)
