"""Per-kernel time and DRAM traffic of the LAST whole step in an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
(steps are delimited by patchify16 launches).  Usage: analyze_traffic.py list.csv [out.json]"""
import collections
import csv
import json
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = launches.setdefault(r["ID"], dict(name=r["Kernel Name"], grid=r["Grid Size"]))
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT[r["Metric Unit"]]
    return list(launches.values())


def short(n):
    n = re.sub(r"\(.*", "", re.sub(r"void |rsp::|v2::", "", n))
    return n[:64] if "gemm" in n else re.sub(r"<.*", "", n)[:64]


def main(path, out=None):
    L = load(path)
    idx = [i for i, r in enumerate(L) if "patchify" in r["name"]]
    seg = L[idx[-1]:]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in seg:
        a = agg[short(r["name"])]
        a[0] += 1
        a[1] += r["gpu__time_duration.sum"]
        a[2] += r["dram__bytes_read.sum"]
        a[3] += r["dram__bytes_write.sum"]
    tot = sum(a[1] for a in agg.values())
    print(f"last step: {len(seg)} launches, {tot / 1e3:.2f} ms kernel time (serialised, cold-cache), "
          f"{sum(a[2] + a[3] for a in agg.values()) / 1e9:.2f} GB DRAM traffic")
    print(f"{'ms':>8} {'%':>5} {'n':>5} {'rd GB':>7} {'wr GB':>7} {'GB/s':>7}  kernel")
    rows = []
    for n, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        rows.append(dict(kernel=n, launches=c, ms=t / 1e3, dram_read_bytes=rd, dram_write_bytes=wr))
        if len(rows) <= 30:
            print(f"{t / 1e3:8.3f} {100 * t / tot:5.1f} {c:5d} {rd / 1e9:7.3f} {wr / 1e9:7.3f} {(rd + wr) / t / 1e3:7.0f}  {n}")
    if out:
        with open(out, "w") as f:
            json.dump(dict(source=path, launches=len(seg), kernel_ms=tot / 1e3, kernels=rows), f, indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:3])
