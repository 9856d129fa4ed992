"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
    python tools/summarize_ncu.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches.md"""
import csv
import io
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n += 1
        if n <= skip:
            continue
        name = r["Kernel Name"].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"^void\s+", "", name)
        m = re.match(r"([\w:]+(?:<[\d, a-z]+>)?)", name)   # keep small integer/bool template args (e.g. gemm<0, 1>)
        name = m.group(1) if m else name[:60]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(unit, 1)
        agg[name][0] += 1
        agg[name][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"launches: {sum(v[0] for v in agg.values())}, total device time {tot / 1e6:.2f} ms (ncu: serialised, cold caches)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.2f} |")


if __name__ == "__main__":
    main()
