"""Sum an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name: python tests/diag/ncu_by_kernel.py <csv> [top]"""
import csv, re, sys
lines = open(sys.argv[1], newline="").readlines()
start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
unit = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}
agg, total, n = {}, 0.0, 0
for r in csv.DictReader(lines[start:]):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    us = float(r["Metric Value"].replace(",", "")) * unit.get(r["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"].split("/")[-1]).replace("void ", "").strip()
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    total += us; n += 1
print(f"{n} launches, {total / 1e3:.2f} ms of kernel time")
for name, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{us / 1e3:8.3f} ms {100 * us / total:5.1f} % {c:5d} x {us / c:8.1f} us  {name}")
