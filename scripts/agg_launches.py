"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last forward pass)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
per_fw = int(sys.argv[2]) if len(sys.argv) > 2 else 0
skip_tail = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}[row["Metric Unit"]]
    rows.append((re.sub(r"\(.*", "", row["Kernel Name"]), v / 1e3, row["Grid Size"], row["Block Size"]))
print("launches captured:", len(rows))
fw = rows[-(per_fw + skip_tail):len(rows) - skip_tail] if per_fw else rows
tot = sum(r[1] for r in fw)
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, g, b in fw:
    agg[n][0] += 1
    agg[n][1] += v
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} n={c:5d} {v/1e3:9.3f} ms {100*v/tot:5.1f}%")
print(f"total {tot/1e3:.3f} ms over {len(fw)} launches")
if "--gemm" in sys.argv:
    sh = collections.defaultdict(lambda: [0, 0.0])
    for n, v, g, b in fw:
        if "gemm_tc" in n or "attn_flash" in n:
            sh[(n[-20:], g)][0] += 1
            sh[(n[-20:], g)][1] += v
    for k, (c, v) in sorted(sh.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {k[0]:22s} grid={k[1]:18s} n={c:3d} total={v/1e3:8.3f} ms avg={v/c:8.1f} us")
