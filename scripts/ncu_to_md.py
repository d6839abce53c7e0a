"""ncu --csv launch list (gpu__time_duration.sum, optional tensor-pipe / dram / L2 metrics) -> markdown table on stdout.
usage: ncu_to_md.py <csv> [skip_first_n_launches]"""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
col = {n: i for i, n in enumerate(hdr)}
by = collections.OrderedDict()
for r in rows:
    key = r[col["ID"]]
    d = by.setdefault(key, {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    try:
        v = float(r[col["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[col["Metric Unit"]]
    name = r[col["Metric Name"]]
    if name == "gpu__time_duration.sum":
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    d[name] = v
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
items = list(by.values())[skip:]
tot = sum(d.get("gpu__time_duration.sum", 0) for d in items)
print(f"{len(items)} launches, {tot / 1e3:.3f} ms\n")
extra = [m for m in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum") if any(m in d for d in items)]
print("| # | kernel | µs | share % |" + "".join(f" {m.split('.')[0].replace('sm__pipe_tensor_cycles_active','tensor %').replace('dram__bytes_','dram ')} |" for m in extra) + " grid |")
print("|---|---|---|---|" + "---|" * len(extra) + "---|")
agg = collections.OrderedDict()
for i, d in enumerate(items):
    t = d.get("gpu__time_duration.sum", 0)
    nm = d["name"][:60]
    print(f"| {i} | `{nm}` | {t:.1f} | {100 * t / tot:.1f} |" + "".join(f" {d.get(m, 0):.3g} |" for m in extra) + f" {d['grid']} |")
    a = agg.setdefault(d["name"].split("(")[0][:60], [0, 0.0]); a[0] += 1; a[1] += t
print("\n| kernel | launches | total µs | share % |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} |")
