"""Turn gpurun_out/launches*.csv (+ optional .ncu-rep full capture) into a markdown summary under profiles/.
usage: python scripts/ncu_summary.py <launches.csv> <out.md> [prof.ncu-rep] [title]"""
import collections, csv, subprocess, sys

launch_csv, out_md = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
title = sys.argv[4] if len(sys.argv) > 4 else "ncu summary"
lines = [l for l in open(launch_csv) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
md = [f"# {title}", "", f"Source: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
      f"(`{launch_csv}`); per-launch times are cold-cache and serialised — compare SHARES, not absolutes.", "",
      "| kernel | launches | total µs | avg µs | share |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    md.append(f"| `{k[:110]}` | {len(v)} | {sum(v):.1f} | {sum(v)/len(v):.1f} | {sum(v)/tot:.3f} |")
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
    idx = [hdr.index(w) for w in want if w in hdr]
    md += ["", f"## `ncu --set full` capture (`{rep}`)", "",
           "| kernel | " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx) + " |", "|---|" + "---:|" * len(idx)]
    kn = hdr.index("Kernel Name")
    for r in rows[2:]:
        md.append(f"| `{r[kn][:70]}` | " + " | ".join(r[i] for i in idx) + " |")
open(out_md, "w").write("\n".join(md) + "\n")
print("\n".join(md))
