"""Summarise ncu outputs into profiles/ (development aid).
  launches <csv> <out.md> <command>       : per-kernel totals/shares from a `--metrics gpu__time_duration.sum --csv` log
  raw <raw.csv> <out.md> <title>          : selected metrics per launch from `ncu -i rep --page raw --csv`"""
import collections, csv, re, sys

def launches(path, out, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "second": 1e9}.get(u, 1)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n\nCommand (B200, one GPU): `%s`\n\nPer-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n" % (out, cmd))
        f.write("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.2f | %.1f | %.1f%% |\n" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot))
        f.write("\nTotal kernel time %.1f ms over %d launches.\n" % (tot / 1e6, sum(v[0] for v in agg.values())))

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]

def raw(path, out, title):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    kn = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write("# %s\n\n%s\n\n" % (out, title))
        f.write("| kernel | " + " | ".join("%s [%s]" % (w, units[i]) for w, i in cols) + " |\n")
        f.write("|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            f.write("| `%s` | " % re.sub(r"\(.*", "", r[kn]) + " | ".join(r[i] for _, i in cols) + " |\n")

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](*sys.argv[2:5])
