"""Aggregate warp-stall samples per reason (and the top instructions) from `ncu -i rep --page source --csv` (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = r
        continue
    cur["rows"].append(r)
agg, tot, n_k = {}, 0, 0
top = []
for b in blocks:
    if want not in b["name"]:
        continue
    n_k += 1
    h = b["hdr"]
    ss, si = h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
    cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    for r in b["rows"]:
        if not r[ss].isdigit():
            continue
        tot += int(r[ss])
        top.append((int(r[ss]), r[si].strip()[:60]))
        for i in cols:
            if r[i].isdigit():
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print("kernels matched: %d, stall samples: %d" % (n_k, tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-28s %6d  %5.1f%%" % (k, v, 100.0 * v / max(tot, 1)))
print("top instructions:")
for c, s in sorted(top, reverse=True)[:8]:
    print("  %6d  %s" % (c, s))
