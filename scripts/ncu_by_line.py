"""Join `ncu --page source --csv` (SASS rows with addresses) with `nvdisasm --print-line-info` of the same cubin and
aggregate executed instructions / stall samples per source line (development aid).
usage: ncu_by_line.py <ncu_source.csv> <nvdisasm.sass> <kernel substring> [top]"""
import csv, re, sys
src_csv, sass, want = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file, line) inside the wanted function
line_of, cur, infn = {}, None, False
for l in open(sass, errors="replace"):
    if l.startswith("//---") and ".text." in l:
        infn = want in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
agg = {}
hdr = None
active = False
base = None
for r in rows:
    if r and r[0] == "Kernel Name":
        active = want in r[1]; hdr = None; base = None
        continue
    if not active:
        continue
    if hdr is None:
        hdr = r
        ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    try:
        a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = a
    key = line_of.get(a - base, ("?", 0))
    ins = int(r[ii]) if r[ii].isdigit() else 0
    sm = int(r[isamp]) if r[isamp].isdigit() else 0
    x = agg.setdefault(key, [0, 0])
    x[0] += ins; x[1] += sm
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-14s %5d  inst %5.1f%%  samples %5.1f%%" % (k[0], k[1], 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s))
