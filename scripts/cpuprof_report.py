"""Symbolise a RTK_CPU_PROFILE dump: self and inclusive sample counts per function (development aid)."""
import bisect, collections, subprocess, sys
maps, samples = [], []
for line in open(sys.argv[1]):
    if line.startswith("M "):
        p = line.split()
        lo, hi = [int(x, 16) for x in p[1].split("-")]
        off = int(p[3], 16)
        path = p[6] if len(p) > 6 else ""
        if path.endswith("ratatosk_b200/librtk_b200.so"):
            import os
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ratatosk_b200", "librtk_b200.so")
        maps.append((lo, hi, off, path))
    elif line.startswith("S"):
        samples.append([int(x, 16) for x in line.split()[1:]])
maps.sort()
los = [m[0] for m in maps]
def locate(a):
    i = bisect.bisect_right(los, a) - 1
    if i >= 0 and a < maps[i][1]:
        return maps[i][3], a - maps[i][0] + maps[i][2]
    return "?", a
by_file = collections.defaultdict(set)
for s in samples:
    for j, a in enumerate(s):
        f, o = locate(a - (1 if j else 0))
        by_file[f].add(o)
sym = {}
for f, offs in by_file.items():
    offs = sorted(offs)
    if not f.startswith("/") :
        for o in offs: sym[(f, o)] = f
        continue
    try:
        out = subprocess.run(["addr2line", "-f", "-C", "-e", f] + [hex(o) for o in offs], capture_output=True, text=True).stdout.split("\n")
        for i, o in enumerate(offs):
            name = out[2 * i] if 2 * i < len(out) else "??"
            sym[(f, o)] = name if name != "??" else f.split("/")[-1]
    except Exception:
        for o in offs: sym[(f, o)] = f.split("/")[-1]
selfc, incl = collections.Counter(), collections.Counter()
for s in samples:
    names = []
    for j, a in enumerate(s):
        f, o = locate(a - (1 if j else 0))
        names.append(sym.get((f, o), "?"))
    if names:
        selfc[names[0]] += 1
        for n in set(names): incl[n] += 1
tot = len(samples)
print("samples", tot)
print("---- self")
for n, c in selfc.most_common(45): print("%6.2f%%  %s" % (100.0 * c / tot, n[:150]))
print("---- inclusive")
for n, c in incl.most_common(70): print("%6.2f%%  %s" % (100.0 * c / tot, n[:150]))
