import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ratatosk_b200 as rb
import refseams as R
import bench
from concurrent.futures import ThreadPoolExecutor
haps = bench.load_haplotypes()
seq, qual, off = bench.make_reads(haps, int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000, seed=20261017)
reads = [(seq[int(off[i]):int(off[i + 1])].tobytes().decode(), qual[int(off[i]):int(off[i + 1])].tobytes().decode()) for i in range(len(off) - 1)]
d = os.path.join(ROOT, "bench_data", "F3")
fa, rt = os.path.join(d, "index.k31.fasta.gz"), os.path.join(d, "index.k31.rtsk")
ref = R.RefGraph(fa, rt, 31, threads=8)
wants = []
for rep in range(2):
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        wants.append(list(ex.map(lambda r: ref.correct_read(r[0], r[1], False), reads)))
print("reference deterministic:", wants[0] == wants[1], [i for i in range(len(reads)) if wants[0][i] != wants[1][i]])
want = wants[0]
g = rb.Graph.load(fa, rt, 31)
ctx = rb.Context(0); ctx.upload(g)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 8):
    got = ctx.correct([r[0] for r in reads], [r[1] for r in reads])
    bad = [i for i in range(len(reads)) if got[i] != want[i]]
    msg = ""
    for i in bad[:3]:
        a, b = got[i], want[i]
        fs = next((x for x in range(min(len(a[0]), len(b[0]))) if a[0][x] != b[0][x]), None)
        fq = next((x for x in range(min(len(a[1]), len(b[1]))) if a[1][x] != b[1][x]), None)
        msg += " [read %d len %d/%d seqdiff@%s qualdiff@%s]" % (i, len(a[0]), len(b[0]), fs, fq)
    print("rep", rep, "reads", len(reads), "mismatches", bad, msg)
