"""Quick end-to-end correction timing on the golden F2 index (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import ratatosk_b200 as rb
from common import GOLDEN, golden_paths, load_golden_reads, read_fastq

recipe = sys.argv[1] if len(sys.argv) > 1 else "F2"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 8
fa, rt = golden_paths(recipe)
g = rb.Graph.load(fa, rt, 31)
ctx = rb.Context(0)
ctx.upload(g)
reads = load_golden_reads(recipe)
gold = read_fastq(os.path.join(GOLDEN, recipe, "corrected_pass1.fastq.gz"))
seqs = [r[1] for r in reads] * rep
quals = [r[2] for r in reads] * rep
nb = sum(len(s) for s in seqs)
for it in range(2):
    st = []
    t = time.time()
    out = ctx.correct(seqs, quals, stats=st)
    dt = time.time() - t
    ok = all((out[i][0], out[i][1]) == (gold[i % len(reads)][1], gold[i % len(reads)][2]) for i in range(len(out)))
    print("correct: reads=%d bases=%d time=%.2fs -> %.3f Mbases/s  waves=%d jobs=%d  identical_to_reference=%s" %
          (len(seqs), nb, dt, nb / dt / 1e6, st[5], st[6], ok))
