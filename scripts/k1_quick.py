"""Quick K1 timing on the golden F2 graph (development aid, not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import ratatosk_b200 as rb
from common import golden_paths, load_golden_reads

fa, rt = golden_paths("F2")
g = rb.Graph.load(fa, rt, 31)
ctx = rb.Context(0)
ctx.upload(g)
reads = [s for _, s, _ in load_golden_reads("F2")]
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 40
reads = reads * rep
nb = sum(len(r) for r in reads)
for it in range(3):
    st = []
    t = time.time()
    ctx.search_sequence(reads, exact=False, insertion=True, deletion=True, substitution=True, or_exclusive_match=True, stats=st)
    dt = time.time() - t
    print("inexact: bases=%d probes=%d raw_hits=%d kernel_ms=%.3f total_s=%.3f  -> %.2f G lookups/s (kernel), %.1f probes/base" %
          (nb, st[0], st[1], st[2] / 1e6, dt, st[0] / (st[2] / 1e9) / 1e9, st[0] / nb))
for it in range(2):
    st = []
    t = time.time()
    ctx.search_sequence(reads, stats=st)
    print("exact: kernel_ms=%.3f total_s=%.3f raw_hits=%d" % (st[2] / 1e6, time.time() - t, st[1]))
st = []
t = time.time()
ctx.get_seeds(reads, stats=st)
print("get_seeds total_s=%.3f stats=%s" % (time.time() - t, st))
