"""Quick K4 timing (development aid): seeded pairs shaped like pass-1 leaf alignments."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import ratatosk_b200 as rb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = np.random.default_rng(0x4D594552)
tl = rng.integers(41, 1032, size=n)
ts, qs = [], []
base = rng.integers(0, 4, size=int(tl.sum()) + 16).astype(np.uint8)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
off = 0
for i in range(n):
    t = lut[base[off:off + tl[i]]]
    off += tl[i]
    keep = rng.random(tl[i]) > 0.05
    q = t[keep].copy()
    sub = rng.random(len(q)) < 0.04
    q[sub] = lut[rng.integers(0, 4, size=int(sub.sum()))]
    ts.append(t.tobytes()); qs.append(q.tobytes() if len(q) else b"A")
modes = (np.arange(n) % 4).clip(0, 2).astype(np.uint8)  # NW:SHW:HW = 1:1:2
cells = float(sum(len(q) * len(t) for q, t in zip(qs, ts)))
ctx = rb.Context(0)
for it in range(3):
    st = []
    t0 = time.time()
    d, e = ctx.edlib_batch(qs, ts, modes, stats=st)
    dt = time.time() - t0
    print("myers: n=%d cells=%.3e kernel_ms=%.2f total_s=%.2f -> %.1f GCUPS (kernel), %.2f M aln/s; mean dist %.1f" %
          (n, cells, st[2] / 1e6, dt, cells / (st[2] / 1e9) / 1e9, n / (st[2] / 1e9) / 1e6, d.mean()))
