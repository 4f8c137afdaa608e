#!/usr/bin/env python3
"""F4: chr20-scale fixture (BASELINE.json configs[2]) — TEST / BENCH DATA generator, numpy-vectorised.

Same recipe as tests/golden/make_fixtures.py F2/F3 (diploid, 0.1 % het SNPs, repeat families, tandem repeats,
homopolymers, 30x PE150 at 0.2 % substitutions, ONT-like long reads at 10 % error) scaled to a 64 Mbp genome.
The pure-Python generator of the small fixtures would take hours at this size, so the draws are vectorised;
everything is a function of --seed.

Outputs under --out-dir:  genome.npz (hap0 2-bit packed + het SNPs, what bench.py draws its reads from),
sr.fastq (short reads, deleted after indexing), lr_sample.fastq (the long reads used for golden outputs and for
colouring the k = 63 graph).
"""
import argparse
import os
import sys

import numpy as np

COMP = np.array([3, 2, 1, 0], dtype=np.uint8)
LET = np.frombuffer(b"ACGT", dtype=np.uint8)


def genome_complex(rng, n, n_fam, n_tandem, n_homo):
    g = rng.integers(0, 4, n, dtype=np.uint8)
    for _ in range(n_fam):
        ln = int(rng.integers(300, 2501))
        unit = rng.integers(0, 4, ln, dtype=np.uint8)
        for _c in range(int(rng.integers(3, 8))):
            pos = int(rng.integers(0, n - ln - 1))
            cp = unit.copy()
            m = rng.random(ln) < 0.02
            cp[m] = (cp[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
            if rng.random() < 0.5:
                cp = COMP[cp][::-1]
            g[pos:pos + ln] = cp
    for _ in range(n_tandem):
        ul = int(rng.integers(2, 41))
        unit = rng.integers(0, 4, ul, dtype=np.uint8)
        cnt = int(rng.integers(3, 31))
        t = np.tile(unit, cnt)[:1200]
        pos = int(rng.integers(0, n - len(t) - 1))
        g[pos:pos + len(t)] = t
    for _ in range(n_homo):
        ln = int(rng.integers(6, 26))
        pos = int(rng.integers(0, n - ln - 1))
        g[pos:pos + ln] = rng.integers(0, 4)
    return g


def write_short_reads(rng, haps, cov, rl, ins_lo, ins_hi, sub, path, chunk=200_000):
    glen = len(haps[0])
    n_pairs = int(cov * glen / (2 * rl))
    qline = b"I" * rl
    ar = np.arange(rl, dtype=np.int64)
    with open(path, "wb") as f:
        for c0 in range(0, n_pairs, chunk):
            m = min(chunk, n_pairs - c0)
            hsel = rng.integers(0, len(haps), m)
            frag = rng.integers(ins_lo, ins_hi + 1, m)
            p = (rng.random(m) * (glen - frag)).astype(np.int64)
            flip = rng.random(m) < 0.5
            out = []
            for h in range(len(haps)):
                idx = np.nonzero(hsel == h)[0]
                if len(idx) == 0:
                    continue
                H = haps[h]
                st, fr, fl = p[idx], frag[idx], flip[idx]
                # mate 1 = first rl bases of the (possibly reverse-complemented) fragment, mate 2 = first rl of its revcomp
                fw1 = H[st[:, None] + ar[None, :]]                                   # fragment forward, from the left end
                rc_end = COMP[H[(st + fr - 1)[:, None] - ar[None, :]]]               # revcomp of the fragment's right end
                r1 = np.where(fl[:, None], rc_end, fw1)
                r2 = np.where(fl[:, None], fw1, rc_end)
                for r in (r1, r2):
                    mm = rng.random(r.shape) < sub
                    r[mm] = (r[mm] + rng.integers(1, 4, int(mm.sum()), dtype=np.uint8)) & 3
                out.append((idx, LET[r1], LET[r2]))
            buf = []
            order = np.concatenate([o[0] for o in out])
            r1s = np.concatenate([o[1] for o in out])
            r2s = np.concatenate([o[2] for o in out])
            inv = np.argsort(order)
            for j in inv:
                name = b"@sr%d\n" % (c0 + int(order[j]))
                buf.append(name + r1s[j].tobytes() + b"\n+\n" + qline + b"\n" + name + r2s[j].tobytes() + b"\n+\n" + qline + b"\n")
            f.write(b"".join(buf))
            print("  short reads: %d / %d pairs" % (c0 + m, n_pairs), file=sys.stderr, flush=True)
    return n_pairs


def noisy_long(rng, s):
    """3 % sub / 2.5 % ins / 4.5 % del, qualities uniform Q5..Q29 (make_fixtures.noisy_long, vectorised)"""
    ln = len(s)
    r = rng.random(ln)
    keep = r >= 0.045
    sub = (r >= 0.045) & (r < 0.075)
    s = s.copy()
    s[sub] = (s[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
    s = s[keep]
    n_ins = rng.geometric(1 - 0.025, len(s)) - 1
    tot = len(s) + int(n_ins.sum())
    out = rng.integers(0, 4, tot, dtype=np.uint8)
    pos = np.cumsum(n_ins + 1) - (n_ins + 1)          # start of each kept base in the output
    out[pos] = s
    q = rng.integers(5, 30, tot).astype(np.uint8) + 33
    return LET[out].tobytes(), q.tobytes()


def write_long_reads(rng, haps, total_bases, path):
    glen = len(haps[0])
    tot = i = 0
    with open(path, "wb") as f:
        while tot < total_bases:
            ln = int(min(max(1000, rng.lognormal(9.0, 0.6)), glen))
            h = haps[int(rng.integers(0, len(haps)))]
            p = int(rng.integers(0, glen - ln + 1))
            s = h[p:p + ln]
            if rng.random() < 0.5:
                s = COMP[s][::-1]
            sq, qq = noisy_long(rng, s)
            f.write(b"@lr%d\n" % i + sq + b"\n+\n" + qq + b"\n")
            tot += len(sq)
            i += 1
    return i, tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--genome-len", type=int, default=64_000_000)
    ap.add_argument("--seed", type=int, default=6400)
    ap.add_argument("--short-cov", type=float, default=30.0)
    ap.add_argument("--long-bases", type=float, default=5.0, help="long-read sample, in multiples of the genome length")
    a = ap.parse_args()
    os.makedirs(a.out_dir, exist_ok=True)
    rng = np.random.default_rng(a.seed)
    n = a.genome_len
    scale = n / 300000.0
    g = genome_complex(rng, n, int(12 * scale), int(40 * scale), int(200 * scale))
    snp = np.nonzero(rng.random(n) < 0.001)[0]
    h1 = g.copy()
    h1[snp] = (g[snp] + rng.integers(1, 4, len(snp), dtype=np.uint8)) & 3
    pad = (-n) % 4
    gp = np.concatenate([g, np.zeros(pad, np.uint8)])
    packed = (gp[0::4] << 6) | (gp[1::4] << 4) | (gp[2::4] << 2) | gp[3::4]
    np.savez_compressed(os.path.join(a.out_dir, "genome.npz"), n=n, hap0_2bit=packed.astype(np.uint8), snp_pos=snp.astype(np.int64), snp_base=h1[snp])
    print("genome written: %d bp, %d het SNPs" % (n, len(snp)), file=sys.stderr, flush=True)
    haps = [g, h1]
    n_long, tot = write_long_reads(rng, haps, a.long_bases * n, os.path.join(a.out_dir, "lr_sample.fastq"))
    print("long reads: %d reads, %d bases" % (n_long, tot), file=sys.stderr, flush=True)
    npairs = write_short_reads(rng, haps, a.short_cov, 150, 350, 450, 0.002, os.path.join(a.out_dir, "sr.fastq"))
    print("recipe=F4 genome=%d pairs=%d long_reads=%d long_bases=%d" % (n, npairs, n_long, tot))


if __name__ == "__main__":
    main()
