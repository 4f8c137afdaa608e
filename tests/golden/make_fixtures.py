#!/usr/bin/env python3
"""Seeded synthetic fixtures for the Ratatosk correction hot path.

TEST INFRASTRUCTURE.  Writes a genome, paired-end short reads (mates share a FASTQ
name, reference README "Before starting") and ONT-like long reads.  Everything is a
pure function of (--recipe, --seed), Python `random` only, so the files can be
regenerated anywhere; the *index* built from them by the reference is NOT
reproducible (std::random_device in src/Graph.cpp:2093,2327) which is why the
index files themselves are committed next to the golden outputs (see
make_golden.sh).

Recipes
  F1  60 kb uniform random genome, 30x PE150 (insert 400, 0.1% subs),
      100 long reads 3-12 kb, 3% sub / 3% ins / 4% del, Q5-29          seed 12345
  F2  300 kb diploid (0.1% het SNPs), 12 repeat families x3-7 copies
      (300-2500 bp, 2% divergence), 40 tandem repeats, 200 homopolymers,
      30x PE150 (0.2% subs, insert 350-450), 10x long reads
      lognormal(9.0,0.6) >= 1 kb, 10% error                           seed 777
  F3  F2 recipe scaled to an E. coli-like 4.64 Mbp genome, 30x/30x      seed 4640
"""
import argparse, math, random, sys

ACGT = "ACGT"
COMP = str.maketrans("ACGT", "TGCA")

def revcomp(s):
    return s.translate(COMP)[::-1]

def rand_seq(rng, n):
    return "".join(rng.choice(ACGT) for _ in range(n))

def mutate_subs(rng, s, rate):
    if rate <= 0: return s
    out = list(s)
    n = len(out)
    # number of substitutions ~ binomial approximated by per-base draw (kept simple & deterministic)
    for i in range(n):
        if rng.random() < rate:
            c = out[i]
            out[i] = rng.choice([x for x in ACGT if x != c])
    return "".join(out)

def genome_plain(rng, n):
    return rand_seq(rng, n)

def genome_complex(rng, n, n_fam=12, n_tandem=40, n_homo=200):
    g = list(rand_seq(rng, n))
    # repeat families
    for _ in range(n_fam):
        ln = rng.randint(300, 2500)
        unit = rand_seq(rng, ln)
        for _c in range(rng.randint(3, 7)):
            pos = rng.randint(0, n - ln - 1)
            cp = mutate_subs(rng, unit, 0.02)
            if rng.random() < 0.5: cp = revcomp(cp)
            g[pos:pos + ln] = cp
    for _ in range(n_tandem):
        ul = rng.randint(2, 40)
        unit = rand_seq(rng, ul)
        cnt = rng.randint(3, 30)
        t = (unit * cnt)[:1200]
        pos = rng.randint(0, n - len(t) - 1)
        g[pos:pos + len(t)] = t
    for _ in range(n_homo):
        ln = rng.randint(6, 25)
        pos = rng.randint(0, n - ln - 1)
        g[pos:pos + ln] = rng.choice(ACGT) * ln
    return "".join(g)

def write_short_reads(rng, haps, cov, rl, ins_lo, ins_hi, sub, path):
    glen = len(haps[0])
    n_pairs = int(cov * glen / (2 * rl))
    with open(path, "w") as f:
        for i in range(n_pairs):
            h = haps[rng.randrange(len(haps))]
            frag = rng.randint(ins_lo, ins_hi)
            p = rng.randint(0, len(h) - frag)
            fr = h[p:p + frag]
            if rng.random() < 0.5: fr = revcomp(fr)
            r1 = mutate_subs(rng, fr[:rl], sub)
            r2 = mutate_subs(rng, revcomp(fr)[:rl], sub)
            q = "I" * rl
            f.write("@sr%d\n%s\n+\n%s\n@sr%d\n%s\n+\n%s\n" % (i, r1, q, i, r2, q))
    return n_pairs

def noisy_long(rng, s, psub, pins, pdel, qlo, qhi):
    out = []; q = []
    for c in s:
        r = rng.random()
        if r < pdel:
            continue
        if r < pdel + psub:
            c = rng.choice([x for x in ACGT if x != c])
        out.append(c); q.append(chr(33 + rng.randint(qlo, qhi)))
        while rng.random() < pins:
            out.append(rng.choice(ACGT)); q.append(chr(33 + rng.randint(qlo, qhi)))
    return "".join(out), "".join(q)

def write_long_reads(rng, haps, n_reads, cov, len_fn, err, path):
    psub, pins, pdel = err
    tot = 0; i = 0
    glen = len(haps[0])
    with open(path, "w") as f:
        while (n_reads is not None and i < n_reads) or (n_reads is None and tot < cov * glen):
            h = haps[rng.randrange(len(haps))]
            ln = min(len_fn(rng), len(h))
            p = rng.randint(0, len(h) - ln)
            s = h[p:p + ln]
            if rng.random() < 0.5: s = revcomp(s)
            s, q = noisy_long(rng, s, psub, pins, pdel, 5, 29)
            f.write("@lr%d\n%s\n+\n%s\n" % (i, s, q))
            tot += len(s); i += 1
    return i, tot

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--recipe", required=True, choices=["F1", "F2", "F3"])
    ap.add_argument("--out", required=True, help="output prefix")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--genome-len", type=int, default=None)
    ap.add_argument("--long-cov", type=float, default=None)
    a = ap.parse_args()
    if a.recipe == "F1":
        rng = random.Random(a.seed if a.seed is not None else 12345)
        g = genome_plain(rng, a.genome_len or 60000)
        haps = [g]
        npairs = write_short_reads(rng, haps, 30, 150, 400, 400, 0.001, a.out + ".sr.fastq")
        n, tot = write_long_reads(rng, haps, 100, None, lambda r: r.randint(3000, 12000),
                                  (0.03, 0.03, 0.04), a.out + ".lr.fastq")
    else:
        glen = a.genome_len or (300000 if a.recipe == "F2" else 4640000)
        scale = glen / 300000.0
        rng = random.Random(a.seed if a.seed is not None else (777 if a.recipe == "F2" else 4640))
        g = genome_complex(rng, glen, int(12 * scale), int(40 * scale), int(200 * scale))
        h2 = mutate_subs(rng, g, 0.001)
        haps = [g, h2]
        npairs = write_short_reads(rng, haps, 30, 150, 350, 450, 0.002, a.out + ".sr.fastq")
        cov = a.long_cov if a.long_cov is not None else (10 if a.recipe == "F2" else 30)
        # 10% error split sub:ins:del = 3:2.5:4.5
        n, tot = write_long_reads(rng, haps, None, cov,
                                  lambda r: max(1000, int(r.lognormvariate(9.0, 0.6))),
                                  (0.03, 0.025, 0.045), a.out + ".lr.fastq")
    with open(a.out + ".genome.fasta", "w") as f:
        for i, h in enumerate(haps):
            f.write(">hap%d\n%s\n" % (i, h))
    print("recipe=%s genome=%d pairs=%d long_reads=%d long_bases=%d" % (a.recipe, len(haps[0]), npairs, n, tot))

if __name__ == "__main__":
    main()
