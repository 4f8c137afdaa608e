#!/usr/bin/env python3
"""Golden vectors for K5 (alignment path / traceback): the reference's edlibAlign with EDLIB_TASK_PATH
(src/edlib.cpp:141-296, obtainAlignment :1164, obtainAlignmentTraceback :945), modes NW and SHW with the IUPAC
equalities, via oracle/_ref/libref_seams.so.  Alignment ops: 0 match, 1 insert (query base unaligned), 2 delete
(target base unaligned), 3 mismatch.  Cases whose traceback state would exceed edlib's 1 MiB switch to Hirschberg
(:1191-1193) are tagged `hirschberg`."""
import gzip, json, os, random, sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import refseams as R  # noqa: E402

IUPAC = "MRSVWYHKDBN"


def mutate(rng, s, rate):
    out = []
    for c in s:
        r = rng.random()
        if r < rate * 0.45:
            continue
        if r < rate * 0.75:
            c = rng.choice("ACGT")
        out.append(c)
        if rng.random() < rate * 0.25:
            out.append(rng.choice("ACGT"))
    return "".join(out)


def main():
    rng = random.Random(0x50415448)
    cases = [("", "", 0), ("", "ACGT", 0), ("ACGT", "", 0), ("ACGT", "", 1), ("A", "C", 0), ("A", "C", 1), ("A", "A", 0), ("ACGT", "ACGT", 0),
             ("ACGT", "AGT", 0), ("AGT", "ACGT", 0), ("AAAA", "AAAAAAAA", 1), ("AAAAAAAA", "AAAA", 0), ("ACGTACGT", "TTTT", 0),
             ("NNNN", "ACGT", 0), ("RYKM", "ACGT", 0), ("A" * 64, "A" * 64, 0), ("A" * 65, "A" * 64, 0), ("A" * 64, "A" * 65, 1),
             ("ACGT" * 33, "ACGT" * 32 + "TTTT", 0), ("GATTACA" * 20, "GATACA" * 20, 0), ("GATTACA" * 20, "GATACA" * 25, 1),
             ("AC" * 100, "CA" * 100, 0), ("AC" * 100, "CA" * 100, 1)]
    for i in range(260):
        tl = rng.choice([rng.randint(1, 120), rng.randint(120, 700), rng.randint(700, 1400)])
        if i % 40 == 39:
            tl = rng.randint(2500, 3200)  # beyond edlib's 1 MiB traceback limit -> Hirschberg
        t = "".join(rng.choice("ACGT") for _ in range(tl))
        q = mutate(rng, t, rng.choice([0.0, 0.03, 0.08, 0.15]))
        if i % 6 == 0:
            q = q[:max(1, len(q) * 2 // 3)]
        if i % 7 == 0:
            q = "".join(rng.choice(IUPAC) if rng.random() < 0.02 else c for c in q)
        if i % 11 == 0:
            t = "".join(rng.choice(IUPAC) if rng.random() < 0.01 else c for c in t)
        if i % 13 == 0:  # low-complexity: many equally optimal alignments -> exercises the move priorities
            t = "".join(rng.choice("AC") for _ in range(tl))
            q = mutate(rng, t, 0.1)
        cases.append((q or "A", t, rng.choice([0, 1])))
    out = []
    for q, t, mode in cases:
        d, ends, starts, aln = R.edlib(q, t, mode, 2, -1, True)
        nb = (len(q) + 63) // 64
        tl = (ends[0] + 1) if (mode == 1 and ends) else len(t)
        hirsch = (20 * nb * tl + 8 * tl) >= (1 << 20)
        out.append({"q": q, "t": t, "mode": mode, "dist": d, "end": ends[0] if ends else None, "aln": list(aln), "hirschberg": hirsch})
    with gzip.open(os.path.join(HERE, "edlib_path_vectors.json.gz"), "wt") as f:
        json.dump(out, f)
    print(len(out), "vectors;", sum(1 for c in out if c["hirschberg"]), "hirschberg-sized")


if __name__ == "__main__":
    main()
