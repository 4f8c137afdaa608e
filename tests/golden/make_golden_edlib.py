#!/usr/bin/env python3
"""Golden vectors for K4 from the UNMODIFIED reference edlib (src/edlib.cpp via oracle/_ref/libref_seams.so),
configured as every hot-path call site does (IUPAC equalities, src/Common.hpp:262-276).
Seeded (`micro-myers` recipe of SURVEY.md §8d, scaled down): target ~ U[1,700], query = target mutated at
0-15% (sub:ins:del = 3:2.5:4.5), 1% IUPAC letters, modes NW/SHW/HW = 1:1:2, plus hand-picked edge cases."""
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


def sprinkle(rng, s, p):
    return "".join(rng.choice(IUPAC) if rng.random() < p else c for c in s)


def main():
    rng = random.Random(0x4D594552)
    cases = []
    edge = [("", "", 0), ("", "ACGT", 0), ("ACGT", "", 1), ("", "ACGT", 2), ("A", "C", 1), ("A", "C", 2), ("A", "A", 0),
            ("ACGT", "ACGT", 0), ("ACGT", "TTACGTTT", 2), ("ACGT", "TTACGTTTACGT", 2), ("AAAA", "AAAAAAAA", 1),
            ("NNNN", "ACGT", 0), ("ACGT", "NNNN", 0), ("RYKM", "ACGT", 0), ("N", "N", 0), ("R", "Y", 0), ("R", "N", 0),
            ("A" * 64, "A" * 64, 0), ("A" * 65, "A" * 64, 0), ("A" * 64, "A" * 65, 1), ("ACGT" * 40, "ACGT" * 41, 2),
            ("ACGT" * 16, "TGCA" * 16, 0), ("ACGT" * 16, "TGCA" * 16, 2), ("A" * 128, "C" * 10, 1), ("A" * 129, "A" * 300, 2)]
    for q, t, m in edge:
        cases.append((q, t, m, -1))
    for i in range(420):
        tl = rng.randint(1, 700) if i % 7 else rng.randint(1500, 2600)
        t = "".join(rng.choice("ACGT") for _ in range(tl))
        q = mutate(rng, t, rng.choice([0.0, 0.02, 0.05, 0.1, 0.15]))
        if i % 5 == 0:
            q = q[:max(1, len(q) * 3 // 4)]
        if i % 9 == 0:
            t = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 50))) + t
        q = sprinkle(rng, q, 0.01)
        t = sprinkle(rng, t, 0.01 if i % 3 == 0 else 0.0)
        mode = rng.choice([0, 1, 2, 2])
        k = -1
        if i % 4 == 0:
            d, _, _, _ = R.edlib(q, t, mode, 0, -1, True)
            k = max(0, d + rng.choice([-3, -1, 0, 1, 10]))
        cases.append((q, t, mode, k))
    out = []
    for q, t, mode, k in cases:
        d, ends, _, _ = R.edlib(q, t, mode, 0, k, True)
        out.append({"q": q, "t": t, "mode": mode, "k": k, "dist": d, "ends": ends})
    with gzip.open(os.path.join(HERE, "edlib_vectors.json.gz"), "wt") as f:
        json.dump(out, f)
    print(len(out), "vectors;", sum(1 for c in out if c["dist"] < 0), "above k")


if __name__ == "__main__":
    main()
