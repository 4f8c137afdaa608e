#!/usr/bin/env python3
"""Golden vectors for the anchor-to-anchor path search: the reference's explorePathsBFS2 / explorePathsBFS
(src/GraphTraversal.cpp:212-454, :3-210) through oracle/_ref/libref_seams.so on seeded synthetic calls over the
committed F2 index.  A call = two anchor k-mers a random walk apart (or one anchor + open end), ref = the true
spelling between them mutated at 0-12 %, colour set = colours seen along the walk / empty."""
import gzip, json, os, random, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, ".."))
import refseams as R  # noqa: E402
import ratatosk_b200 as rb  # noqa: E402
from common import SIM_LIB, golden_paths  # noqa: E402
from make_golden_subgraph import mutate, rc, K  # noqa: E402


def main():
    recipe = "F2"
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, K, lib=SIM_LIB)
    n = g.info()["n_unitigs"]
    seqs = [g.unitig_seq(u) for u in range(n)]
    words = [g.unitig_words(u) for u in range(n)]
    adj = [w[2] for w in words]
    rg = R.RefGraph(fa, rt, K)
    dump = "/tmp/rtk_paths_dump.txt"
    rg.dump(dump)
    seq2id = {s: i for i, s in enumerate(seqs)}
    id2key, key2id = {}, {}
    for line in open(dump):
        c = line.split("\t")
        u = seq2id[c[1]]
        id2key[u] = int(c[0]); key2id[int(c[0])] = u

    def succs(u, strand):
        out = []
        for c in range(4):
            v = adj[u][c] if strand else adj[u][4 + (3 - c)]
            if v == 0xFFFFFFFF:
                continue
            # follow only edges whose flag is set (UnitigData::getSharedPids), like the traversal does
            bit = (1 << c) << 4 if strand else (1 << c)
            if not (words[u][1] & bit):
                continue
            vs = (v >> 31) & 1
            out.append((v & 0x7FFFFFFF, vs if strand else 1 - vs))
        return out

    rng = random.Random(20261019)
    cases = []
    tries = 0
    while len(cases) < 90 and tries < 20000:
        tries += 1
        u0, s0 = rng.randrange(n), rng.randrange(2)
        walk = [(u0, s0)]
        for _ in range(rng.randint(0, 9)):
            nx = succs(*walk[-1])
            if not nx:
                break
            walk.append(rng.choice(nx))
        o0 = seqs[u0] if s0 else rc(seqs[u0])
        a0 = rng.randint(0, len(o0) - K)                     # start anchor offset in traversal orientation
        open_end = rng.random() < 0.3
        if len(walk) == 1:
            a1 = rng.randint(a0, len(o0) - K)
            true = o0[a0:a1 + K]
        else:
            true = o0[a0:]
            for (u, s) in walk[1:-1]:
                o = seqs[u] if s else rc(seqs[u])
                true += o[K - 1:]
            ue, se = walk[-1]
            oe = seqs[ue] if se else rc(seqs[ue])
            a1 = rng.randint(0, len(oe) - K)
            true += oe[K - 1:a1 + K]
        if len(true) < K + 20 or len(true) > 1000:
            continue
        ue, se = walk[-1]
        oe = seqs[ue] if se else rc(seqs[ue])
        start = (u0, s0, a0 if s0 else len(o0) - K - a0)
        end = None if open_end else (ue, se, a1 if se else len(oe) - K - a1)
        ref = true[:K] + mutate(rng, true[K:-K], rng.choice([0.0, 0.04, 0.08, 0.12])) + (true[-K:] if len(true) >= 2 * K else "")
        if len(ref) <= K + 5:
            continue
        pids = set()
        if rng.random() < 0.8:
            for (u, s) in walk:
                gi, li = g.unitig_colors(u)
                ids = gi + li
                rng.shuffle(ids)
                pids.update(ids[:30])
        paths = rg.explore_paths((id2key[start[0]], start[1], start[2]), None if end is None else (id2key[end[0]], end[1], end[2]),
                                 ref, sorted(pids))
        conv = [{"um": [[key2id[k_], st, d, l] for (k_, st, d, l) in ums], "qual": q} for ums, q in paths]
        cases.append({"start": list(start), "end": list(end) if end else None, "ref": ref, "pids": sorted(pids), "paths": conv,
                      "hops": len(walk) - 1})
    with gzip.open(os.path.join(HERE, "paths_vectors.json.gz"), "wt") as f:
        json.dump({"recipe": recipe, "k": K, "cases": cases}, f)
    print(len(cases), "calls;", sum(1 for c in cases if c["paths"]), "found a path;", sum(1 for c in cases if c["end"] is None), "open-ended;",
          sum(1 for c in cases if c["paths"] and len(c["paths"][0]["um"]) > 1), "multi-unitig")


if __name__ == "__main__":
    main()
