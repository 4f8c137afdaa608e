#!/usr/bin/env python3
"""Golden vectors for K2/K3: the reference's exploreSubGraph (src/GraphTraversal.cpp:456-587) called through
oracle/_ref/libref_seams.so on seeded synthetic calls over the committed F2 index: random walks of 1-6 hops from
a random oriented unitig, ref = the walk's true spelling mutated at 0-12 %, colour set = colours seen along the
walk (subsampled) / empty / random ids, target = last node of the walk at a random offset (or none)."""
import gzip, json, os, random, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, ".."))
import refseams as R  # noqa: E402
import ratatosk_b200 as rb  # noqa: E402
from common import SIM_LIB, golden_paths  # noqa: E402

K = 31
rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))


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
    recipe = "F2"
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, K, lib=SIM_LIB)
    n = g.info()["n_unitigs"]
    seqs = [g.unitig_seq(u) for u in range(n)]
    adj = [g.unitig_words(u)[2] for u in range(n)]
    rg = R.RefGraph(fa, rt, K)
    dump = "/tmp/rtk_subgraph_dump.txt"
    rg.dump(dump)
    seq2id = {}
    for i, s in enumerate(seqs):
        seq2id[s] = i
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
            vs = (v >> 31) & 1
            out.append((v & 0x7FFFFFFF, vs if strand else 1 - vs))
        return out

    rng = random.Random(20261018)
    cases = []
    while len(cases) < 160:
        u0, s0 = rng.randrange(n), rng.randrange(2)
        if not succs(u0, s0):
            continue
        walk = [(u0, s0)]
        for _ in range(rng.randint(1, 6)):
            nx = succs(*walk[-1])
            if not nx:
                break
            walk.append(rng.choice(nx))
        if len(walk) < 2:
            continue
        o0 = seqs[u0] if s0 else rc(seqs[u0])
        keep0 = rng.randint(K, min(len(o0), 220))
        true = o0[len(o0) - keep0:]
        for (u, s) in walk[1:-1]:
            o = seqs[u] if s else rc(seqs[u])
            true += o[K - 1:]
        ue, se = walk[-1]
        oe = seqs[ue] if se else rc(seqs[ue])
        cut = rng.randint(0, len(oe) - K)           # target k-mer offset in traversal orientation
        true += oe[K - 1:cut + K]
        if len(true) > 1400:
            continue
        end_dist = cut if se else (len(oe) - K - cut)
        ref = mutate(rng, true, rng.choice([0.0, 0.03, 0.08, 0.12]))
        if len(ref) <= K:
            continue
        has_end = rng.random() < 0.8
        pids = set()
        mode = rng.random()
        if mode < 0.7:
            for (u, s) in walk:
                gi, li = g.unitig_colors(u)
                ids = gi + li
                rng.shuffle(ids)
                pids.update(ids[:30])
        elif mode < 0.85:
            pids = set(rng.randrange(0, 30000) for _ in range(40))
        level = 3 if rng.random() < 0.8 else rng.randint(1, 2)
        L = len(ref) - K
        max_len = max(int(max(L + L * 0.25, 1.0)), 10) + K
        sc_t, sc_nt, term, nonterm = rg.explore_subgraph(id2key[u0], s0, id2key[ue] if has_end else None, se, end_dist, ref,
                                                         level, max_len, sorted(pids))
        conv = lambda paths: [{"um": [[key2id[k_], st, d, l] for (k_, st, d, l) in ums], "qual": q} for ums, q in paths]
        cases.append({"start": [u0, s0], "end": [ue, se, end_dist] if has_end else None, "ref": ref, "level": level,
                      "max_len_path": max_len, "pids": sorted(pids), "score_t": sc_t, "score_nt": sc_nt,
                      "terminal": conv(term), "nonterminal": conv(nonterm)})
    with gzip.open(os.path.join(HERE, "subgraph_vectors.json.gz"), "wt") as f:
        json.dump({"recipe": recipe, "k": K, "cases": cases}, f)
    nt = sum(1 for c in cases if c["terminal"])
    print(len(cases), "calls;", nt, "with terminal paths;", sum(len(c["nonterminal"]) for c in cases), "non-terminal paths;",
          "mean term score %.3f" % (sum(c["score_t"] for c in cases) / len(cases)))


if __name__ == "__main__":
    main()
