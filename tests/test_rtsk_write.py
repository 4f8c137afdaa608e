"""The .rtsk writer (rtk_rtsk_write / rtk_graph_recolor: writeGraphData src/Graph.cpp:786-801, UnitigData::write src/UnitigData.hpp:493-517,
PairID::write src/PairID.cpp:1135-1172) on synthetic colourings that force every PairID container kind the writer emits: 61-bit
vector (ids < 61, and the empty set), single id, CRoaring portable blob with array containers (several 64 Ki chunks) and with a
bitset container (> 4096 ids in one chunk).  Host code only (no kernel): the file is read back by this library's parser and, where
the seam library is built, by the REFERENCE's readGraphData (per-unitig dump)."""
import os

import numpy as np
import pytest

import ratatosk_b200 as rb
import refseams
from common import GOLDEN


def _sets(n):
    rng = np.random.RandomState(7)
    sets = [[] for _ in range(n)]
    sets[0] = list(range(3, 10003))                                          # bitset container (10000 ids in chunk 0)
    sets[1] = [5, 70000, 70001, 140000, 1 << 20, (1 << 24) + 17]             # array containers in five chunks
    sets[2] = [100]                                                          # single id
    sets[3] = [0, 7, 60]                                                     # 61-bit vector
    sets[4] = []                                                             # empty
    sets[5] = [61]                                                           # just past the vector
    sets[6] = sorted(set(int(x) for x in rng.randint(0, 300000, 6000)))      # mixed: arrays, dense-ish
    sets[7] = list(range(65536 - 2, 65536 + 5000))                           # straddles a chunk border, bitset + array
    for u in range(8, n):
        sets[u] = sorted(set(int(x) for x in rng.randint(0, 2000, rng.randint(0, 6))))
    return sets


def test_rtsk_writer_container_kinds_roundtrip_and_reference_reader(sim_lib, tmp_path):
    d = os.path.join(GOLDEN, "F2")
    fa = os.path.join(d, "index.k63.fasta.gz")
    g = rb.Graph.load(fa, "", 63, lib=sim_lib)
    n = g.info()["n_unitigs"]
    sets = _sets(n)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in sets])
    ids = np.array([x for s in sets for x in s], dtype=np.uint32)
    kmcov = (np.arange(n, dtype=np.uint64) * np.uint64(977)) << np.uint64(31)
    kmcov[::3] |= np.uint64(1) << np.uint64(63)
    shared = np.arange(n, dtype=np.uint64) & np.uint64(0xff)
    amb = [[] for _ in range(n)]
    amb[0] = [(5 << 4) | 3]
    amb[1] = [(2 << 4) | 5, (30 << 4) | 10, (60 << 4) | 12]       # several ids: roaring array container
    amb[9] = [(1 << 4) | 3]                                                  # id 19 < 61: vector kind
    aoff = np.zeros(n + 1, dtype=np.uint64)
    aoff[1:] = np.cumsum([len(a) for a in amb])
    aids = np.array([x for a in amb for x in a], dtype=np.uint32)
    cyc = [b"" for _ in range(n)]
    cyc[2] = b"\0"
    cyc[3] = b"ACG\0T\0"
    coff = np.zeros(n + 1, dtype=np.uint64)
    coff[1:] = np.cumsum([len(c) for c in cyc])
    is_cycle = np.array([1 if c else 0 for c in cyc], dtype=np.uint8)
    g2 = g.recolor(kmcov, shared, off, ids)
    path = str(tmp_path / "synthetic.rtsk")
    g2.write_rtsk(path, aoff, aids, is_cycle, coff, b"".join(cyc))
    # this library's parser
    back = rb.Graph.load(fa, path, 63, lib=sim_lib)
    for u in range(n):
        a, b = back.unitig_colors(u)
        assert sorted(a + b) == sets[u], u
        kc, sh, _ = back.unitig_words(u)
        assert kc == int(kmcov[u]) and sh == (int(shared[u]) | (0x100 if cyc[u] else 0)), u
        assert back.unitig_annotations(u) == (amb[u], cyc[u]), u
    # the reference's reader
    if refseams.available():
        ref = refseams.RefGraph(fa, path, 63)
        dump = str(tmp_path / "dump.tsv")
        ref.dump(dump)
        ref.close()
        by_seq = {g.unitig_seq(u): u for u in range(n)}
        seen = 0
        for line in open(dump):
            f = line.rstrip("\n").split("\t")
            u = by_seq[f[1]]
            assert int(f[2]) == int(kmcov[u]) and int(f[3]) == (int(shared[u]) | (0x100 if cyc[u] else 0)), u
            assert sorted(int(x) for x in (f[4] + f[5]).split(",") if x) == sets[u], u
            assert len([x for x in f[6].split(",") if x]) == len(amb[u]), u
            assert f[8] == "".join(c.decode() + ";" for c in cyc[u].split(b"\0")[:-1]), u
            seen += 1
        assert seen == n
    for x in (g, g2, back):
        x.close()
