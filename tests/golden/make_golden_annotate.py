#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — golden vectors for detectSNPs / detectShortCycles under options the stored indexes do not cover,
recorded from the UNMODIFIED reference through the seam probe ref_annotate (oracle/ref_seams.cpp): the annotations of the
loaded index are cleared and both functions re-run with min_cov_vertices in {1, 2, 3, 5}, single- and multi-threaded (the
two branches of src/Graph.cpp:498 / :577 and :4740 / :4774).  Run in the build container only.

  tests/golden/annotate_vectors.json.gz   {fixture: {k: {min_cov: {unitig id: [ambiguity ids, flag, blob as latin-1]}}}}
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ratatosk_b200 as rb  # noqa: E402  (unitig ids = order of the index FASTA, read through the simulator library)
from refseams import RefGraph  # noqa: E402

SIM = os.path.join(ROOT, "tests", "hostsim", "_build", "librtk_hostsim.so")
out = {}
for fx in ("F1", "F2"):
    out[fx] = {}
    for k in (31, 63):
        d = os.path.join(HERE, fx)
        fa, rt = os.path.join(d, "index.k%d.fasta.gz" % k), os.path.join(d, "index.k%d.rtsk" % k)
        g = rb.Graph.load(fa, rt, k, lib=SIM)
        ids = {g.unitig_seq(u): u for u in range(g.info()["n_unitigs"])}
        ref = RefGraph(fa, rt, k)
        out[fx][str(k)] = {}
        for mc in (1, 2, 3, 5):
            a1 = ref.annotate(mc, threads=1)
            a4 = ref.annotate(mc, threads=4)
            assert a1 == a4, "single- and multi-thread branches of the reference disagree"
            out[fx][str(k)][str(mc)] = {str(ids[s]): [v[0], v[1], v[2].decode("latin1")] for s, v in a1.items()}
            print(fx, k, "min_cov", mc, "unitigs annotated", len(a1), "marks", sum(len(v[0]) for v in a1.values()),
                  "cycle unitigs", sum(v[1] for v in a1.values()))
        if True:   # min_cov = 2 is what the index stores
            stored = {}
            for u in range(g.info()["n_unitigs"]):
                amb, blob = g.unitig_annotations(u)
                flag = (g.unitig_words(u)[1] >> 8) & 1
                if amb or blob or flag:
                    stored[str(u)] = [amb, flag, blob.decode("latin1")]
            assert stored == out[fx][str(k)]["2"], "re-run with the index's own min_cov differs from what the index stores"
        ref.close()
        g.close()
with gzip.GzipFile(os.path.join(HERE, "annotate_vectors.json.gz"), "wb", 9, mtime=0) as f:
    f.write(json.dumps(out, sort_keys=True).encode())
