#!/usr/bin/env python
"""DEVELOPMENT AID: detectSNPs + detectShortCycles of the product library on a reference-built index, timed, and checked against
what that index stores.  Usage: annotate_quick.py <dir with index.k<K>.fasta.gz/.rtsk> <K> [lib]  -> one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ratatosk_b200 as rb

d, k = sys.argv[1], int(sys.argv[2])
lib = sys.argv[3] if len(sys.argv) > 3 else None
t = time.time()
g = rb.Graph.load(os.path.join(d, "index.k%d.fasta.gz" % k), os.path.join(d, "index.k%d.rtsk" % k), k, lib=lib)
t_load = time.time() - t
ctx = rb.Context(0, lib=lib)
ctx.upload(g)
info = g.info()
n = info["n_unitigs"]
out = {"dir": d, "k": k, "n_unitigs": n, "n_kmers": info["n_kmers"], "slab_bytes": info["slab_bytes"], "load_s": round(t_load, 2)}
for rep in range(2):   # second repetition: buffers grown, code resident
    st1, st2 = [0] * 10, [0] * 10
    t = time.time(); off, ids = ctx.detect_snps(stats=st1); t_snp = time.time() - t
    t = time.time(); flags, coff, pool = ctx.detect_short_cycles(stats=st2); t_cyc = time.time() - t
out.update({"detect_snps_s": round(t_snp, 3), "k1_probes": st1[0], "k1_kernel_ms": st1[2] / 1e6, "k1_stage_ms": st1[3] / 1e6,
            "candidates": st1[4], "unitigs_with_candidates": st1[5], "traversals": st1[6], "snp_kernel_ms": st1[7] / 1e6,
            "snp_rerun": st1[8], "marks": int(len(ids)), "detect_short_cycles_s": round(t_cyc, 3), "cycles_kernel_ms": st2[7] / 1e6,
            "cycle_unitigs": int(flags.sum()), "cycles_rerun": st2[8]})
if st1[2]:
    out["k1_lookups_per_s"] = st1[0] / (st1[2] / 1e9)
# parity against the stored annotations, over the slab sections directly (fast)
slab = g.slab()
hdr = np.frombuffer(slab[:512].tobytes(), dtype=np.uint64)
names = ["off_unitig_off", "off_pool", "off_table", "off_blk2unitig", "off_kmcov", "off_shared", "off_adj", "off_gset_of", "off_gset_off",
         "off_gset_ids", "off_loc_off", "off_loc_ids", "off_amb_off", "off_amb_ids", "off_hap_off", "off_hap_ids", "off_cyc_off", "off_cyc_pool"]
o = {nm: int(hdr[10 + i]) for i, nm in enumerate(names)}
w_amb_off = np.frombuffer(slab[o["off_amb_off"]:o["off_amb_off"] + 8 * (n + 1)].tobytes(), dtype=np.uint64)
w_amb = np.frombuffer(slab[o["off_amb_ids"]:o["off_amb_ids"] + 4 * int(w_amb_off[-1])].tobytes(), dtype=np.uint32)
w_cyc_off = np.frombuffer(slab[o["off_cyc_off"]:o["off_cyc_off"] + 8 * (n + 1)].tobytes(), dtype=np.uint64)
w_cyc = slab[o["off_cyc_pool"]:o["off_cyc_pool"] + int(w_cyc_off[-1])].tobytes()
w_shared = np.frombuffer(slab[o["off_shared"]:o["off_shared"] + 8 * n].tobytes(), dtype=np.uint64)
out["snps_identical"] = bool(np.array_equal(w_amb_off, off) and np.array_equal(w_amb, ids))
out["cycles_identical"] = bool(np.array_equal(w_cyc_off, coff) and w_cyc == pool and np.array_equal(((w_shared >> np.uint64(8)) & np.uint64(1)).astype(np.uint8), flags))
print(json.dumps(out))
