"""Device-resident region engine (ratatosk_b200/csrc/region.cuh, C ABI rtk_region_paths_batch) against the path vectors
recorded from the unmodified reference (tests/golden/make_golden_paths.py: explorePathsBFS2 / explorePathsBFS winners with
their per-base qualities, fixRepeats included).  A golden case is one hop; as a region it is a call without weak anchors whose
window is the hop's read window, so the engine's chain (start k-mer merged with the hop's path) must equal the hop's path."""
import gzip
import json
import os

import pytest

import ratatosk_b200 as rb
from common import GOLDEN, golden_paths


def _cases():
    with gzip.open(os.path.join(GOLDEN, "paths_vectors.json.gz"), "rt") as f:
        d = json.load(f)
    return d["recipe"], d["cases"]


def _calls(cases, k=31):
    calls = []
    for c in cases:
        ref = c["ref"]
        call = {"window": ref, "start": (0, c["start"][0], c["start"][1], c["start"][2]), "pids": c["pids"], "weak": []}
        if c["end"]:
            call["end"] = (len(ref) - k, c["end"][0], c["end"][1], c["end"][2])
        else:
            call["end"] = None
            call["s_len"] = len(ref)
        calls.append(call)
    return calls


def _check(ctx, cases, lib=None):
    opt = rb.default_opt(1, lib=lib)
    opt.max_len_weak_region1 = 1000000   # the vectors were recorded by calling explorePathsBFS* directly, whatever the window length
    res = ctx.region_paths(_calls(cases), opt=opt)
    bad, declined = [], 0
    for i, (c, r) in enumerate(zip(cases, res)):
        exp = c["paths"][0] if c["paths"] else None
        if r["status"] == 2:
            declined += 1
            continue
        if exp is None:
            ok = r["status"] == 1
        else:
            ok = r["status"] == 0 and r["nodes"] == [tuple(u) for u in exp["um"]] and r["qual"] == exp["qual"] and len(r["seq"]) == len(exp["qual"])
        if not ok:
            bad.append(i)
    assert not bad, bad
    assert declined <= len(cases) // 10, declined   # the engine may decline rare shapes, not the bulk of them


def test_region_engine_kernel_source_matches_reference(sim_lib):
    recipe, cases = _cases()
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, 31, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    _check(ctx, cases, lib=sim_lib)
    ctx.close()
    g.close()


def test_region_engine_declines_instead_of_overflowing(sim_lib, monkeypatch):
    """scratch too small for almost any region: calls come back declined (status 2) - or, if they happen to fit, right"""
    recipe, cases = _cases()
    fa, rt = golden_paths(recipe)
    monkeypatch.setenv("RTK_RG_ARENA_CAP", "256")
    g = rb.Graph.load(fa, rt, 31, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    opt = rb.default_opt(1, lib=sim_lib)
    opt.max_len_weak_region1 = 1000000
    res = ctx.region_paths(_calls(cases[:12]), opt=opt)
    assert sum(1 for r in res if r["status"] == 2 and r["bail"] != 0) >= 8, [(r["status"], r["bail"]) for r in res]
    for c, r in zip(cases[:12], res):
        if r["status"] != 2 and c["paths"]:
            assert r["nodes"] == [tuple(u) for u in c["paths"][0]["um"]] and r["qual"] == c["paths"][0]["qual"]
    ctx.close()
    g.close()


@pytest.mark.gpu
def test_region_engine_cuda_matches_reference_golden():
    recipe, cases = _cases()
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    _check(ctx, cases)
    # batching is transparent: many copies of the cases in one launch give the same answers
    opt = rb.default_opt(1)
    opt.max_len_weak_region1 = 1000000
    one = ctx.region_paths(_calls(cases), opt=opt)
    many = ctx.region_paths(_calls(cases) * 8, opt=opt)
    for i in range(len(many)):
        a, b = one[i % len(cases)], many[i]
        assert (a["status"], a["nodes"], a["qual"]) == (b["status"], b["nodes"], b["qual"]), i
    ctx.close()
    g.close()


def test_engine_divide_and_conquer_traceback_matches_the_k5_driver(sim_lib, tmp_path):
    """edlib switches from the direct traceback to its divide-and-conquer at 1 MiB of alignment state (src/edlib.cpp:1191-1193);
    the two give different co-optimal paths, so the switch is part of the output.  The golden reads only reach it in pass 2, so
    here RTK_TB_LIMIT lowers the switch until every path-quality alignment is split several levels deep, and the device engine
    (region.cuh: rg_nw_quality) must produce the same corrected reads as the request path, whose K5 host driver
    (traceback_host.hpp) is pinned by the golden PATH vectors.  Separate processes: the limit is read once."""
    import pickle
    import subprocess
    import sys
    script = r'''
import os, pickle, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ratatosk_b200 as rb
from common import GOLDEN, load_golden_reads
lib = %r
g = rb.Graph.load(os.path.join(GOLDEN, "F2", "index.k31.fasta.gz"), os.path.join(GOLDEN, "F2", "index.k31.rtsk"), 31, lib=lib)
ctx = rb.Context(0, lib=lib); ctx.upload(g)
reads = load_golden_reads("F2")[:10]
pickle.dump(ctx.correct([s for _, s, _ in reads], [q for _, _, q in reads]), open(sys.argv[1], "wb"))
''' % (os.path.join(os.path.dirname(__file__), ".."), os.path.dirname(__file__), sim_lib)
    outs = []
    for tag, extra in (("engine", {}), ("requests", {"RTK_NO_REGION_ENGINE": "1"}), ("normal", {"RTK_TB_LIMIT": str(1 << 20)})):
        env = dict(os.environ, RTK_TB_LIMIT="3000")
        env.update(extra)
        dst = str(tmp_path / (tag + ".pkl"))
        subprocess.check_call([sys.executable, "-c", script, dst], env=env)
        outs.append(pickle.load(open(dst, "rb")))
    assert outs[0] == outs[1]
    assert outs[0] != outs[2]   # the lowered switch does change co-optimal choices: the comparison above is not vacuous
