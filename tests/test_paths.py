"""explorePathsBFS2 / explorePathsBFS parity (host orchestration over the K2/K3/K4/K5 kernels) against vectors recorded
from the unmodified reference (tests/golden/make_golden_paths.py): winning path vertices and per-base qualities."""
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


def _check(ctx, cases):
    bad = []
    for i, c in enumerate(cases):
        r = ctx.explore_paths(tuple(c["start"]), tuple(c["end"]) if c["end"] else None, c["ref"], c["pids"])
        exp = c["paths"][0] if c["paths"] else None
        if exp is None:
            ok = r is None
        else:
            ok = r is not None and r["nodes"] == [tuple(u) for u in exp["um"]] and r["qual"] == exp["qual"]
        if not ok:
            bad.append(i)
    assert not bad, bad


def test_path_search_kernel_sources_match_reference(sim_lib):
    recipe, cases = _cases()
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, 31, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    _check(ctx, cases)
    ctx.close()
    g.close()


@pytest.mark.gpu
def test_path_search_cuda_matches_reference_golden():
    recipe, cases = _cases()
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    _check(ctx, cases)
    ctx.close()
    g.close()
