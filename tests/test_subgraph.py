"""K2/K3 (exploreSubGraph) parity against vectors recorded from the unmodified reference
(tests/golden/make_golden_subgraph.py): kept terminal / non-terminal paths in discovery order and best scores."""
import gzip
import json
import os

import pytest

import ratatosk_b200 as rb
from common import GOLDEN, golden_paths


def _cases():
    with gzip.open(os.path.join(GOLDEN, "subgraph_vectors.json.gz"), "rt") as f:
        d = json.load(f)
    return d["recipe"], d["cases"]


def _check(ctx, cases):
    calls = [{"start": tuple(c["start"]), "end": tuple(c["end"]) if c["end"] else None, "ref": c["ref"], "level": c["level"],
              "max_len_path": c["max_len_path"], "pids": c["pids"]} for c in cases]
    res = ctx.explore_subgraph(calls)
    for i, (c, r) in enumerate(zip(cases, res)):
        assert r["scores"][0] == c["score_t"], (i, "terminal score", r["scores"], c["score_t"])
        assert r["scores"][2] == c["score_nt"], (i, "non-terminal score", r["scores"], c["score_nt"])
        assert [p["nodes"] for p in r["terminal"]] == [[tuple(u) for u in p["um"]] for p in c["terminal"]], (i, "terminal paths")
        assert [p["nodes"] for p in r["nonterminal"]] == [[tuple(u) for u in p["um"]] for p in c["nonterminal"]], (i, "non-terminal paths")


def test_subgraph_kernel_source_matches_reference(sim_lib):
    recipe, cases = _cases()
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, 31, lib=sim_lib)
    ctx = rb.Context(0, lib=sim_lib)
    ctx.upload(g)
    _check(ctx, cases)
    ctx.close()
    g.close()


@pytest.mark.gpu
def test_subgraph_cuda_matches_reference_golden():
    recipe, cases = _cases()
    fa, rt = golden_paths(recipe)
    g = rb.Graph.load(fa, rt, 31)
    ctx = rb.Context(0)
    ctx.upload(g)
    _check(ctx, cases)
    # batching is transparent
    one = ctx.explore_subgraph([{"start": tuple(cases[3]["start"]), "end": tuple(cases[3]["end"]) if cases[3]["end"] else None,
                                 "ref": cases[3]["ref"], "level": cases[3]["level"], "max_len_path": cases[3]["max_len_path"],
                                 "pids": cases[3]["pids"]}])
    assert one[0]["scores"][0] == cases[3]["score_t"]
    ctx.close()
    g.close()
